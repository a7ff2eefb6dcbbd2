// Fused activation passes of the refiner's ConvBlock for the TRAINING step:
//
//   forward   y  = row padded ? 0 : aptx(u; beta, gamma) [+ res]            (preencoder.py:97-101)
//   backward  du = row padded ? 0 : dy * aptx'(u),   dres = row padded ? 0 : dy
//
// u is the fp32 convolution output (kept for the backward pass: APTx is not invertible), y / res / dy / du /
// dres are bf16 so the next tcgen05 convolution (forward, data- or weight-gradient) consumes them as they
// are.  One pass each instead of the ~15 element-wise PyTorch kernels autograd would run per activation;
// HBM-bound, 8 channels (16-byte bf16 / two 16-byte fp32 accesses) per thread.
#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

template <bool kBackward>
__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ u, const uint4* __restrict__ in_bf16,
                                                  const uint8_t* __restrict__ row_mask, long long groups, int cgroups,
                                                  int pix_per_row, float beta, float gamma, uint4* __restrict__ out,
                                                  uint4* __restrict__ dres) {
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // group of 8 channels
  if (g >= groups) return;
  const long long pixel = g / cgroups;
  const bool padded = row_mask != nullptr && row_mask[pixel / pix_per_row] != 0;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (padded) {
    out[g] = zero;
    if (kBackward && dres != nullptr) dres[g] = zero;
    return;
  }
  const float4 u0 = reinterpret_cast<const float4*>(u)[2 * g];
  const float4 u1 = reinterpret_cast<const float4*>(u)[2 * g + 1];
  const float uv[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
  float o[8];
  if (!kBackward) {
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (in_bf16 != nullptr) unpack8(in_bf16[g], r);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = aptx<false>(uv[i], beta, gamma) + r[i];
  } else {
    const uint4 dyv = in_bf16[g];
    float dy[8];
    unpack8(dyv, dy);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float t = tanh_precise(beta * uv[i]);
      // d/du [gamma u (1 + t)] = gamma (1 + t) + gamma beta u (1 - t^2)
      o[i] = dy[i] * gamma * fmaf(beta * uv[i], fmaf(-t, t, 1.0f), 1.0f + t);
    }
    if (dres != nullptr) dres[g] = dyv;
  }
  out[g] = pack8(o);
}

static int act_launch(bool backward, const float* u, const void* in_bf16, const uint8_t* row_mask, int64_t pixels, int C,
                      int pix_per_row, float beta, float gamma, void* out, void* dres, cudaStream_t stream, const char* who) {
  MQ_REQUIRE(u && out, "%s: null pointer argument", who);
  MQ_REQUIRE(!backward || in_bf16, "%s: dy is required", who);
  MQ_REQUIRE(pixels >= 0 && C >= 8 && C % 8 == 0 && pix_per_row >= 1, "%s: pixels=%lld C=%d (multiple of 8) pix_per_row=%d", who,
             (long long)pixels, C, pix_per_row);
  if (pixels == 0) return 0;
  const long long groups = static_cast<long long>(pixels) * (C / 8);
  const long long grid = (groups + 255) / 256;
  MQ_REQUIRE(grid < (1LL << 31), "%s: tensor too large", who);
  if (backward)
    act_kernel<true><<<static_cast<unsigned>(grid), 256, 0, stream>>>(u, reinterpret_cast<const uint4*>(in_bf16), row_mask, groups,
                                                                       C / 8, pix_per_row, beta, gamma,
                                                                       reinterpret_cast<uint4*>(out), reinterpret_cast<uint4*>(dres));
  else
    act_kernel<false><<<static_cast<unsigned>(grid), 256, 0, stream>>>(u, reinterpret_cast<const uint4*>(in_bf16), row_mask, groups,
                                                                        C / 8, pix_per_row, beta, gamma,
                                                                        reinterpret_cast<uint4*>(out), nullptr);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mq

using namespace mq;

extern "C" int mq_act_forward(const float* u, const void* res_bf16, const uint8_t* row_mask, int64_t pixels, int C,
                              int pix_per_row, float beta, float gamma, void* out_bf16, mq_stream_t stream) {
  return act_launch(false, u, res_bf16, row_mask, pixels, C, pix_per_row, beta, gamma, out_bf16, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), "mq_act_forward");
}

extern "C" int mq_act_backward(const void* dy_bf16, const float* u, const uint8_t* row_mask, int64_t pixels, int C,
                               int pix_per_row, float beta, float gamma, void* du_bf16, void* dres_bf16, mq_stream_t stream) {
  return act_launch(true, u, dy_bf16, row_mask, pixels, C, pix_per_row, beta, gamma, du_bf16, dres_bf16,
                    reinterpret_cast<cudaStream_t>(stream), "mq_act_backward");
}
