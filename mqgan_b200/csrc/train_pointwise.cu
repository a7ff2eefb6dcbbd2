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

// kAct 0: APTx(beta, gamma); 1: LeakyReLU(slope = beta) (discriminators.py:187, 234).  kUBf16: u is bf16 (cuDNN's
// autocast output) instead of fp32.
template <bool kBackward, int kAct, bool kUBf16>
__global__ void __launch_bounds__(256) act_kernel(const void* __restrict__ u, const uint4* __restrict__ in_bf16,
                                                  const uint8_t* __restrict__ row_mask, long long groups, int cgroups,
                                                  int pix_per_row, float beta, float gamma, uint4* __restrict__ out,
                                                  uint4* __restrict__ dres, float* __restrict__ dbias_part) {
  // backward only, optional: per-block column sums of du (the convolution's bias gradient), [gridDim.x][C]
  __shared__ float red[kBackward ? 256 * 9 : 1];
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // group of 8 channels
  const bool want_bias = kBackward && dbias_part != nullptr;
  const bool in_range = g < groups;
  const long long pixel = in_range ? g / cgroups : 0;
  const bool padded = in_range && row_mask != nullptr && row_mask[pixel / pix_per_row] != 0;
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  if (!in_range || padded) {
    if (in_range) {
      out[g] = zero;
      if (kBackward && dres != nullptr) dres[g] = zero;
    }
    if (!want_bias) return;
  }
  if (want_bias) {
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (in_range && !padded) {
      float uv[8], dy[8];
      if (kUBf16) {
        unpack8(reinterpret_cast<const uint4*>(u)[g], uv);
      } else {
        const float4 u0 = reinterpret_cast<const float4*>(u)[2 * g];
        const float4 u1 = reinterpret_cast<const float4*>(u)[2 * g + 1];
        uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
      }
      const uint4 dyv = in_bf16[g];
      unpack8(dyv, dy);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kAct == 0) {
          const float t = tanh_precise(beta * uv[i]);
          o[i] = dy[i] * gamma * fmaf(beta * uv[i], fmaf(-t, t, 1.0f), 1.0f + t);
        } else {
          o[i] = uv[i] > 0.0f ? dy[i] : beta * dy[i];
        }
      }
      if (dres != nullptr) dres[g] = dyv;
      out[g] = pack8(o);
    }
    // block-local column sums: the block covers 256 / cgroups whole pixels (the host checks 256 % cgroups == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 9 + i] = o[i];
    __syncthreads();
    const int C = cgroups * 8;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int cg = c >> 3, i = c & 7;
      float acc = 0.0f;
      for (int p = cg; p < 256; p += cgroups) acc += red[p * 9 + i];
      dbias_part[static_cast<size_t>(blockIdx.x) * C + c] = acc;
    }
    return;
  }
  float uv[8];
  if (kUBf16) {
    unpack8(reinterpret_cast<const uint4*>(u)[g], uv);
  } else {
    const float4 u0 = reinterpret_cast<const float4*>(u)[2 * g];
    const float4 u1 = reinterpret_cast<const float4*>(u)[2 * g + 1];
    uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
  }
  float o[8];
  if (!kBackward) {
    float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (in_bf16 != nullptr) unpack8(in_bf16[g], r);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      o[i] = (kAct == 0 ? aptx<false>(uv[i], beta, gamma) : (uv[i] > 0.0f ? uv[i] : beta * uv[i])) + r[i];
  } else {
    const uint4 dyv = in_bf16[g];
    float dy[8];
    unpack8(dyv, dy);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (kAct == 0) {
        const float t = tanh_precise(beta * uv[i]);
        // d/du [gamma u (1 + t)] = gamma (1 + t) + gamma beta u (1 - t^2)
        o[i] = dy[i] * gamma * fmaf(beta * uv[i], fmaf(-t, t, 1.0f), 1.0f + t);
      } else {
        o[i] = uv[i] > 0.0f ? dy[i] : beta * dy[i];
      }
    }
    if (dres != nullptr) dres[g] = dyv;
  }
  out[g] = pack8(o);
}

static int act_launch(bool backward, int act, bool u_bf16, const void* u, const void* in_bf16, const uint8_t* row_mask, int64_t pixels, int C,
                      int pix_per_row, float beta, float gamma, void* out, void* dres, cudaStream_t stream, const char* who,
                      float* dbias_part = nullptr) {
  MQ_REQUIRE(u && out, "%s: null pointer argument", who);
  MQ_REQUIRE(!backward || in_bf16, "%s: dy is required", who);
  MQ_REQUIRE(pixels >= 0 && C >= 8 && C % 8 == 0 && pix_per_row >= 1, "%s: pixels=%lld C=%d (multiple of 8) pix_per_row=%d", who,
             (long long)pixels, C, pix_per_row);
  MQ_REQUIRE(dbias_part == nullptr || 256 % (C / 8) == 0, "%s: the fused bias gradient needs 256 %% (C / 8) == 0 (C=%d)", who, C);
  if (pixels == 0) return 0;
  const long long groups = static_cast<long long>(pixels) * (C / 8);
  const long long grid = (groups + 255) / 256;
  MQ_REQUIRE(grid < (1LL << 31), "%s: tensor too large", who);
#define MQ_ACT_LAUNCH(BW, ACT, UB)                                                                                         \
  act_kernel<BW, ACT, UB><<<static_cast<unsigned>(grid), 256, 0, stream>>>(                                                \
      u, reinterpret_cast<const uint4*>(in_bf16), row_mask, groups, C / 8, pix_per_row, beta, gamma,                       \
      reinterpret_cast<uint4*>(out), reinterpret_cast<uint4*>(dres), dbias_part)
  if (backward) {
    if (act == 0) { if (u_bf16) MQ_ACT_LAUNCH(true, 0, true); else MQ_ACT_LAUNCH(true, 0, false); }
    else { if (u_bf16) MQ_ACT_LAUNCH(true, 1, true); else MQ_ACT_LAUNCH(true, 1, false); }
  } else {
    if (act == 0) { if (u_bf16) MQ_ACT_LAUNCH(false, 0, true); else MQ_ACT_LAUNCH(false, 0, false); }
    else { if (u_bf16) MQ_ACT_LAUNCH(false, 1, true); else MQ_ACT_LAUNCH(false, 1, false); }
  }
#undef MQ_ACT_LAUNCH
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace mq

using namespace mq;

extern "C" int mq_act_forward(const float* u, const void* res_bf16, const uint8_t* row_mask, int64_t pixels, int C,
                              int pix_per_row, float beta, float gamma, void* out_bf16, mq_stream_t stream) {
  return act_launch(false, 0, false, u, res_bf16, row_mask, pixels, C, pix_per_row, beta, gamma, out_bf16, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), "mq_act_forward");
}

extern "C" int mq_act_backward(const void* dy_bf16, const float* u, const uint8_t* row_mask, int64_t pixels, int C,
                               int pix_per_row, float beta, float gamma, void* du_bf16, void* dres_bf16, float* dbias_part,
                               mq_stream_t stream) {
  return act_launch(true, 0, false, u, dy_bf16, row_mask, pixels, C, pix_per_row, beta, gamma, du_bf16, dres_bf16,
                    reinterpret_cast<cudaStream_t>(stream), "mq_act_backward", dbias_part);
}

extern "C" int mq_act_bias_blocks(int64_t pixels, int C) {
  if (C < 8 || C % 8 || 256 % (C / 8)) return 0;           // 0: the fused bias gradient is not available for this width
  return static_cast<int>((pixels * (C / 8) + 255) / 256);
}

extern "C" int mq_leaky_mask_forward(const void* u, int u_is_bf16, const uint8_t* pix_mask, int64_t pixels, int C, float slope,
                                     void* out_bf16, mq_stream_t stream) {
  return act_launch(false, 1, u_is_bf16 != 0, u, nullptr, pix_mask, pixels, C, 1, slope, 0.0f, out_bf16, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), "mq_leaky_mask_forward");
}

extern "C" int mq_leaky_mask_backward(const void* dy_bf16, const void* u, int u_is_bf16, const uint8_t* pix_mask, int64_t pixels,
                                      int C, float slope, void* du_bf16, mq_stream_t stream) {
  return act_launch(true, 1, u_is_bf16 != 0, u, dy_bf16, pix_mask, pixels, C, 1, slope, 0.0f, du_bf16, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), "mq_leaky_mask_backward");
}
