// Fused activation passes of the TRAINING step over channel-last (pixels, C) tensors:
//
//   refiner ConvBlock (preencoder.py:97-101):   y  = row padded ? 0 : aptx(u; beta, gamma) [+ res]
//                                               du = row padded ? 0 : dy * aptx'(u),  dres = row padded ? 0 : dy
//   discriminators (discriminators.py:234,247): y  = patch padded ? 0 : LeakyReLU_slope(u + bias)
//                                               du = patch padded ? 0 : dy * (u + bias > 0 ? 1 : slope)
//
// u is the convolution output kept for the backward pass (fp32 from the library's convolutions, bf16 from cuDNN's
// autocast ones); y / res / dy / du / dres are bf16 so the next tensor-core convolution (forward, data- or
// weight-gradient) consumes them as they are.  The backward pass can also emit per-block column sums of du - summed
// over blocks they are the bias gradient of the convolution that produced u - so du is not read a second time.
// One pass each instead of the ~15 element-wise PyTorch kernels autograd would run per activation; HBM-bound,
// 8 channels (one 16-byte bf16 / two 16-byte fp32 accesses) per thread.
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

struct ActArgs {
  const void* u;            // fp32 or bf16 (pixels, C)
  const uint4* in_bf16;     // forward: res (optional); backward: dy
  const float* bias;        // [C], added to u before the activation (optional)
  const uint8_t* row_mask;  // [pixels / pix_per_row], 1 = padded (optional)
  long long groups;         // pixels * C / 8
  int cgroups;              // C / 8
  int pix_per_row;
  float beta, gamma;        // APTx (beta, gamma) or LeakyReLU slope in beta
  uint4* out;               // forward: y; backward: du
  uint4* dres;              // backward, optional
  float* dbias_part;        // backward, optional: [gridDim.x][C]
};

// kAct 0: APTx; 1: LeakyReLU.  kUBf16: u is bf16.
template <bool kBackward, int kAct, bool kUBf16>
__global__ void __launch_bounds__(256) act_kernel(const ActArgs a) {
  __shared__ float red[kBackward ? 256 * 9 : 1];
  const long long g = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;   // group of 8 channels
  const bool want_bias = kBackward && a.dbias_part != nullptr;
  const bool in_range = g < a.groups;
  if (!in_range && !want_bias) return;
  const long long pixel = in_range ? g / a.cgroups : 0;
  const int cg = in_range ? static_cast<int>(g - pixel * a.cgroups) : 0;
  const bool live = in_range && !(a.row_mask != nullptr && a.row_mask[pixel / a.pix_per_row] != 0);
  float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  uint4 dyv = make_uint4(0u, 0u, 0u, 0u);
  if (live) {
    float uv[8];
    if (kUBf16) {
      unpack8(reinterpret_cast<const uint4*>(a.u)[g], uv);
    } else {
      const float4 u0 = reinterpret_cast<const float4*>(a.u)[2 * g];
      const float4 u1 = reinterpret_cast<const float4*>(a.u)[2 * g + 1];
      uv[0] = u0.x; uv[1] = u0.y; uv[2] = u0.z; uv[3] = u0.w; uv[4] = u1.x; uv[5] = u1.y; uv[6] = u1.z; uv[7] = u1.w;
    }
    if (a.bias != nullptr) {
      const float4 b0 = reinterpret_cast<const float4*>(a.bias)[2 * cg];
      const float4 b1 = reinterpret_cast<const float4*>(a.bias)[2 * cg + 1];
      uv[0] += b0.x; uv[1] += b0.y; uv[2] += b0.z; uv[3] += b0.w; uv[4] += b1.x; uv[5] += b1.y; uv[6] += b1.z; uv[7] += b1.w;
    }
    if (!kBackward) {
      float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (a.in_bf16 != nullptr) unpack8(a.in_bf16[g], r);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        o[i] = (kAct == 0 ? aptx<false>(uv[i], a.beta, a.gamma) : (uv[i] > 0.0f ? uv[i] : a.beta * uv[i])) + r[i];
    } else {
      dyv = a.in_bf16[g];
      float dy[8];
      unpack8(dyv, dy);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (kAct == 0) {
          const float t = tanh_precise(a.beta * uv[i]);
          // d/du [gamma u (1 + t)] = gamma (1 + t) + gamma beta u (1 - t^2)
          o[i] = dy[i] * a.gamma * fmaf(a.beta * uv[i], fmaf(-t, t, 1.0f), 1.0f + t);
        } else {
          o[i] = uv[i] > 0.0f ? dy[i] : a.beta * dy[i];
        }
      }
    }
  }
  if (in_range) {
    a.out[g] = pack8(o);
    if (kBackward && a.dres != nullptr) a.dres[g] = dyv;       // zero at padded rows
  }
  if (want_bias) {
    // block-local column sums of the unrounded du: the block covers 256 / cgroups whole pixels
    // (the host checks 256 % cgroups == 0)
#pragma unroll
    for (int i = 0; i < 8; ++i) red[threadIdx.x * 9 + i] = o[i];
    __syncthreads();
    const int C = a.cgroups * 8;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const int cc = c >> 3, i = c & 7;
      float acc = 0.0f;
      for (int p = cc; p < 256; p += a.cgroups) acc += red[p * 9 + i];
      a.dbias_part[static_cast<size_t>(blockIdx.x) * C + c] = acc;
    }
  }
}

static int act_launch(bool backward, int act, bool u_bf16, const void* u, const void* in_bf16, const float* bias,
                      const uint8_t* row_mask, int64_t pixels, int C, int pix_per_row, float beta, float gamma, void* out,
                      void* dres, float* dbias_part, cudaStream_t stream, const char* who) {
  MQ_REQUIRE(u && out, "%s: null pointer argument", who);
  MQ_REQUIRE(!backward || in_bf16, "%s: dy is required", who);
  MQ_REQUIRE(pixels >= 0 && C >= 8 && C % 8 == 0 && pix_per_row >= 1, "%s: pixels=%lld C=%d (multiple of 8) pix_per_row=%d", who,
             (long long)pixels, C, pix_per_row);
  MQ_REQUIRE(dbias_part == nullptr || 256 % (C / 8) == 0, "%s: the fused bias gradient needs 256 %% (C / 8) == 0 (C=%d)", who, C);
  MQ_REQUIRE(bias == nullptr || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "%s: bias must be 16-byte aligned", who);
  if (pixels == 0) return 0;
  ActArgs a;
  a.u = u; a.in_bf16 = reinterpret_cast<const uint4*>(in_bf16); a.bias = bias; a.row_mask = row_mask;
  a.groups = static_cast<long long>(pixels) * (C / 8); a.cgroups = C / 8; a.pix_per_row = pix_per_row;
  a.beta = beta; a.gamma = gamma; a.out = reinterpret_cast<uint4*>(out); a.dres = reinterpret_cast<uint4*>(dres);
  a.dbias_part = dbias_part;
  const long long grid = (a.groups + 255) / 256;
  MQ_REQUIRE(grid < (1LL << 31), "%s: tensor too large", who);
#define MQ_ACT_LAUNCH(BW, ACT, UB) act_kernel<BW, ACT, UB><<<static_cast<unsigned>(grid), 256, 0, stream>>>(a)
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
  return act_launch(false, 0, false, u, res_bf16, nullptr, row_mask, pixels, C, pix_per_row, beta, gamma, out_bf16, nullptr,
                    nullptr, reinterpret_cast<cudaStream_t>(stream), "mq_act_forward");
}

extern "C" int mq_act_backward(const void* dy_bf16, const float* u, const uint8_t* row_mask, int64_t pixels, int C,
                               int pix_per_row, float beta, float gamma, void* du_bf16, void* dres_bf16, float* dbias_part,
                               mq_stream_t stream) {
  return act_launch(true, 0, false, u, dy_bf16, nullptr, row_mask, pixels, C, pix_per_row, beta, gamma, du_bf16, dres_bf16,
                    dbias_part, reinterpret_cast<cudaStream_t>(stream), "mq_act_backward");
}

extern "C" int mq_act_bias_blocks(int64_t pixels, int C) {
  if (C < 8 || C % 8 || 256 % (C / 8)) return 0;           // 0: the fused bias gradient is not available for this width
  return static_cast<int>((pixels * (C / 8) + 255) / 256);
}

extern "C" int mq_leaky_mask_forward(const void* u, int u_is_bf16, const float* bias, const uint8_t* pix_mask, int64_t pixels,
                                     int C, float slope, void* out_bf16, mq_stream_t stream) {
  return act_launch(false, 1, u_is_bf16 != 0, u, nullptr, bias, pix_mask, pixels, C, 1, slope, 0.0f, out_bf16, nullptr, nullptr,
                    reinterpret_cast<cudaStream_t>(stream), "mq_leaky_mask_forward");
}

extern "C" int mq_leaky_mask_backward(const void* dy_bf16, const void* u, int u_is_bf16, const float* bias,
                                      const uint8_t* pix_mask, int64_t pixels, int C, float slope, void* du_bf16,
                                      float* dbias_part, mq_stream_t stream) {
  return act_launch(true, 1, u_is_bf16 != 0, u, dy_bf16, bias, pix_mask, pixels, C, 1, slope, 0.0f, du_bf16, nullptr, dbias_part,
                    reinterpret_cast<cudaStream_t>(stream), "mq_leaky_mask_backward");
}
