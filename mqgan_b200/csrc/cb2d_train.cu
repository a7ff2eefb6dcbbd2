// ConvBlock2D point-wise stage for the TRAINING step, forward and backward, without ever
// materialising the reference's (B, C, C, T) expansion (preencoder.py:288-295):
//
//   y[p]   = bout + sum_k wout_k * a(u_k),   u_k = wpw_k * s[p] + bpw_k,   a(u) = (1 + tanh u) * u / 2
//   ds[p]  = dy[p] * sum_k wout_k * wpw_k * a'(u_k),          a'(u) = (1 + t)/2 + u (1 - t^2)/2,  t = tanh u
//   dwout_k = sum_p dy[p] a(u_k)     dwpw_k = sum_p dy[p] wout_k a'(u_k) s[p]     dbpw_k = sum_p dy[p] wout_k a'(u_k)
//
// over valid rows; at padded rows (row_mask != 0) the reference zeroes u after the point-wise conv (:292), so
// y = bout and nothing but bout receives a gradient.  s is the masked depth-wise 5x5 output (:286-287), which
// stays a 25-tap library convolution on the host side; dbout = sum dy is a plain reduction done by the caller.
//
// All three kernels are MUFU-bound (C tanh per pixel: ex2 + rcp), not HBM-bound: pixel-parallel for y and ds
// (parameters broadcast from shared memory), and K-PARALLEL for the parameter gradients - thread k owns its
// three running sums and walks the block's pixels, which are broadcast from shared memory, so there is no
// cross-thread reduction in the inner loop.  Block partials [blocks][3][C] are summed by the caller.
#include <string.h>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

constexpr int kCbPixPerThread = 4;
constexpr int kCbThreads = 256;
constexpr int kCbGradPix = 1024;        // pixels per block of the parameter-gradient kernel

// kFast: tanh.approx.f32 (one MUFU op, ~2^-11 relative error - below the bf16 rounding of the convolutions around it)
// instead of ex2 + rcp (two MUFU ops, ~1e-7 absolute): these kernels are MUFU-bound, so it is nearly 2x.
template <bool kFast>
__device__ __forceinline__ void cb_eval(float u, float& a, float& da) {
  const float t = kFast ? tanh_fast(u) : tanh_precise(u);
  const float h = 0.5f * (1.0f + t);
  a = h * u;
  da = fmaf(0.5f * u, fmaf(-t, t, 1.0f), h);
}

// mode 0: y = g(s);  mode 1: ds = dy * g'(s)
template <int kMode, bool kFast>
__global__ void __launch_bounds__(kCbThreads) cb2d_point_kernel(const float* __restrict__ s, const float* __restrict__ dy,
                                                               const uint8_t* __restrict__ row_mask, long long rows, int C,
                                                               const float* __restrict__ wpw, const float* __restrict__ bpw,
                                                               const float* __restrict__ wout, const float* __restrict__ bout,
                                                               float* __restrict__ out) {
  extern __shared__ float4 prm[];                       // (wpw, bpw, wout, wout*wpw) per k
  for (int k = threadIdx.x; k < C; k += blockDim.x) prm[k] = make_float4(wpw[k], bpw[k], wout[k], wout[k] * wpw[k]);
  __syncthreads();
  const long long total = rows * C;
  const long long base = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) * kCbPixPerThread;
  if (base >= total) return;
  float sv[kCbPixPerThread], acc[kCbPixPerThread];
  bool live[kCbPixPerThread];
#pragma unroll
  for (int i = 0; i < kCbPixPerThread; ++i) {
    const long long p = base + i;
    live[i] = p < total && (row_mask == nullptr || row_mask[p / C] == 0);
    sv[i] = (p < total) ? s[p] : 0.0f;
    acc[i] = 0.0f;
  }
  for (int k = 0; k < C; ++k) {
    const float4 q = prm[k];
#pragma unroll
    for (int i = 0; i < kCbPixPerThread; ++i) {
      float a, da;
      cb_eval<kFast>(fmaf(q.x, sv[i], q.y), a, da);
      acc[i] = kMode == 0 ? fmaf(q.z, a, acc[i]) : fmaf(q.w, da, acc[i]);
    }
  }
  const float b0 = bout[0];
#pragma unroll
  for (int i = 0; i < kCbPixPerThread; ++i) {
    const long long p = base + i;
    if (p < total) {
      if (kMode == 0) out[p] = live[i] ? acc[i] + b0 : b0;
      else out[p] = live[i] ? acc[i] * dy[p] : 0.0f;
    }
  }
}

// parameter gradients: thread k, pixels [blockIdx.x * kCbGradPix, +kCbGradPix) broadcast from shared memory
template <bool kFast>
__global__ void __launch_bounds__(1024) cb2d_param_grad_kernel(const float* __restrict__ s, const float* __restrict__ dy,
                                                                const uint8_t* __restrict__ row_mask, long long rows, int C,
                                                                const float* __restrict__ wpw, const float* __restrict__ bpw,
                                                                const float* __restrict__ wout, float* __restrict__ part) {
  __shared__ float2 px[kCbGradPix];                     // (s, dy or 0 at padded rows)
  const long long total = rows * C;
  const long long p0 = static_cast<long long>(blockIdx.x) * kCbGradPix;
  for (int i = threadIdx.x; i < kCbGradPix; i += blockDim.x) {
    const long long p = p0 + i;
    float2 v = make_float2(0.0f, 0.0f);
    if (p < total && (row_mask == nullptr || row_mask[p / C] == 0)) v = make_float2(s[p], dy[p]);
    px[i] = v;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < C; k += blockDim.x) {
    const float w = wpw[k], b = bpw[k], wo = wout[k];
    float g_wout = 0.0f, g_wpw = 0.0f, g_bpw = 0.0f;
#pragma unroll 4
    for (int i = 0; i < kCbGradPix; ++i) {
      const float2 v = px[i];
      float a, da;
      cb_eval<kFast>(fmaf(w, v.x, b), a, da);
      g_wout = fmaf(v.y, a, g_wout);
      const float g = v.y * da;
      g_bpw += g;
      g_wpw = fmaf(g, v.x, g_wpw);
    }
    float* o = part + static_cast<size_t>(blockIdx.x) * 3 * C;
    o[k] = g_wpw * wo;
    o[C + k] = g_bpw * wo;
    o[2 * C + k] = g_wout;
  }
}

}  // namespace mq

using namespace mq;

extern "C" int mq_cb2d_point_forward(const float* s, const uint8_t* row_mask, int64_t rows, int C, const float* wpw,
                                     const float* bpw, const float* wout, const float* bout, int fast_tanh, float* y,
                                     mq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MQ_REQUIRE(s && wpw && bpw && wout && bout && y, "mq_cb2d_point_forward: null pointer argument");
  MQ_REQUIRE(rows >= 0 && C >= 1 && C <= 2048, "mq_cb2d_point_forward: rows=%lld C=%d", (long long)rows, C);
  if (rows == 0) return 0;
  const long long total = rows * C;
  const long long per_block = static_cast<long long>(kCbThreads) * kCbPixPerThread;
  const long long grid = (total + per_block - 1) / per_block;
  MQ_REQUIRE(grid < (1LL << 31), "mq_cb2d_point_forward: too many pixels");
  if (fast_tanh)
    cb2d_point_kernel<0, true><<<static_cast<unsigned>(grid), kCbThreads, C * sizeof(float4), stream>>>(s, nullptr, row_mask, rows,
                                                                                                    C, wpw, bpw, wout, bout, y);
  else
    cb2d_point_kernel<0, false><<<static_cast<unsigned>(grid), kCbThreads, C * sizeof(float4), stream>>>(s, nullptr, row_mask, rows,
                                                                                                     C, wpw, bpw, wout, bout, y);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_cb2d_grad_blocks(int64_t rows, int C) {
  const long long total = static_cast<long long>(rows) * C;
  return static_cast<int>((total + kCbGradPix - 1) / kCbGradPix);
}

extern "C" int mq_cb2d_backward(const float* s, const float* dy, const uint8_t* row_mask, int64_t rows, int C,
                                const float* wpw, const float* bpw, const float* wout, int fast_tanh, float* ds, float* part,
                                mq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MQ_REQUIRE(s && dy && wpw && bpw && wout && ds && part, "mq_cb2d_backward: null pointer argument");
  MQ_REQUIRE(rows >= 0 && C >= 1 && C <= 2048, "mq_cb2d_backward: rows=%lld C=%d", (long long)rows, C);
  if (rows == 0) return 0;
  const long long total = rows * C;
  const long long per_block = static_cast<long long>(kCbThreads) * kCbPixPerThread;
  const long long grid = (total + per_block - 1) / per_block;
  MQ_REQUIRE(grid < (1LL << 31), "mq_cb2d_backward: too many pixels");
  // bout is unused by the ds pass; wout stands in for the pointer
  if (fast_tanh)
    cb2d_point_kernel<1, true><<<static_cast<unsigned>(grid), kCbThreads, C * sizeof(float4), stream>>>(s, dy, row_mask, rows, C, wpw,
                                                                                                    bpw, wout, wout, ds);
  else
    cb2d_point_kernel<1, false><<<static_cast<unsigned>(grid), kCbThreads, C * sizeof(float4), stream>>>(s, dy, row_mask, rows, C, wpw,
                                                                                                     bpw, wout, wout, ds);
  MQ_CUDA_OK(cudaGetLastError());
  const int gblocks = mq_cb2d_grad_blocks(rows, C);
  const int threads = C >= 1024 ? 1024 : ((C + 31) / 32 * 32);
  if (fast_tanh)
    cb2d_param_grad_kernel<true><<<gblocks, threads, 0, stream>>>(s, dy, row_mask, rows, C, wpw, bpw, wout, part);
  else
    cb2d_param_grad_kernel<false><<<gblocks, threads, 0, stream>>>(s, dy, row_mask, rows, C, wpw, bpw, wout, part);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}
