// Memory-/MUFU-bound kernels of the PreEncoder path:
// fused, coalesced, 128-bit vectorised passes over channel-last tensors.
#include <math.h>
#include <string.h>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

static inline int sm_count_cached() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

// grid sized as a multiple of the SM count for grid-stride kernels
static inline int grid_for(int64_t work_items, int per_block, int waves = 8) {
  int64_t need = (work_items + per_block - 1) / per_block;
  int64_t cap = static_cast<int64_t>(sm_count_cached()) * waves;
  if (need < 1) need = 1;
  return static_cast<int>(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------
// fp32 -> bf16 / bf16x3
// ---------------------------------------------------------------------------
__global__ void split_bf16_kernel(const float* __restrict__ x, uint16_t* __restrict__ out,
                                  int64_t rows, int C, int nterms) {
  // nterms 1: plain bf16; 3: bf16x3; 2: f16x2.  Term j at [j*C, (j+1)*C) of each output row.
  const int64_t total = rows * (C / 4);
  const int c4 = C / 4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    const float4 v = *reinterpret_cast<const float4*>(x + r * C + c);
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint16_t* o = out + r * (static_cast<int64_t>(nterms) * C) + c;
    if (nterms == 1) {
      uint2 u;
      u.x = pack_bf16x2(f[0], f[1]);
      u.y = pack_bf16x2(f[2], f[3]);
      *reinterpret_cast<uint2*>(o) = u;
    } else {
      store_terms4(o, C, nterms == 2 ? 1 : 0, f);
    }
  }
}

// ---------------------------------------------------------------------------
// K2: ConvBlock2D pre/post.  Block = (b, 8 frames, 64 channels); 256 threads,
// two pixels per thread.  MUFU-bound: C tanh per pixel.
// ---------------------------------------------------------------------------
constexpr int kCbT = 8, kCbC = 64;

template <bool kFast, bool kInBf16>
__global__ void __launch_bounds__(256)
convblock2d_kernel(const void* __restrict__ xin, int B, int T, int C, const float* __restrict__ dw,
                   const float4* __restrict__ pw, float bout, const uint8_t* __restrict__ row_mask,
                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                   uint16_t* __restrict__ out_split, int split_kind) {
  extern __shared__ float4 smem_pw[];                       // [C] {wpw, bpw, 0.5*wout, 0}
  __shared__ float tile[kCbT + 4][kCbC + 4];
  __shared__ float dws[26];
  const int ctiles = (C + kCbC - 1) / kCbC;
  const int ttiles = (T + kCbT - 1) / kCbT;
  int bid = blockIdx.x;
  const int ct = bid % ctiles; bid /= ctiles;
  const int tt = bid % ttiles;
  const int b = bid / ttiles;
  const int t0 = tt * kCbT, c0 = ct * kCbC;

  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    float4 p = pw[i];
    p.z *= 0.5f;                                            // fold APTx gamma into wout
    smem_pw[i] = p;
  }
  if (threadIdx.x < 26) dws[threadIdx.x] = dw[threadIdx.x];
  for (int i = threadIdx.x; i < (kCbT + 4) * (kCbC + 4); i += blockDim.x) {
    const int lt = i / (kCbC + 4), lc = i - lt * (kCbC + 4);
    const int t = t0 + lt - 2, c = c0 + lc - 2;
    float v = 0.0f;
    if (t >= 0 && t < T && c >= 0 && c < C) {
      const int64_t off = (static_cast<int64_t>(b) * T + t) * C + c;
      v = kInBf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xin)[off])
                  : reinterpret_cast<const float*>(xin)[off];
    }
    tile[lt][lc] = v;
  }
  __syncthreads();

  // thread -> pixels (lt, lc) and (lt + 4, lc): lc = tid % 64, lt = tid / 64
  const int lc = threadIdx.x & 63;
  const int lt0 = threadIdx.x >> 6;
  float s[2];
  bool live[2], masked[2];
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    const int lt = lt0 + 4 * p;
    const int t = t0 + lt, c = c0 + lc;
    live[p] = (t < T) && (c < C);
    masked[p] = live[p] && row_mask != nullptr && row_mask[static_cast<int64_t>(b) * T + t] != 0;
    float acc = dws[25];
#pragma unroll
    for (int i = 0; i < 5; ++i)        // i: channel offset, j: time offset (Conv2d on the (C, T) image)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc = fmaf(dws[i * 5 + j], tile[lt + j][lc + i], acc);
    s[p] = masked[p] ? 0.0f : acc;
  }
  float acc0 = 0.0f, acc1 = 0.0f;
#pragma unroll 4
  for (int k = 0; k < C; ++k) {
    const float4 p = smem_pw[k];
    const float u0 = fmaf(p.x, s[0], p.y);
    const float u1 = fmaf(p.x, s[1], p.y);
    const float th0 = kFast ? tanh_fast(u0) : tanh_precise(u0);
    const float th1 = kFast ? tanh_fast(u1) : tanh_precise(u1);
    acc0 = fmaf(fmaf(u0, th0, u0), p.z, acc0);             // (1 + tanh u) * u * (0.5 wout)
    acc1 = fmaf(fmaf(u1, th1, u1), p.z, acc1);
  }
  const float y[2] = {acc0 + bout, acc1 + bout};
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (!live[p]) continue;
    const int t = t0 + lt0 + 4 * p, c = c0 + lc;
    const float v = masked[p] ? bout : y[p];
    const int64_t row = static_cast<int64_t>(b) * T + t;
    if (out_f32) out_f32[row * C + c] = v;
    if (out_bf16) out_bf16[row * C + c] = __float2bfloat16_rn(v);
    if (out_split) store_terms1(out_split + row * split_nterms(split_kind) * C + c, C, split_kind, v);
  }
}

// K2, table mode: g(s) from per-interval cubics (HBM-bound instead of MUFU-bound).
// Block = (b, 32 frames, 128 channels); thread = 4 consecutive channels of four frames (8 apart):
// the 2-row / 2-column stencil halo costs 16 % instead of 55 % of the tile loads, and the per-block
// set-up is amortised over 4x the outputs.
constexpr int kCtT = 32, kCtC = 128;

template <bool kFast, bool kInBf16>
__global__ void __launch_bounds__(256)
convblock2d_table_kernel(const void* __restrict__ xin, int B, int T, int C, const float* __restrict__ dw,
                         const float4* __restrict__ pw, float bout, const uint8_t* __restrict__ row_mask,
                         const float4* __restrict__ table, int table_n, int table_off, float inv_h,
                         float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                         uint16_t* __restrict__ out_split, int split_kind) {
  __shared__ float tile[kCtT + 4][kCtC + 4];
  __shared__ float dws[26];
  const int ctiles = (C + kCtC - 1) / kCtC;
  const int ttiles = (T + kCtT - 1) / kCtT;
  int bid = blockIdx.x;
  const int ct = bid % ctiles; bid /= ctiles;
  const int tt = bid % ttiles;
  const int b = bid / ttiles;
  const int t0 = tt * kCtT, c0 = ct * kCtC;
  if (threadIdx.x < 26) dws[threadIdx.x] = dw[threadIdx.x];
  for (int i = threadIdx.x; i < (kCtT + 4) * (kCtC + 4); i += blockDim.x) {
    const int lt = i / (kCtC + 4), lc = i - lt * (kCtC + 4);
    const int t = t0 + lt - 2, c = c0 + lc - 2;
    float v = 0.0f;
    if (t >= 0 && t < T && c >= 0 && c < C) {
      const int64_t off = (static_cast<int64_t>(b) * T + t) * C + c;
      v = kInBf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(xin)[off])
                  : reinterpret_cast<const float*>(xin)[off];
    }
    tile[lt][lc] = v;
  }
  __syncthreads();
  const int lc = (threadIdx.x & 31) * 4;        // 32 x 4 channels
  const int c = c0 + lc;
  if (c >= C) return;
#pragma unroll 1
  for (int lt = threadIdx.x >> 5; lt < kCtT; lt += 8) {
  const int t = t0 + lt;
  if (t >= T) break;
  const int64_t row = static_cast<int64_t>(b) * T + t;
  const bool masked = row_mask != nullptr && row_mask[row] != 0;
  float y[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float acc = dws[25];
#pragma unroll
    for (int i = 0; i < 5; ++i)
#pragma unroll
      for (int j = 0; j < 5; ++j) acc = fmaf(dws[i * 5 + j], tile[lt + j][lc + e + i], acc);
    const float s = acc;
    const float u = s * inv_h;                  // inv_h is a power of two: exact
    const float fl = floorf(u);
    const int idx = static_cast<int>(fl) + table_off;
    float g;
    if (idx >= 0 && idx < table_n) {
      const float tt_ = u - fl;                 // exact
      const float4 cf = __ldg(table + idx);
      g = fmaf(tt_, fmaf(tt_, fmaf(tt_, cf.w, cf.z), cf.y), cf.x);
    } else {                                    // outside the grid (or NaN): exact C-term sum
      float a = 0.0f;
      for (int k = 0; k < C; ++k) {
        const float4 p = __ldg(pw + k);
        const float uu = fmaf(p.x, s, p.y);
        const float th = kFast ? tanh_fast(uu) : tanh_precise(uu);
        a = fmaf(fmaf(uu, th, uu), 0.5f * p.z, a);
      }
      g = a + bout;
    }
    y[e] = masked ? bout : g;
  }
  const int nv = min(4, C - c);
  if (nv == 4) {
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_bf16) {
      uint2 u2;
      u2.x = pack_bf16x2(y[0], y[1]);
      u2.y = pack_bf16x2(y[2], y[3]);
      *reinterpret_cast<uint2*>(out_bf16 + row * C + c) = u2;
    }
    if (out_split) store_terms4(out_split + row * split_nterms(split_kind) * C + c, C, split_kind, y);
  } else {
    for (int e = 0; e < nv; ++e) {
      if (out_f32) out_f32[row * C + c + e] = y[e];
      if (out_bf16) out_bf16[row * C + c + e] = __float2bfloat16_rn(y[e]);
      if (out_split) store_terms1(out_split + row * split_nterms(split_kind) * C + c + e, C, split_kind, y[e]);
    }
  }
  }
}

// ---------------------------------------------------------------------------
// K5: CAM reduce / gate
// ---------------------------------------------------------------------------
constexpr int kCamRows = 64;

__global__ void __launch_bounds__(256)
cam_reduce_kernel(const float* __restrict__ o, const uint8_t* __restrict__ row_mask, int T, int C,
                  int nchunk, float* __restrict__ part) {
  const int chunk = blockIdx.x, b = blockIdx.y;
  const int t_beg = chunk * kCamRows;
  const int t_end = min(T, t_beg + kCamRows);
  float* pmax = part + ((static_cast<int64_t>(b) * nchunk + chunk) * 2) * C;
  float* psum = pmax + C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float mx = -INFINITY, sm = 0.0f;
    for (int t = t_beg; t < t_end; ++t) {
      const int64_t row = static_cast<int64_t>(b) * T + t;
      const float v = o[row * C + c];
      mx = fmaxf(mx, v);                                   // max sees padded rows too (App. B1)
      const bool pad = row_mask != nullptr && row_mask[row] != 0;
      sm += pad ? 0.0f : v;
    }
    pmax[c] = mx;
    psum[c] = sm;
  }
}

__global__ void __launch_bounds__(256)
cam_gate_kernel(const float* __restrict__ part, const uint8_t* __restrict__ row_mask, int T, int C,
                int R, int nchunk, const float* __restrict__ w0, const float* __restrict__ b0,
                const float* __restrict__ w2, const float* __restrict__ b2,
                float* __restrict__ gate) {
  extern __shared__ float sm[];           // mx[C], av[C], hmx[R], hav[R]
  float* mx = sm;
  float* av = sm + C;
  float* hmx = av + C;
  float* hav = hmx + R;
  __shared__ int cnt_s;
  const int b = blockIdx.x;
  if (threadIdx.x == 0) cnt_s = 0;
  __syncthreads();
  int local = 0;
  for (int t = threadIdx.x; t < T; t += blockDim.x)
    local += (row_mask == nullptr || row_mask[static_cast<int64_t>(b) * T + t] == 0) ? 1 : 0;
  atomicAdd(&cnt_s, local);
  __syncthreads();
  const float cnt = fmaxf(static_cast<float>(cnt_s), 1.0f);     // clamp(min=1), attentions.py:129
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float m = -INFINITY, s = 0.0f;
    for (int k = 0; k < nchunk; ++k) {
      const float* p = part + ((static_cast<int64_t>(b) * nchunk + k) * 2) * C;
      m = fmaxf(m, p[c]);
      s += p[C + c];
    }
    mx[c] = m;
    av[c] = s / cnt;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int r = warp; r < R; r += nwarp) {
    float a0 = 0.0f, a1 = 0.0f;
    for (int c = lane; c < C; c += 32) {
      const float w = w0[static_cast<int64_t>(r) * C + c];
      a0 = fmaf(w, mx[c], a0);
      a1 = fmaf(w, av[c], a1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a0 += __shfl_xor_sync(0xffffffffu, a0, off);
      a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    }
    if (lane == 0) {
      hmx[r] = fmaxf(a0 + b0[r], 0.0f);
      hav[r] = fmaxf(a1 + b0[r], 0.0f);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a0 = b2[c], a1 = b2[c];
    for (int r = 0; r < R; ++r) {
      const float w = w2[static_cast<int64_t>(c) * R + r];
      a0 = fmaf(w, hmx[r], a0);
      a1 = fmaf(w, hav[r], a1);
    }
    gate[static_cast<int64_t>(b) * C + c] = sigmoid_precise(a0 + a1);
  }
}

// ---------------------------------------------------------------------------
// K6: SAM pools + conv7 + sigmoid + CBAM residual + block residual + mask + APTx
// ---------------------------------------------------------------------------
constexpr int kSamT = 32;

__global__ void __launch_bounds__(256)
cbam_apply_kernel(const float* __restrict__ o, const float* __restrict__ gate,
                  const float* __restrict__ res, const uint8_t* __restrict__ row_mask, int T, int C,
                  const float* __restrict__ sam_w, float beta, float gamma,
                  float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                  uint16_t* __restrict__ out_split, int split_kind) {
  extern __shared__ float gs[];                   // gate[C]
  __shared__ float pmax[kSamT + 6], pavg[kSamT + 6], sam[kSamT];
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kSamT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  for (int c = threadIdx.x; c < C; c += blockDim.x) gs[c] = gate[static_cast<int64_t>(b) * C + c];
  __syncthreads();
  for (int i = warp; i < kSamT + 6; i += nwarp) {
    const int t = t0 + i - 3;
    float m = -INFINITY, s = 0.0f;
    if (t >= 0 && t < T) {
      const float* row = o + (static_cast<int64_t>(b) * T + t) * C;
      for (int c = lane * 4; c < C; c += 128) {
        const float4 v = *reinterpret_cast<const float4*>(row + c);
        const float a0 = v.x * gs[c], a1 = v.y * gs[c + 1], a2 = v.z * gs[c + 2], a3 = v.w * gs[c + 3];
        m = fmaxf(fmaxf(m, fmaxf(a0, a1)), fmaxf(a2, a3));
        s += (a0 + a1) + (a2 + a3);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
        s += __shfl_xor_sync(0xffffffffu, s, off);
      }
      s /= static_cast<float>(C);
    } else {
      m = 0.0f;                                   // zero padding of the k=7 conv input
      s = 0.0f;
    }
    if (lane == 0) {
      pmax[i] = m;
      pavg[i] = s;
    }
  }
  __syncthreads();
  if (threadIdx.x < kSamT) {
    float l = 0.0f;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      l = fmaf(sam_w[j], pmax[threadIdx.x + j], fmaf(sam_w[7 + j], pavg[threadIdx.x + j], l));
    sam[threadIdx.x] = sigmoid_precise(l);
  }
  __syncthreads();
  const int c4 = C / 4;
  const int rows = min(kSamT, T - t0);
  for (int i = threadIdx.x; i < rows * c4; i += blockDim.x) {
    const int lt = i / c4;
    const int c = (i - lt * c4) * 4;
    const int64_t row = static_cast<int64_t>(b) * T + t0 + lt;
    const float4 ov = *reinterpret_cast<const float4*>(o + row * C + c);
    const float4 rv = *reinterpret_cast<const float4*>(res + row * C + c);
    const bool pad = row_mask != nullptr && row_mask[row] != 0;
    const float sv = sam[lt];
    const float oo[4] = {ov.x, ov.y, ov.z, ov.w};
    const float rr[4] = {rv.x, rv.y, rv.z, rv.w};
    float y[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float o1 = gs[c + e] * oo[e];
      float v = fmaf(sv, o1, oo[e]) + rr[e];      // SAM(CAM(o)) + o + residual
      v = pad ? 0.0f : v;
      y[e] = aptx<false>(v, beta, gamma);
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(y[0], y[1], y[2], y[3]);
    if (out_bf16) {
      uint2 u;
      u.x = pack_bf16x2(y[0], y[1]);
      u.y = pack_bf16x2(y[2], y[3]);
      *reinterpret_cast<uint2*>(out_bf16 + row * C + c) = u;
    }
    if (out_split) store_terms4(out_split + row * split_nterms(split_kind) * C + c, C, split_kind, y);
  }
}

// ---------------------------------------------------------------------------
// K7: q_in_proj + FSQ.  One warp per frame; fp64 accumulation of the D dots.
// ---------------------------------------------------------------------------
struct FsqDev {
  int D;
  float half_l[8], shift[8], offset[8];
  int half_w[8], basis[8], levels[8];
};

__device__ __forceinline__ int64_t fsq_index(const float* z, const FsqDev& f, float* codes) {
  int idx = 0;
  for (int d = 0; d < f.D; ++d) {
    const float bounded = tanhf(z[d] + f.shift[d]) * f.half_l[d] - f.offset[d];   // quantizer.py:114
    const float q = rintf(bounded);                                               // round half to even
    if (codes) codes[d] = q / static_cast<float>(f.half_w[d]);
    idx += (static_cast<int>(q) + f.half_w[d]) * f.basis[d];                      // :166, :181
  }
  return static_cast<int64_t>(idx);
}

// q_in_proj + FSQ, HBM-bound form: the weights live in shared memory as DOUBLES (converted once per CTA - the first
// version converted every weight for every frame and was bound by the F2F.F64 pipe at 0.22 of HBM speed), a warp
// handles four consecutive frames per pass so each shared-memory weight read feeds four DFMAs, and the activations
// are converted once per element.  Same arithmetic as before: exact fp64 products of fp32 values, fp64 accumulation,
// one rounding to fp32 at the end - the index gate is untouched.
constexpr int kQinRows = 4;

__global__ void __launch_bounds__(256)
qin_fsq_kernel4(const float* __restrict__ y, int64_t rows, int C, const float* __restrict__ w,
                const float* __restrict__ bias, const FsqDev f, int64_t* __restrict__ idx,
                float* __restrict__ z_out) {
  extern __shared__ double wd[];                       // [D][C]
  for (int i = threadIdx.x; i < f.D * C; i += blockDim.x) wd[i] = static_cast<double>(w[i]);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  const int64_t groups = (rows + kQinRows - 1) / kQinRows;
  for (int64_t g = warp0; g < groups; g += nwarps) {
    const int64_t r0 = g * kQinRows;
    double acc[kQinRows][8];
#pragma unroll
    for (int j = 0; j < kQinRows; ++j)
#pragma unroll
      for (int d = 0; d < 8; ++d) acc[j][d] = 0.0;
    for (int c = lane * 4; c < C; c += 128) {
      double v[kQinRows][4];
#pragma unroll
      for (int j = 0; j < kQinRows; ++j) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r0 + j < rows) t = __ldcs(reinterpret_cast<const float4*>(y + (r0 + j) * C + c));   // streamed once
        v[j][0] = t.x; v[j][1] = t.y; v[j][2] = t.z; v[j][3] = t.w;
      }
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        if (d < f.D) {
          const double2 w01 = *reinterpret_cast<const double2*>(wd + d * C + c);
          const double2 w23 = *reinterpret_cast<const double2*>(wd + d * C + c + 2);
#pragma unroll
          for (int j = 0; j < kQinRows; ++j) {
            // same summation order as the one-row kernel: ((x0 w0 + x1 w1) + x2 w2) + x3 w3, then += acc
            acc[j][d] += v[j][0] * w01.x + v[j][1] * w01.y + v[j][2] * w23.x + v[j][3] * w23.y;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kQinRows; ++j)
#pragma unroll
      for (int d = 0; d < 8; ++d)
        if (d < f.D) {
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) acc[j][d] += __shfl_xor_sync(0xffffffffu, acc[j][d], off);
        }
    if (lane < kQinRows && r0 + lane < rows) {
      float z[8];
#pragma unroll
      for (int j = 0; j < kQinRows; ++j)
        if (j == lane)
#pragma unroll
          for (int d = 0; d < 8; ++d) z[d] = d < f.D ? static_cast<float>(acc[j][d] + static_cast<double>(bias[d])) : 0.0f;
      const int64_t r = r0 + lane;
      if (z_out)
        for (int d = 0; d < f.D; ++d) z_out[r * f.D + d] = z[d];
      idx[r] = fsq_index(z, f, nullptr);
    }
  }
}

__global__ void __launch_bounds__(256)
qin_fsq_kernel(const float* __restrict__ y, int64_t rows, int C, const float* __restrict__ w,
               const float* __restrict__ bias, const FsqDev f, int64_t* __restrict__ idx,
               float* __restrict__ z_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < rows; r += nwarps) {
    double acc[8];
#pragma unroll
    for (int d = 0; d < 8; ++d) acc[d] = 0.0;
    const float* yr = y + r * C;
    for (int c = lane * 4; c < C; c += 128) {
      const float4 v = *reinterpret_cast<const float4*>(yr + c);
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        if (d < f.D) {
          const float4 wv = *reinterpret_cast<const float4*>(w + static_cast<int64_t>(d) * C + c);
          acc[d] += static_cast<double>(v.x) * wv.x + static_cast<double>(v.y) * wv.y +
                    static_cast<double>(v.z) * wv.z + static_cast<double>(v.w) * wv.w;
        }
      }
    }
#pragma unroll
    for (int d = 0; d < 8; ++d)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], off);
    if (lane == 0) {
      float z[8];
      for (int d = 0; d < f.D; ++d) z[d] = static_cast<float>(acc[d] + static_cast<double>(bias[d]));
      if (z_out)
        for (int d = 0; d < f.D; ++d) z_out[r * f.D + d] = z[d];
      idx[r] = fsq_index(z, f, nullptr);
    }
  }
}

__global__ void fsq_quantize_kernel(const float* __restrict__ z, int64_t rows, const FsqDev f,
                                    int64_t* __restrict__ idx, float* __restrict__ codes) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float zz[8], cc[8];
    for (int d = 0; d < f.D; ++d) zz[d] = z[r * f.D + d];
    idx[r] = fsq_index(zz, f, codes ? cc : nullptr);
    if (codes)
      for (int d = 0; d < f.D; ++d) codes[r * f.D + d] = cc[d];
  }
}

// ---------------------------------------------------------------------------
// K8: code-table gather
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
code_gather_kernel(const int64_t* __restrict__ idx, int64_t rows, const float* __restrict__ table,
                   int n_codes, int C, __nv_bfloat16* __restrict__ out_bf16,
                   float* __restrict__ out_f32, int* __restrict__ bad) {
  const int c4 = C / 4;
  const int64_t total = rows * c4;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / c4;
    const int c = static_cast<int>(i - r * c4) * 4;
    const int64_t k = idx[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k >= 0 && k < n_codes) {
      v = *reinterpret_cast<const float4*>(table + k * C + c);
    } else if (bad != nullptr && c == 0) {
      atomicExch(bad, 1);
    }
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * C + c) = v;
    if (out_bf16) {
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out_bf16 + r * C + c) = u;
    }
  }
}

// ---------------------------------------------------------------------------
// masks
// ---------------------------------------------------------------------------
__global__ void sequence_mask_kernel(const int64_t* __restrict__ lengths, int B, int T,
                                     uint8_t* __restrict__ mask) {
  const int64_t total = static_cast<int64_t>(B) * T;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / T);
    const int t = static_cast<int>(i - static_cast<int64_t>(b) * T);
    mask[i] = t >= lengths[b] ? 1 : 0;
  }
}

// one block per batch element; levels are processed in order with block syncs
__global__ void refiner_masks_kernel(const uint8_t* __restrict__ mask, int B, int T, int T8, int depth,
                                     uint8_t* __restrict__ down, uint8_t* __restrict__ up) {
  const int b = blockIdx.x;
  int64_t off = 0;
  // level 0
  for (int t = threadIdx.x; t < T8; t += blockDim.x)
    down[static_cast<int64_t>(b) * T8 + t] =
        (t < T) ? (mask != nullptr ? mask[static_cast<int64_t>(b) * T + t] : 0) : 1;
  __syncthreads();
  int64_t offs[16];
  offs[0] = 0;
  for (int l = 1; l <= depth; ++l) {
    const int Hp = T8 >> (l - 1), Hc = T8 >> l;
    offs[l] = offs[l - 1] + static_cast<int64_t>(B) * Hp;
    const uint8_t* src = down + offs[l - 1] + static_cast<int64_t>(b) * Hp;
    uint8_t* dst = down + offs[l] + static_cast<int64_t>(b) * Hc;
    for (int t = threadIdx.x; t < Hc; t += blockDim.x) dst[t] = src[2 * t] | src[2 * t + 1];
    __syncthreads();
  }
  (void)off;
  {
    const int Hd = T8 >> depth;
    const uint8_t* src = down + offs[depth] + static_cast<int64_t>(b) * Hd;
    uint8_t* dst = up + offs[depth] + static_cast<int64_t>(b) * Hd;
    for (int t = threadIdx.x; t < Hd; t += blockDim.x) dst[t] = src[t];
    __syncthreads();
  }
  for (int l = depth - 1; l >= 0; --l) {
    const int Hc = T8 >> l, Hn = T8 >> (l + 1);
    const uint8_t* src = up + offs[l + 1] + static_cast<int64_t>(b) * Hn;
    uint8_t* dst = up + offs[l] + static_cast<int64_t>(b) * Hc;
    for (int t = threadIdx.x; t < Hc; t += blockDim.x) dst[t] = src[t >> 1];
    __syncthreads();
  }
}

// zero the rows that are padded under `mask_new` but were valid under `mask_old` (the up-path mask
// of the refiner is coarser than the down-path mask the skip tensor was written with)
__global__ void __launch_bounds__(256)
zero_rows_kernel(uint4* __restrict__ x, const uint8_t* __restrict__ mask_new,
                 const uint8_t* __restrict__ mask_old, int64_t row_u4) {
  const int64_t row = blockIdx.x;
  if (mask_new[row] == 0 || (mask_old != nullptr && mask_old[row] != 0)) return;
  uint4* p = x + row * row_u4;
  for (int64_t i = threadIdx.x; i < row_u4; i += blockDim.x) p[i] = make_uint4(0, 0, 0, 0);
}

// ---------------------------------------------------------------------------
// K13: pooling / upsample+concat (bf16, 8 channels per thread)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint4 avg_bf16x8(const uint4& a, const uint4& b) {
  const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
  const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
  uint4 r;
  uint32_t* pr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 fa = __bfloat1622float2(pa[e]);
    const float2 fb = __bfloat1622float2(pb[e]);
    pr[e] = pack_bf16x2(0.5f * (fa.x + fb.x), 0.5f * (fa.y + fb.y));
  }
  return r;
}

__global__ void __launch_bounds__(256)
avgpool_mask_kernel(const uint4* __restrict__ x, uint4* __restrict__ y,
                    const uint8_t* __restrict__ mask_out, int B, int Ho, int F, int C8) {
  const int64_t rowlen = static_cast<int64_t>(F) * C8;          // uint4 per (b, h) row
  const int64_t total = static_cast<int64_t>(B) * Ho * rowlen;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / rowlen;                              // b * Ho + ho
    const int64_t in = i - row * rowlen;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (mask_out == nullptr || mask_out[row] == 0) {
      const uint4 a = x[(2 * row) * rowlen + in];
      const uint4 b = x[(2 * row + 1) * rowlen + in];
      r = avg_bf16x8(a, b);
    }
    y[i] = r;
  }
}

__global__ void __launch_bounds__(256)
upcat_mask_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip, uint4* __restrict__ y,
                  const uint8_t* __restrict__ mask_out, int B, int H, int F, int Cx8, int Cs8) {
  const int Cy8 = Cx8 + Cs8;
  const int64_t total = static_cast<int64_t>(B) * H * F * Cy8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % Cy8);
    const int64_t pix = i / Cy8;                                 // (b*H + h)*F + f
    const int f = static_cast<int>(pix % F);
    const int64_t row = pix / F;                                 // b*H + h
    uint4 r = make_uint4(0, 0, 0, 0);
    if (mask_out == nullptr || mask_out[row] == 0) {
      if (c < Cx8) {
        const int64_t b = row / H;
        const int h = static_cast<int>(row - b * H);
        const int64_t src_row = b * (H / 2) + (h >> 1);
        r = x[(src_row * F + f) * Cx8 + c];
      } else {
        r = skip[pix * Cs8 + (c - Cx8)];
      }
    }
    y[i] = r;
  }
}

// fp32-grade decoder mode: activations are two fp16 terms [h0 | h1] along channels (2C per pixel).
// AvgPool2d((2,1)) in fp32 on h0 + h1 of both rows (the reference pools in fp32, preencoder.py:112), re-split.
__global__ void __launch_bounds__(256)
avgpool_mask_split_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                          const uint8_t* __restrict__ mask_out, int B, int Ho, int F, int C) {
  const int C4 = C / 4;
  const int64_t rowlen = static_cast<int64_t>(F) * C4;          // 4-channel groups per (b, h) row
  const int64_t total = static_cast<int64_t>(B) * Ho * rowlen;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t row = i / rowlen;
    const int64_t in = i - row * rowlen;
    const int64_t f = in / C4;
    const int c = static_cast<int>(in - f * C4) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (mask_out == nullptr || mask_out[row] == 0) {
      const uint16_t* a = x + ((2 * row) * F + f) * (2 * C) + c;
      const uint16_t* b = x + ((2 * row + 1) * F + f) * (2 * C) + c;
      const uint2 a0 = *reinterpret_cast<const uint2*>(a), a1 = *reinterpret_cast<const uint2*>(a + C);
      const uint2 b0 = *reinterpret_cast<const uint2*>(b), b1 = *reinterpret_cast<const uint2*>(b + C);
      const __half2* pa0 = reinterpret_cast<const __half2*>(&a0);
      const __half2* pa1 = reinterpret_cast<const __half2*>(&a1);
      const __half2* pb0 = reinterpret_cast<const __half2*>(&b0);
      const __half2* pb1 = reinterpret_cast<const __half2*>(&b1);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float2 fa0 = __half22float2(pa0[e]), fa1 = __half22float2(pa1[e]);
        const float2 fb0 = __half22float2(pb0[e]), fb1 = __half22float2(pb1[e]);
        v[2 * e] = 0.5f * ((fa0.x + fa1.x) + (fb0.x + fb1.x));
        v[2 * e + 1] = 0.5f * ((fa0.y + fa1.y) + (fb0.y + fb1.y));
      }
    }
    store_terms4(y + (row * F + f) * (2 * C) + c, C, 1, v);
  }
}

// nearest Upsample((2,1)) of x + concat with skip + mask on split tensors:
// [x_h0 | x_h1] (2Cx), [s_h0 | s_h1] (2Cs) -> [x_h0 | s_h0 | x_h1 | s_h1] (2(Cx+Cs)); pure 16-byte copies.
__global__ void __launch_bounds__(256)
upcat_mask_split_kernel(const uint4* __restrict__ x, const uint4* __restrict__ skip, uint4* __restrict__ y,
                        const uint8_t* __restrict__ mask_out, int B, int H, int F, int Cx8, int Cs8) {
  const int Cy8 = Cx8 + Cs8;
  const int64_t total = static_cast<int64_t>(B) * H * F * 2 * Cy8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c2 = static_cast<int>(i % (2 * Cy8));
    const int term = c2 / Cy8, c = c2 - term * Cy8;
    const int64_t pix = i / (2 * Cy8);
    const int f = static_cast<int>(pix % F);
    const int64_t row = pix / F;
    uint4 r = make_uint4(0, 0, 0, 0);
    if (mask_out == nullptr || mask_out[row] == 0) {
      if (c < Cx8) {
        const int64_t b = row / H;
        const int h = static_cast<int>(row - b * H);
        const int64_t src_row = b * (H / 2) + (h >> 1);
        r = x[(src_row * F + f) * (2 * Cx8) + term * Cx8 + c];
      } else {
        r = skip[pix * (2 * Cs8) + term * Cs8 + (c - Cx8)];
      }
    }
    y[i] = r;
  }
}

// ---------------------------------------------------------------------------
// K12a: refiner.pre.conv1 (1 -> C) + APTx.  Thread = (pixel, 8 channels).
// ---------------------------------------------------------------------------
constexpr int kStemRows = 8;

// kSplit: write the two fp16 terms [h0 | h1] (2C channels per pixel) of the fp32-grade decoder mode instead of bf16.
template <bool kFast, bool kSplit>
__global__ void __launch_bounds__(256)
refiner_stem_kernel(const float* __restrict__ r, const uint8_t* __restrict__ mask, int B, int T, int T8,
                    int F, int C, const float* __restrict__ w, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ y) {
  // Block = kStemRows output rows of one batch element: the masked / zero-padded input rows go to
  // shared memory; thread = (pixel, group of 8 output channels) with that group's 72 weights + 8
  // biases in registers (group = threadIdx.x % (C/8), constant because blockDim.x % (C/8) == 0).
  extern __shared__ float rows[];               // [kStemRows + 2][F + 2], columns shifted by one
  const int t0 = blockIdx.x * kStemRows, b = blockIdx.y;
  const int Fp = F + 2;
  for (int i = threadIdx.x; i < (kStemRows + 2) * Fp; i += blockDim.x) {
    const int rr = i / Fp, ff = i - rr * Fp - 1;
    const int tt = t0 + rr - 1;
    float v = 0.0f;
    if (ff >= 0 && ff < F && tt >= 0 && tt < T && (mask == nullptr || mask[b * T + tt] == 0))
      v = r[(static_cast<int64_t>(b) * T + tt) * F + ff];
    rows[i] = v;
  }
  const int C8 = C / 8;
  const int cg = threadIdx.x % C8;
  float wr[8][9], br[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    br[e] = bias[cg * 8 + e];
#pragma unroll
    for (int k = 0; k < 9; ++k) wr[e][k] = w[(cg * 8 + e) * 9 + k];
  }
  __syncthreads();
  const int nrow = min(kStemRows, T8 - t0);
  const int ldy = kSplit ? 2 * C : C;
  __nv_bfloat16* ybase = y + (static_cast<int64_t>(b) * T8 + t0) * F * ldy;
  // thread = (pixel lane, channel group); pixels advance by blockDim.x / C8 per iteration with (lt, f) kept
  // incrementally (no integer divisions in the loop: they cost as much as the nine FMAs per output)
  const int ppb = blockDim.x / C8;
  int px = threadIdx.x / C8;
  int lt = px / F, f = px - lt * F;
  for (; px < nrow * F; px += ppb) {
    float in[9];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr)
#pragma unroll
      for (int df = 0; df < 3; ++df) in[rr * 3 + df] = rows[(lt + rr) * Fp + f + df];
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float acc = br[e];
#pragma unroll
      for (int k = 0; k < 9; ++k) acc = fmaf(wr[e][k], in[k], acc);
      v[e] = aptx<kFast>(acc, 1.0f, 0.5f);
    }
    if (kSplit) {
      uint16_t* o = reinterpret_cast<uint16_t*>(ybase) + static_cast<int64_t>(px) * ldy + cg * 8;
      const float lo4[4] = {v[0], v[1], v[2], v[3]}, hi4[4] = {v[4], v[5], v[6], v[7]};
      store_terms4(o, C, 1, lo4);
      store_terms4(o + 4, C, 1, hi4);
    } else {
      *reinterpret_cast<uint4*>(ybase + static_cast<int64_t>(px) * C + cg * 8) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
    f += ppb;
    while (f >= F) { f -= F; ++lt; }
  }
}

// ---------------------------------------------------------------------------
// K12b/K14: refiner.post (C -> 1) + crop + mask + reproj + x_recon add.
// Block = kTailT frames of one batch element.
// ---------------------------------------------------------------------------
constexpr int kTailT = 8;      // frames per block (24 frames / 8 per thread measured 8 % slower: fewer blocks in flight to overlap the two phases)
constexpr int kTailFr = 4;     // frames per thread in the reproj phase

// taps: (B, T8, F, ldp) fp32, channel k = 3*(dt+1)+(df+1) holds sum_c x[.., c] * w[k][c] (a 1x1
// tcgen05 GEMM, C -> 9).  post(x)[t,f] = bias + sum_k taps[t+dt, f+df, k].  ldp == 1: taps is post(x) itself
// (without the bias), produced by the 3x3 halo kernel with one output channel: 4 bytes per pixel instead of 48.
// Phase 2 (reproj, F -> M per frame) was one global load per FMA; now a thread owns one output channel m of four
// frames and walks f four at a time: 4 coalesced weight loads + 4 shared-memory float4 reads feed 16 FMAs.
__global__ void __launch_bounds__(256)
refiner_tail_kernel(const float* __restrict__ taps, int ldp, const uint8_t* __restrict__ mask, int T, int T8,
                    int F, float bias, const float* __restrict__ reproj_t, int M,
                    const float* __restrict__ r, float* __restrict__ out) {
  extern __shared__ __align__(16) float osm[];  // o[kTailT][Fp], Fp = F rounded up to 4 (zero tail)
  const int Fp = (F + 3) & ~3;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kTailT;
  for (int i = threadIdx.x; i < kTailT * Fp; i += blockDim.x) {
    const int lt = i / Fp, f = i - lt * Fp;
    const int t = t0 + lt;
    float acc = 0.0f;
    if (f < F && t < T && !(mask != nullptr && mask[static_cast<int64_t>(b) * T + t] != 0)) {
      acc = bias;
      if (ldp == 1) {             // post already contracted over its 3x3 taps (the halo conv kernel, C -> 1)
        osm[i] = acc + __ldg(taps + (static_cast<int64_t>(b) * T8 + t) * F + f);
        continue;
      }
      if (ldp == 4) {             // three row sums per pixel (channel dt + 1 = sum over df and c of row t's taps): add down T
        const float4* rec = reinterpret_cast<const float4*>(taps) + (static_cast<int64_t>(b) * T8 + t) * F + f;
        if (t > 0) acc += __ldg(rec - F).x;          // row t-1 evaluated with the dt = -1 weights
        acc += __ldg(rec).y;
        if (t + 1 < T8) acc += __ldg(rec + F).z;
        osm[i] = acc;
        continue;
      }
#pragma unroll
      for (int dt = -1; dt <= 1; ++dt) {
        const int tt = t + dt;
        if (tt < 0 || tt >= T8) continue;
#pragma unroll
        for (int df = -1; df <= 1; ++df) {
          const int ff = f + df;
          if (ff < 0 || ff >= F) continue;
          acc += __ldg(taps + ((static_cast<int64_t>(b) * T8 + tt) * F + ff) * ldp + (dt + 1) * 3 + (df + 1));
        }
      }
    }
    osm[i] = acc;                                  // masked rows -> 0 (preencoder.py:198)
  }
  __syncthreads();
  // a thread owns one output channel m of kTailFr consecutive frames: 4 coalesced weight loads + kTailFr shared-memory
  // float4 reads feed 4 kTailFr FMAs (summation order per output unchanged: f ascending)
  for (int idx = threadIdx.x; idx < M * (kTailT / kTailFr); idx += blockDim.x) {
    const int tg = idx / M, m = idx - tg * M;
    const float* o0 = osm + (tg * kTailFr) * Fp;
    float acc[kTailFr];
#pragma unroll
    for (int j = 0; j < kTailFr; ++j) acc[j] = 0.0f;
    for (int f = 0; f < Fp; f += 4) {
      float w[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) w[e] = (f + e < F) ? __ldg(reproj_t + static_cast<int64_t>(f + e) * M + m) : 0.0f;
#pragma unroll
      for (int j = 0; j < kTailFr; ++j) {
        const float4 ov = *reinterpret_cast<const float4*>(o0 + j * Fp + f);
        acc[j] = fmaf(w[0], ov.x, acc[j]);
        acc[j] = fmaf(w[1], ov.y, acc[j]);
        acc[j] = fmaf(w[2], ov.z, acc[j]);
        acc[j] = fmaf(w[3], ov.w, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < kTailFr; ++j) {
      const int t = t0 + tg * kTailFr + j;
      if (t < T) {
        const int64_t row = static_cast<int64_t>(b) * T + t;
        out[row * M + m] = r[row * F + m] + acc[j];  // x_post = x_recon + residual (preencoder.py:499)
      }
    }
  }
}

}  // namespace mq

using namespace mq;

#define STREAM(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mq_split_bf16(const float* x, void* out, int64_t rows, int C, int nterms,
                             mq_stream_t stream) {
  MQ_REQUIRE(x && out && rows > 0 && C > 0 && C % 4 == 0, "mq_split_bf16: bad args (C=%d must be a multiple of 4)", C);
  MQ_REQUIRE(nterms >= 1 && nterms <= 3, "mq_split_bf16: nterms=%d", nterms);
  split_bf16_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, STREAM(stream)>>>(
      x, reinterpret_cast<uint16_t*>(out), rows, C, nterms);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_convblock2d(const mq_cb2d_params* p, mq_stream_t stream) {
  MQ_REQUIRE(p && p->x && p->dw && p->pw, "mq_convblock2d: null argument");
  MQ_REQUIRE(p->B > 0 && p->T > 0 && p->C > 0, "mq_convblock2d: bad shape");
  MQ_REQUIRE(p->out_f32 || p->out_bf16 || p->out_split, "mq_convblock2d: no output");
  const int ctiles = (p->C + kCbC - 1) / kCbC, ttiles = (p->T + kCbT - 1) / kCbT;
  const long long blocks = 1LL * p->B * ttiles * ctiles;
  MQ_REQUIRE(blocks < (1LL << 31), "mq_convblock2d: grid too large");
  const size_t smem = static_cast<size_t>(p->C) * sizeof(float4);
  MQ_REQUIRE(smem <= 40 * 1024, "mq_convblock2d: C=%d too large", p->C);
  auto* ob = reinterpret_cast<__nv_bfloat16*>(p->out_bf16);
  auto* os = reinterpret_cast<uint16_t*>(p->out_split);
  const int sk = p->split_kind;
  const float4* pw = reinterpret_cast<const float4*>(p->pw);
  if (p->table != nullptr) {
    MQ_REQUIRE(p->table_n > 0 && p->table_inv_h > 0.0f, "mq_convblock2d: bad table");
    MQ_REQUIRE(p->C % 4 == 0, "mq_convblock2d: table mode needs C %% 4 == 0");
    const int ct2 = (p->C + kCtC - 1) / kCtC, tt2 = (p->T + kCtT - 1) / kCtT;
    const long long nb = 1LL * p->B * tt2 * ct2;
    MQ_REQUIRE(nb < (1LL << 31), "mq_convblock2d: grid too large");
    const float4* tb = reinterpret_cast<const float4*>(p->table);
#define LAUNCH_CT(FAST, INBF)                                                                     \
  convblock2d_table_kernel<FAST, INBF><<<static_cast<unsigned>(nb), 256, 0, STREAM(stream)>>>(    \
      p->x, p->B, p->T, p->C, p->dw, pw, p->bout, p->row_mask, tb, p->table_n, p->table_off,      \
      p->table_inv_h, p->out_f32, ob, os, sk)
    if (p->fast_tanh) {
      if (p->x_is_bf16) LAUNCH_CT(true, true); else LAUNCH_CT(true, false);
    } else {
      if (p->x_is_bf16) LAUNCH_CT(false, true); else LAUNCH_CT(false, false);
    }
#undef LAUNCH_CT
    MQ_CUDA_OK(cudaGetLastError());
    return 0;
  }
  dim3 grid(static_cast<unsigned>(blocks));
#define LAUNCH_CB(FAST, INBF)                                                                     \
  convblock2d_kernel<FAST, INBF><<<grid, 256, smem, STREAM(stream)>>>(                            \
      p->x, p->B, p->T, p->C, p->dw, pw, p->bout, p->row_mask, p->out_f32, ob, os, sk)
  if (p->fast_tanh) {
    if (p->x_is_bf16) LAUNCH_CB(true, true); else LAUNCH_CB(true, false);
  } else {
    if (p->x_is_bf16) LAUNCH_CB(false, true); else LAUNCH_CB(false, false);
  }
#undef LAUNCH_CB
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_cam_chunks(int T) { return (T + kCamRows - 1) / kCamRows; }

extern "C" int mq_cam_reduce(const float* o, const uint8_t* row_mask, int B, int T, int C, float* part,
                             mq_stream_t stream) {
  MQ_REQUIRE(o && part && B > 0 && T > 0 && C > 0, "mq_cam_reduce: bad args");
  const int nchunk = mq_cam_chunks(T);
  cam_reduce_kernel<<<dim3(nchunk, B), 256, 0, STREAM(stream)>>>(o, row_mask, T, C, nchunk, part);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_cam_gate(const float* part, const uint8_t* row_mask, int B, int T, int C, int R,
                           const float* w0, const float* b0, const float* w2, const float* b2,
                           float* gate, mq_stream_t stream) {
  MQ_REQUIRE(part && w0 && b0 && w2 && b2 && gate && B > 0 && C > 0 && R > 0, "mq_cam_gate: bad args");
  const size_t smem = (2 * static_cast<size_t>(C) + 2 * R) * sizeof(float);
  cam_gate_kernel<<<B, 256, smem, STREAM(stream)>>>(part, row_mask, T, C, R, mq_cam_chunks(T), w0, b0,
                                                    w2, b2, gate);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_cbam_apply(const mq_cbam_apply_params* p, mq_stream_t stream) {
  MQ_REQUIRE(p && p->o && p->gate && p->res && p->sam_w, "mq_cbam_apply: null argument");
  MQ_REQUIRE(p->C % 4 == 0 && p->B > 0 && p->T > 0, "mq_cbam_apply: bad shape (C %% 4)");
  MQ_REQUIRE(p->out_f32 || p->out_bf16 || p->out_split, "mq_cbam_apply: no output");
  const size_t smem = static_cast<size_t>(p->C) * sizeof(float);
  dim3 grid((p->T + kSamT - 1) / kSamT, p->B);
  cbam_apply_kernel<<<grid, 256, smem, STREAM(stream)>>>(
      p->o, p->gate, p->res, p->row_mask, p->T, p->C, p->sam_w, p->beta, p->gamma, p->out_f32,
      reinterpret_cast<__nv_bfloat16*>(p->out_bf16), reinterpret_cast<uint16_t*>(p->out_split), p->split_kind);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

static int fill_fsq(const mq_fsq_params* f, FsqDev* d) {
  if (!f || f->D < 1 || f->D > 8) return 1;
  memset(d, 0, sizeof(*d));
  d->D = f->D;
  for (int i = 0; i < f->D; ++i) {
    d->half_l[i] = f->half_l[i]; d->shift[i] = f->shift[i]; d->offset[i] = f->offset[i];
    d->half_w[i] = f->half_w[i]; d->basis[i] = f->basis[i]; d->levels[i] = f->levels[i];
  }
  return 0;
}

extern "C" int mq_qin_fsq(const float* y, int64_t rows, int C, const float* w, const float* b,
                          const mq_fsq_params* fsq, int64_t* idx, float* z_out, mq_stream_t stream) {
  MQ_REQUIRE(y && w && b && idx && rows > 0 && C > 0 && C % 4 == 0, "mq_qin_fsq: bad args");
  FsqDev f;
  MQ_REQUIRE(fill_fsq(fsq, &f) == 0, "mq_qin_fsq: bad fsq params");
  const size_t wbytes = static_cast<size_t>(f.D) * C * sizeof(double);
  if (wbytes <= 96 * 1024) {
    int dev = 0, sms = 0;
    MQ_CUDA_OK(cudaGetDevice(&dev));
    MQ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MQ_CUDA_OK(cudaFuncSetAttribute(qin_fsq_kernel4, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(wbytes)));
    const int64_t groups = (rows + kQinRows - 1) / kQinRows;
    int64_t blocks = (groups + 7) / 8;                       // 8 warps per CTA, one 4-frame group per warp per pass
    const int64_t cap = static_cast<int64_t>(sms) * 2 * 4;   // a few waves of resident CTAs: the weight conversion amortises
    if (blocks > cap) blocks = cap;
    qin_fsq_kernel4<<<static_cast<unsigned>(blocks), 256, wbytes, STREAM(stream)>>>(y, rows, C, w, b, f, idx, z_out);
  } else {
    qin_fsq_kernel<<<grid_for(rows, 8), 256, 0, STREAM(stream)>>>(y, rows, C, w, b, f, idx, z_out);
  }
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_fsq_quantize(const float* z, int64_t rows, const mq_fsq_params* fsq, int64_t* idx,
                               float* codes_out, mq_stream_t stream) {
  MQ_REQUIRE(z && idx && rows > 0, "mq_fsq_quantize: bad args");
  FsqDev f;
  MQ_REQUIRE(fill_fsq(fsq, &f) == 0, "mq_fsq_quantize: bad fsq params");
  fsq_quantize_kernel<<<grid_for(rows, 256), 256, 0, STREAM(stream)>>>(z, rows, f, idx, codes_out);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_code_gather(const int64_t* idx, int64_t rows, const float* table, int n_codes, int C,
                              void* out_bf16, float* out_f32, int* bad, mq_stream_t stream) {
  MQ_REQUIRE(idx && table && rows > 0 && C > 0 && C % 4 == 0 && n_codes > 0, "mq_code_gather: bad args");
  MQ_REQUIRE(out_bf16 || out_f32, "mq_code_gather: no output");
  code_gather_kernel<<<grid_for(rows * (C / 4), 256), 256, 0, STREAM(stream)>>>(
      idx, rows, table, n_codes, C, reinterpret_cast<__nv_bfloat16*>(out_bf16), out_f32, bad);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_sequence_mask(const int64_t* lengths, int B, int T, uint8_t* mask, mq_stream_t stream) {
  MQ_REQUIRE(lengths && mask && B > 0 && T > 0, "mq_sequence_mask: bad args");
  sequence_mask_kernel<<<grid_for(static_cast<int64_t>(B) * T, 256), 256, 0, STREAM(stream)>>>(lengths, B, T, mask);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_refiner_masks(const uint8_t* mask, int B, int T, int depth, uint8_t* down, uint8_t* up,
                                mq_stream_t stream) {
  MQ_REQUIRE(down && up && B > 0 && T > 0 && depth >= 0 && depth < 15, "mq_refiner_masks: bad args");
  const int mult = 1 << depth;
  const int T8 = (T + mult - 1) / mult * mult;
  refiner_masks_kernel<<<B, 256, 0, STREAM(stream)>>>(mask, B, T, T8, depth, down, up);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_zero_rows(void* x, const uint8_t* mask_new, const uint8_t* mask_old, int64_t rows,
                            int64_t row_bytes, mq_stream_t stream) {
  MQ_REQUIRE(x && mask_new && rows > 0 && rows < (1LL << 31) && row_bytes > 0 && row_bytes % 16 == 0,
             "mq_zero_rows: bad args");
  zero_rows_kernel<<<static_cast<unsigned>(rows), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<uint4*>(x), mask_new, mask_old, row_bytes / 16);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_avgpool_mask(const void* x, void* y, const uint8_t* mask_out, int B, int H, int F, int C,
                               mq_stream_t stream) {
  MQ_REQUIRE(x && y && B > 0 && H > 0 && H % 2 == 0 && F > 0 && C % 8 == 0, "mq_avgpool_mask: bad args");
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * F * (C / 8);
  avgpool_mask_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), mask_out, B, H / 2, F, C / 8);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_upcat_mask(const void* x, const void* skip, void* y, const uint8_t* mask_out, int B, int H,
                             int F, int Cx, int Cs, mq_stream_t stream) {
  MQ_REQUIRE(x && skip && y && B > 0 && H > 0 && H % 2 == 0 && F > 0 && Cx % 8 == 0 && Cs % 8 == 0,
             "mq_upcat_mask: bad args");
  const int64_t total = static_cast<int64_t>(B) * H * F * ((Cx + Cs) / 8);
  upcat_mask_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(skip),
      reinterpret_cast<uint4*>(y), mask_out, B, H, F, Cx / 8, Cs / 8);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_avgpool_mask_split(const void* x, void* y, const uint8_t* mask_out, int B, int H, int F, int C,
                                     mq_stream_t stream) {
  MQ_REQUIRE(x && y && B > 0 && H > 0 && H % 2 == 0 && F > 0 && C % 4 == 0, "mq_avgpool_mask_split: bad args");
  const int64_t total = static_cast<int64_t>(B) * (H / 2) * F * (C / 4);
  avgpool_mask_split_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const uint16_t*>(x), reinterpret_cast<uint16_t*>(y), mask_out, B, H / 2, F, C);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_upcat_mask_split(const void* x, const void* skip, void* y, const uint8_t* mask_out, int B, int H,
                                   int F, int Cx, int Cs, mq_stream_t stream) {
  MQ_REQUIRE(x && skip && y && B > 0 && H > 0 && H % 2 == 0 && F > 0 && Cx % 8 == 0 && Cs % 8 == 0,
             "mq_upcat_mask_split: bad args");
  const int64_t total = static_cast<int64_t>(B) * H * F * 2 * ((Cx + Cs) / 8);
  upcat_mask_split_kernel<<<grid_for(total, 256), 256, 0, STREAM(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<const uint4*>(skip), reinterpret_cast<uint4*>(y), mask_out,
      B, H, F, Cx / 8, Cs / 8);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_refiner_stem(const float* r, const uint8_t* mask, int B, int T, int T8, int F, int C,
                               const float* w, const float* b, int fast_tanh, void* y, mq_stream_t stream) {
  MQ_REQUIRE(r && w && b && y && B > 0 && T > 0 && T8 >= T && F > 0 && C % 8 == 0 && C / 8 <= 256,
             "mq_refiner_stem: bad args");
  MQ_REQUIRE(B <= 65535, "mq_refiner_stem: B too large for one launch");
  const int threads = 256 / (C / 8) * (C / 8);       // multiple of C/8 so a thread keeps its channel group
  const size_t smem = (kStemRows + 2) * static_cast<size_t>(F + 2) * sizeof(float);
  dim3 grid((T8 + kStemRows - 1) / kStemRows, B);
  if (fast_tanh)
    refiner_stem_kernel<true, false><<<grid, threads, smem, STREAM(stream)>>>(r, mask, B, T, T8, F, C, w, b, reinterpret_cast<__nv_bfloat16*>(y));
  else
    refiner_stem_kernel<false, false><<<grid, threads, smem, STREAM(stream)>>>(r, mask, B, T, T8, F, C, w, b, reinterpret_cast<__nv_bfloat16*>(y));
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_refiner_stem_split(const float* r, const uint8_t* mask, int B, int T, int T8, int F, int C,
                                     const float* w, const float* b, void* y, mq_stream_t stream) {
  MQ_REQUIRE(r && w && b && y && B > 0 && T > 0 && T8 >= T && F > 0 && C % 8 == 0 && C / 8 <= 256,
             "mq_refiner_stem_split: bad args");
  MQ_REQUIRE(B <= 65535, "mq_refiner_stem_split: B too large for one launch");
  const int threads = 256 / (C / 8) * (C / 8);
  const size_t smem = (kStemRows + 2) * static_cast<size_t>(F + 2) * sizeof(float);
  dim3 grid((T8 + kStemRows - 1) / kStemRows, B);
  refiner_stem_kernel<false, true><<<grid, threads, smem, STREAM(stream)>>>(r, mask, B, T, T8, F, C, w, b,
                                                                           reinterpret_cast<__nv_bfloat16*>(y));
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int mq_refiner_tail(const float* taps, int ldp, const uint8_t* mask, int B, int T, int T8, int F,
                               float bias, const float* reproj_t, int M, const float* r, float* out,
                               mq_stream_t stream) {
  MQ_REQUIRE(taps && reproj_t && r && out && B > 0 && T > 0 && T8 >= T && F >= M && (ldp >= 9 || ldp == 1 || ldp == 4),
             "mq_refiner_tail: bad args (ldp = 9.. tap planes, 4 = row sums, 1 = post already summed)");
  MQ_REQUIRE(ldp != 4 || (reinterpret_cast<uintptr_t>(taps) & 15) == 0, "mq_refiner_tail: row sums must be 16-byte aligned");
  const size_t smem = kTailT * static_cast<size_t>((F + 3) & ~3) * sizeof(float);
  dim3 grid((T + kTailT - 1) / kTailT, B);
  refiner_tail_kernel<<<grid, 256, smem, STREAM(stream)>>>(taps, ldp, mask, T, T8, F, bias, reproj_t, M, r, out);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}
