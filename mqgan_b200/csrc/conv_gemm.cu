// Implicit-GEMM convolution on tcgen05 / TMEM, fed by TMA.  One kernel serves every
// dense contraction of the PreEncoder path:
// a convolution is a sum over filter taps of shifted GEMMs; the shift is a TMA box
// coordinate and the zero padding is TMA's out-of-bounds fill, so no im2col buffer
// and no padded copy of the activation ever exists.
//
// Shape of one CTA tile: M = 128 output pixels (a bh x bw patch of the (H, W) grid
// of one batch element, pixels on TMEM lanes), N = bn output channels (TMEM
// columns), K = taps * nseg * kchunks * 64 bf16.  Persistent CTAs (one per SM),
// warp-specialised: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread
// tcgen05.mma issuer, warps 2-5 = epilogue (tcgen05.ld -> bias / mask / APTx /
// residual -> global).  The accumulator is double-buffered in TMEM (2 x 256
// columns) so the epilogue of tile i overlaps the main loop of tile i+1.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

constexpr int kBlockK = 64;                 // bf16 per K block = one 128-byte swizzle row
constexpr int kUmmaK = 16;                  // K per tcgen05.mma (kind::f16)
constexpr int kTileM = 128;                 // UMMA M, cta_group::1
constexpr int kThreads = 384;               // WG0: producer, MMA issuer (+2 idle warps); WG1-2: 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;
constexpr int kMaxAccBufs = 4;              // TMEM accumulator stages: 512 / (msub * bn) columns each, at most 4
constexpr int kATileBytes = kTileM * kBlockK * 2;   // 16 KiB
constexpr int kSmemBudget = 227 * 1024;
constexpr int kBiasSmemFloats = 1024;       // the whole (padded) bias vector is staged in shared memory once per CTA
constexpr int kStageBytesPerWarp = 32 * 64; // lean epilogue store staging: 32 pixels x 32 bf16 channels per warp
constexpr int kStageBytes = (kEpiThreads / 32) * kStageBytesPerWarp;

// Bench-only bottleneck probes (MQ_CONV_DEBUG bit mask: 1 no epilogue math/stores, 2 no MMA, 4 no TMA, 8 no stores)
// exist only in a -DMQ_CONV_PROBES build; release kernels carry no probe branches in their MMA / TMA / epilogue loops.
#ifdef MQ_CONV_PROBES
#define MQ_PROBE(a, bit) ((a).debug & (bit))
#else
#define MQ_PROBE(a, bit) 0
#endif

#ifdef MQ_CONV_PROBES
// cycle counters of the epilogue role, summed over all epilogue warps' lane 0 (probe build only; tools/conv_cycles.py):
// [0] waiting for the accumulator, [1] tcgen05.ld + wait, [2] epilogue body (math + stores), [3] whole tile loop,
// [4] tiles x warps, [5] MMA warp: waiting for a free accumulator, [6] MMA warp: waiting for operands, [7] MMA warp total
__device__ unsigned long long mq_probe_cycles[12];
#define MQ_CLK() clock64()
#define MQ_ACC(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&mq_probe_cycles[i], static_cast<unsigned long long>(v)); } while (0)
#else
#define MQ_CLK() 0LL
#define MQ_ACC(i, v) do { } while (0)
#endif

struct ConvArgs {
  int N, H, W;
  int tiles_h, tiles_w, tiles_n, num_tiles;
  int bh, bw, bn, cout;
  int msub;                       // 128-pixel sub-tiles per CTA tile (share one B tile)
  int nbuf;                       // TMEM accumulator stages: min(4, 512 / acc_stride)
  int acc_stride;                 // TMEM columns per accumulator stage: msub * bn rounded up to 32
  int taps, nseg, kchunks;
  int tap_dh[MQ_MAX_TAPS], tap_dw[MQ_MAX_TAPS], a_coff[MQ_MAX_SEGS];
  // fused nearest-upsample + concat (UpBlock): output rows 2*H, tiles carry a row parity
  int up_mode, up_taps, kchunks2, par_tiles, cout_pad;
  int tap_dh_odd[MQ_MAX_TAPS];
  int stages;
  int halo_nA, halo_nB, halo_slot_bytes, halo_tx_bytes;   // halo-tile variant (conv_halo_kernel)
  int tile_rows;                  // rows of H one tile covers: bh * msub (x2 for a CTA pair)
  int pair_boxb_off, pair_tx0, pair_tx1;   // conv_pair_kernel: byte offset of the second skip box, A bytes per chunk of group 0 / 1
  int pair_bgrp;                  // taps per weight-ring slot (one barrier round trip and one commit per slot)
  int mma_issuers;                // conv_pair_kernel: 1, or 2 = warps 1 and 2 of the leader each issue the MMAs of half the sub-tiles
  int b_resident;                 // conv_pair_kernel: the whole weight set of the (single) N tile stays in shared memory for the
                                  // kernel's lifetime - loaded once, no weight ring, no weight barriers in the tile loop
  int debug;                      // bench-only bottleneck probes (MQ_CONV_DEBUG): 1 no epilogue math/stores, 2 no MMA, 4 no TMA, 8 no stores
  uint32_t a_tx_bytes, b_tile_bytes;
  // epilogue
  const float* bias;
  const uint8_t* row_mask;
  int mask_pre, mask_post, act, res_mode;
  float beta, gamma;
  const void* res;
  int res_is_bf16, res_ld, res_coff;
  float* out_f32;
  int f32_ld, f32_coff;
  __nv_bfloat16* out_bf16;
  int bf16_ld, bf16_coff;
  __nv_bfloat16* out_pool;        // optional fused AvgPool2d((2,1)) of the bf16 output: (N, H/2, W, pool_ld)
  int pool_ld;
  uint16_t* out_split;
  int split_ld, split_seg, split_kind;
  int slim;                       // lean + APTx(tanh.approx) + no residual: epilogue_slim
  int slim16;                     // ... on the 16-epilogue-warp variant of conv_pair_kernel (bn = 64 / 128)
  int stage_out;                  // lean epilogue: transpose each warp's 32 px x 32 ch chunk through shared memory so a
                                  // store instruction writes 8 pixels x 64 contiguous bytes instead of 32 pixels x 16
  int op_f16;                     // operands are fp16 (f16x2 mode) instead of bf16
  float acc_scale;                // accumulator scale applied before the bias (undoes the f16x2 weight scale)
};

__device__ __forceinline__ void decode_tile(const ConvArgs& a, int tile, int& n_idx, int& h0,
                                            int& w0, int& n0, int& par) {
  int tn = tile % a.tiles_n;
  int tm = tile / a.tiles_n;
  par = tm % a.par_tiles;
  tm /= a.par_tiles;
  int tw = tm % a.tiles_w;
  int t2 = tm / a.tiles_w;
  int th = t2 % a.tiles_h;
  n_idx = t2 / a.tiles_h;
  h0 = th * a.tile_rows;
  w0 = tw * a.bw;
  n0 = tn * a.bn;
}

// ---------------------------------------------------------------------------
// epilogue bodies: one 32-column chunk of one accumulator row (pixel) per call
// ---------------------------------------------------------------------------
template <bool kFast>
__device__ __forceinline__ void epilogue_generic(const ConvArgs& a, const uint32_t (&v)[32], const float* bs,
                                                 int64_t pix, int co0, bool masked) {
    const int nvalid = min(32, a.cout - co0);
    float x[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaf(__uint_as_float(v[j]), a.acc_scale, bs[j]);

    float rr[32];
    if (a.res_mode != 0) {
      if (a.res_is_bf16) {
        const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.res) +
                                  pix * a.res_ld + a.res_coff + co0;
        if (nvalid == 32) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 u = *reinterpret_cast<const uint4*>(rp + 8 * g);
            const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float2 f = __bfloat1622float2(b2[e]);
              rr[8 * g + 2 * e] = f.x;
              rr[8 * g + 2 * e + 1] = f.y;
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rr[j] = j < nvalid ? __bfloat162float(rp[j]) : 0.0f;
        }
      } else {
        const float* rp = reinterpret_cast<const float*>(a.res) + pix * a.res_ld + a.res_coff + co0;
        if (nvalid == 32) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            float4 f = *reinterpret_cast<const float4*>(rp + 4 * g);
            rr[4 * g] = f.x; rr[4 * g + 1] = f.y; rr[4 * g + 2] = f.z; rr[4 * g + 3] = f.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) rr[j] = j < nvalid ? rp[j] : 0.0f;
        }
      }
    }
    if (a.res_mode == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += rr[j];
    }
    if (a.mask_pre && masked) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = 0.0f;
    }
    if (a.act) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = aptx<kFast>(x[j], a.beta, a.gamma);
    }
    if (a.res_mode == 2) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += rr[j];
    }
    if (a.mask_post && masked) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = 0.0f;
    }

    if (a.out_f32 != nullptr) {
      float* op = a.out_f32 + pix * a.f32_ld + a.f32_coff + co0;
      if (nvalid == 32) {
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<float4*>(op + 4 * g) =
              make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
      } else if ((nvalid & 3) == 0) {
        // a partial chunk whose width is a multiple of four channels still goes out as 16-byte stores (the refiner's
        // 64 -> 9 tap-plane GEMM is declared 12 wide for this: nine 4-byte stores per lane at a 48-byte pitch were
        // three times the L1 tag traffic and paced that kernel)
#pragma unroll
        for (int g = 0; g < 8; ++g)
          if (4 * g < nvalid)
            *reinterpret_cast<float4*>(op + 4 * g) = make_float4(x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) op[j] = x[j];
      }
    }
    if (a.out_bf16 != nullptr) {
      __nv_bfloat16* op = a.out_bf16 + pix * a.bf16_ld + a.bf16_coff + co0;
      if (nvalid == 32) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16x2(x[8 * g], x[8 * g + 1]);
          u.y = pack_bf16x2(x[8 * g + 2], x[8 * g + 3]);
          u.z = pack_bf16x2(x[8 * g + 4], x[8 * g + 5]);
          u.w = pack_bf16x2(x[8 * g + 6], x[8 * g + 7]);
          *reinterpret_cast<uint4*>(op + 8 * g) = u;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) op[j] = __float2bfloat16_rn(x[j]);
      }
    }
    if (a.out_split != nullptr) {
      uint16_t* op = a.out_split + pix * a.split_ld + co0;
      if (nvalid == 32) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float y4[4] = {x[4 * g], x[4 * g + 1], x[4 * g + 2], x[4 * g + 3]};
          store_terms4(op + 4 * g, a.split_seg, a.split_kind, y4);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (j < nvalid) store_terms1(op + j, a.split_seg, a.split_kind, x[j]);
      }
    }
}

// Lean variant for the bf16 decoder / refiner layers: cout % 32 == 0, bf16 output only, optional
// bf16 residual.  Dead branches of the generic body are compiled out (narrow layers are
// epilogue-issue bound, ncu profiles/ncu_conv_small_r01).
// Warp-private 32 px x 64 B staging tile of the lean epilogue (narrow layers).  A thread owns one accumulator row
// (pixel), so a direct 16-byte store per lane touches 32 different 128-byte lines per instruction: on the 64-channel
// refiner layers the L1 tag stage (one line per clock) took longer than the MMA main loop (ncu
// profiles/ncu_conv_pair_narrow_r02_summary.md: pre.conv2 tensor pipe 37 % active, l1tex 63 %).  Staged, lane l of pass
// `it` stores vector l % 4 of pixel it * 8 + l / 4: eight pixels x 64 contiguous bytes per instruction, 4x fewer tags.
// Slot swizzle v ^ ((px >> 1) & 3) keeps both the 16-byte writes (row per lane) and reads (4 lanes per row) conflict-free.
__device__ __forceinline__ void stage_write(uint8_t* st, int lane, const uint32_t (&u)[16]) {
#pragma unroll
  for (int g = 0; g < 4; ++g)
    *reinterpret_cast<uint4*>(st + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
        make_uint4(u[4 * g], u[4 * g + 1], u[4 * g + 2], u[4 * g + 3]);
}
__device__ __forceinline__ uint4 stage_read(const uint8_t* st, int px, int vec) {
  return *reinterpret_cast<const uint4*>(st + px * 64 + ((vec ^ ((px >> 1) & 3)) << 4));
}

template <bool kFast>
__device__ __forceinline__ void epilogue_lean(const ConvArgs& a, const uint32_t (&v)[32], const float* bs,
                                              int64_t pix, int co0, bool masked, bool valid, int64_t pix_pool,
                                              uint8_t* stage) {
  float x[32];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * g);
    x[4 * g] = __uint_as_float(v[4 * g]) + b4.x;
    x[4 * g + 1] = __uint_as_float(v[4 * g + 1]) + b4.y;
    x[4 * g + 2] = __uint_as_float(v[4 * g + 2]) + b4.z;
    x[4 * g + 3] = __uint_as_float(v[4 * g + 3]) + b4.w;
  }
  float rr[32];
  const bool has_res = a.res_mode != 0 && valid;       // rows / columns outside the image have no residual to read
  if (has_res) {
    const __nv_bfloat16* rp = reinterpret_cast<const __nv_bfloat16*>(a.res) + pix * a.res_ld + a.res_coff + co0;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const uint4 u = *reinterpret_cast<const uint4*>(rp + 8 * g);
      const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(b2[e]);
        rr[8 * g + 2 * e] = f.x;
        rr[8 * g + 2 * e + 1] = f.y;
      }
    }
    if (a.res_mode == 1) {
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] += rr[j];
    }
  }
  const bool zero_pre = a.mask_pre && masked;
  const bool zero_post = a.mask_post && masked;
  if (MQ_PROBE(a, 16)) {            // probe: activation without the MUFU op
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaf(a.gamma * x[j], a.beta * x[j], a.gamma * x[j]);
  } else if (a.act) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = aptx<kFast>(zero_pre ? 0.0f : x[j], a.beta, a.gamma);
  } else if (zero_pre) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = 0.0f;
  }
  if (has_res && a.res_mode == 2) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] += rr[j];
  }
  uint32_t u[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) u[j] = zero_post ? 0u : pack_bf16x2(x[2 * j], x[2 * j + 1]);
  if (MQ_PROBE(a, 8) && x[0] != 1234.5678f) valid = false;   // probe: math without the global stores
  const int lane_ = threadIdx.x & 31;
  if (stage != nullptr) {
    stage_write(stage, lane_, u);
    __syncwarp();
    const int vec = lane_ & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int px = it * 8 + (lane_ >> 2);
      const int64_t pix_o = __shfl_sync(0xffffffffu, pix, px);
      const bool valid_o = __shfl_sync(0xffffffffu, static_cast<int>(valid), px) != 0;
      const uint4 w4 = stage_read(stage, px, vec);
      if (valid_o) *reinterpret_cast<uint4*>(a.out_bf16 + pix_o * a.bf16_ld + a.bf16_coff + co0 + 8 * vec) = w4;
    }
    __syncwarp();                 // the tile is rewritten by the pooled words / the next chunk
  } else if (valid) {
    __nv_bfloat16* op = a.out_bf16 + pix * a.bf16_ld + a.bf16_coff + co0;
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(op + 8 * g) = make_uint4(u[4 * g], u[4 * g + 1], u[4 * g + 2], u[4 * g + 3]);
  }
  if (a.out_pool != nullptr) {
    // Fused AvgPool2d((2,1)) + pooled-mask fill (preencoder.py:111-114, :63-65, :96): with 8-pixel-wide
    // sub-tiles the two rows of a pooling pair sit 8 lanes apart.  The bf16-rounded outputs are
    // averaged (what a separate pass over the stored tensor computes, bit for bit), the pooled row is
    // padded when either source row is (max-pool of the mask), and the even row's lane writes it.
    const unsigned lane = threadIdx.x & 31u;
    const bool pm = (__shfl_xor_sync(0xffffffffu, static_cast<int>(masked), 8) != 0) || masked;
    uint32_t r[16];
    const __nv_bfloat162 half2 = __floats2bfloat162_rn(0.5f, 0.5f);
    const bool pz = pm && a.mask_post;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      // bf16 add (one rounding, the sum of two bf16 is exact in fp32) then an exact halving ==
      // rn_bf16(0.5f * (a + b)) of the separate pass, at two packed instructions per word
      const uint32_t o = __shfl_xor_sync(0xffffffffu, u[j], 8);
      const __nv_bfloat162 s2 = __hmul2(__hadd2(*reinterpret_cast<const __nv_bfloat162*>(&u[j]),
                                                *reinterpret_cast<const __nv_bfloat162*>(&o)), half2);
      r[j] = pz ? 0u : *reinterpret_cast<const uint32_t*>(&s2);
    }
    if (stage != nullptr) {
      // the 16 pooled pixels of this warp (lanes 0-7 and 16-23) go through rows 0..15 of the same tile
      const int prow = static_cast<int>((lane & 7u) | ((lane >> 4) << 3));
      if ((lane & 8u) == 0u) stage_write(stage, prow, r);
      __syncwarp();
      const int vec = lane_ & 3;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int px = it * 8 + (lane_ >> 2);                     // pooled pixel 0..15
        const int src = (px & 7) | ((px >> 3) << 4);              // the lane that owns it
        const int64_t pp_o = __shfl_sync(0xffffffffu, pix_pool, src);
        const bool valid_o = __shfl_sync(0xffffffffu, static_cast<int>(valid), src) != 0;
        const uint4 w4 = stage_read(stage, px, vec);
        if (valid_o) *reinterpret_cast<uint4*>(a.out_pool + pp_o * a.pool_ld + co0 + 8 * vec) = w4;
      }
      __syncwarp();
    } else if (valid && (lane & 8u) == 0u) {
      __nv_bfloat16* pp = a.out_pool + pix_pool * a.pool_ld + co0;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        *reinterpret_cast<uint4*>(pp + 8 * g) = make_uint4(r[4 * g], r[4 * g + 1], r[4 * g + 2], r[4 * g + 3]);
    }
  }
}

// Slim variant of the lean body for the refiner's ConvBlock layers (bias + APTx with tanh.approx, optional row mask,
// optional fused pool, no residual): what the 64- and 128-channel layers spend their time in.  On those layers the
// epilogue, not the MMA loop, sets the pace (probe build, tools/conv_probe2.sh: pre.conv2 0.36 ms of epilogue against
// 0.28 ms of MMA + loads), and the generic lean body issued 469 instructions per 32 x 32 chunk.  Here the row mask is
// folded into the APTx gain (g = 0 on padded rows: aptx(x) * 0, no per-element selects), beta == 1 costs nothing, the
// pooled mask is folded into the 0.5 of the average, and every flag is a template parameter: ~260 instructions.
template <bool kPool, bool kStage>
__device__ __forceinline__ void epilogue_slim(const ConvArgs& a, const uint32_t (&v)[32], const float* bs, int64_t pix,
                                              int co0, bool masked, bool valid, int64_t pix_pool, uint8_t* stage,
                                              int lane) {
  const float g = masked ? 0.0f : a.gamma;          // the launcher only selects this body when a mask flag is set or no mask is given
  float x[32];
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bs + 4 * q4);
    x[4 * q4] = __uint_as_float(v[4 * q4]) + b4.x;
    x[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) + b4.y;
    x[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) + b4.z;
    x[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) + b4.w;
  }
  uint32_t u[16];
  if (a.beta == 1.0f) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float t0 = tanh_fast(x[2 * j]), t1 = tanh_fast(x[2 * j + 1]);
      const float g0 = g * x[2 * j], g1 = g * x[2 * j + 1];
      u[j] = pack_bf16x2(fmaf(g0, t0, g0), fmaf(g1, t1, g1));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float t0 = tanh_fast(a.beta * x[2 * j]), t1 = tanh_fast(a.beta * x[2 * j + 1]);
      const float g0 = g * x[2 * j], g1 = g * x[2 * j + 1];
      u[j] = pack_bf16x2(fmaf(g0, t0, g0), fmaf(g1, t1, g1));
    }
  }
  if (kStage) {
    stage_write(stage, lane, u);
    __syncwarp();
    const int vec = lane & 3;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int px = it * 8 + (lane >> 2);
      const int64_t pix_o = __shfl_sync(0xffffffffu, pix, px);
      const bool valid_o = __shfl_sync(0xffffffffu, static_cast<int>(valid), px) != 0;
      const uint4 w4 = stage_read(stage, px, vec);
      if (valid_o) *reinterpret_cast<uint4*>(a.out_bf16 + pix_o * a.bf16_ld + a.bf16_coff + co0 + 8 * vec) = w4;
    }
    __syncwarp();
  } else if (valid) {
    __nv_bfloat16* op = a.out_bf16 + pix * a.bf16_ld + a.bf16_coff + co0;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4)
      *reinterpret_cast<uint4*>(op + 8 * q4) = make_uint4(u[4 * q4], u[4 * q4 + 1], u[4 * q4 + 2], u[4 * q4 + 3]);
  }
  if (kPool) {
    // see epilogue_lean: rows of a pooling pair sit 8 lanes apart; bf16 add (one rounding) then an exact scaling by 0.5,
    // or by 0 when either source row is padded (max-pooled mask)
    const bool pm = (__shfl_xor_sync(0xffffffffu, static_cast<int>(masked), 8) != 0) || masked;
    const float hf = pm ? 0.0f : 0.5f;
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(hf, hf);
    uint32_t r[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const uint32_t o = __shfl_xor_sync(0xffffffffu, u[j], 8);
      const __nv_bfloat162 s2 = __hmul2(__hadd2(*reinterpret_cast<const __nv_bfloat162*>(&u[j]),
                                                *reinterpret_cast<const __nv_bfloat162*>(&o)), h2);
      r[j] = *reinterpret_cast<const uint32_t*>(&s2);
    }
    if (kStage) {
      const int prow = (lane & 7) | ((lane >> 4) << 3);
      if ((lane & 8) == 0) stage_write(stage, prow, r);
      __syncwarp();
      const int vec = lane & 3;
#pragma unroll
      for (int it = 0; it < 2; ++it) {
        const int px = it * 8 + (lane >> 2);
        const int src = (px & 7) | ((px >> 3) << 4);
        const int64_t pp_o = __shfl_sync(0xffffffffu, pix_pool, src);
        const bool valid_o = __shfl_sync(0xffffffffu, static_cast<int>(valid), src) != 0;
        const uint4 w4 = stage_read(stage, px, vec);
        if (valid_o) *reinterpret_cast<uint4*>(a.out_pool + pp_o * a.pool_ld + co0 + 8 * vec) = w4;
      }
      __syncwarp();
    } else if (valid && (lane & 8) == 0) {
      __nv_bfloat16* pp = a.out_pool + pix_pool * a.pool_ld + co0;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        *reinterpret_cast<uint4*>(pp + 8 * q4) = make_uint4(r[4 * q4], r[4 * q4 + 1], r[4 * q4 + 2], r[4 * q4 + 3]);
    }
  }
}

// The epilogue role (warps 4..11), shared by both main-loop variants.
// (tile0, tstep): this CTA's tile sequence; hoff: row offset of this CTA inside the tile (CTA pairs);
// tempty_addr: shared::cluster address of the accumulator-drained barriers (the leader's for a pair).
template <bool kFast, bool kLean>
__device__ __forceinline__ void run_epilogue(const ConvArgs& a, uint32_t tmem_base, uint64_t* tfull_bar,
                                             uint32_t tempty_addr, float* bias_s, int warp, int lane,
                                             int tile0, int tstep, int hoff) {
  // optional warp-private store staging tile (2 KB per epilogue warp) right behind the bias vector
  uint8_t* stage = (kLean && a.stage_out)
                       ? reinterpret_cast<uint8_t*>(bias_s + kBiasSmemFloats) + (warp - 4) * kStageBytesPerWarp
                       : nullptr;
  // Warp w may only read TMEM lanes [32*(w%4), +32); warps w and w+4 share a lane quarter and
  // split the 32-column chunks of the accumulator between them.
  const int q = warp & 3;
  const int half = (warp - 4) >> 2;
  const int r = q * 32 + lane;              // accumulator row == pixel within the sub-tile
  const int lh = r / a.bw;
  const int lw = r - lh * a.bw;
  const int et = threadIdx.x - 128;
  // The bias vector is staged once per CTA, and the row masks of tile i+1 are fetched while tile i is
  // processed: no global-load latency and no CTA-wide barrier sit between "accumulator ready" and the
  // first tcgen05.ld of a tile (short tiles of narrow layers were paying ~1 us per tile for both).
  for (int j = et; j < a.cout_pad; j += kEpiThreads)
    bias_s[j] = (a.bias != nullptr && j < a.cout) ? a.bias[j] : 0.0f;
  named_bar_sync(1, kEpiThreads);
  const int hmul = a.up_mode ? 2 : 1;
  const int Hout = a.H * hmul;
  auto fetch_masks = [&](int tile, uint8_t (&m)[4]) {
    int n_idx, h0, w0, n0, par;
    decode_tile(a, tile, n_idx, h0, w0, n0, par);
    h0 += hoff;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int h = h0 + sub * a.bh + lh;
      m[sub] = 0;
      if (sub < a.msub && r < a.bh * a.bw && h < a.H)
        m[sub] = a.row_mask[static_cast<int64_t>(n_idx) * Hout + h * hmul + par];
    }
  };
  uint8_t mnext[4] = {0, 0, 0, 0};
  if (a.row_mask != nullptr && tile0 < a.num_tiles) fetch_masks(tile0, mnext);
  int it = 0;
  long long pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0, pk4 = 0;      // probe build: cycle sums, flushed once at the end
  (void)pk0; (void)pk1; (void)pk2; (void)pk3; (void)pk4;
  for (int tile = tile0; tile < a.num_tiles; tile += tstep, ++it) {
    int n_idx, h0, w0, n0, par;
    decode_tile(a, tile, n_idx, h0, w0, n0, par);
    h0 += hoff;
    const uint32_t buf = it % a.nbuf;
    const float* bs = bias_s + n0;
    uint32_t mbits = 0;                       // bit sub = "this thread's row of sub-tile sub is padded"
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) mbits |= (mnext[sub] != 0 ? 1u : 0u) << sub;
    if (a.row_mask != nullptr && tile + tstep < a.num_tiles) fetch_masks(tile + tstep, mnext);

    const long long pc0 = MQ_CLK();
    mbar_wait_relaxed(&tfull_bar[buf], (it / a.nbuf) & 1);
    tc_fence_after();
    const long long pc1 = MQ_CLK();
    pk0 += pc1 - pc0;
    pk4 += 1;
#pragma unroll 1
    for (int sub = 0; sub < a.msub; ++sub) {
      const int h = h0 + sub * a.bh + lh, w = w0 + lw;
      const bool valid = (r < a.bh * a.bw) && (h < a.H) && (w < a.W);
      const int64_t pix = (static_cast<int64_t>(n_idx) * Hout + h * hmul + par) * a.W + w;
      const int64_t pix_pool = (static_cast<int64_t>(n_idx) * (a.H >> 1) + (h >> 1)) * a.W + w;
      const bool masked = (mbits >> sub) & 1u;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * a.acc_stride + sub * a.bn;
#pragma unroll 1
      for (int c = half * 32; c < a.bn; c += 64) {
        uint32_t v[32];
        __syncwarp();                         // tcgen05.ld is .sync.aligned: reconverge first
        if (MQ_PROBE(a, 64)) {                // probe: no TMEM read
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(0.25f * static_cast<float>(j + lane));
        } else {
          const long long pl0 = MQ_CLK();
          tmem_ld_32x32(t_row + c, v);
          tmem_ld_wait();
          pk1 += MQ_CLK() - pl0;
        }
        const int co0 = n0 + c;
        if (MQ_PROBE(a, 1)) continue;
        const long long pb0 = MQ_CLK();
        if (kLean && kFast && a.slim) {
          if (a.out_pool != nullptr) {
            if (stage != nullptr) epilogue_slim<true, true>(a, v, bs + c, pix, co0, masked, valid, pix_pool, stage, lane);
            else epilogue_slim<true, false>(a, v, bs + c, pix, co0, masked, valid, pix_pool, stage, lane);
          } else {
            if (stage != nullptr) epilogue_slim<false, true>(a, v, bs + c, pix, co0, masked, valid, pix_pool, stage, lane);
            else epilogue_slim<false, false>(a, v, bs + c, pix, co0, masked, valid, pix_pool, stage, lane);
          }
        } else if (kLean) {
          epilogue_lean<kFast>(a, v, bs + c, pix, co0, masked, valid, pix_pool, stage);
        } else {
          if (valid && co0 < a.cout) epilogue_generic<kFast>(a, v, bs + c, pix, co0, masked);
        }
        pk2 += MQ_CLK() - pb0;
      }
    }
    // all TMEM reads of this buffer are complete (tcgen05.wait::ld above)
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(tempty_addr + buf * 8);
    pk3 += MQ_CLK() - pc0;
  }
  MQ_ACC(0, pk0); MQ_ACC(1, pk1); MQ_ACC(2, pk2); MQ_ACC(3, pk3); MQ_ACC(4, pk4);
}

template <bool kFast, bool kLean>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_a2, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B operand tiles need 1024-byte alignment (in the shared window).
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);

  const int stages = a.stages;
  const int a_stage_bytes = a.msub * kATileBytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * a_stage_bytes;
  uint8_t* tail = smem_b + stages * a.b_tile_bytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty_bar = full_bar + kMaxStages;
  uint64_t* tfull_bar = empty_bar + kMaxStages;
  uint64_t* tempty_bar = tfull_bar + kMaxAccBufs;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + kMaxAccBufs);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_s + 4);   // [2][256]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (a.up_mode) tma_prefetch_desc(&map_a2);
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < kMaxAccBufs; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], kEpiThreads / 32);     // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int kblocks = a.up_mode ? a.up_taps * a.kchunks + (a.taps - a.up_taps) * a.kchunks2
                                : a.taps * a.nseg * a.kchunks;

  if (warp == 0) {
    // ===================== TMA producer (one thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int n_idx, h0, w0, n0, par;
        decode_tile(a, tile, n_idx, h0, w0, n0, par);
        const int brow = n0 + par * a.cout_pad;
        int kb = 0;
        // K order: segment-major (bf16x3: the small products x2w0, x1w1, ... come first, so they
        // accumulate while |acc| is small and the tensor core's truncating fp32 adds cost nothing;
        // only the final x0w0 chain sees full-magnitude truncation), then tap, then channel chunk.
        for (int seg = 0; seg < a.nseg; ++seg) {
          const int cbase = a.a_coff[seg];
          for (int tap = 0; tap < a.taps; ++tap) {
            const bool src2 = a.up_mode && tap >= a.up_taps;
            const int nch = src2 ? a.kchunks2 : a.kchunks;
            const int ww = w0 + a.tap_dw[tap];
            int hh, par2 = 0;
            if (!src2) {
              hh = h0 + ((a.up_mode && par) ? a.tap_dh_odd[tap] : a.tap_dh[tap]);
            } else {
              const int orow = par + a.tap_dh[tap];   // output-row offset -> (parity, half-row) of the skip tensor
              par2 = orow & 1;
              hh = h0 + (orow >> 1);
            }
            for (int kc = 0; kc < nch; ++kc, ++kb) {
              mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
              if (MQ_PROBE(a, 4)) {
                mbar_arrive(&full_bar[stage]);
                if (++stage == stages) { stage = 0; phase ^= 1; }
                continue;
              }
              mbar_expect_tx(&full_bar[stage], a.msub * a.a_tx_bytes + a.b_tile_bytes);
              for (int sub = 0; sub < a.msub; ++sub) {
                uint8_t* dst = smem_a + stage * a_stage_bytes + sub * kATileBytes;
                if (!src2)
                  tma_load_4d(&map_a, &full_bar[stage], dst, cbase + kc * kBlockK, ww, hh + sub * a.bh, n_idx);
                else
                  tma_load_5d(&map_a2, &full_bar[stage], dst, kc * kBlockK, ww, par2, hh + sub * a.bh, n_idx);
              }
              tma_load_2d(&map_b, &full_bar[stage], smem_b + stage * a.b_tile_bytes, kb * kBlockK, brow);
              if (++stage == stages) {
                stage = 0;
                phase ^= 1;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the loop (warp-uniform control flow and descriptors); one elected lane
    // issues.  Per MMA the loop costs two 32-bit adds: narrow tiles (N <= 128, <= 64 clocks per MMA)
    // are otherwise bound by this thread's issue rate, not by the tensor pipe.
    {
      const uint32_t idesc = a.op_f16 ? umma_idesc_f16(kTileM, a.bn) : umma_idesc_bf16(kTileM, a.bn);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      const int msub = MQ_PROBE(a, 2) ? 0 : a.msub;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it % a.nbuf;
        mbar_wait(&tempty_bar[buf], ((it / a.nbuf) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * a.acc_stride;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t b_lo = umma_desc_lo(smem_u32(smem_b + stage * a.b_tile_bytes));
          const uint32_t a_lo = umma_desc_lo(smem_u32(smem_a + stage * a_stage_bytes));
          const uint32_t acc0 = kb != 0 ? 1u : 0u;
          if (elect_one_sync()) {
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
              if (sub < msub) {
#pragma unroll
                for (int k = 0; k < kBlockK / kUmmaK; ++k) {
                  // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in the >>4 field
                  umma_bf16(d_tmem + sub * a.bn, umma_desc_make(a_lo + sub * (kATileBytes >> 4) + 2 * k, hi),
                            umma_desc_make(b_lo + 2 * k, hi), idesc, k != 0 ? 1u : acc0);
                }
              }
            }
            umma_commit(&empty_bar[stage]);     // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one_sync()) umma_commit(&tfull_bar[buf]);   // accumulator ready for the epilogue
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    run_epilogue<kFast, kLean>(a, tmem_base, tfull_bar, smem_u32(tempty_bar), bias_s, warp, lane,
                               blockIdx.x, gridDim.x, 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// Halo-tile variant for 3x3 / pad-1 convolutions (single source, nseg == 1).
//
// The tap-shifted kernel above fetches every activation tile nine times (once per tap), and narrow
// layers are bound by that L2->SM traffic.  Here a CTA tile is msub sub-tiles of 16 rows x 8
// columns stacked along H; per 64-channel chunk ONE halo box (16*msub + 2) x 10 pixels is loaded,
// and the nine taps are nine UMMA descriptors into it: pixel (y, x) of the halo sits at row
// y*10 + x (128 B per row, SWIZZLE_128B), so the 8 pixels of an image row are one 8-row group and
// consecutive image rows are 1280 B apart (SBO = 1280).  Tap (dh, dw) only moves the descriptor
// start by ((1+dh)*10 + (1+dw)) rows.  The swizzle XOR is a function of the absolute shared-memory
// address bits, so unaligned starts need no descriptor base offset (probe: tools/umma_offset_probe.cu).
// Activations and weights run in two independent TMA rings (A: per chunk, B: per chunk x tap).
// ---------------------------------------------------------------------------
constexpr int kHaloW = 10, kHaloSubRows = 16, kHaloMaxA = 4, kHaloMaxB = 16;

__device__ __forceinline__ uint64_t umma_desc_sw128_sbo(uint32_t smem_addr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

template <bool kFast, bool kLean>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_b, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int nA = a.halo_nA, nB = a.halo_nB;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nA * a.halo_slot_bytes;
  uint8_t* tail = smem_b + nB * a.b_tile_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(tail);
  uint64_t* emptyA = fullA + kHaloMaxA;
  uint64_t* fullB = emptyA + kHaloMaxA;
  uint64_t* emptyB = fullB + kHaloMaxB;
  uint64_t* tfull_bar = emptyB + kHaloMaxB;
  uint64_t* tempty_bar = tfull_bar + kMaxAccBufs;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + kMaxAccBufs);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_s + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < nA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < nB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int b = 0; b < kMaxAccBufs; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], kEpiThreads / 32); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x) {
        int n_idx, h0, w0, n0, par;
        decode_tile(a, tile, n_idx, h0, w0, n0, par);
        for (int kc = 0; kc < a.kchunks; ++kc) {
          mbar_wait_relaxed(&emptyA[sa], pa ^ 1);
          if (MQ_PROBE(a, 4)) {
            mbar_arrive(&fullA[sa]);
          } else {
            mbar_expect_tx(&fullA[sa], a.halo_tx_bytes);
            tma_load_4d(&map_a, &fullA[sa], smem_a + sa * a.halo_slot_bytes, kc * kBlockK, w0 - 1, h0 - 1, n_idx);
          }
          if (++sa == nA) { sa = 0; pa ^= 1; }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait_relaxed(&emptyB[sb], pb ^ 1);
            if (MQ_PROBE(a, 4)) {
              mbar_arrive(&fullB[sb]);
            } else {
              mbar_expect_tx(&fullB[sb], a.b_tile_bytes);
              tma_load_2d(&map_b, &fullB[sb], smem_b + sb * a.b_tile_bytes, (tap * a.kchunks + kc) * kBlockK, n0);
            }
            if (++sb == nB) { sb = 0; pb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp walks the loop, one elected lane issues (see conv_gemm_kernel)
    {
      const uint32_t idesc = a.op_f16 ? umma_idesc_f16(kTileM, a.bn) : umma_idesc_bf16(kTileM, a.bn);
      constexpr uint32_t hi_a = umma_desc_hi_sw128(kHaloW * 128), hi_b = umma_desc_hi_sw128(1024);
      constexpr uint32_t kSubStep = (kHaloSubRows * kHaloW * 128) >> 4;      // one sub-tile down the halo, 16-byte units
      const int msub = MQ_PROBE(a, 2) ? 0 : a.msub;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < a.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it % a.nbuf;
        mbar_wait(&tempty_bar[buf], ((it / a.nbuf) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * a.acc_stride;
        for (int kc = 0; kc < a.kchunks; ++kc) {
          mbar_wait(&fullA[sa], pa);
          const uint32_t a_lo = umma_desc_lo(smem_u32(smem_a + sa * a.halo_slot_bytes));
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&fullB[sb], pb);
            tc_fence_after();
            const uint32_t b_lo = umma_desc_lo(smem_u32(smem_b + sb * a.b_tile_bytes));
            const uint32_t a_tap = a_lo + (((tap / 3) * kHaloW + (tap % 3)) * 128 >> 4);   // shifted view of the halo
            const uint32_t acc0 = (kc | tap) != 0 ? 1u : 0u;
            if (elect_one_sync()) {
#pragma unroll
              for (int sub = 0; sub < 4; ++sub) {
                if (sub < msub) {
#pragma unroll
                  for (int k = 0; k < kBlockK / kUmmaK; ++k)
                    umma_bf16(d_tmem + sub * a.bn, umma_desc_make(a_tap + sub * kSubStep + 2 * k, hi_a),
                              umma_desc_make(b_lo + 2 * k, hi_b), idesc, k != 0 ? 1u : acc0);
                }
              }
              umma_commit(&emptyB[sb]);
            }
            __syncwarp();
            if (++sb == nB) { sb = 0; pb ^= 1; }
          }
          if (elect_one_sync()) umma_commit(&emptyA[sa]);
          __syncwarp();
          if (++sa == nA) { sa = 0; pa ^= 1; }
        }
        if (elect_one_sync()) umma_commit(&tfull_bar[buf]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    run_epilogue<kFast, kLean>(a, tmem_base, tfull_bar, smem_u32(tempty_bar), bias_s, warp, lane,
                               blockIdx.x, gridDim.x, 0);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// Sixteen epilogue warps for the narrow ConvBlock layers (bn = 64 / 128, slim epilogue conditions).
//
// ncu on pre.conv2 (profiles/ncu_conv_pair_narrow_r02_summary.md): the eight epilogue warps - two per scheduler - issue
// 16 % of the time and wait on their own dependency chains for the rest (TMEM load, MUFU, shuffle and store
// latencies), while the MMA loop of these short-K layers would be done in two thirds of the epilogue's time.  More
// warps hide that latency: four warps share a TMEM lane quarter and take 16-column chunks round robin
// (tcgen05.ld.32x32b.x16, half the registers, so 640 threads fit the register file), everything else as epilogue_slim.
// ---------------------------------------------------------------------------
constexpr int kEpi16Warps = 16;
constexpr int kEpi16Threads = kEpi16Warps * 32;
constexpr int kThreads16 = 128 + kEpi16Threads;
constexpr int kStage16BytesPerWarp = 32 * 32;          // 32 pixels x 16 bf16 channels

__device__ __forceinline__ void stage16_write(uint8_t* st, int row, const uint32_t (&u)[8]) {
#pragma unroll
  for (int g = 0; g < 2; ++g)
    *reinterpret_cast<uint4*>(st + row * 32 + ((g ^ ((row >> 2) & 1)) << 4)) =
        make_uint4(u[4 * g], u[4 * g + 1], u[4 * g + 2], u[4 * g + 3]);
}
__device__ __forceinline__ uint4 stage16_read(const uint8_t* st, int px, int vec) {
  return *reinterpret_cast<const uint4*>(st + px * 32 + ((vec ^ ((px >> 2) & 1)) << 4));
}

// Store addressing without shuffles: a sub-tile of the pair kernel is 16 rows x 8 pixels, so the lane that STORES pixel
// px of the warp's 32 (pass it, px = 16 it + lane / 2) knows that pixel's row and column from its own lane id.  The
// pointers and bounds of both passes (and of the pooled row pair) are computed once per sub-tile, before the TMEM load
// lands; the chunk body is then load-free index arithmetic: math -> stage -> 16-byte predicated stores.
struct Slim16Out {
  __nv_bfloat16* p0;        // pass 0: pixel (row 4 q + lane / 16, column (lane / 2) % 8), channel 8 (lane % 2) of the N tile
  __nv_bfloat16* pp;        // pooled pixel lane / 2 of the warp's 16
  int64_t pass_step;        // elements between pass 0 and pass 1 (two image rows)
  bool ok0, ok1, okp;
};

template <bool kPool>
__device__ __forceinline__ void epilogue_slim16(const ConvArgs& a, const uint32_t (&v)[16], const float4 (&b4)[4], float g,
                                                float pool_half, const Slim16Out& o, int c, uint8_t* stage, int lane) {
  float x[16];
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    x[4 * q4] = __uint_as_float(v[4 * q4]) + b4[q4].x;
    x[4 * q4 + 1] = __uint_as_float(v[4 * q4 + 1]) + b4[q4].y;
    x[4 * q4 + 2] = __uint_as_float(v[4 * q4 + 2]) + b4[q4].z;
    x[4 * q4 + 3] = __uint_as_float(v[4 * q4 + 3]) + b4[q4].w;
  }
  uint32_t u[8];
  if (a.beta == 1.0f) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t0 = tanh_fast(x[2 * j]), t1 = tanh_fast(x[2 * j + 1]);
      const float g0 = g * x[2 * j], g1 = g * x[2 * j + 1];
      u[j] = pack_bf16x2(fmaf(g0, t0, g0), fmaf(g1, t1, g1));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t0 = tanh_fast(a.beta * x[2 * j]), t1 = tanh_fast(a.beta * x[2 * j + 1]);
      const float g0 = g * x[2 * j], g1 = g * x[2 * j + 1];
      u[j] = pack_bf16x2(fmaf(g0, t0, g0), fmaf(g1, t1, g1));
    }
  }
  // staged store: lane l of pass `it` writes vector l % 2 of pixel it * 16 + l / 2 (16 pixels x 32 contiguous bytes)
  stage16_write(stage, lane, u);
  __syncwarp();
  {
    const int vec = lane & 1, px = lane >> 1;
    const uint4 w0 = stage16_read(stage, px, vec);
    const uint4 w1 = stage16_read(stage, 16 + px, vec);
    if (o.ok0) *reinterpret_cast<uint4*>(o.p0 + c) = w0;
    if (o.ok1) *reinterpret_cast<uint4*>(o.p0 + o.pass_step + c) = w1;
  }
  if (kPool) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(pool_half, pool_half);
    uint32_t r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t ot = __shfl_xor_sync(0xffffffffu, u[j], 8);
      const __nv_bfloat162 s2 = __hmul2(__hadd2(*reinterpret_cast<const __nv_bfloat162*>(&u[j]),
                                                *reinterpret_cast<const __nv_bfloat162*>(&ot)), h2);
      r[j] = *reinterpret_cast<const uint32_t*>(&s2);
    }
    __syncwarp();                                              // pass reads of the stage tile are done
    const int prow = (lane & 7) | ((lane >> 4) << 3);          // pooled pixel 0..15 owned by lanes 0-7, 16-23
    if ((lane & 8) == 0) stage16_write(stage, prow, r);
    __syncwarp();
    const uint4 w4 = stage16_read(stage, lane >> 1, lane & 1);
    if (o.okp) *reinterpret_cast<uint4*>(o.pp + c) = w4;
  }
  __syncwarp();
}

// The epilogue role with sixteen warps (warps 4..19): warp (q, part) owns TMEM lanes [32 q, +32) and the 16-column
// chunks part, part + 4, ... of every sub-tile.  Tile walk, mask prefetch and barrier protocol as run_epilogue.
__device__ __forceinline__ void run_epilogue_slim16(const ConvArgs& a, uint32_t tmem_base, uint64_t* tfull_bar,
                                                    uint32_t tempty_addr, float* bias_s, int warp, int lane,
                                                    int tile0, int tstep, int hoff) {
  uint8_t* stage = reinterpret_cast<uint8_t*>(bias_s + kBiasSmemFloats) + (warp - 4) * kStage16BytesPerWarp;
  const int q = warp & 3;
  const int part = (warp - 4) >> 2;
  const int lh = q * 4 + (lane >> 3);                 // own accumulator row: pixel (lh, lane % 8) of the 16 x 8 sub-tile
  const int slh = q * 4 + (lane >> 4), slw = (lane >> 1) & 7;      // pixel this lane stores in pass 0 (pass 1: two rows down)
  const int plh = q * 4 + ((lane >> 4) << 1);                       // first row of the pooled pair this lane stores
  const int et = threadIdx.x - 128;
  for (int j = et; j < a.cout_pad; j += kEpi16Threads)
    bias_s[j] = (a.bias != nullptr && j < a.cout) ? a.bias[j] : 0.0f;
  named_bar_sync(1, kEpi16Threads);
  auto fetch_masks = [&](int tile, uint8_t (&m)[4]) {
    int n_idx, h0, w0, n0, par;
    decode_tile(a, tile, n_idx, h0, w0, n0, par);
    h0 += hoff;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
      const int h = h0 + sub * a.bh + lh;
      m[sub] = 0;
      if (sub < a.msub && h < a.H) m[sub] = a.row_mask[static_cast<int64_t>(n_idx) * a.H + h];
    }
  };
  uint8_t mnext[4] = {0, 0, 0, 0};
  if (a.row_mask != nullptr && tile0 < a.num_tiles) fetch_masks(tile0, mnext);
  const int64_t pass_step = static_cast<int64_t>(2) * a.W * a.bf16_ld;
  int it = 0;
  for (int tile = tile0; tile < a.num_tiles; tile += tstep, ++it) {
    int n_idx, h0, w0, n0, par;
    decode_tile(a, tile, n_idx, h0, w0, n0, par);
    h0 += hoff;
    const uint32_t buf = it % a.nbuf;
    const float* bs = bias_s + n0;
    uint32_t mbits = 0;
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) mbits |= (mnext[sub] != 0 ? 1u : 0u) << sub;
    if (a.row_mask != nullptr && tile + tstep < a.num_tiles) fetch_masks(tile + tstep, mnext);
    // rows of the neighbour in the pooled pair (lane ^ 8 holds row lh ^ 1)
    const uint32_t mbits_pair = mbits | __shfl_xor_sync(0xffffffffu, mbits, 8);
    const bool okw = (w0 + slw) < a.W;
    // store-lane pointers of sub-tile 0; a sub-tile further down is 16 image rows (8 pooled rows) away
    const int64_t spix = (static_cast<int64_t>(n_idx) * a.H + h0 + slh) * a.W + w0 + slw;
    __nv_bfloat16* const out0 = a.out_bf16 + spix * a.bf16_ld + a.bf16_coff + n0 + 8 * (lane & 1);
    const int64_t sub_step = static_cast<int64_t>(a.bh) * a.W * a.bf16_ld;
    __nv_bfloat16* pool0 = nullptr;
    int64_t pool_step = 0;
    if (a.out_pool != nullptr) {
      const int64_t ppix = (static_cast<int64_t>(n_idx) * (a.H >> 1) + ((h0 + plh) >> 1)) * a.W + w0 + slw;
      pool0 = a.out_pool + ppix * a.pool_ld + n0 + 8 * (lane & 1);
      pool_step = static_cast<int64_t>(a.bh >> 1) * a.W * a.pool_ld;
    }
    mbar_wait_relaxed(&tfull_bar[buf], (it / a.nbuf) & 1);
    tc_fence_after();
#pragma unroll 1
    for (int sub = 0; sub < a.msub; ++sub) {
      const int hs = h0 + sub * a.bh;
      Slim16Out o;
      o.p0 = out0 + sub * sub_step;
      o.pp = pool0 + sub * pool_step;
      o.pass_step = pass_step;
      o.ok0 = okw && (hs + slh) < a.H;
      o.ok1 = okw && (hs + slh + 2) < a.H;
      o.okp = okw && (hs + plh) < a.H;
      const float g = ((mbits >> sub) & 1u) ? 0.0f : a.gamma;
      const float pool_half = ((mbits_pair >> sub) & 1u) ? 0.0f : 0.5f;
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * a.acc_stride + sub * a.bn;
#pragma unroll 1
      for (int c = part * 16; c < a.bn; c += 64) {
        uint32_t v[16];
        float4 b4[4];
        __syncwarp();
        tmem_ld_32x16(t_row + c, v);
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) b4[q4] = *reinterpret_cast<const float4*>(bs + c + 4 * q4);   // under the TMEM load
        tmem_ld_wait();
        if (a.out_pool != nullptr) epilogue_slim16<true>(a, v, b4, g, pool_half, o, c, stage, lane);
        else epilogue_slim16<false>(a, v, b4, g, pool_half, o, c, stage, lane);
      }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(tempty_addr + buf * 8);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair halo kernel (cta_group::2) for the refiner's 3x3 convolutions, plain and fused
// nearest-upsample + concat.
//
// Every refiner layer is limited by L2->SM operand traffic before it is limited by the tensor pipe
// (the LTS cap is ~43 B/clk/SM; the single-CTA kernels need 34-49).  Two CTAs on the two SMs of a
// TPC run one M = 256 MMA: each stages the halo of its own msub sub-tiles (16 rows x 8 columns each,
// stacked along H, rank 1 below rank 0) and only HALF of every weight tile (bn/2 rows), so weight
// traffic per pixel halves in L2->SM and in shared-memory reads.  The leader (rank 0) issues the
// MMAs; TMA loads of both CTAs count on the leader's full barriers (cp.async.bulk.tensor
// .cta_group::2); tcgen05.commit multicasts "slot free" / "accumulator ready" to both CTAs; the
// epilogue warps of both CTAs arrive on the leader's "accumulator drained" barrier.
//
// K loop = chunk-major over up to two operand groups:
//   group 0: `in` (for the fused up-conv: the LOW-resolution tensor), one (R+2) x 10 halo per
//            64-channel chunk (R = 16*msub), taps [0, 9) or [0, up_taps) as shifted descriptors;
//   group 1 (fused up-conv only): the skip tensor viewed as (C, W, parity, H, N).  An output tile has
//            row parity p and rows 2i+p; tap dh = 0 reads skip parity p rows i (box A, origin i0), taps
//            dh = -1 / +1 read parity 1-p rows i-1+p / i+p (box B, origin i0-1+p, row offsets 0 / 1).
// ---------------------------------------------------------------------------
constexpr int kPairMaxA = 4, kPairMaxB = 16;

template <bool kFast, bool kLean, int kEpiWarps>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + 32 * kEpiWarps, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap map_a,
                 const __grid_constant__ CUtensorMap map_b,
                 const __grid_constant__ CUtensorMap map_a2, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int nA = a.halo_nA, nB = a.halo_nB;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nA * a.halo_slot_bytes;
  uint8_t* tail = smem_b + (a.b_resident ? a.taps * a.kchunks : nB * a.pair_bgrp) * a.b_tile_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(tail);        // used in the leader only
  uint64_t* emptyA = fullA + kPairMaxA;                       // per CTA (multicast commit)
  uint64_t* fullB = emptyA + kPairMaxA;                       // leader only
  uint64_t* emptyB = fullB + kPairMaxB;                       // per CTA
  uint64_t* tfull_bar = emptyB + kPairMaxB;                   // per CTA (multicast commit)
  uint64_t* tempty_bar = tfull_bar + kMaxAccBufs;                       // leader only, both CTAs' epilogues arrive
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + kMaxAccBufs);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_s + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (a.up_mode) tma_prefetch_desc(&map_a2);
    // every issuing warp commits once per slot / per accumulator
    for (int i = 0; i < nA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], a.mma_issuers); }
    for (int i = 0; i < nB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], a.mma_issuers); }
    for (int b = 0; b < kMaxAccBufs; ++b) { mbar_init(&tfull_bar[b], a.mma_issuers); mbar_init(&tempty_bar[b], 2 * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_s, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();          // both CTAs' barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int tile0 = blockIdx.x >> 1, tstep = gridDim.x >> 1;
  const int R = kHaloSubRows * a.msub;          // rows of this CTA's share of a tile
  const int hoff = static_cast<int>(rank) * R;
  const int ntap0 = a.up_mode ? a.up_taps : a.taps;
  const int nsegs = a.up_mode ? 1 : a.nseg;

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int brow_half = static_cast<int>(rank) * (a.bn >> 1);
      if (a.b_resident && tile0 < a.num_tiles) {
        // Weights-stationary layers (one N tile, K loop of one or two chunks): all taps x chunks of this CTA's half of the
        // weight tile are fetched once.  With a ring, the 64-channel layers had room for exactly one tile's worth of
        // weight slots next to their two 85 KB halo slots, so every tile started by waiting for its own weights.
        const uint32_t fb = smem_u32(&fullB[0]) & kPeerBitMask;
        const int nkb = a.taps * a.kchunks;
        if (rank == 0) mbar_expect_tx(&fullB[0], 2u * a.b_tile_bytes * nkb);
        for (int kb = 0; kb < nkb; ++kb)
          tma_load_2d_2cta(&map_b, fb, smem_b + kb * a.b_tile_bytes, kb * kBlockK, brow_half);
      }
      for (int tile = tile0; tile < a.num_tiles; tile += tstep) {
        int n_idx, h0, w0, n0, par;
        decode_tile(a, tile, n_idx, h0, w0, n0, par);
        const int hc = h0 + hoff;
        const int brow = n0 + par * a.cout_pad + brow_half;
        // fp32-grade split operands (nseg > 1, plain 3x3 only): segment-major K order as packed (smallest products
        // first); a segment pairs the activation term at channel offset a_coff[seg] with its own weight block
        for (int seg = 0; seg < nsegs; ++seg)
        for (int grp = 0; grp < (a.up_mode ? 2 : 1); ++grp) {
          const int nch = grp ? a.kchunks2 : a.kchunks;
          const int t0 = grp ? a.up_taps : 0, t1 = grp ? a.taps : ntap0;
          const int kbase = grp ? a.up_taps * a.kchunks : seg * a.taps * a.kchunks;
          const int cseg = a.a_coff[seg];
          for (int kc = 0; kc < nch; ++kc) {
            mbar_wait_relaxed(&emptyA[sa], pa ^ 1);
            const uint32_t fa = smem_u32(&fullA[sa]) & kPeerBitMask;
            uint8_t* dst = smem_a + sa * a.halo_slot_bytes;
            if (MQ_PROBE(a, 4)) {
              if (rank == 0) mbar_arrive(&fullA[sa]);
            } else if (rank == 0) {
              mbar_expect_tx(&fullA[sa], 2u * static_cast<uint32_t>(grp ? a.pair_tx1 : a.pair_tx0));
            }
            if (MQ_PROBE(a, 4)) {
            } else if (!grp) {
              tma_load_4d_2cta(&map_a, fa, dst, cseg + kc * kBlockK, w0 - 1, hc - 1, n_idx);
            } else {
              tma_load_5d_2cta(&map_a2, fa, dst, kc * kBlockK, w0 - 1, par, hc, n_idx);
              tma_load_5d_2cta(&map_a2, fa, dst + a.pair_boxb_off, kc * kBlockK, w0 - 1, par ^ 1, hc - 1 + par, n_idx);
            }
            if (++sa == nA) { sa = 0; pa ^= 1; }
            if (a.b_resident) continue;
            for (int tap = t0; tap < t1; tap += a.pair_bgrp) {
              mbar_wait_relaxed(&emptyB[sb], pb ^ 1);
              if (MQ_PROBE(a, 4)) {
                if (rank == 0) mbar_arrive(&fullB[sb]);
              } else {
                if (rank == 0) mbar_expect_tx(&fullB[sb], 2u * a.b_tile_bytes * a.pair_bgrp);
                const uint32_t fb = smem_u32(&fullB[sb]) & kPeerBitMask;
                for (int j = 0; j < a.pair_bgrp; ++j)
                  tma_load_2d_2cta(&map_b, fb, smem_b + (sb * a.pair_bgrp + j) * a.b_tile_bytes,
                                   (kbase + (tap + j - t0) * nch + kc) * kBlockK, brow);
              }
              if (++sb == nB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1 || (warp == 2 && a.mma_issuers == 2)) {
    // The leader's warp walks the loop, one elected lane issues for both CTAs.  A tcgen05.mma with N = 64 runs for 32
    // clocks but costs the issuing thread ~48 (two 64-bit descriptors through uniform registers per instruction, plus the
    // per-tap offsets and barrier traffic: profiles/conv_cycles_r02.log), so layers with several sub-tiles per tile split
    // them between two issuing warps: accumulators are disjoint, both warps wait on the same "operands landed"
    // barriers and each commits its own MMAs to the "slot free" / "accumulator ready" barriers (count = issuers).
    if (rank == 0) {
      const uint32_t idesc = a.op_f16 ? umma_idesc_f16(2 * kTileM, a.bn) : umma_idesc_bf16(2 * kTileM, a.bn);
      constexpr uint32_t hi_a = umma_desc_hi_sw128(kHaloW * 128), hi_b = umma_desc_hi_sw128(1024);
      constexpr uint32_t kSubStep = (kHaloSubRows * kHaloW * 128) >> 4;
      const int msub = MQ_PROBE(a, 2) ? 0 : a.msub;
      const int sub_lo = (warp - 1) * msub / a.mma_issuers, sub_hi = (warp * msub) / a.mma_issuers;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      const bool resident = a.b_resident != 0;
      long long pm5 = 0, pm6 = 0, pm7 = 0, pm8 = 0, pm9 = 0, pm10 = 0;
      (void)pm5; (void)pm6; (void)pm7; (void)pm8; (void)pm9; (void)pm10;
      if (resident && tile0 < a.num_tiles) {
        mbar_wait(&fullB[0], 0);                  // the one-off weight load
        tc_fence_after();
      }
      const uint32_t b_res_lo = umma_desc_lo(smem_u32(smem_b));
      // lane t holds the offset (16-byte units) of tap t's shifted view inside a halo slot, for even / odd output rows
      uint32_t toff_even = 0, toff_odd = 0;
      if (lane < a.taps) {
        const int tx = a.tap_dw[lane] + 1;
        if (!a.up_mode || lane < a.up_taps) {
          toff_even = static_cast<uint32_t>(((a.tap_dh[lane] + 1) * kHaloW + tx) * 8);
          toff_odd = a.up_mode ? static_cast<uint32_t>(((a.tap_dh_odd[lane] + 1) * kHaloW + tx) * 8) : toff_even;
        } else {
          const int dh = a.tap_dh[lane];
          toff_even = toff_odd =
              static_cast<uint32_t>(((dh != 0 ? a.pair_boxb_off : 0) >> 4) + ((dh == 1 ? kHaloW : 0) + tx) * 8);
        }
      }
      for (int tile = tile0; tile < a.num_tiles; tile += tstep, ++it) {
        const int par = (tile / a.tiles_n) % a.par_tiles;
        const uint32_t buf = it % a.nbuf;
        const long long pm0 = MQ_CLK();
        mbar_wait(&tempty_bar[buf], ((it / a.nbuf) & 1) ^ 1);
        tc_fence_after();
        pm5 += MQ_CLK() - pm0;
        const uint32_t d_tmem = tmem_base + buf * a.acc_stride;
        uint32_t acc = 0;
        for (int seg = 0; seg < nsegs; ++seg)
        for (int grp = 0; grp < (a.up_mode ? 2 : 1); ++grp) {
          const int nch = grp ? a.kchunks2 : a.kchunks;
          const int t0 = grp ? a.up_taps : 0, t1 = grp ? a.taps : ntap0;
          for (int kc = 0; kc < nch; ++kc) {
            const long long pa0 = MQ_CLK();
            mbar_wait(&fullA[sa], pa);
            pm6 += MQ_CLK() - pa0;
            if (resident) tc_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_u32(smem_a + sa * a.halo_slot_bytes));
            for (int tg = t0; tg < t1; tg += a.pair_bgrp) {
              if (!resident) {
                const long long pb0 = MQ_CLK();
                mbar_wait(&fullB[sb], pb);
                tc_fence_after();
                pm8 += MQ_CLK() - pb0;
              }
              // ring slot of this tap group, or the group's place in the resident weight image (tap-major, then chunk)
              const uint32_t b_lo = resident ? b_res_lo + ((tg * nch + kc) * (a.b_tile_bytes >> 4))
                                             : umma_desc_lo(smem_u32(smem_b + sb * a.pair_bgrp * a.b_tile_bytes));
              const uint32_t b_step = resident ? nch * (a.b_tile_bytes >> 4) : (a.b_tile_bytes >> 4);
              // this slot's tap offsets from the lanes that hold them (no indexed constant loads in the loop)
              uint32_t toffs[3];
#pragma unroll
              for (int j = 0; j < 3; ++j) toffs[j] = __shfl_sync(0xffffffffu, par ? toff_odd : toff_even, (tg + j) & 31);
              const long long pi0 = MQ_CLK();
              // one elected lane issues the whole slot - taps x its sub-tiles x four K steps - and commits it
              if (elect_one_sync()) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  if (j < a.pair_bgrp) {
                    const uint32_t a_tap = a_lo + toffs[j];
                    const uint32_t b_tap = b_lo + j * b_step;
#pragma unroll 1
                    for (int sub = sub_lo; sub < sub_hi; ++sub) {
                      const uint32_t d_sub = d_tmem + sub * a.bn, a_sub = a_tap + sub * kSubStep;
#pragma unroll
                      for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16_2cta(d_sub, umma_desc_make(a_sub + 2 * k, hi_a), umma_desc_make(b_tap + 2 * k, hi_b), idesc,
                                       k != 0 ? 1u : acc);
                    }
                    acc = 1;
                  }
                }
                if (!resident) umma_commit_2cta(&emptyB[sb]);
              }
              __syncwarp();
              acc = 1;
              pm9 += MQ_CLK() - pi0;
              if (!resident) {
                if (++sb == nB) { sb = 0; pb ^= 1; }
              }
            }
            const long long pc1 = MQ_CLK();
            if (elect_one_sync()) umma_commit_2cta(&emptyA[sa]);
            __syncwarp();
            pm10 += MQ_CLK() - pc1;
            if (++sa == nA) { sa = 0; pa ^= 1; }
          }
        }
        if (elect_one_sync()) umma_commit_2cta(&tfull_bar[buf]);
        __syncwarp();
        pm7 += MQ_CLK() - pm0;
      }
      if (warp == 1) { MQ_ACC(5, pm5); MQ_ACC(6, pm6); MQ_ACC(7, pm7); MQ_ACC(8, pm8); MQ_ACC(9, pm9); MQ_ACC(10, pm10); }
    }
  } else if (warp >= 4) {
    if constexpr (kEpiWarps == kEpi16Warps)
      run_epilogue_slim16(a, tmem_base, tfull_bar, smem_u32(tempty_bar) & kPeerBitMask, bias_s, warp, lane, tile0, tstep, hoff);
    else
      run_epilogue<kFast, kLean>(a, tmem_base, tfull_bar, smem_u32(tempty_bar) & kPeerBitMask, bias_s, warp, lane,
                               tile0, tstep, hoff);
  }

  // the leader's barriers receive arrivals from the peer until its last tile: leave together
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// CTA-pair kernel for the 1-D convolutions and linears (W == 1: encoder / decoder blocks, projections).
//
// Same pairing as conv_pair_kernel (M = 256 across two CTAs, half a weight tile per CTA, leader issues),
// with a ROW halo: per (segment, 64-channel chunk) a CTA loads the 128*msub + taps - 1 consecutive rows
// its sub-tiles need ONCE, and tap j is the descriptor shifted by j rows (rows are 128 bytes, so a
// shift is 8 sixteen-byte units; the 128-byte swizzle depends on absolute address bits, so unaligned
// starts need no base offset).  The tap-shifted kernel above re-fetches the activation tile for every
// tap (k = 3..7 times the L2->SM traffic).  K order = (segment, chunk, tap): segment-major like the
// packed weights, so the f16x2 / bf16x3 "small products first" accumulation order is unchanged.
// Weight ring slots hold up to three taps (ragged last group) to amortise the barrier round trip.
// ---------------------------------------------------------------------------
constexpr int kPair1dGroup = 3;

template <bool kFast, bool kLean>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
conv_pair1d_kernel(const __grid_constant__ CUtensorMap map_a,
                   const __grid_constant__ CUtensorMap map_b, const ConvArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  const int nA = a.halo_nA, nB = a.halo_nB;
  const int bslot = kPair1dGroup * static_cast<int>(a.b_tile_bytes);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nA * a.halo_slot_bytes;
  uint8_t* tail = smem_b + nB * bslot;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(tail);        // leader only
  uint64_t* emptyA = fullA + kPairMaxA;                       // per CTA (multicast commit)
  uint64_t* fullB = emptyA + kPairMaxA;                       // leader only
  uint64_t* emptyB = fullB + kPairMaxB;                       // per CTA
  uint64_t* tfull_bar = emptyB + kPairMaxB;                   // per CTA (multicast commit)
  uint64_t* tempty_bar = tfull_bar + kMaxAccBufs;             // leader only, both CTAs' epilogues arrive
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tempty_bar + kMaxAccBufs);
  float* bias_s = reinterpret_cast<float*>(tmem_ptr_s + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int i = 0; i < nA; ++i) { mbar_init(&fullA[i], 1); mbar_init(&emptyA[i], 1); }
    for (int i = 0; i < nB; ++i) { mbar_init(&fullB[i], 1); mbar_init(&emptyB[i], 1); }
    for (int b = 0; b < kMaxAccBufs; ++b) { mbar_init(&tfull_bar[b], 1); mbar_init(&tempty_bar[b], 2 * (kEpiThreads / 32)); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_ptr_s, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  const int tile0 = blockIdx.x >> 1, tstep = gridDim.x >> 1;
  const int R = kTileM * a.msub;                // rows of this CTA's share of a tile
  const int hoff = static_cast<int>(rank) * R;
  const int dh0 = a.tap_dh[0];                  // taps are consecutive row offsets dh0, dh0 + 1, ...

  if (warp == 0) {
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      const int brow_half = static_cast<int>(rank) * (a.bn >> 1);
      for (int tile = tile0; tile < a.num_tiles; tile += tstep) {
        int n_idx, h0, w0, n0, par;
        decode_tile(a, tile, n_idx, h0, w0, n0, par);
        const int hc = h0 + hoff;
        for (int seg = 0; seg < a.nseg; ++seg) {
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait_relaxed(&emptyA[sa], pa ^ 1);
            if (rank == 0) mbar_expect_tx(&fullA[sa], 2u * static_cast<uint32_t>(a.pair_tx0));
            tma_load_4d_2cta(&map_a, smem_u32(&fullA[sa]) & kPeerBitMask, smem_a + sa * a.halo_slot_bytes,
                             a.a_coff[seg] + kc * kBlockK, 0, hc + dh0, n_idx);
            if (++sa == nA) { sa = 0; pa ^= 1; }
            for (int tap = 0; tap < a.taps; tap += kPair1dGroup) {
              const int g = min(kPair1dGroup, a.taps - tap);
              mbar_wait_relaxed(&emptyB[sb], pb ^ 1);
              if (rank == 0) mbar_expect_tx(&fullB[sb], 2u * a.b_tile_bytes * g);
              const uint32_t fb = smem_u32(&fullB[sb]) & kPeerBitMask;
              for (int j = 0; j < g; ++j)
                tma_load_2d_2cta(&map_b, fb, smem_b + sb * bslot + j * a.b_tile_bytes,
                                 ((seg * a.taps + tap + j) * a.kchunks + kc) * kBlockK, n0 + brow_half);
              if (++sb == nB) { sb = 0; pb ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      const uint32_t idesc = a.op_f16 ? umma_idesc_f16(2 * kTileM, a.bn) : umma_idesc_bf16(2 * kTileM, a.bn);
      constexpr uint32_t hi = umma_desc_hi_sw128(1024);
      constexpr uint32_t kSubStep = kATileBytes >> 4;           // 128 rows of 128 bytes, 16-byte units
      const int msub = a.msub;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int it = 0;
      for (int tile = tile0; tile < a.num_tiles; tile += tstep, ++it) {
        const uint32_t buf = it % a.nbuf;
        mbar_wait(&tempty_bar[buf], ((it / a.nbuf) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * a.acc_stride;
        uint32_t acc = 0;
        for (int seg = 0; seg < a.nseg; ++seg) {
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&fullA[sa], pa);
            const uint32_t a_lo = umma_desc_lo(smem_u32(smem_a + sa * a.halo_slot_bytes));
            for (int tap = 0; tap < a.taps; tap += kPair1dGroup) {
              const int g = min(kPair1dGroup, a.taps - tap);
              mbar_wait(&fullB[sb], pb);
              tc_fence_after();
              const uint32_t b_lo = umma_desc_lo(smem_u32(smem_b + sb * bslot));
              for (int j = 0; j < g; ++j) {
                const uint32_t a_tap = a_lo + static_cast<uint32_t>((tap + j) * 8);     // j rows down the halo
                const uint32_t b_tap = b_lo + j * (a.b_tile_bytes >> 4);
                if (elect_one_sync()) {
#pragma unroll
                  for (int sub = 0; sub < 4; ++sub) {
                    if (sub < msub) {
#pragma unroll
                      for (int k = 0; k < kBlockK / kUmmaK; ++k)
                        umma_bf16_2cta(d_tmem + sub * a.bn, umma_desc_make(a_tap + sub * kSubStep + 2 * k, hi),
                                       umma_desc_make(b_tap + 2 * k, hi), idesc, k != 0 ? 1u : acc);
                    }
                  }
                }
                __syncwarp();
                acc = 1;
              }
              if (elect_one_sync()) umma_commit_2cta(&emptyB[sb]);
              __syncwarp();
              if (++sb == nB) { sb = 0; pb ^= 1; }
            }
            if (elect_one_sync()) umma_commit_2cta(&emptyA[sa]);
            __syncwarp();
            if (++sa == nA) { sa = 0; pa ^= 1; }
          }
        }
        if (elect_one_sync()) umma_commit_2cta(&tfull_bar[buf]);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    run_epilogue<kFast, kLean>(a, tmem_base, tfull_bar, smem_u32(tempty_bar) & kPeerBitMask, bias_s, warp, lane,
                               tile0, tstep, hoff);
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// Experiment knobs from the environment, read ONCE per process (not per launch).
struct ConvEnv {
  int nbuf = 0, debug = 0, stages = 0, bgrp = 0, stage_out = -1, slim = 1, slim16 = 1, b_resident = 1, issuers = 2;
  ConvEnv() {
    auto geti = [](const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; };
    nbuf = geti("MQ_CONV_NBUF");
    debug = geti("MQ_CONV_DEBUG");
    stages = geti("MQ_CONV_STAGES");
    bgrp = geti("MQ_PAIR_BGRP");
    if (getenv("MQ_B_RESIDENT")) b_resident = geti("MQ_B_RESIDENT");   // 0 = always the weight ring
    if (getenv("MQ_MMA_ISSUERS")) issuers = geti("MQ_MMA_ISSUERS");   // 1 = a single MMA-issuing warp everywhere
    if (getenv("MQ_SLIM16")) slim16 = geti("MQ_SLIM16");              // 0 = eight epilogue warps everywhere
    if (getenv("MQ_SLIM")) slim = geti("MQ_SLIM");                    // 0 = always the generic lean body
    if (getenv("MQ_STAGE_OUT")) stage_out = geti("MQ_STAGE_OUT");     // 0 = never, 1 = bn <= 128 (default), 2 = always
  }
};
static const ConvEnv& conv_env() {
  static const ConvEnv env;
  return env;
}

static int conv_smem_bytes(int stages, int a_stage_bytes, int b_tile_bytes) {
  return 1024 /*alignment slack*/ + stages * (a_stage_bytes + b_tile_bytes) +
         (2 * kMaxStages + 2 * kMaxAccBufs) * 8 + 16 + kBiasSmemFloats * 4;
}

}  // namespace mq

using namespace mq;

#ifdef MQ_CONV_PROBES
extern "C" int mq_conv_probe_cycles(unsigned long long* out12, int reset) {
  MQ_CUDA_OK(cudaDeviceSynchronize());
  MQ_CUDA_OK(cudaMemcpyFromSymbol(out12, mq_probe_cycles, sizeof(unsigned long long) * 12));
  if (reset) {
    unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    MQ_CUDA_OK(cudaMemcpyToSymbol(mq_probe_cycles, z, sizeof(z)));
  }
  return 0;
}
#endif

extern "C" int mq_conv_gemm(const mq_conv_params* p, mq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MQ_REQUIRE(p != nullptr, "mq_conv_gemm: null params");
  MQ_REQUIRE(p->in && p->wpack, "mq_conv_gemm: null in/wpack");
  MQ_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "mq_conv_gemm: bad N/H/W %d %d %d", p->N, p->H, p->W);
  MQ_REQUIRE(p->bn >= 32 && p->bn <= 256 && p->bn % 32 == 0, "mq_conv_gemm: bn=%d must be a multiple of 32 in [32,256]", p->bn);
  MQ_REQUIRE(p->cout > 0 && p->cout_pad >= p->cout && p->cout_pad % p->bn == 0, "mq_conv_gemm: cout=%d cout_pad=%d bn=%d", p->cout, p->cout_pad, p->bn);
  MQ_REQUIRE(p->cout_pad <= kBiasSmemFloats, "mq_conv_gemm: cout_pad=%d exceeds %d", p->cout_pad, kBiasSmemFloats);
  MQ_REQUIRE(p->taps >= 1 && p->taps <= MQ_MAX_TAPS, "mq_conv_gemm: taps=%d", p->taps);
  const bool up = p->in2 != nullptr;
  if (up) {
    MQ_REQUIRE(p->nseg == 1 && p->up_taps >= 1 && p->up_taps < p->taps && p->kchunks2 >= 1 && p->in2_ld % 8 == 0 &&
               (reinterpret_cast<uintptr_t>(p->in2) & 15) == 0, "mq_conv_gemm: bad upsample-concat arguments");
  }
  MQ_REQUIRE(p->nseg >= 1 && p->nseg <= MQ_MAX_SEGS, "mq_conv_gemm: nseg=%d", p->nseg);
  MQ_REQUIRE(p->kchunks >= 1, "mq_conv_gemm: kchunks=%d", p->kchunks);
  MQ_REQUIRE(p->bh >= 1 && p->bw >= 1 && p->bh * p->bw <= kTileM && p->bh <= 256 && p->bw <= 256, "mq_conv_gemm: tile %dx%d", p->bh, p->bw);
  MQ_REQUIRE(p->in_ld % 8 == 0, "mq_conv_gemm: in_ld=%d must be a multiple of 8 (16-byte TMA strides)", p->in_ld);
  MQ_REQUIRE((reinterpret_cast<uintptr_t>(p->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->wpack) & 15) == 0, "mq_conv_gemm: in/wpack must be 16-byte aligned");
  MQ_REQUIRE(p->out_f32 || p->out_bf16 || p->out_split, "mq_conv_gemm: no output");
  MQ_REQUIRE(p->res_mode == 0 || p->res != nullptr, "mq_conv_gemm: res_mode without res");
  MQ_REQUIRE(!(p->mask_pre || p->mask_post) || p->row_mask != nullptr, "mq_conv_gemm: mask flag without row_mask");
  // (fewer than four output channels never take a 16-byte store: the refiner's 3x3 post conv writes one float per pixel)
  if (p->out_f32) MQ_REQUIRE((p->f32_ld % 4 == 0 && p->f32_coff % 4 == 0) || p->cout < 4, "mq_conv_gemm: f32 ld/coff must be multiples of 4");
  if (p->out_bf16) MQ_REQUIRE(p->bf16_ld % 8 == 0 && p->bf16_coff % 8 == 0, "mq_conv_gemm: bf16 ld/coff must be multiples of 8");
  if (p->out_split) MQ_REQUIRE(p->split_ld % 8 == 0 && p->split_seg % 8 == 0, "mq_conv_gemm: split ld/seg must be multiples of 8");
  if (p->res_mode) MQ_REQUIRE(p->res_ld % 8 == 0 && p->res_coff % 8 == 0, "mq_conv_gemm: res ld/coff must be multiples of 8");

  EncodeTiledFn encode = get_encode_fn();
  MQ_REQUIRE(encode != nullptr, "mq_conv_gemm: cuTensorMapEncodeTiled not available from the driver");

  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.N = p->N; a.H = p->H; a.W = p->W;
  a.bh = p->bh; a.bw = p->bw; a.bn = p->bn; a.cout = p->cout;
  a.msub = p->msub > 0 ? p->msub : 1;
  MQ_REQUIRE(a.msub * p->bn <= kTmemCols && a.msub <= 4, "mq_conv_gemm: msub=%d * bn=%d exceeds %d TMEM columns", a.msub, p->bn, kTmemCols);
  a.acc_stride = (a.msub * p->bn + 31) / 32 * 32;
  a.nbuf = kTmemCols / a.acc_stride;               // 1 (msub*bn > 256), 2, or up to 4 stages for small tiles
  if (a.nbuf > kMaxAccBufs) a.nbuf = kMaxAccBufs;
  const ConvEnv& env = conv_env();
  if (env.nbuf >= 1 && env.nbuf < a.nbuf) a.nbuf = env.nbuf;
  const bool pair = p->pair != 0;
  a.tile_rows = p->bh * a.msub * (pair ? 2 : 1);
  a.tiles_h = (p->H + a.tile_rows - 1) / a.tile_rows;
  a.tiles_w = (p->W + p->bw - 1) / p->bw;
  a.tiles_n = p->cout_pad / p->bn;
  a.up_mode = up ? 1 : 0;
  a.up_taps = up ? p->up_taps : 0;
  a.kchunks2 = up ? p->kchunks2 : 0;
  a.par_tiles = up ? 2 : 1;
  a.cout_pad = p->cout_pad;
  for (int i = 0; i < MQ_MAX_TAPS; ++i) a.tap_dh_odd[i] = p->tap_dh_odd[i];
  const long long nt = 1LL * p->N * a.tiles_h * a.tiles_w * a.tiles_n * a.par_tiles;
  MQ_REQUIRE(nt < (1LL << 31), "mq_conv_gemm: too many tiles");
  a.num_tiles = static_cast<int>(nt);
  a.taps = p->taps; a.nseg = p->nseg; a.kchunks = p->kchunks;
  for (int i = 0; i < MQ_MAX_TAPS; ++i) { a.tap_dh[i] = p->tap_dh[i]; a.tap_dw[i] = p->tap_dw[i]; }
  for (int i = 0; i < MQ_MAX_SEGS; ++i) a.a_coff[i] = p->a_coff[i];
  const bool halo = p->halo != 0;
  if (halo) {
    MQ_REQUIRE(!up && p->nseg == 1 && p->taps == 9 && p->bw == 8 && p->bh == kHaloSubRows,
               "mq_conv_gemm: halo mode needs a single-source 3x3 conv with an 16x8 sub-tile");
    for (int t = 0; t < 9; ++t)
      MQ_REQUIRE(p->tap_dh[t] == t / 3 - 1 && p->tap_dw[t] == t % 3 - 1, "mq_conv_gemm: halo mode needs the standard 3x3 tap order");
  }
  const bool pair1d = pair && p->W == 1;          // row-halo variant for 1-D convolutions / linears
  if (pair1d) {
    MQ_REQUIRE(!halo && !up && p->bw == 1 && p->bh == kTileM && a.msub == 1 && p->bn % 32 == 0,
               "mq_conv_gemm: 1-D pair mode needs a 128x1 sub-tile, msub == 1, bn a multiple of 32 and no in2");
    for (int t = 0; t < p->taps; ++t)
      MQ_REQUIRE(p->tap_dw[t] == 0 && p->tap_dh[t] == p->tap_dh[0] + t, "mq_conv_gemm: 1-D pair mode needs consecutive row taps");
  } else if (pair) {
    MQ_REQUIRE(!halo && (p->nseg == 1 || !up) && p->bw == 8 && p->bh == kHaloSubRows && p->bn % 32 == 0,
               "mq_conv_gemm: pair mode needs a 16x8 sub-tile, bn a multiple of 32 and nseg == 1 for the fused up-conv");
    if (!up) {
      // the 3x3 convolution in its standard tap order, or one row of it (three taps along W: the refiner's post conv
      // computes its three row sums as three output channels); either way every tap is a shifted view of the halo box
      MQ_REQUIRE(p->taps == 9 || p->taps == 3, "mq_conv_gemm: pair mode needs a 3x3 convolution or one row of it");
      for (int t = 0; t < p->taps; ++t)
        MQ_REQUIRE(p->taps == 9 ? (p->tap_dh[t] == t / 3 - 1 && p->tap_dw[t] == t % 3 - 1)
                                : (p->tap_dh[t] == 0 && p->tap_dw[t] == t - 1),
                   "mq_conv_gemm: pair mode needs the standard 3x3 tap order (or dh = 0, dw = -1, 0, 1)");
    } else {
      for (int t = 0; t < p->taps; ++t) {
        MQ_REQUIRE(p->tap_dw[t] >= -1 && p->tap_dw[t] <= 1 && p->tap_dh[t] >= -1 && p->tap_dh[t] <= 1,
                   "mq_conv_gemm: pair mode taps must stay inside a 1-pixel halo");
        if (t < p->up_taps)
          MQ_REQUIRE(p->tap_dh_odd[t] >= -1 && p->tap_dh_odd[t] <= 1, "mq_conv_gemm: pair mode taps must stay inside a 1-pixel halo");
      }
    }
  }
  a.a_tx_bytes = static_cast<uint32_t>(p->bh * p->bw * kBlockK * 2);
  a.b_tile_bytes = static_cast<uint32_t>((pair ? p->bn / 2 : p->bn) * kBlockK * 2);
  const int a_stage_bytes = a.msub * kATileBytes;
  int stages = (kSmemBudget - 1024 - 6144) / (a_stage_bytes + static_cast<int>(a.b_tile_bytes));
  if (stages > kMaxStages) stages = kMaxStages;
  MQ_REQUIRE(stages >= 2, "mq_conv_gemm: not enough shared memory for 2 stages");
  a.stages = stages;
  a.debug = env.debug;
  if (env.stages >= 2 && env.stages <= stages) a.stages = stages = env.stages;
  a.bias = p->bias; a.row_mask = p->row_mask;
  a.mask_pre = p->mask_pre; a.mask_post = p->mask_post; a.act = p->act; a.res_mode = p->res_mode;
  a.beta = p->beta; a.gamma = p->gamma;
  a.res = p->res; a.res_is_bf16 = p->res_is_bf16; a.res_ld = p->res_ld; a.res_coff = p->res_coff;
  a.out_f32 = p->out_f32; a.f32_ld = p->f32_ld; a.f32_coff = p->f32_coff;
  a.out_bf16 = reinterpret_cast<__nv_bfloat16*>(p->out_bf16); a.bf16_ld = p->bf16_ld; a.bf16_coff = p->bf16_coff;
  a.out_split = reinterpret_cast<uint16_t*>(p->out_split); a.split_ld = p->split_ld; a.split_seg = p->split_seg;
  a.out_pool = reinterpret_cast<__nv_bfloat16*>(p->out_pool); a.pool_ld = p->pool_ld;
  a.split_kind = p->split_kind; a.op_f16 = p->op_f16;
  a.acc_scale = p->acc_scale != 0.0f ? p->acc_scale : 1.0f;
  const CUtensorMapDataType op_dt = p->op_f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const bool lean = p->out_bf16 != nullptr && p->out_f32 == nullptr && p->out_split == nullptr &&
                    p->cout % 32 == 0 && (p->res_mode == 0 || p->res_is_bf16) && a.acc_scale == 1.0f;
  {
    const int mode = env.stage_out < 0 ? 1 : env.stage_out;
    // halo / CTA-pair kernels only (the tap-loop kernel sizes its ring before this point and serves 1x1 layers)
    a.stage_out = (lean && (pair || halo) && (mode == 2 || (mode == 1 && p->bn <= 128))) ? 1 : 0;
  }
  // `masked` rows are zeroed through the gain, which is right whenever a row mask is given at all (mask_pre before
  // APTx and mask_post after it both give 0 = aptx(0)); a mask pointer without either flag must not zero anything
  a.slim = (env.slim && lean && p->act && p->fast_tanh && p->res_mode == 0 &&
            (p->row_mask == nullptr || p->mask_pre || p->mask_post)) ? 1 : 0;
  // sixteen epilogue warps: the 3x3 pair kernel's narrow layers (plain or fused up-conv output rows are handled by the
  // eight-warp body: the 16-warp tile walk assumes one output row per input row)
  a.slim16 = (env.slim16 && a.slim && pair && !pair1d && !up && p->bn % 64 == 0 && p->bn <= 128) ? 1 : 0;
  if (a.slim16) a.stage_out = 1;
  // two MMA-issuing warps where a tile has several sub-tiles of a narrow N (the issue rate, not the tensor pipe, paces those)
  a.mma_issuers = (pair && !pair1d && env.issuers == 2 && a.msub >= 2 && p->bn <= 128) ? 2 : 1;
  int stage_bytes = a.stage_out ? kStageBytes : 0;   // 8 warps x 2 KB == 16 warps x 1 KB

  // --- tensor maps ---
  CUtensorMap map_a, map_b, map_a2;
  {
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(p->in_ld), static_cast<cuuint64_t>(p->W),
                          static_cast<cuuint64_t>(p->H), static_cast<cuuint64_t>(p->N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(p->in_ld) * 2,
                             static_cast<cuuint64_t>(p->W) * p->in_ld * 2,
                             static_cast<cuuint64_t>(p->H) * p->W * p->in_ld * 2};
    cuuint32_t box[4] = {kBlockK, static_cast<cuuint32_t>(p->bw), static_cast<cuuint32_t>(p->bh), 1};
    if (pair1d) {                     // one (128*msub + taps - 1)-row halo per channel chunk
      box[1] = 1;
      box[2] = static_cast<cuuint32_t>(kTileM * a.msub + p->taps - 1);
    } else if (halo || pair) {        // one (16*msub + 2) x 10 pixel halo box per channel chunk
      box[1] = kHaloW;
      box[2] = static_cast<cuuint32_t>(kHaloSubRows * a.msub + 2);
    }
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map_a, op_dt, 4, const_cast<void*>(p->in), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MQ_REQUIRE(r == CUDA_SUCCESS, "mq_conv_gemm: cuTensorMapEncodeTiled(A) failed with %d (in_ld=%d W=%d H=%d N=%d)", (int)r, p->in_ld, p->W, p->H, p->N);
  }
  {
    const cuuint64_t K = up ? (static_cast<cuuint64_t>(p->up_taps) * p->kchunks +
                               static_cast<cuuint64_t>(p->taps - p->up_taps) * p->kchunks2) * kBlockK
                            : static_cast<cuuint64_t>(p->taps) * p->nseg * p->kchunks * kBlockK;
    cuuint64_t dims[2] = {K, static_cast<cuuint64_t>(p->cout_pad) * (up ? 2 : 1)};
    cuuint64_t strides[1] = {K * 2};
    cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(pair ? p->bn / 2 : p->bn)};   // a pair CTA stages half the rows
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&map_b, op_dt, 2, const_cast<void*>(p->wpack), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MQ_REQUIRE(r == CUDA_SUCCESS, "mq_conv_gemm: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }

  if (up) {
    // skip tensor (N, 2H, W, C2) viewed as (N, H, 2, W, C2): the row parity is its own dimension
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(p->in2_ld), static_cast<cuuint64_t>(p->W), 2,
                          static_cast<cuuint64_t>(p->H), static_cast<cuuint64_t>(p->N)};
    const cuuint64_t rowb = static_cast<cuuint64_t>(p->W) * p->in2_ld * 2;
    cuuint64_t strides[4] = {static_cast<cuuint64_t>(p->in2_ld) * 2, rowb, 2 * rowb, 2 * rowb * p->H};
    cuuint32_t box[5] = {kBlockK, static_cast<cuuint32_t>(p->bw), 1, static_cast<cuuint32_t>(p->bh), 1};
    if (pair) {                       // one-parity halo boxes of 16*msub + 1 half-rows x 10 pixels
      box[1] = kHaloW;
      box[3] = static_cast<cuuint32_t>(kHaloSubRows * a.msub + 1);
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = encode(&map_a2, op_dt, 5, const_cast<void*>(p->in2), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MQ_REQUIRE(r == CUDA_SUCCESS, "mq_conv_gemm: cuTensorMapEncodeTiled(A2) failed with %d", (int)r);
  } else {
    map_a2 = map_a;
  }

  int dev = 0, sms = 0;
  MQ_CUDA_OK(cudaGetDevice(&dev));
  MQ_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int smem = conv_smem_bytes(stages, a_stage_bytes, static_cast<int>(a.b_tile_bytes));
  if (halo) {
    const int halo_px = (kHaloSubRows * a.msub + 2) * kHaloW;
    a.halo_tx_bytes = halo_px * 128;
    a.halo_slot_bytes = (a.halo_tx_bytes + 1023) / 1024 * 1024;
    const int tail_bytes = (2 * kHaloMaxA + 2 * kHaloMaxB + 2 * kMaxAccBufs) * 8 + 16 + kBiasSmemFloats * 4 + stage_bytes;
    const int budget = kSmemBudget - 1024 - tail_bytes - 256;
    int nA = a.kchunks >= 3 ? 3 : 2;
    while (nA > 2 && budget - nA * a.halo_slot_bytes < 4 * static_cast<int>(a.b_tile_bytes)) --nA;
    int nB = (budget - nA * a.halo_slot_bytes) / static_cast<int>(a.b_tile_bytes);
    if (nB > kHaloMaxB) nB = kHaloMaxB;
    MQ_REQUIRE(nB >= 3, "mq_conv_gemm: halo mode does not fit shared memory (msub=%d bn=%d)", a.msub, p->bn);
    a.halo_nA = nA; a.halo_nB = nB;
    smem = 1024 + nA * a.halo_slot_bytes + nB * static_cast<int>(a.b_tile_bytes) + tail_bytes;
  }
  int grid = a.num_tiles < sms ? a.num_tiles : sms;
  if (pair1d) {
    a.pair_tx0 = (kTileM * a.msub + p->taps - 1) * 128;
    a.halo_slot_bytes = (a.pair_tx0 + 1023) / 1024 * 1024;
    const int tail_bytes = (2 * kPairMaxA + 2 * kPairMaxB + 2 * kMaxAccBufs) * 8 + 16 + kBiasSmemFloats * 4 + stage_bytes;
    const int budget = kSmemBudget - 1024 - tail_bytes - 256;
    const int bslot = kPair1dGroup * static_cast<int>(a.b_tile_bytes);
    int nA = (a.nseg * a.kchunks >= 3) ? 3 : 2;
    int nB = (budget - nA * a.halo_slot_bytes) / bslot;
    if (nB > kPairMaxB) nB = kPairMaxB;
    MQ_REQUIRE(nB >= 2, "mq_conv_gemm: 1-D pair mode does not fit shared memory (bn=%d)", p->bn);
    a.halo_nA = nA; a.halo_nB = nB;
    smem = 1024 + nA * a.halo_slot_bytes + nB * bslot + tail_bytes;
    const int pairs = sms / 2;
    grid = 2 * (a.num_tiles < pairs ? a.num_tiles : pairs);
  } else if (pair) {
    const int R = kHaloSubRows * a.msub;
    a.pair_tx0 = (R + 2) * kHaloW * 128;
    const int box1 = (R + 1) * kHaloW * 128;
    a.pair_tx1 = 2 * box1;
    a.pair_boxb_off = (box1 + 1023) / 1024 * 1024;
    int slot = a.pair_tx0;
    if (up && a.pair_boxb_off + box1 > slot) slot = a.pair_boxb_off + box1;
    a.halo_slot_bytes = (slot + 1023) / 1024 * 1024;
    // weight ring: slots of pair_bgrp taps (3 = one filter row; every tap count on this path is a
    // multiple of 3) -> one full/empty barrier round trip and one multicast commit per slot
    int bgrp = 3, nA = 0, nB = 0, tail_bytes = 0, budget = 0;
    auto plan_ring = [&](int stage_b) {
      tail_bytes = (2 * kPairMaxA + 2 * kPairMaxB + 2 * kMaxAccBufs) * 8 + 16 + kBiasSmemFloats * 4 + stage_b;
      budget = kSmemBudget - 1024 - tail_bytes - 256;
      bgrp = 3;
      if (env.bgrp == 1 || env.bgrp == 3) bgrp = env.bgrp;
      if (p->taps % 3 != 0 || (up && p->up_taps % 3 != 0)) bgrp = 1;
      for (;;) {
        const int bslot = bgrp * static_cast<int>(a.b_tile_bytes);
        nA = (a.nseg * a.kchunks + a.kchunks2 >= 3) ? 3 : 2;
        while (nA > 2 && budget - nA * a.halo_slot_bytes < 2 * bslot) --nA;
        nB = (budget - nA * a.halo_slot_bytes) / bslot;
        if (nB >= 2 || bgrp == 1) break;
        bgrp = 1;
      }
      return nB * bgrp;                                   // weight taps in flight
    };
    int taps_in_flight = plan_ring(stage_bytes);
    // The staging tile costs 16 KB of ring.  Where the activation slots leave little room (fused up-concat, bn = 128:
    // two 85 KB slots) that drops the weight ring from two 3-tap slots to four 1-tap slots and the layer from 1.08 to
    // 1.52 ms (profiles/conv_probe3_r02.log); coalesced stores are worth nothing next to that, so they go first.
    if (a.stage_out && !a.slim16 && env.stage_out < 0 && taps_in_flight < 9) {
      const int staged = taps_in_flight, staged_grp = bgrp;
      const int plain = plan_ring(0);
      if (plain > staged || bgrp > staged_grp) {
        a.stage_out = 0;
        stage_bytes = 0;
        taps_in_flight = plain;
      } else {
        taps_in_flight = plan_ring(stage_bytes);
      }
    }
    (void)taps_in_flight;
    if (nB > kPairMaxB) nB = kPairMaxB;
    MQ_REQUIRE(nB >= 2, "mq_conv_gemm: pair mode does not fit shared memory (msub=%d bn=%d)", a.msub, p->bn);
    a.halo_nA = nA; a.halo_nB = nB; a.pair_bgrp = bgrp;
    smem = 1024 + nA * a.halo_slot_bytes + nB * bgrp * static_cast<int>(a.b_tile_bytes) + tail_bytes;
    // weights-stationary: a single N tile whose taps x chunks fit beside two halo slots (the 64 -> 64 and 64 -> 128
    // layers) keeps its weights in shared memory for the whole kernel instead of cycling them through the ring
    if (env.b_resident && !up && a.nseg == 1 && a.tiles_n == 1 && p->taps % bgrp == 0) {
      const int resident = p->taps * a.kchunks * static_cast<int>(a.b_tile_bytes);
      const int nA2 = 2;
      if (nA2 * a.halo_slot_bytes + resident <= budget) {
        a.b_resident = 1;
        a.halo_nA = nA2;
        a.halo_nB = 1;
        smem = 1024 + nA2 * a.halo_slot_bytes + resident + tail_bytes;
      }
    }
    const int pairs = sms / 2;
    grid = 2 * (a.num_tiles < pairs ? a.num_tiles : pairs);
  }
  if (p->out_pool != nullptr)
    MQ_REQUIRE(lean && !up && p->bw == 8 && p->bh % 2 == 0 && p->H % 2 == 0 && p->pool_ld % 8 == 0 &&
                   (reinterpret_cast<uintptr_t>(p->out_pool) & 15) == 0,
               "mq_conv_gemm: out_pool needs the bf16-only epilogue, an 8-pixel-wide sub-tile with an even number of rows, even H");
#define MQ_LAUNCH_CONV(FAST, LEAN)                                                                          \
  do {                                                                                                      \
    MQ_CUDA_OK(cudaFuncSetAttribute(conv_gemm_kernel<FAST, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    conv_gemm_kernel<FAST, LEAN><<<grid, kThreads, smem, stream>>>(map_a, map_b, map_a2, a);                       \
  } while (0)
#define MQ_LAUNCH_HALO(FAST, LEAN)                                                                          \
  do {                                                                                                      \
    MQ_CUDA_OK(cudaFuncSetAttribute(conv_halo_kernel<FAST, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    conv_halo_kernel<FAST, LEAN><<<grid, kThreads, smem, stream>>>(map_a, map_b, a);                        \
  } while (0)
#define MQ_LAUNCH_PAIR(FAST, LEAN)                                                                          \
  do {                                                                                                      \
    MQ_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<FAST, LEAN, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    conv_pair_kernel<FAST, LEAN, 8><<<grid, kThreads, smem, stream>>>(map_a, map_b, map_a2, a);             \
  } while (0)
#define MQ_LAUNCH_PAIR1D(FAST, LEAN)                                                                        \
  do {                                                                                                      \
    MQ_CUDA_OK(cudaFuncSetAttribute(conv_pair1d_kernel<FAST, LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    conv_pair1d_kernel<FAST, LEAN><<<grid, kThreads, smem, stream>>>(map_a, map_b, a);                      \
  } while (0)
  if (pair1d) {
    if (p->fast_tanh) {
      if (lean) MQ_LAUNCH_PAIR1D(true, true); else MQ_LAUNCH_PAIR1D(true, false);
    } else {
      if (lean) MQ_LAUNCH_PAIR1D(false, true); else MQ_LAUNCH_PAIR1D(false, false);
    }
  } else if (pair && a.slim16) {
    MQ_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<true, true, kEpi16Warps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    conv_pair_kernel<true, true, kEpi16Warps><<<grid, kThreads16, smem, stream>>>(map_a, map_b, map_a2, a);
  } else if (pair) {
    if (p->fast_tanh) {
      if (lean) MQ_LAUNCH_PAIR(true, true); else MQ_LAUNCH_PAIR(true, false);
    } else {
      if (lean) MQ_LAUNCH_PAIR(false, true); else MQ_LAUNCH_PAIR(false, false);
    }
  } else if (halo) {
    if (p->fast_tanh) {
      if (lean) MQ_LAUNCH_HALO(true, true); else MQ_LAUNCH_HALO(true, false);
    } else {
      if (lean) MQ_LAUNCH_HALO(false, true); else MQ_LAUNCH_HALO(false, false);
    }
  } else if (p->fast_tanh) {
    if (lean) MQ_LAUNCH_CONV(true, true); else MQ_LAUNCH_CONV(true, false);
  } else {
    if (lean) MQ_LAUNCH_CONV(false, true); else MQ_LAUNCH_CONV(false, false);
  }
#undef MQ_LAUNCH_CONV
#undef MQ_LAUNCH_HALO
#undef MQ_LAUNCH_PAIR
#undef MQ_LAUNCH_PAIR1D
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}
