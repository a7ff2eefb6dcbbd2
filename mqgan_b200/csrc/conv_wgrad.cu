// Weight gradient of the implicit-GEMM convolutions (training step, SURVEY 8-f4 / BASELINE configs[4]):
//
//   dW[tap][co][ci] = sum_{n,h,w} dY[n, h, w, co] * X[n, h + dh[tap], w + dw[tap], ci]
//
// i.e. what autograd computes for the weights of F.conv2d (preencoder.py:97-98), F.conv1d
// (attentions.py:471-474, 532-541) and F.linear (preencoder.py:433, 486, 490) in the reference's
// training step (train.py:380-501 -> loss.backward()).  Zero padding of the convolution is TMA
// out-of-bounds fill on the shifted X box, exactly as in the forward kernel.
//
// As a GEMM the contraction runs over PIXELS: D (M = 128 output channels) x (N = bn input channels),
// K = pixels.  Both operands are channel-last in HBM, so a [64 pixels][64 channels] TMA box lands in
// shared memory as 64 rows of 128 bytes with the 128-byte swizzle, which is the canonical *MN-major*
// tcgen05 operand tile (K along rows): no transposes anywhere.  Per 64-pixel K block a CTA stages
// 2 dY atoms (128 co) + bn/64 X atoms and issues four K = 16 MMAs into one TMEM accumulator; a CTA
// owns one (tap, co tile, ci tile, K split) and writes its fp32 partial tile once at the end
// (partials [split][taps][cout][cin], summed by the caller: deterministic, no atomics).
//
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (owns TMEM), warps 2-5 epilogue.
#include <stdio.h>
#include <string.h>

#include <mutex>

#include "../../include/mqgan_b200.h"
#include "common.cuh"

namespace mq {

constexpr int kWgThreads = 192;
constexpr int kWgStages = 4;
constexpr int kWgKPix = 64;                      // pixels per K block (4 MMAs of K = 16)
constexpr int kWgAtomBytes = kWgKPix * 128;      // [64 pixel rows][64 channels] bf16
constexpr int kWgTileM = 128;
constexpr int kWgMaxBn = 256;
constexpr int kWgAStage = 2 * kWgAtomBytes;
constexpr int kWgBStage = (kWgMaxBn / 64) * kWgAtomBytes;

struct WgArgs {
  int taps;
  int dh[MQ_MAX_TAPS], dw[MQ_MAX_TAPS];
  int cout, cin;
  int co_tiles, ci_tiles, bn;
  int tiles_w, tiles_h, n_img;
  int bh, bw;
  int kblocks, split, kb_per_split;
  float* out;
};

// MN-major, 128-byte-swizzled operand: 64-channel atoms [K rows][128 B]; LBO = distance between atoms
// (next 64 channels), SBO = distance between 8-row groups along K.
__device__ __forceinline__ uint32_t wg_desc_hi() {
  return (1024u >> 4) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint32_t wg_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (static_cast<uint32_t>(kWgAtomBytes >> 4) << 16);
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x,
                  const WgArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;                                   // [stages][2 atoms]
  uint8_t* smem_b = smem_a + kWgStages * kWgAStage;         // [stages][bn/64 atoms]
  const int b_stage = (a.bn / 64) * kWgAtomBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + kWgStages * kWgBStage);
  uint64_t* empty_bar = full_bar + kWgStages;
  uint64_t* tfull_bar = empty_bar + kWgStages;
  uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // work item
  int idx = blockIdx.x;
  const int sp = idx % a.split; idx /= a.split;
  const int ci_t = idx % a.ci_tiles; idx /= a.ci_tiles;
  const int co_t = idx % a.co_tiles;
  const int tap = idx / a.co_tiles;
  const int kb0 = sp * a.kb_per_split;
  const int kb1 = min(a.kblocks, kb0 + a.kb_per_split);
  const int co0 = co_t * kWgTileM;
  const int ci0 = ci_t * a.bn;
  const int a_atoms = (a.cout - co0 > 64) ? 2 : 1;                       // atoms that hold real channels
  const int b_atoms = min(a.bn / 64, (a.cin - ci0 + 63) / 64);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < kWgStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr_s, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_s;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int per_img = a.tiles_h * a.tiles_w;
      const int ddh = a.dh[tap], ddw = a.dw[tap];
      for (int kb = kb0; kb < kb1; ++kb) {
        const int n = kb / per_img;
        const int r = kb - n * per_img;
        const int th = r / a.tiles_w;
        const int tw = r - th * a.tiles_w;
        const int h0 = th * a.bh, w0 = tw * a.bw;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>((a_atoms + b_atoms) * kWgAtomBytes));
        uint8_t* sa = smem_a + stage * kWgAStage;
        uint8_t* sb = smem_b + stage * b_stage;
        for (int i = 0; i < a_atoms; ++i)
          tma_load_4d(&map_dy, &full_bar[stage], sa + i * kWgAtomBytes, co0 + 64 * i, w0, h0, n);
        for (int i = 0; i < b_atoms; ++i)
          tma_load_4d(&map_x, &full_bar[stage], sb + i * kWgAtomBytes, ci0 + 64 * i, w0 + ddw, h0 + ddh, n);
        if (++stage == kWgStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // both operands MN-major: a_major (bit 15) and b_major (bit 16) set
    const uint32_t idesc = umma_idesc_bf16(kWgTileM, static_cast<uint32_t>(a.bn)) | (1u << 15) | (1u << 16);
    const uint32_t hi = wg_desc_hi();
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = kb0; kb < kb1; ++kb) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      const uint32_t a_lo = wg_desc_lo(smem_u32(smem_a + stage * kWgAStage));
      const uint32_t b_lo = wg_desc_lo(smem_u32(smem_b + stage * b_stage));
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < kWgKPix / 16; ++k)          // 16 pixel rows = 2048 bytes per K step
          umma_bf16(tmem_base, umma_desc_make(a_lo + k * (2048 >> 4), hi), umma_desc_make(b_lo + k * (2048 >> 4), hi),
                    idesc, (kb != kb0 || k != 0) ? 1u : 0u);
        umma_commit(&empty_bar[stage]);
      }
      __syncwarp();
      if (++stage == kWgStages) { stage = 0; phase ^= 1; }
    }
    if (elect_one_sync()) umma_commit(tfull_bar);
    __syncwarp();
  } else {
    // epilogue: TMEM lane = output channel, columns = input channels
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    const int co = co0 + q * 32 + lane;
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    float* orow = a.out + ((static_cast<size_t>(sp) * a.taps + tap) * a.cout + co) * a.cin + ci0;
    const bool vec = (a.cin & 3) == 0;
    for (int c = 0; c < a.bn; c += 32) {
      if (ci0 + c >= a.cin) break;                // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c, v);
      tmem_ld_wait();
      if (co < a.cout) {
        if (kb1 <= kb0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0u;
        }
        if (vec && ci0 + c + 32 <= a.cin) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(orow + c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (ci0 + c + j < a.cin) orow[c + j] = __uint_as_float(v[j]);
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 256);
  }
}

typedef CUresult (*WgEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static WgEncodeFn wg_encode_fn() {
  static WgEncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<WgEncodeFn>(p);
  });
  return fn;
}

static int wg_choose_bn(int cin) {
  int best = 64, best_cost = 1 << 30;
  for (int bn = 64; bn <= kWgMaxBn; bn += 64) {
    const int tiles = (cin + bn - 1) / bn;
    const int cost = tiles * bn * 8 + tiles;          // padded columns first, then fewer tiles
    if (cost < best_cost || (cost == best_cost && bn > best)) { best = bn; best_cost = cost; }
  }
  return best;
}

}  // namespace mq

using namespace mq;

extern "C" int mq_conv_wgrad_split(const mq_wgrad_params* p) {
  if (p == nullptr || p->N <= 0 || p->H <= 0 || p->W <= 0 || p->cout <= 0 || p->cin <= 0 || p->taps <= 0 ||
      p->bh * p->bw != kWgKPix)
    return 0;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int bn = wg_choose_bn(p->cin);
  const long long tiles = static_cast<long long>(p->taps) * ((p->cout + kWgTileM - 1) / kWgTileM) * ((p->cin + bn - 1) / bn);
  const long long kblocks = static_cast<long long>(p->N) * ((p->H + p->bh - 1) / p->bh) * ((p->W + p->bw - 1) / p->bw);
  // One CTA per SM is resident (197 KB of shared memory), so time ~ waves x (K blocks per CTA + fixed cost): pick the split
  // that minimises it.  (The first version took ceil(2 * SMs / tiles): 360 CTAs = 2.4 waves on mid.conv1, a third of the
  // last wave's SMs idle - profiles/ncu_conv_wgrad_mid_r01.csv.)
  const long long fixed = 8;                               // prologue + epilogue in K-block units
  long long split = 1, best_cost = -1;
  for (long long sp = 1; sp <= 64; ++sp) {
    const long long per = (kblocks + sp - 1) / sp;
    if (sp > 1 && per < 16) break;
    const long long waves = (tiles * sp + sms - 1) / sms;
    const long long cost = waves * (per + fixed);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; split = sp; }
  }
  return static_cast<int>(split);
}

extern "C" int mq_conv_wgrad(const mq_wgrad_params* p, mq_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  MQ_REQUIRE(p != nullptr, "mq_conv_wgrad: null params");
  MQ_REQUIRE(p->dy && p->x && p->dw, "mq_conv_wgrad: null pointer argument");
  MQ_REQUIRE(p->N > 0 && p->H > 0 && p->W > 0, "mq_conv_wgrad: N=%d H=%d W=%d", p->N, p->H, p->W);
  MQ_REQUIRE(p->cout > 0 && p->cin > 0 && p->cout <= p->dy_ld && p->cin <= p->x_ld, "mq_conv_wgrad: cout=%d (ld %d) cin=%d (ld %d)",
             p->cout, p->dy_ld, p->cin, p->x_ld);
  MQ_REQUIRE(p->dy_ld % 8 == 0 && p->x_ld % 8 == 0, "mq_conv_wgrad: channel pitches must be multiples of 8 (16-byte TMA strides)");
  MQ_REQUIRE(p->taps >= 1 && p->taps <= MQ_MAX_TAPS, "mq_conv_wgrad: taps=%d", p->taps);
  MQ_REQUIRE(p->bh >= 1 && p->bw >= 1 && p->bh * p->bw == kWgKPix && p->bh <= 256 && p->bw <= 256,
             "mq_conv_wgrad: bh*bw must be %d (bh=%d bw=%d)", kWgKPix, p->bh, p->bw);
  MQ_REQUIRE(p->split >= 1, "mq_conv_wgrad: split=%d", p->split);
  MQ_REQUIRE((reinterpret_cast<uintptr_t>(p->dy) & 15) == 0 && (reinterpret_cast<uintptr_t>(p->x) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(p->dw) & 15) == 0,
             "mq_conv_wgrad: buffers must be 16-byte aligned");
  WgEncodeFn encode = wg_encode_fn();
  MQ_REQUIRE(encode != nullptr, "mq_conv_wgrad: cuTensorMapEncodeTiled not available from the driver");

  WgArgs a;
  memset(&a, 0, sizeof(a));
  a.taps = p->taps;
  for (int i = 0; i < p->taps; ++i) { a.dh[i] = p->tap_dh[i]; a.dw[i] = p->tap_dw[i]; }
  a.cout = p->cout; a.cin = p->cin;
  a.bn = wg_choose_bn(p->cin);
  a.co_tiles = (p->cout + kWgTileM - 1) / kWgTileM;
  a.ci_tiles = (p->cin + a.bn - 1) / a.bn;
  a.bh = p->bh; a.bw = p->bw;
  a.tiles_h = (p->H + p->bh - 1) / p->bh;
  a.tiles_w = (p->W + p->bw - 1) / p->bw;
  a.n_img = p->N;
  const long long kblocks = static_cast<long long>(p->N) * a.tiles_h * a.tiles_w;
  MQ_REQUIRE(kblocks < (1LL << 31), "mq_conv_wgrad: too many pixel blocks");
  a.kblocks = static_cast<int>(kblocks);
  a.split = p->split;
  a.kb_per_split = static_cast<int>((kblocks + p->split - 1) / p->split);
  a.out = p->dw;

  CUtensorMap map_dy, map_x;
  for (int which = 0; which < 2; ++which) {
    const void* base = which == 0 ? p->dy : p->x;
    const int ld = which == 0 ? p->dy_ld : p->x_ld;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(ld), static_cast<cuuint64_t>(p->W), static_cast<cuuint64_t>(p->H),
                          static_cast<cuuint64_t>(p->N)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(p->W) * ld * 2,
                             static_cast<cuuint64_t>(p->H) * p->W * ld * 2};
    cuuint32_t box[4] = {64, static_cast<cuuint32_t>(p->bw), static_cast<cuuint32_t>(p->bh), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(which == 0 ? &map_dy : &map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims,
                        strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    MQ_REQUIRE(r == CUDA_SUCCESS, "mq_conv_wgrad: cuTensorMapEncodeTiled(%s) failed with %d (ld=%d W=%d H=%d N=%d)",
               which == 0 ? "dy" : "x", (int)r, ld, p->W, p->H, p->N);
  }

  const int smem = 1024 + kWgStages * (kWgAStage + kWgBStage) + (2 * kWgStages + 1) * 8 + 16;
  MQ_CUDA_OK(cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const long long grid = static_cast<long long>(a.taps) * a.co_tiles * a.ci_tiles * a.split;
  MQ_REQUIRE(grid < (1LL << 31), "mq_conv_wgrad: grid too large");
  conv_wgrad_kernel<<<static_cast<unsigned>(grid), kWgThreads, smem, stream>>>(map_dy, map_x, a);
  MQ_CUDA_OK(cudaGetLastError());
  return 0;
}
