// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX
// wrappers and small math.  Hand-written inline PTX; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mq {

// ---------------------------------------------------------------------------
// error plumbing shared by the C-ABI entry points (never throw across the ABI)
// ---------------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define MQ_CUDA_OK(expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      mq::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                \
                         cudaGetErrorString(_e));                                     \
      return 2;                                                                       \
    }                                                                                 \
  } while (0)
#define MQ_REQUIRE(cond, ...)                                                         \
  do {                                                                                \
    if (!(cond)) {                                                                    \
      mq::set_last_error(__VA_ARGS__);                                                \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

// ---------------------------------------------------------------------------
// shared-memory address / mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// try_wait suspends the thread until the phase completes or a time limit passes.  With the default limit a waiting warp
// comes back every ~150 clocks: in the wide refiner layers, whose epilogue warps wait for an accumulator 70 % of the time,
// these polls were half of all executed instructions (ncu source page of mid.conv1) - issue slots and power on a
// power-capped step.  The hint asks for a longer sleep; completion still wakes the thread at once.
#ifndef MQ_MBAR_SUSPEND_NS
#define MQ_MBAR_SUSPEND_NS 2000
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(static_cast<uint32_t>(MQ_MBAR_SUSPEND_NS))
      : "memory");
  return ok != 0;
}

// Bounded spin: a mis-programmed pipeline traps (reported as a launch failure) instead of
// hanging the GPU.  2^24 polls of a try_wait that sleeps 0.1 - 2 us each is 1 - 30 seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}

// Waits that are long by design - an epilogue warp waiting for the next accumulator of a wide layer (70 % of its time),
// the producer waiting for a free ring slot: after a few polls, sleep between polls.  The hardware caps try_wait's own
// suspend time at ~150 clocks whatever the hint, so without this the polls stay 40 % of a wide layer's instructions.
#ifndef MQ_RELAXED_WAIT_NS
#define MQ_RELAXED_WAIT_NS 256
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
    if (MQ_RELAXED_WAIT_NS > 0 && spins > 4) __nanosleep(MQ_RELAXED_WAIT_NS);
  }
}

// barrier among a subset of the CTA's warps (id 1..15; id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA, commit, load, fences
// ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem], bf16 operands, fp32 accumulate, M = 128.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrive on an mbarrier once every previously issued tcgen05.mma has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane
// (lane_base + i), registers v[0..31] = columns col .. col+31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
// 32 lanes x 16 consecutive fp32 columns (half the registers of the x32 form: the 16-epilogue-warp kernels)
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile (rows of 128 B, 8-row groups 1024 B
// apart).  Field layout: start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;            // LBO (ignored for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;    // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;            // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;            // SWIZZLE_128B
  return d;
}

// The same descriptor split into 32-bit halves: only the low word (start address, LBO) changes
// between MMAs of one main loop, so the issue loop advances a 32-bit value and keeps the high word
// (SBO, version, swizzle mode) constant.  Shared-memory addresses stay below 2^18, so adding tap /
// sub-tile / K-step offsets (in 16-byte units) never carries out of the 14-bit address field.
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
}
__host__ __device__ constexpr uint32_t umma_desc_hi_sw128(uint32_t sbo_bytes) {
  return (sbo_bytes >> 4) | (1u << 14) | (2u << 29);
}
__device__ __forceinline__ uint64_t umma_desc_make(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

// One lane of a fully active warp (the tcgen05 issue sites are warp-uniform code guarded by this, so
// descriptors live in uniform registers and ptxas emits no per-lane waterfall loop around UTCHMMA).
__device__ __forceinline__ bool elect_one_sync() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}

// kind::f16 instruction descriptor: C=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16
// [10,13)=1, both K-major, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// same with FP16 operands (A/B format 0)
__host__ __device__ __forceinline__ uint32_t umma_idesc_f16(uint32_t m, uint32_t n) {
  return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ---------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster on the two SMs of a TPC run one M = 256 MMA.
// Each CTA stages its own 128 accumulator rows of A and HALF of the B tile; the leader (cluster
// rank 0) issues the MMAs, and every barrier the two CTAs share lives in the leader's shared memory.
// A shared::cta address is also a valid shared::cluster address of the own CTA; clearing bit 24
// (the CTA rank inside a pair) turns it into the leader's copy of the same variable.
// ---------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on a barrier given by its shared::cluster address (possibly in the peer CTA)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta): a .cluster-scope release would put a MEMBAR.ALL.GPU in front of
  // the arrive and stall the epilogue warp on its own outstanding global stores
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// M = 256 (128 rows per CTA), issued by the leader only; descriptors are CTA-local offsets that
// the hardware applies in both CTAs (same shared-memory layout on both sides).
__device__ __forceinline__ void umma_bf16_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all prior MMAs retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}
// TMA loads whose completion bytes are counted on the LEADER's barrier (bar_leader is the leader's
// shared::cluster address); the data lands in the issuing CTA's own shared memory.
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap* m, uint32_t bar_leader, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2cta(const CUtensorMap* m, uint32_t bar_leader, void* dst, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_2cta(const CUtensorMap* m, uint32_t bar_leader, void* dst, int c0, int c1,
                                                 int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------------------
// math
// ---------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// tanh with ~1e-7 absolute error: 1 - 2 / (exp(2x) + 1), via ex2.approx + rcp.
// APTx uses 1 + tanh(.), so absolute (not relative) accuracy is what matters.
__device__ __forceinline__ float tanh_precise(float x) {
  float xc = fminf(fmaxf(x, -15.0f), 15.0f);
  float e = __expf(2.0f * xc);
  return 1.0f - __fdividef(2.0f, e + 1.0f);
}

template <bool kFast>
__device__ __forceinline__ float aptx(float v, float beta, float gamma) {
  float t = kFast ? tanh_fast(beta * v) : tanh_precise(beta * v);
  float gv = gamma * v;
  return fmaf(gv, t, gv);
}

__device__ __forceinline__ float sigmoid_precise(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// fp32 -> three bf16 terms x0 + x1 + x2 == x (24 significant bits, exact unless
// the low terms underflow).
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b,
                                       __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  float r = x - __bfloat162float(a);
  b = __float2bfloat16_rn(r);
  r -= __bfloat162float(b);
  c = __float2bfloat16_rn(r);
}

// Multi-term operand splits for fp32-grade tensor-core GEMMs.  kind 0: three bf16 terms (24 bits);
// kind 1: two fp16 terms (22 bits; |x| is clamped to the fp16 range, tiny terms bottom out at the
// fp16 subnormal spacing 2^-24, an absolute error that is irrelevant next to O(1) dot products).
__device__ __forceinline__ int split_nterms(int kind) { return kind == 1 ? 2 : 3; }

__device__ __forceinline__ void split_terms(float x, int kind, uint16_t (&t)[3]) {
  if (kind == 1) {
    const float xs = fminf(fmaxf(x, -65504.0f), 65504.0f);
    const __half h0 = __float2half_rn(xs);
    const __half h1 = __float2half_rn(xs - __half2float(h0));
    t[0] = __half_as_ushort(h0);
    t[1] = __half_as_ushort(h1);
    t[2] = 0;
  } else {
    __nv_bfloat16 a, b, c;
    split3(x, a, b, c);
    t[0] = __bfloat16_as_ushort(a);
    t[1] = __bfloat16_as_ushort(b);
    t[2] = __bfloat16_as_ushort(c);
  }
}

// four consecutive channels -> nterms 8-byte stores, term j at element offset j*seg
__device__ __forceinline__ void store_terms4(uint16_t* base, int64_t seg, int kind, const float (&y)[4]) {
  uint16_t t[4][3];
#pragma unroll
  for (int e = 0; e < 4; ++e) split_terms(y[e], kind, t[e]);
  const int nt = split_nterms(kind);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (j < nt) {
      uint2 u;
      u.x = static_cast<uint32_t>(t[0][j]) | (static_cast<uint32_t>(t[1][j]) << 16);
      u.y = static_cast<uint32_t>(t[2][j]) | (static_cast<uint32_t>(t[3][j]) << 16);
      *reinterpret_cast<uint2*>(base + j * seg) = u;
    }
  }
}

__device__ __forceinline__ void store_terms1(uint16_t* base, int64_t seg, int kind, float v) {
  uint16_t t[3];
  split_terms(v, kind, t);
  const int nt = split_nterms(kind);
  for (int j = 0; j < nt; ++j) base[j * seg] = t[j];
}

}  // namespace mq
