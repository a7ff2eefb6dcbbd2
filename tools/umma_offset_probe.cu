// Probe: can a tcgen05.mma A descriptor start at a row that is NOT 1024-byte aligned inside a
// SWIZZLE_128B tile (matrix base_offset field), and can the 8-row-group stride (SBO) be 1280 B?
// This is the addressing the halo-tile convolution needs (one smem halo, 9 shifted descriptors).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_offset_probe umma_offset_probe.cu -lcuda? (driver entry via runtime)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include "../mqgan_b200/csrc/common.cuh"
using namespace mq;

namespace mq { void set_last_error(const char*, ...) {} }

struct Args { int row_shift; int sbo_bytes; int use_base_offset; int rows_per_group_pitch; };

// A region: ROWS x 64 bf16 loaded by TMA (SW128) as a 2-D box {64, ROWS}; B: 64 x 64.
constexpr int ROWS = 200;
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
             Args args, float* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* sa = smem;                       // ROWS*128 B (<= 25600) -> 26 KB
  uint8_t* sb = smem + 26 * 1024;           // 64*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 36 * 1024);
  uint32_t* tptr = reinterpret_cast<uint32_t*>(bars + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); fence_barrier_init(); }
  if (warp == 1) { tmem_alloc(tptr, 64); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = *tptr;
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bars[0], ROWS * 128 + 64 * 128);
    tma_load_2d(&map_a, &bars[0], sa, 0, 0);
    tma_load_2d(&map_b, &bars[0], sb, 0, 0);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t a_start = smem_u32(sa) + args.row_shift * 128;
    uint64_t da = 0;
    da |= static_cast<uint64_t>((a_start & 0x3FFFFu) >> 4);
    da |= static_cast<uint64_t>(1) << 16;
    da |= static_cast<uint64_t>(args.sbo_bytes >> 4) << 32;
    da |= static_cast<uint64_t>(1) << 46;
    if (args.use_base_offset) da |= static_cast<uint64_t>((a_start >> 7) & 7) << 49;
    da |= static_cast<uint64_t>(2) << 61;
    const uint64_t db = umma_desc_sw128(smem_u32(sb));
    const uint32_t idesc = umma_idesc_bf16(128, 64);
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, da + 2 * k, db + 2 * k, idesc, k != 0);
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  uint32_t v[32];
  for (int c = 0; c < 64; c += 32) {
    __syncwarp();
    tmem_ld_32x32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tc_fence_before(); __syncthreads();
  if (warp == 1) { __syncwarp(); tmem_dealloc(tmem, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  EncodeTiledFn encode = (EncodeTiledFn)fp;
  std::vector<__nv_bfloat16> ha(ROWS * 64), hb(64 * 64);
  std::vector<float> fa(ROWS * 64), fb(64 * 64);
  srand(1);
  for (int i = 0; i < ROWS * 64; ++i) { float v = (rand() % 17 - 8) / 8.0f; ha[i] = __float2bfloat16(v); fa[i] = v; }
  for (int i = 0; i < 64 * 64; ++i) { float v = (rand() % 13 - 6) / 8.0f; hb[i] = __float2bfloat16(v); fb[i] = v; }
  __nv_bfloat16 *da_, *db_; float* dout;
  cudaMalloc(&da_, ha.size() * 2); cudaMalloc(&db_, hb.size() * 2); cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da_, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db_, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap ma, mb;
  { cuuint64_t d[2] = {64, ROWS}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, ROWS}; cuuint32_t e[2] = {1, 1};
    CUresult r = encode(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da_, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode A failed %d\n", (int)r); return 1; } }
  { cuuint64_t d[2] = {64, 64}; cuuint64_t s[1] = {128}; cuuint32_t b[2] = {64, 64}; cuuint32_t e[2] = {1, 1};
    CUresult r = encode(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db_, d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode B failed %d\n", (int)r); return 1; } }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
  std::vector<float> hout(128 * 64);
  // cases: (row_shift, sbo, use_base_offset)
  int cases[][3] = {{0, 1024, 0}, {8, 1024, 0}, {3, 1024, 1}, {3, 1024, 0}, {11, 1024, 1}, {0, 1280, 0}, {8, 1280, 0},
                    {1, 1280, 1}, {11, 1280, 1}, {11, 1280, 0}, {21, 1280, 1}};
  for (auto& c : cases) {
    Args a{c[0], c[1], c[2], 0};
    cudaMemset(dout, 0, 128 * 64 * 4);
    probe_kernel<<<1, 128, 40 * 1024>>>(ma, mb, a, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("case shift=%d sbo=%d bo=%d: CUDA error %s\n", c[0], c[1], c[2], cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hout.data(), dout, hout.size() * 4, cudaMemcpyDeviceToHost);
    // expected: D[m][n] = sum_k A[row(m)][k] * B[n][k], row(m) = shift + (m/8)*(sbo/128) + m%8
    double maxerr = 0;
    for (int m = 0; m < 128; ++m) {
      int row = c[0] + (m / 8) * (c[1] / 128) + (m % 8);
      for (int n = 0; n < 64; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += (double)fa[row * 64 + k] * fb[n * 64 + k];
        maxerr = fmax(maxerr, fabs(acc - hout[m * 64 + n]));
      }
    }
    printf("row_shift=%2d sbo=%4d base_offset=%d -> max |err| = %.4g  %s\n", c[0], c[1], c[2], maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  }
  return 0;
}
