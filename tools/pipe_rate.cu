// Microbenchmark: TMA -> smem ring -> tcgen05.mma pipeline throughput (1 producer thread, 1 MMA thread).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

struct P { int tiles, stages, rows, mma_per_tile, N, do_mma, do_tma, box2, spinners, a_shift, sleep_ns; };

__global__ void __launch_bounds__(384, 1) pipe(const __grid_constant__ CUtensorMap map, P p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[8], empty[8], done;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t tile_bytes = p.rows * 128;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
    mbar_init(smem_u32(&done), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base_smem;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < p.tiles; ++t) {
      while (!mbar_try_wait(smem_u32(&empty[s]), ph ^ 1)) {}
      if (p.do_tma) {
        mbar_arrive_expect_tx(smem_u32(&full[s]), tile_bytes);
        const int row = ((t * 37 + blockIdx.x * 11) % 64) * p.rows;
        if (p.box2) {
          tma_load_2d(&map, smem_u32(&full[s]), base + s * tile_bytes, 0, row);
          tma_load_2d(&map, smem_u32(&full[s]), base + s * tile_bytes + tile_bytes / 2, 64, row);
        } else {
          tma_load_2d(&map, smem_u32(&full[s]), base + s * tile_bytes, (t & 3) * 64, row);
        }
      } else {
        mbar_arrive(smem_u32(&full[s]));
      }
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, (uint32_t)p.N);
    int s = 0; uint32_t ph = 0;
    for (int t = 0; t < p.tiles; ++t) {
      while (!mbar_try_wait(smem_u32(&full[s]), ph)) {}
      tc_fence_after();
      if (p.do_mma) {
        const uint64_t da = make_smem_desc(base + s * tile_bytes + p.a_shift * 128, 16, 1024, kLayoutSW128);
        for (int i = 0; i < p.mma_per_tile; ++i) umma_f16(tm + (i & 1) * 256, da + 2 * (i & 3), da + 2 * (i & 3), idesc, 1u);
      }
      umma_commit(smem_u32(&empty[s]));
      if (++s == p.stages) { s = 0; ph ^= 1; }
    }
    umma_commit(smem_u32(&done));
    while (!mbar_try_wait(smem_u32(&done), 0)) {}
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  else if (threadIdx.x >= 128 && (int)threadIdx.x < 128 + p.spinners) {
    // epilogue-like warps spinning on a barrier that completes at the end
    while (!mbar_try_wait(smem_u32(&done), 0)) { if (p.sleep_ns) __nanosleep(p.sleep_ns); }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)fnp;
  const int K = 256, ROWS = 64 * 256 + 256;
  void* g; cudaMalloc(&g, (size_t)ROWS * K * 2); cudaMemset(g, 0, (size_t)ROWS * K * 2);
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int rows : {256, 128}) {
    CUtensorMap map; cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)ROWS}; cuuint64_t str[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r) { printf("encode failed %d\n", (int)r); return 1; }
    struct { int stages, mma, spinners, a_shift, sleep; } cf[] = {
      {4, 8, 0, 0, 0}, {4, 8, 256, 0, 0}, {4, 8, 256, 0, 500}, {4, 8, 0, 1, 0}, {4, 8, 0, 2, 0}, {4, 8, 256, 1, 0}};
    for (auto& c : cf) {
      if (rows != 256) continue;
      P p{512, c.stages, rows - 8, c.mma, 256, 1, 1, 0, c.spinners, c.a_shift, c.sleep};
      p.rows = rows;
      pipe<<<148, 384, 200 * 1024>>>(map, p, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
      printf("stages=%d mma/tile=%2d spinners=%3d a_shift=%d sleep=%d: %.0f cyc/tile %s\n", c.stages, c.mma, c.spinners, c.a_shift,
             c.sleep, (double)h / p.tiles, cudaGetErrorString(e));
    }
  }
  return 0;
}
