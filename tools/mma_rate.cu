// Microbenchmark: issue rate of tcgen05.mma (kind::f16, M=128, cta_group::1) from resident shared memory.
#include <cstdio>
#include <cuda_runtime.h>
#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

__global__ void __launch_bounds__(128, 1) mma_rate(int n_mma, int N, int two_acc, int kadv, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (48 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tmem_base_smem), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, (uint32_t)N);
    const uint64_t da = make_smem_desc(base, 16, 1024, kLayoutSW128);
    const uint64_t db = make_smem_desc(base + 16384, 16, 1024, kLayoutSW128);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tm + ((two_acc && (i & 1)) ? 256u : 0u);
      const uint32_t k = kadv ? (uint32_t)((i >> (two_acc ? 1 : 0)) & 3) * 2u : 0u;
      umma_f16(d, da + k, db + k, idesc, i > 1 ? 1u : 0u);
    }
    umma_commit(smem_u32(&bar));
    long long t1 = clock64();
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int cfgs[][3] = {{256, 0, 1}, {256, 1, 1}, {128, 0, 1}, {128, 1, 1}, {64, 0, 1}, {256, 0, 0}, {16, 0, 1}};
  for (auto& c : cfgs) for (int grid : {1, 148}) {
    const int n = 2048;
    mma_rate<<<grid, 128, 64 * 1024>>>(n, c[0], c[1], c[2], d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("N=%3d two_acc=%d kadv=%d grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s)\n", c[0], c[1], c[2], grid,
           (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(e));
  }
  return 0;
}
