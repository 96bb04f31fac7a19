// Microbenchmark: rate of tcgen05.mma.cta_group::2 (kind::f16, M = 256 over a CTA pair) from resident shared memory,
// against the single-CTA M = 128 instruction (tools/mma_rate.cu).  a_shift: A descriptor start moved by whole 128-byte
// rows (the flat kernel's tap addressing).
#include <cstdio>
#include <cuda_runtime.h>
#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

__device__ __forceinline__ void umma2_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma2_f16_mask(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  const uint32_t z = 0;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred;
}
// mode 4: disable-output-lane mask form; mode 5: issued by an elected lane of a converged warp, peer CTA's warps exit early
// mode: 0 = SW128 K-major A and B; 1 = no-swizzle K-major A and B; 2 = A from TMEM; 3 = M = 128 over the pair (64 rows per CTA)
__global__ void __launch_bounds__(128, 1) mma_rate2(int n_mma, int N, int a_shift, int n_acc, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t rank = cluster_ctarank();
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (60 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  if (threadIdx.x < 32) { tmem_alloc_pair(smem_u32(&tmem_base_smem), 512); tmem_relinquish_pair(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
  const uint32_t tm = tmem_base_smem;
  if (mode == 5) {
    if (threadIdx.x < 32 && rank == 0) {
      const uint32_t idesc = make_idesc(1u, 0u, 0u, 256u, (uint32_t)N);
      const uint64_t db = make_smem_desc(base + 36864, 16, 1024, kLayoutSW128);
      const uint64_t da = make_smem_desc(base, 16, 1024, kLayoutSW128);
      long long t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const uint32_t k = (uint32_t)(i & 3) * 2u;
        if (elect_one()) umma2_f16(tm, da + k, db + k, idesc, i >= 1 ? 1u : 0u);
        __syncwarp();
      }
      if (elect_one()) umma2_commit(smem_u32(&bar));
      __syncwarp();
      long long t1 = clock64();
      while (!mbar_try_wait(smem_u32(&bar), 0)) {}
      long long t2 = clock64();
      if (blockIdx.x == 0 && threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  } else if (threadIdx.x == 0 && rank == 0) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, mode == 3 ? 128u : 256u, (uint32_t)N);
    const uint64_t db = mode == 1 ? make_smem_desc(base + 36864, 128, 256, 0) : make_smem_desc(base + 36864, 16, 1024, kLayoutSW128);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t d = tm + (uint32_t)(i % n_acc) * 256u;
      const uint32_t k = (uint32_t)((i / n_acc) & 3) * 2u;
      const uint32_t sh = a_shift ? (uint32_t)((i / (4 * n_acc)) % 3) * 128u : 0u;
      const uint64_t da = mode == 1 ? make_smem_desc(base + sh, 128, 256, 0) : make_smem_desc(base + sh, 16, 1024, kLayoutSW128);
      if (mode == 2) umma2_f16_ts(d, tm + 448u + (k >> 1) * 8u, db + k, idesc, i >= n_acc ? 1u : 0u);
      else if (mode == 4) umma2_f16_mask(d, da + k, db + k, idesc, i >= n_acc ? 1u : 0u);
      else umma2_f16(d, da + (mode == 1 ? k * 8u : k), db + (mode == 1 ? k * 8u : k), idesc, i >= n_acc ? 1u : 0u);
    }
    umma2_commit(smem_u32(&bar));
    long long t1 = clock64();
    while (!mbar_try_wait(smem_u32(&bar), 0)) {}
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads(); cluster_sync_all();
  if (threadIdx.x < 32) tmem_dealloc_pair(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 16);
  cudaFuncSetAttribute(mma_rate2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  int cfgs[][4] = {{256, 0, 1, 0}, {256, 1, 1, 0}, {256, 0, 1, 1}, {64, 0, 1, 1}, {256, 0, 1, 2}, {128, 0, 1, 2}, {64, 0, 1, 2},
                   {256, 0, 1, 3}, {64, 0, 1, 3}, {32, 0, 1, 0}, {256, 0, 1, 4}, {256, 0, 1, 5}};
  for (auto& c : cfgs) for (int grid : {2, 148}) {
    const int n = 2048;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 64 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e0 = cudaLaunchKernelEx(&cfg, mma_rate2, n, c[0], c[1], c[2], c[3], d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("pair mode=%d N=%3d a_shift=%d n_acc=%d grid=%3d: issue %.1f cyc/mma, complete %.1f cyc/mma  (%s / %s)\n", c[3], c[0], c[1],
           c[2], grid, (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(e0), cudaGetErrorString(e));
  }
  return 0;
}
