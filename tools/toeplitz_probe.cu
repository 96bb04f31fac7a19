// Probe (round 1, verified on B200: relative L2 error 5.3e-8 with LBO = 16 B / SBO = 128 B, 1.4 with the two swapped —
// profiles/r01_toeplitz_probe.txt): tcgen05 CAN read the im2col ("Toeplitz") operand of a few-channel convolution row
// straight from a CONTIGUOUS shared-memory copy of the input pixels.
//
// Today the image layers (c7s1-64: 3 -> 64 channels, 7x7) use the row-packed operand: pixel q of the A tile is the
// 128-byte window [16 q, 16 q + 128) of the input row (8 pixels x 8 bf16 channels), fetched by TMA as 128 overlapping
// rows — every input byte crosses L2 -> SM eight times (measured 325 us per launch at batch 24 for ~35 us of traffic,
// DESIGN.md "Known next steps").  With the NO-SWIZZLE K-major canonical layout the address of element
// (row m, 16-byte K chunk j) is  base + (m / 8) * SBO + j * LBO + (m % 8) * 16.  Choosing SBO = 128 B and LBO = 16 B
// gives  base + 16 * (m + j)  — exactly P[m + j] of a plain array P of 16-byte pixels.  The core matrices then overlap
// in memory, which is legal for an operand that is only read.
//
// The probe builds P (135 pixels x 8 channels) and a no-swizzle B tile (64 x 64), issues the four K = 16 instructions
// of one 128 x 64 x 64 tile and compares TMEM with a host evaluation of  D[m][n] = sum_{j,c} P[m+j][c] * B[n][8 j + c].
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/toeplitz_probe tools/toeplitz_probe.cu && tools/bin/toeplitz_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

constexpr int kM = 128, kN = 64, kTaps = 8, kCh = 8, kK = kTaps * kCh;   // K = 64
constexpr int kPix = kM + kTaps - 1;                                      // 135 input pixels

__global__ void __launch_bounds__(128, 1)
toeplitz_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ b, float* __restrict__ d,
                uint32_t lbo_a, uint32_t sbo_a) {
  __shared__ __align__(1024) uint8_t smem[4096 + kN * kK * 2];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = smem_u32(smem);
  // P: pixel i at byte 16 i
  for (int i = threadIdx.x; i < kPix * kCh; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smem)[i] = x[i];
  for (int i = kPix * kCh + threadIdx.x; i < 2048; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smem)[i] = __float2bfloat16(0.f);
  // B, no-swizzle K-major: core matrix (n / 8, k / 8) = 8 rows x 16 B, contiguous 128 B; K-adjacent core matrices
  // 128 B apart (LBO), N-adjacent groups of 8 rows 1024 B apart (SBO)
  __nv_bfloat16* sb = reinterpret_cast<__nv_bfloat16*>(smem + 4096);
  for (int i = threadIdx.x; i < kN * kK; i += blockDim.x) {
    const int n = i / kK, k = i % kK;
    sb[(n / 8) * 512 + (k / 8) * 64 + (n % 8) * 8 + (k % 8)] = b[i];
  }
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    fence_mbar_init();
  }
  if (threadIdx.x < 32) {
    tmem_alloc(smem_u32(&tmem_base_smem), 64);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, static_cast<uint32_t>(kN));
    const uint64_t da = make_smem_desc(base, lbo_a, sbo_a, 0u);             // Toeplitz A: LBO 16 B, SBO 128 B
    const uint64_t db = make_smem_desc(base + 4096, 128, 1024, 0u);
#pragma unroll
    for (int k = 0; k < 4; ++k)   // K = 16 per instruction = two 16-byte chunks: A advances 32 B, B two core matrices
      umma_f16(tm, da + 2u * k, db + 16u * k, idesc, k ? 1u : 0u);
    umma_commit(smem_u32(&bar));
  }
  while (!mbar_try_wait(smem_u32(&bar), 0)) {
  }
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < kN; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(tm + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) d[row * kN + c0 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tm, 64);
}

int main() {
  std::vector<__nv_bfloat16> hx(kPix * kCh), hb(kN * kK);
  std::vector<float> fx(kPix * kCh), fb(kN * kK);
  srand(7);
  for (size_t i = 0; i < hx.size(); ++i) {
    hx[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    fx[i] = __bfloat162float(hx[i]);
  }
  for (size_t i = 0; i < hb.size(); ++i) {
    hb[i] = __float2bfloat16((rand() % 2001 - 1000) / 4000.f);
    fb[i] = __bfloat162float(hb[i]);
  }
  __nv_bfloat16 *dx, *db;
  float* dd;
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dd, kM * kN * 4);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  std::vector<double> ref(kM * kN, 0.0);
  for (int m = 0; m < kM; ++m)
    for (int n = 0; n < kN; ++n) {
      double acc = 0.0;
      for (int j = 0; j < kTaps; ++j)
        for (int c = 0; c < kCh; ++c) acc += (double)fx[(m + j) * kCh + c] * fb[n * kK + j * kCh + c];
      ref[m * kN + n] = acc;
    }
  // (LBO, SBO) candidates in bytes: the intended assignment first, then the swapped one as a control
  const uint32_t cand[][2] = {{16, 128}, {128, 16}};
  for (auto& c : cand) {
    cudaMemset(dd, 0, kM * kN * 4);
    toeplitz_kernel<<<1, 128>>>(dx, db, dd, c[0], c[1]);
    const cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(kM * kN);
    cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
    double num = 0.0, den = 0.0;
    for (int i = 0; i < kM * kN; ++i) {
      num += (out[i] - ref[i]) * (out[i] - ref[i]);
      den += ref[i] * ref[i];
    }
    printf("LBO %3u B, SBO %3u B: relative L2 error %.3e  %s  (%s)\n", c[0], c[1], std::sqrt(num / den),
           std::sqrt(num / den) < 1e-5 ? "PASS: the Toeplitz operand reads correctly" : "fail", cudaGetErrorString(e));
  }
  return 0;
}
