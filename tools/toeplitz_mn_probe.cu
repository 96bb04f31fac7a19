// Probe for round 2 (compiled, NOT yet run on hardware): the MN-major form of the Toeplitz operand that
// tools/toeplitz_probe.cu verified for the forward pass — needed by the WEIGHT GRADIENT of the image layers, whose
// reduction runs over pixels:   dW[o][j*8 + c] = sum_p dy[p][o] * x[p + j][c]   (one filter row, taps j = 0..7).
// The shifted operand B[n = j*8 + c][k = p] = P[p + j][c] over a contiguous array P of 16-byte pixels is, in the
// NO-SWIZZLE MN-major canonical layout  addr(n-chunk j, k) = j * SBO + (k % 8) * 16 + (k / 8) * LBO,  obtained with
// SBO = 16 B and LBO = 128 B:  base + 16 * (j + k).  (For no-swizzle MN-major operands the descriptor's SBO is the
// stride between 16-byte chunks along MN and LBO the stride between groups of 8 along K — the roles are swapped with
// respect to the swizzled layouts; the control run below swaps them back.)
// A = dy is staged in plain canonical order (core matrix = 8 pixels x 16 B, K groups 128 B apart, channel chunks 1 KB).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tools/bin/toeplitz_mn_probe tools/toeplitz_mn_probe.cu && tools/bin/toeplitz_mn_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

constexpr int kM = 128, kN = 64, kK = 64, kTaps = 8, kCh = 8;   // M: dy channels, N: tap x channel, K: pixels
constexpr int kPix = kK + kTaps - 1;                             // 71 input pixels

__global__ void __launch_bounds__(128, 1)
toeplitz_mn_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, float* __restrict__ d,
                   uint32_t lbo_b, uint32_t sbo_b) {
  __shared__ __align__(1024) uint8_t smem[16384 + 2048];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = smem_u32(smem);
  // A (dy, [pixel][128 channels] in global): channel chunk i (8 channels), pixel k -> i * 1024 + (k / 8) * 128 + (k % 8) * 16
  __nv_bfloat16* sa = reinterpret_cast<__nv_bfloat16*>(smem);
  for (int e = threadIdx.x; e < kK * kM; e += blockDim.x) {
    const int k = e / kM, m = e % kM;
    sa[(m / 8) * 512 + (k / 8) * 64 + (k % 8) * 8 + (m % 8)] = dy[e];
  }
  // P: pixel i at byte 16 i (71 pixels, zero tail)
  __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(smem + 16384);
  for (int e = threadIdx.x; e < 1024; e += blockDim.x) sp[e] = e < kPix * kCh ? x[e] : __float2bfloat16(0.f);
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
    const uint32_t idesc = make_idesc(1u, 1u, 1u, 128u, static_cast<uint32_t>(kN));   // both operands MN-major
    const uint64_t da = make_smem_desc(base, 128, 1024, 0u);                 // LBO: K groups, SBO: channel chunks
    const uint64_t db = make_smem_desc(base + 16384, lbo_b, sbo_b, 0u);      // Toeplitz: LBO 128 B, SBO 16 B
#pragma unroll
    for (int k = 0; k < 4; ++k)   // K = 16 pixels per instruction = two groups of 8: 256 B along K for both operands
      umma_f16(tm, da + 16u * k, db + 16u * k, idesc, k ? 1u : 0u);
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
  std::vector<__nv_bfloat16> hdy(kK * kM), hx(kPix * kCh);
  std::vector<float> fdy(kK * kM), fx(kPix * kCh);
  srand(9);
  for (size_t i = 0; i < hdy.size(); ++i) {
    hdy[i] = __float2bfloat16((rand() % 2001 - 1000) / 2000.f);
    fdy[i] = __bfloat162float(hdy[i]);
  }
  for (size_t i = 0; i < hx.size(); ++i) {
    hx[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    fx[i] = __bfloat162float(hx[i]);
  }
  __nv_bfloat16 *ddy, *dx;
  float* dd;
  cudaMalloc(&ddy, hdy.size() * 2);
  cudaMalloc(&dx, hx.size() * 2);
  cudaMalloc(&dd, kM * kN * 4);
  cudaMemcpy(ddy, hdy.data(), hdy.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
  std::vector<double> ref(kM * kN, 0.0);
  for (int m = 0; m < kM; ++m)
    for (int j = 0; j < kTaps; ++j)
      for (int c = 0; c < kCh; ++c) {
        double acc = 0.0;
        for (int p = 0; p < kK; ++p) acc += (double)fdy[p * kM + m] * fx[(p + j) * kCh + c];
        ref[m * kN + j * kCh + c] = acc;
      }
  const uint32_t cand[][2] = {{128, 16}, {16, 128}};   // (LBO, SBO) in bytes: intended first, swapped as a control
  for (auto& c : cand) {
    cudaMemset(dd, 0, kM * kN * 4);
    toeplitz_mn_kernel<<<1, 128>>>(ddy, dx, dd, c[0], c[1]);
    const cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> out(kM * kN);
    cudaMemcpy(out.data(), dd, out.size() * 4, cudaMemcpyDeviceToHost);
    double num = 0.0, den = 0.0;
    for (int i = 0; i < kM * kN; ++i) {
      num += (out[i] - ref[i]) * (out[i] - ref[i]);
      den += ref[i] * ref[i];
    }
    printf("B: LBO %3u B, SBO %3u B: relative L2 error %.3e  %s  (%s)\n", c[0], c[1], std::sqrt(num / den),
           std::sqrt(num / den) < 1e-5 ? "PASS: the MN-major Toeplitz operand reads correctly" : "fail",
           cudaGetErrorString(e));
  }
  return 0;
}
