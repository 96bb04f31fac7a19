// Round-2 preparation (standalone, not part of libcdb200.so): the c7s1-64 image layer (7x7, 3 -> 64 channels, stride 1,
// reflect padding materialised by the caller, fused InstanceNorm statistics) with the TOEPLITZ operand that
// tools/toeplitz_probe.cu verified on B200 — each input row segment is staged ONCE as a contiguous array of 16-byte
// pixels (one cp.async.bulk of 135 x 16 B per filter row) and read by tcgen05 through a no-swizzle K-major
// descriptor with LBO = 16 B, SBO = 128 B, so the 8 taps of a filter row are the 8 K chunks of one 128 x 64 x 64 block.
// The library's row-packed path fetches 128 overlapping 128-byte rows per K block instead (8x read amplification,
// 325 us per launch at batch 24).  The 7 x 8 KB weight blocks stay resident in shared memory.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/toeplitz_conv_bench tools/toeplitz_conv_bench.cu
//   tools/bin/toeplitz_conv_bench [batch]
// -DSTAGED_STORE builds the variant with a coalesced (shared-memory staged) epilogue (written after the measured run).
// Prints the error against a host evaluation on sampled outputs, the error of the statistics, and the time per launch.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "../cycle_depth_estimation_b200/csrc/ptx.cuh"
using namespace cdb;

constexpr int kR = 7, kTaps = 8, kCh = 8, kCout = 64, kBM = 128;
constexpr int kH = 256, kW = 256, kHp = kH + kR - 1, kPitch = 264;        // input rows of 264 pixels (262 real + 2 zero)
constexpr int kSeg = kBM + kTaps - 1;                                     // 135 pixels per row segment
constexpr int kSegBytes = kSeg * 16;                                      // 2160
constexpr int kRowStride = 2176;                                          // per filter row inside a stage (16B multiple)
constexpr int kStageBytes = 15360;                                        // 7 x 2176 = 15232 -> 15 KB
constexpr int kStages = 4;
constexpr int kWBytes = kR * kCout * kTaps * kCh * 2;                     // 7 x 8 KB
constexpr int kSlabBytes = 4 * 32 * 17 * 4;                               // epilogue transposes (4 warps)
constexpr int kStoreStageBytes = 4 * 4096;                                // -DSTAGED_STORE: 32 x 128 B per epilogue warp
constexpr int kSmem = kWBytes + kStages * kStageBytes + kSlabBytes + kStoreStageBytes + 1024;

__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// x: [n][kHp][kPitch][8] bf16 (channels 3..7 zero); wblk: [7][8 KB] bf16 already in the no-swizzle core-matrix order;
// y: [n][kH][kW][64] bf16; stats: [n][64][2] fp32 (zeroed by the caller)
__global__ void __launch_bounds__(256, 1)
toeplitz_conv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wblk,
                     __nv_bfloat16* __restrict__ y, float* __restrict__ stats, int n_img) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kStages], bar_empty[kStages], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_base = base, st_base = base + kWBytes;
  float* slab_all = reinterpret_cast<float*>(gen + kWBytes + kStages * kStageBytes);
  uint8_t* stage_all = gen + kWBytes + kStages * kStageBytes + kSlabBytes;   // 16-byte aligned (all sizes are)
  (void)stage_all;
  const int total_tiles = n_img * kH * (kW / kBM);

  for (int i = threadIdx.x; i < kWBytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(wblk)[i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), 128);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_smem), 128);
    tmem_relinquish();
  }
  fence_proxy_async_smem();   // the weight blocks were written by ordinary stores and are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0 && lane == 0) {
    // ---------------------------------------------------------------- producer: 7 row segments per tile
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int half = tile % (kW / kBM);
      const int p = (tile / (kW / kBM)) % kH;
      const int img = tile / (kW / kBM) / kH;
      while (!mbar_try_wait(smem_u32(&bar_empty[stage]), phase ^ 1u)) {
      }
      const uint32_t full = smem_u32(&bar_full[stage]);
      mbar_arrive_expect_tx(full, kR * kSegBytes);
      for (int r = 0; r < kR; ++r) {
        const __nv_bfloat16* src = x + ((static_cast<int64_t>(img) * kHp + p + r) * kPitch + half * kBM) * kCh;
        bulk_load(st_base + stage * kStageBytes + r * kRowStride, src, kSegBytes, full);
      }
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ---------------------------------------------------------------- MMA issuer: 7 filter rows x 4 instructions
    const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, static_cast<uint32_t>(kCout));
    int stage = 0, local = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t tphase = (local >> 1) & 1u;
      while (!mbar_try_wait(smem_u32(&bar_tempty[buf]), tphase ^ 1u)) {
      }
      while (!mbar_try_wait(smem_u32(&bar_full[stage]), phase)) {
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * 64u;
#pragma unroll
      for (int r = 0; r < kR; ++r) {
        const uint64_t da = make_smem_desc(st_base + stage * kStageBytes + r * kRowStride, 16, 128, 0u);   // Toeplitz
        const uint64_t db = make_smem_desc(w_base + r * (kCout * kTaps * kCh * 2), 128, 1024, 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2u * k, db + 16u * k, idesc, (r | k) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bar_empty[stage]));
      umma_commit(smem_u32(&bar_tfull[buf]));
      if (++stage == kStages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: TMEM -> bf16 -> global, IN statistics
    const int ew = warp - 4;
    float* slab = slab_all + ew * 32 * 17;
    int local = 0;
    // per-channel sums stay in registers while the CTA's tiles belong to the same image (its tiles are 148 apart: 74
    // output rows): two or three flushes of 2 x 64 atomics per warp instead of 128 atomics per tile
    float a1[4] = {0.f, 0.f, 0.f, 0.f}, a2[4] = {0.f, 0.f, 0.f, 0.f};
    int cur_img = -1;
    auto flush = [&]() {
      if (cur_img >= 0 && lane < 16) {
        float* st = stats + static_cast<int64_t>(cur_img) * kCout * 2;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          atomicAdd(st + (qd * 16 + lane) * 2, a1[qd]);
          atomicAdd(st + (qd * 16 + lane) * 2 + 1, a2[qd]);
          a1[qd] = a2[qd] = 0.f;
        }
      }
    };
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int buf = local & 1;
      const uint32_t tphase = (local >> 1) & 1u;
      const int half = tile % (kW / kBM);
      const int p = (tile / (kW / kBM)) % kH;
      const int img = tile / (kW / kBM) / kH;
      while (!mbar_try_wait(smem_u32(&bar_tfull[buf]), tphase)) {
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(buf) * 64u;
      uint32_t v[64];
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&bar_tempty[buf]));   // the accumulator is in registers: the next tile may overwrite it
#ifdef STAGED_STORE
      // (not yet run on hardware) the warp's 32 pixels x 128 B go through a swizzled staging tile so that every store
      // instruction writes 512 contiguous bytes (4 pixels) instead of 32 separate 16-byte pieces
      {
        uint4* stg = reinterpret_cast<uint4*>(stage_all + ew * 4096);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 o;
          o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
          o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
          o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
          o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
          stg[lane * 8 + (j ^ (lane & 7))] = o;          // row `lane`, 16-byte chunk j at position j ^ (row & 7)
        }
        __syncwarp();
        uint4* dst4 = reinterpret_cast<uint4*>(
            y + ((static_cast<int64_t>(img) * kH + p) * kW + half * kBM + ew * 32) * kCout);
        const int c = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = (lane >> 3) + 4 * i;
          dst4[row * 8 + c] = stg[row * 8 + (c ^ (row & 7))];
        }
        __syncwarp();
      }
#else
      const int q = half * kBM + ew * 32 + lane;
      __nv_bfloat16* dst = y + ((static_cast<int64_t>(img) * kH + p) * kW + q) * kCout;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
        o.y = pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
        o.z = pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
        o.w = pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
        reinterpret_cast<uint4*>(dst)[j] = o;
      }
#endif
      if (img != cur_img) {
        flush();
        cur_img = img;
      }
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
#pragma unroll
        for (int j = 0; j < 16; ++j) slab[lane * 17 + j] = __uint_as_float(v[qd * 16 + j]);
        __syncwarp();
        if (lane < 16) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const float t = slab[i * 17 + lane];
            s1 += t;
            s2 = fmaf(t, t, s2);
          }
          a1[qd] += s1;
          a2[qd] += s2;
        }
        __syncwarp();
      }
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 128);
}

int main(int argc, char** argv) {
  const int n_img = argc > 1 ? atoi(argv[1]) : 8;
  const size_t x_elems = (size_t)n_img * kHp * kPitch * kCh, y_elems = (size_t)n_img * kH * kW * kCout;
  std::vector<__nv_bfloat16> hx(x_elems, __float2bfloat16(0.f));
  std::vector<float> fx(x_elems, 0.f);
  srand(11);
  for (int n = 0; n < n_img; ++n)
    for (int h = 0; h < kHp; ++h)
      for (int w = 0; w < kW + kR - 1; ++w)
        for (int c = 0; c < 3; ++c) {
          const size_t i = (((size_t)n * kHp + h) * kPitch + w) * kCh + c;
          hx[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
          fx[i] = __bfloat162float(hx[i]);
        }
  // weights W[o][c][r][s] -> per filter row r a 64 x 64 K-major block, k = s * 8 + c, in no-swizzle core-matrix order
  std::vector<float> fw((size_t)kCout * 3 * kR * kR);
  for (auto& v : fw) v = (rand() % 2001 - 1000) / 8000.f;
  std::vector<__nv_bfloat16> hw((size_t)kR * kCout * kTaps * kCh, __float2bfloat16(0.f));
  std::vector<float> fwq(fw.size());
  for (int o = 0; o < kCout; ++o)
    for (int c = 0; c < 3; ++c)
      for (int r = 0; r < kR; ++r)
        for (int s = 0; s < kR; ++s) {
          const size_t i = (((size_t)o * 3 + c) * kR + r) * kR + s;
          const __nv_bfloat16 q = __float2bfloat16(fw[i]);
          fwq[i] = __bfloat162float(q);
          const int k = s * kCh + c;
          hw[(size_t)r * kCout * kTaps * kCh + (o / 8) * 512 + (k / 8) * 64 + (o % 8) * 8 + (k % 8)] = q;
        }
  __nv_bfloat16 *dx, *dw, *dy;
  float* dstats;
  cudaMalloc(&dx, x_elems * 2);
  cudaMalloc(&dw, hw.size() * 2);
  cudaMalloc(&dy, y_elems * 2);
  cudaMalloc(&dstats, (size_t)n_img * kCout * 2 * 4);
  cudaMemcpy(dx, hx.data(), x_elems * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dy, 0xff, y_elems * 2);
  cudaMemset(dstats, 0, (size_t)n_img * kCout * 2 * 4);
  cudaFuncSetAttribute(toeplitz_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  toeplitz_conv_kernel<<<sms, 256, kSmem>>>(dx, dw, dy, dstats, n_img);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  std::vector<__nv_bfloat16> hy(y_elems);
  std::vector<float> hs((size_t)n_img * kCout * 2);
  cudaMemcpy(hy.data(), dy, y_elems * 2, cudaMemcpyDeviceToHost);
  cudaMemcpy(hs.data(), dstats, hs.size() * 4, cudaMemcpyDeviceToHost);
  auto ref_at = [&](int n, int p, int q, int o) {
    double acc = 0.0;
    for (int r = 0; r < kR; ++r)
      for (int s = 0; s < kR; ++s)
        for (int c = 0; c < 3; ++c)
          acc += (double)fx[(((size_t)n * kHp + p + r) * kPitch + q + s) * kCh + c] *
                 fwq[(((size_t)o * 3 + c) * kR + r) * kR + s];
    return acc;
  };
  double num = 0.0, den = 0.0;
  for (int t = 0; t < 20000; ++t) {
    const int n = rand() % n_img, p = rand() % kH, q = (t % 4 == 0) ? (rand() % 2 ? 0 : kW - 1) : rand() % kW, o = rand() % kCout;
    const double r = ref_at(n, p, q, o);
    const double g = __bfloat162float(hy[(((size_t)n * kH + p) * kW + q) * kCout + o]);
    num += (g - r) * (g - r);
    den += r * r;
  }
  printf("sampled outputs: relative L2 error %.3e (bf16 storage: expect ~2e-3)\n", std::sqrt(num / den));
  // statistics of image 0, channel 5 against the stored outputs
  double s1 = 0.0, s2 = 0.0;
  for (size_t i = 0; i < (size_t)kH * kW; ++i) {
    const double r = ref_at(0, (int)(i / kW), (int)(i % kW), 5);
    s1 += r;
    s2 += r * r;
  }
  printf("statistics (image 0, channel 5): sum %.4f vs %.4f, sum of squares %.4f vs %.4f\n", hs[5 * 2], s1, hs[5 * 2 + 1], s2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < iters; ++i) toeplitz_conv_kernel<<<sms, 256, kSmem>>>(dx, dw, dy, dstats, n_img);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / iters;
  printf("batch %d: %.1f us per launch, output %.1f MB -> %.0f GB/s written (%s)\n", n_img, us, y_elems * 2 / 1e6,
         y_elems * 2 / us * 1e-3, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
