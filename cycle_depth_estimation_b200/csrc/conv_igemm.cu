// K1/K2 — implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05), sm_100a only.
//
// GEMM view:  D[m, n] = sum_k A[m, k] * B[n, k]
//   m : 128 output pixels of one tile  (tile_n images x tile_h rows x tile_w columns)
//   n : output channels (one tile of bn <= 256)
//   k : (filter tap, 64-channel chunk) pairs — one pipeline stage per pair
// A is never materialised: for every tap the TMA engine fetches the shifted NHWC window
// {64 ch, tile_w, tile_h, tile_n} of the input straight into 128B-swizzled shared memory, where it
// is exactly the K-major operand tcgen05.mma expects; out-of-range coordinates are zero-filled by the
// TMA unit, which implements zero padding. Stride-2 and transposed convolutions are expressed through
// parity views of the input / output, so the main loop is the same for all of them.
//
// Warp roles (256 threads, persistent over tiles):
//   warp 0   TMA producer (one lane)        warp 1   tcgen05.mma issuer (one lane)
//   warp 2   TMEM allocator                 warps 4-7  epilogue: tcgen05.ld -> bias/act -> global
// Accumulators are double buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
// main loop of tile i+1.
#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

constexpr int kMaxTaps = 64;
constexpr int kMaxStages = 8;
constexpr int kABytes = 128 * 128;  // 128 pixels x 64 bf16

struct IgemmTap {
  int16_t map, dh, dw, pad_;
  int32_t wk;
};

struct IgemmParams {
  int32_t n_taps, k_chunks;
  int32_t tile_w, tile_h, tile_n;
  int32_t tiles_w, tiles_h, tiles_n;
  int32_t n_tiles_n;
  int32_t dom_n, dom_h, dom_w;
  int32_t cout, cstore, bn;
  int32_t out_dtype, act;
  float slope;
  int32_t stages;
  int32_t kgroup;    // (tap, chunk) K blocks per pipeline stage: one barrier round trip of the producer / MMA threads per group
  int32_t stats_on;
  int32_t stats_batch;  // 1: one statistics group for the whole batch (BatchNorm)
  int32_t fast_out;
  int32_t tf32;      // operands are fp32 in memory, multiplied as TF32 (kind::tf32); K block = 32 channels
  int32_t kelems;    // channels per 128-byte K block: 64 (bf16) or 32 (tf32)
  int32_t vec_out;   // fp32 NHWC output with 16-byte aligned pixels: float4 stores
  int32_t round_out; // round the fp32 output to TF32 (nearest) so that the next TF32 convolution reads exact operands
  const float* bias;
  void* out;
  int64_t o_sn, o_sh, o_sw, o_sc;
  float* stats;
  int* abort_flag;
  long long* dbg;  // optional per-role wait cycles of CTA 0 (CDB_IGEMM_DEBUG=1)
  IgemmTap taps[kMaxTaps];
};

struct IgemmMaps {
  CUtensorMap a[4];
  CUtensorMap b;
  CUtensorMap out;  // fast output path: bf16 NHWC 4-D map, box {64, bw, bh, bn} = 32 rows, 128B swizzle
};

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case CDB_ACT_RELU: return v > 0.f ? v : 0.f;
    case CDB_ACT_LEAKY: return v > 0.f ? v : v * slope;
    case CDB_ACT_TANH: return tanhf(v);
    case CDB_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// kPair: a CTA pair (cluster of 2, cta_group::2) works on two neighbouring pixel tiles with the same channel tile:
// each CTA stages ITS 128-pixel window and HALF of the weight tile per (tap, chunk), the even CTA issues M = 256
// instructions for both, each CTA drains its own 128 x bn accumulator (this kernel is bound by L2 -> SM operand
// traffic: one stage of a 256-channel layer is 16 + 32 KB for 4.2 MFLOP; the pair needs 16 + 16 KB per SM).
template <bool kPair>
__global__ void __launch_bounds__(256, 1)
igemm_kernel(const __grid_constant__ IgemmMaps maps, const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t b_bytes = static_cast<uint32_t>(kPair ? p.bn / 2 : p.bn) * 128u;
  const uint32_t kblock_bytes = kABytes + b_bytes;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.kgroup) * kblock_bytes;
  const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
  // work units: (pixel tile, channel tile); pair: (two neighbouring pixel tiles, channel tile)
  const int total_tiles = (kPair ? (m_tiles + 1) / 2 : m_tiles) * p.n_tiles_n;
  // every CTA (pair) walks a CONTIGUOUS range of work units: consecutive pixel tiles belong to the same image, so the
  // InstanceNorm sums are kept in registers across tiles and reach memory once per image (the per-tile atomics of 592
  // warps onto 64 cache lines took 40 % of the u64 layer)
  const int n_workers = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int worker = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int per_worker = (total_tiles + n_workers - 1) / n_workers;
  const int tile_first = worker * per_worker;
  const int tile_last = min(total_tiles, tile_first + per_worker);
  const int k_blocks = p.n_taps * p.k_chunks;

  if (threadIdx.x == 0) {
    abort_smem = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), kPair ? 256 : 128);  // pair: the epilogue threads of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.b);
    prefetch_tmap(&maps.a[0]);
  }
  if (warp == 2) {
    if (kPair) {
      tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(smem_u32(&tmem_base_smem), 512);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // the peer's barriers exist before any TMA / commit / arrive reaches them
  tc_fence_after();
  pdl_trigger();
  pdl_wait();   // everything above overlapped the tail of the previous kernel in the stream
  const uint32_t tmem_base = tmem_base_smem;
  volatile int* abort_flag = &abort_smem;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      long long prod_wait = 0;
      const long long tp0 = dbg ? clock64() : 0;
      for (int tile = tile_first; tile < tile_last && ok; ++tile) {
        const int n_tile = tile % p.n_tiles_n;
        // an odd tile count leaves the last pair's second CTA with a tile beyond the batch: TMA zero-fills it
        int m_tile = kPair ? 2 * (tile / p.n_tiles_n) + static_cast<int>(rank) : tile / p.n_tiles_n;
        const int tw = m_tile % p.tiles_w;
        m_tile /= p.tiles_w;
        const int th = m_tile % p.tiles_h;
        const int tn = m_tile / p.tiles_h;
        const int q0 = tw * p.tile_w, p0 = th * p.tile_h, img0 = tn * p.tile_n;
        const int n0 = n_tile * p.bn + (kPair ? static_cast<int>(rank) * (p.bn / 2) : 0);
        // The single producer thread and the single MMA thread each spend ~500-700 cycles per pipeline stage on the
        // barrier round trip (measured, tools/igemm_dbg.py), more than the 272 cycles of tensor work a 64 / 128-channel
        // K block carries: a stage therefore holds kgroup K blocks.
        for (int g0 = 0; g0 < k_blocks && ok; g0 += p.kgroup) {
          const int cnt = min(p.kgroup, k_blocks - g0);
          const long long tw0 = dbg ? clock64() : 0;
          if (!mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u, abort_flag)) {
            ok = false;
            break;
          }
          if (dbg) prod_wait += clock64() - tw0;
          const uint32_t full = smem_u32(&bar_full[stage]);
          const uint32_t full0 = kPair ? mapa_shared(full, 0) : full;
          // pair: both CTAs' loads complete on the issuing (even) CTA's barrier, armed by it with the bytes of both
          if (rank == 0) mbar_arrive_expect_tx(full, (kPair ? 2u : 1u) * static_cast<uint32_t>(cnt) * kblock_bytes);
          for (int j = 0; j < cnt; ++j) {
            const int kb = g0 + j;
            const int t = kb / p.k_chunks, c = kb - t * p.k_chunks;
            const IgemmTap tap = p.taps[t];
            const CUtensorMap* amap = &maps.a[tap.map];
            const uint32_t sa = smem_base + stage * stage_bytes + static_cast<uint32_t>(j) * kblock_bytes;
            if (kPair) {
              tma_load_4d_pair(amap, full0, sa, c * p.kelems, q0 + tap.dw, p0 + tap.dh, img0);
              tma_load_2d_pair(&maps.b, full0, sa + kABytes, tap.wk + c * p.kelems, n0);
            } else {
              tma_load_4d(amap, full, sa, c * p.kelems, q0 + tap.dw, p0 + tap.dh, img0);
              tma_load_2d(&maps.b, full, sa + kABytes, tap.wk + c * p.kelems, n0);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
      if (dbg) {
        p.dbg[1] = prod_wait;
        p.dbg[2] = clock64() - tp0;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // pair: the whole warp of the even CTA walks the loop and an elected lane issues (a cta_group::2 instruction
    // issued from a lone thread of a diverged warp takes 185-283 cycles, tools/mma_rate2.cu)
    if (kPair ? (rank == 0) : (lane == 0)) {
      const bool tf32 = p.tf32 != 0;
      const uint32_t idesc = make_idesc(tf32 ? 2u : 1u, 0u, 0u, kPair ? 256u : 128u, static_cast<uint32_t>(p.bn));
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      bool ok = true;
      long long w_tempty = 0, w_full = 0;
      const bool dbgl = dbg && lane == 0;
      const long long tm0 = dbgl ? clock64() : 0;
      for (int tile = tile_first; tile < tile_last && ok; ++tile, ++local) {
        const int buf = local & 1;
        const uint32_t tphase = (local >> 1) & 1u;
        long long tw0 = dbgl ? clock64() : 0;
        ok = mbar_wait(smem_u32(&bar_tempty[buf]), tphase ^ 1u, abort_flag);
        if (kPair) ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        if (dbgl) w_tempty += clock64() - tw0;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * 256u;
        for (int g0 = 0; g0 < k_blocks; g0 += p.kgroup) {
          const int cnt = min(p.kgroup, k_blocks - g0);
          tw0 = dbgl ? clock64() : 0;
          ok = mbar_wait(smem_u32(&bar_full[stage]), phase, abort_flag);
          if (kPair) ok = __all_sync(0xffffffffu, ok);
          if (!ok) break;
          if (dbgl) w_full += clock64() - tw0;
          tc_fence_after();
          // one instruction consumes 32 bytes of K per row in both precisions: 16 bf16 or 8 tf32
          if (!kPair || elect_one()) {
            for (int j = 0; j < cnt; ++j) {
              const uint32_t sa = smem_base + stage * stage_bytes + static_cast<uint32_t>(j) * kblock_bytes;
              const uint64_t da = make_smem_desc(sa, 16, 1024, kLayoutSW128);
              const uint64_t db = make_smem_desc(sa + kABytes, 16, 1024, kLayoutSW128);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (g0 | j | k) != 0 ? 1u : 0u;
                if (kPair) {
                  if (tf32) umma2_tf32(d_tmem, da + 2u * k, db + 2u * k, idesc, acc);
                  else umma2_f16(d_tmem, da + 2u * k, db + 2u * k, idesc, acc);
                } else {
                  if (tf32) umma_tf32(d_tmem, da + 2u * k, db + 2u * k, idesc, acc);
                  else umma_f16(d_tmem, da + 2u * k, db + 2u * k, idesc, acc);
                }
              }
            }
            if (kPair) umma2_commit(smem_u32(&bar_empty[stage]));
            else umma_commit(smem_u32(&bar_empty[stage]));
          }
          if (kPair) __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (ok) {
          if (kPair) {
            if (elect_one()) umma2_commit(smem_u32(&bar_tfull[buf]));
            __syncwarp();
          } else {
            umma_commit(smem_u32(&bar_tfull[buf]));
          }
        }
      }
      if (dbgl) {
        p.dbg[3] = w_tempty;
        p.dbg[4] = w_full;
        p.dbg[5] = clock64() - tm0;
        p.dbg[9] = local;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const int r_w = row % p.tile_w;
    const int r_h = (row / p.tile_w) % p.tile_h;
    const int r_n = row / (p.tile_w * p.tile_h);
    // the 32 rows of this warp as a sub-box of the tile (all tile dimensions are powers of two)
    const int m0 = ew * 32;
    const int w_off = m0 % p.tile_w, h_off = (m0 / p.tile_w) % p.tile_h, n_off = m0 / (p.tile_w * p.tile_h);
    const bool one_image = p.tile_w * p.tile_h >= 32;  // all 32 rows of a warp belong to one image
    const uint32_t stage_addr = smem_base + p.stages * stage_bytes + ew * 8192;   // two 4 KB staging buffers per warp
    float* slab0 = reinterpret_cast<float*>(smem_raw + (stage_addr - smem_u32(smem_raw)));
    float* slab = slab0;
    const bool has_bias = p.bias != nullptr;
    const int act = p.act;
    const float slope = p.slope;
    const int cout = p.cout, cstore = p.cstore;
    const bool stats_on = p.stats_on != 0;
    const uint32_t tempty_remote = kPair ? mapa_shared(smem_u32(&bar_tempty[0]), 0) : 0u;
    auto mbar_arrive_tempty = [&](int b) {
      if (kPair) mbar_arrive_cluster(tempty_remote + static_cast<uint32_t>(b) * 8u);
      else mbar_arrive(smem_u32(&bar_tempty[b]));
    };
    // InstanceNorm / BatchNorm sums of this warp's rows, per 64-column slab and 16-column quarter (lanes 0-15 hold one
    // channel each), carried across the tiles of one image and added to memory when the image (or channel tile) changes
    float acc1[4][4], acc2[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc1[i][j] = acc2[i][j] = 0.f;
    int acc_img = -1, acc_ntile = -1;
    auto flush_stats = [&]() {
      if (acc_img >= 0 && lane < 16 && acc_img < p.dom_n) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = acc_ntile * p.bn + i * 64 + j * 16 + lane;
            if (i * 64 + j * 16 < p.bn && ch < cout) {
              atomicAdd(p.stats + (static_cast<int64_t>(acc_img) * cout + ch) * 2, acc1[i][j]);
              atomicAdd(p.stats + (static_cast<int64_t>(acc_img) * cout + ch) * 2 + 1, acc2[i][j]);
            }
          }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc1[i][j] = acc2[i][j] = 0.f;
    };
    uint32_t slab_count = 0;
    int local = 0;
    long long w_tfull = 0, w_store = 0;
    const bool dbge = dbg && threadIdx.x == 128;
    const long long te0 = dbge ? clock64() : 0;
    for (int tile = tile_first; tile < tile_last; ++tile, ++local) {
      bool released = false;
      const long long twf = dbge ? clock64() : 0;
      const int buf = local & 1;
      const uint32_t tphase = (local >> 1) & 1u;
      if (!mbar_wait(smem_u32(&bar_tfull[buf]), tphase, abort_flag)) break;
      if (dbge) w_tfull += clock64() - twf;
      tc_fence_after();
      const int n_tile = tile % p.n_tiles_n;
      int m_tile = kPair ? 2 * (tile / p.n_tiles_n) + static_cast<int>(rank) : tile / p.n_tiles_n;
      const int tw = m_tile % p.tiles_w;
      m_tile /= p.tiles_w;
      const int th = m_tile % p.tiles_h;
      const int tn = m_tile / p.tiles_h;
      const int q = tw * p.tile_w + r_w, pp = th * p.tile_h + r_h, img = tn * p.tile_n + r_n;
      const bool valid = (q < p.dom_w) && (pp < p.dom_h) && (img < p.dom_n);
      const int n0 = n_tile * p.bn;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                             static_cast<uint32_t>(buf) * 256u;
      if (p.fast_out) {
        // ---- bf16 NHWC: 64-column slabs -> swizzled staging -> TMA store (clips the tile edges)
        const int simg = p.stats_batch ? 0 : tn * p.tile_n + n_off;
        const bool reg_stats = stats_on && (one_image || p.stats_batch);
        if (reg_stats && (simg != acc_img || n_tile != acc_ntile)) {
          flush_stats();
          acc_img = simg;
          acc_ntile = n_tile;
        }
#pragma unroll
        for (int si = 0; si < 4; ++si) {
          const int c0 = si * 64;
          if (c0 >= p.bn || n0 + c0 >= cstore) break;
          uint32_t v[64];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, v + 32);
          tmem_ld_wait();
          if (c0 + 64 >= p.bn || n0 + c0 + 64 >= cstore) {
            // last slab: the accumulator is in registers, the MMA warp may refill this TMEM buffer
            tc_fence_before();
            mbar_arrive_tempty(buf);
            released = true;
          }
          if (has_bias) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const int ch = n0 + c0 + j;
              if (ch < cout) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + ch));
            }
          }
          // the switch stays OUTSIDE the unrolled loops: inside, every copy carries the predicated tanh / sigmoid paths
          // (~140 instructions per element, measured on the Toeplitz kernel: 14 k cycles per 128 x 64 tile)
          if (act == CDB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
          } else if (act == CDB_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const float t = __uint_as_float(v[j]);
              v[j] = __float_as_uint(t > 0.f ? t : t * slope);
            }
          } else if (act == CDB_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(tanhf(__uint_as_float(v[j])));
          } else if (act == CDB_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(1.f / (1.f + __expf(-__uint_as_float(v[j]))));
          }
          if (n0 + c0 + 64 > cout || c0 + 64 > p.bn) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (n0 + c0 + j >= cout || c0 + j >= p.bn) v[j] = 0u;
          }
          // two staging buffers per warp: the TMA store of the previous slab may still be reading the other one
          const uint32_t sbuf = stage_addr + (slab_count & 1u) * 4096u;
          float* slab = slab0 + (slab_count & 1u) * 1024u;
          ++slab_count;
          const long long tws = dbge ? clock64() : 0;
          if (lane == 0) bulk_wait_read1();
          __syncwarp();
          if (dbge) w_store += clock64() - tws;
          if (stats_on) {
            if (reg_stats) {
#pragma unroll
              for (int qd = 0; qd < 4; ++qd) {
                if (c0 + qd * 16 >= p.bn) break;
#pragma unroll
                for (int j = 0; j < 16; ++j) slab[lane * 17 + j] = valid ? __uint_as_float(v[qd * 16 + j]) : 0.f;
                __syncwarp();
                if (lane < 16) {
                  float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
                  for (int i = 0; i < 32; ++i) {
                    const float t = slab[i * 17 + lane];
                    s1 += t;
                    s2 = fmaf(t, t, s2);
                  }
                  acc1[si][qd] += s1;
                  acc2[si][qd] += s2;
                }
                __syncwarp();
              }
            } else if (valid) {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const int ch = n0 + c0 + j;
                if (ch < cout && c0 + j < p.bn) {
                  const float t = __uint_as_float(v[j]);
                  atomicAdd(p.stats + (static_cast<int64_t>(img) * cout + ch) * 2, t);
                  atomicAdd(p.stats + (static_cast<int64_t>(img) * cout + ch) * 2 + 1, t * t);
                }
              }
            }
          }
          const uint32_t rbase = sbuf + lane * 128;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t a = rbase + ((static_cast<uint32_t>(j) ^ (lane & 7u)) << 4);
            st_shared_v4(a, pack_bf16x2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1])),
                         pack_bf16x2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3])),
                         pack_bf16x2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5])),
                         pack_bf16x2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7])));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&maps.out, sbuf, n0 + c0, tw * p.tile_w + w_off, th * p.tile_h + h_off,
                         tn * p.tile_n + n_off);
            bulk_commit();
          }
        }
      } else {
        // ---- generic path: strided / fp32 / NCHW outputs, 16 columns at a time
        const int64_t obase = img * p.o_sn + pp * p.o_sh + q * p.o_sw;
        for (int c0 = 0; c0 < p.bn; c0 += 16) {
          if (n0 + c0 >= cstore) break;  // uniform across the CTA
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          float f[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ch = n0 + c0 + j;
            float x = __uint_as_float(v[j]);
            if (has_bias && ch < cout) x += __ldg(p.bias + ch);
            f[j] = ch < cout ? x : 0.f;
          }
          // activation applied by uniform branches OUTSIDE the unrolled loop (see the fast path)
          if (act == CDB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
          } else if (act == CDB_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = f[j] > 0.f ? f[j] : f[j] * slope;
          } else if (act == CDB_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = n0 + c0 + j < cout ? tanhf(f[j]) : 0.f;
          } else if (act == CDB_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = n0 + c0 + j < cout ? 1.f / (1.f + __expf(-f[j])) : 0.f;
          }
          if (stats_on && valid) {
            const int simg = p.stats_batch ? 0 : img;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int ch = n0 + c0 + j;
              if (ch < cout) {
                atomicAdd(p.stats + (static_cast<int64_t>(simg) * cout + ch) * 2, f[j]);
                atomicAdd(p.stats + (static_cast<int64_t>(simg) * cout + ch) * 2 + 1, f[j] * f[j]);
              }
            }
          }
          if (valid) {
            if (p.out_dtype == CDB_BF16) {
              __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + obase;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int ch = n0 + c0 + j;
                if (ch < cstore) o[ch * p.o_sc] = __float2bfloat16(f[j]);
              }
            } else {
              if (p.round_out) {
#pragma unroll
                for (int j = 0; j < 16; ++j) f[j] = round_tf32(f[j]);
              }
              float* o = static_cast<float*>(p.out) + obase;
              if (p.vec_out) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const int ch = n0 + c0 + j;
                  if (ch < cstore) *reinterpret_cast<float4*>(o + ch) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int ch = n0 + c0 + j;
                  if (ch < cstore) o[ch * p.o_sc] = f[j];
                }
              }
            }
          }
        }
      }
      if (!released) {
        tc_fence_before();
        mbar_arrive_tempty(buf);
      }
    }
    flush_stats();
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    if (dbge) {
      p.dbg[6] = w_tfull;
      p.dbg[7] = w_store;
      p.dbg[8] = clock64() - te0;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // neither CTA leaves while the other may still read its shared memory / signal it
  if (threadIdx.x == 0 && abort_smem && p.abort_flag) atomicExch(p.abort_flag, 1);
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// -------------------------------------------------------------------------------------------------
// Host side: geometry -> tap tables, TMA maps, launches
// -------------------------------------------------------------------------------------------------
static void choose_tile(int dom_n, int dom_h, int dom_w, int* tw, int* th, int* tn) {
  int w = 1;
  while (w < dom_w && w < 128) w <<= 1;
  int h = 1;
  while (h < dom_h && h * w < 128) h <<= 1;
  int n = 128 / (w * h);
  *tw = w;
  *th = h;
  *tn = n;
  (void)dom_n;
}

struct SrcView {
  void* ptr;
  int64_t dims[4];     // c, w, h, n
  int64_t strides[3];  // w, h, n strides in elements
};

static int launch_igemm(const SrcView* views, int n_views, const void* wpacked, int w_rows_pad,
                        int w_ktotal, IgemmParams& prm, cudaStream_t stream) {
  IgemmMaps maps;
  memset(&maps, 0, sizeof(maps));
  const int m_tiles = prm.tiles_n * prm.tiles_h * prm.tiles_w;
  // CTA pairs (cta_group::2): OFF by default.  Measured on B200 (tools/time_igemm.py, batch 24): with 256-channel weight
  // tiles the pair is 1-6 % faster in isolation (dgrad of u128 60.0 -> 56.4 us) and neutral in the step; with bn <= 128 it
  // is SLOWER (d128 114 -> 138 us: an M = 256 instruction costs ~128+ cycles whatever N is, and these layers are bound by
  // the A-operand traffic, which the pair does not reduce).  CDB_IGEMM_PAIR=1: bn == 256 layers; =2: every eligible layer.
  static const int pair_env = getenv("CDB_IGEMM_PAIR") ? atoi(getenv("CDB_IGEMM_PAIR")) : 0;
  const bool pair = pair_env != 0 && (pair_env > 1 ? (prm.bn % 16 == 0 && prm.bn >= 64)
                                                   : (prm.bn == 256 && (int64_t)m_tiles * prm.n_tiles_n > sm_count()));
  const uint64_t esz = prm.tf32 ? 4 : 2;
  const CUtensorMapDataType dt = prm.tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  prm.kelems = prm.tf32 ? 32 : 64;
  const uint32_t abox[4] = {(uint32_t)prm.kelems, (uint32_t)prm.tile_w, (uint32_t)prm.tile_h, (uint32_t)prm.tile_n};
  for (int i = 0; i < 4; ++i) {
    const SrcView& v = views[i < n_views ? i : 0];
    uint64_t dims[4] = {(uint64_t)v.dims[0], (uint64_t)v.dims[1], (uint64_t)v.dims[2], (uint64_t)v.dims[3]};
    uint64_t str[3] = {(uint64_t)v.strides[0] * esz, (uint64_t)v.strides[1] * esz, (uint64_t)v.strides[2] * esz};
    for (int d = 0; d < 4; ++d)
      if (dims[d] == 0) dims[d] = 1;
    int rc = make_tmap(&maps.a[i], dt, 4, v.ptr, dims, str, abox);
    if (rc != CDB_OK) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)w_ktotal, (uint64_t)w_rows_pad};
    uint64_t str[1] = {(uint64_t)w_ktotal * esz};
    uint32_t box[2] = {(uint32_t)prm.kelems, (uint32_t)(pair ? prm.bn / 2 : prm.bn)};
    int rc = make_tmap(&maps.b, dt, 2, const_cast<void*>(wpacked), dims, str, box);
    if (rc != CDB_OK) return rc;
  }
  prm.fast_out = (prm.out_dtype == CDB_BF16 && prm.o_sc == 1) ? 1 : 0;
  prm.vec_out = (prm.out_dtype == CDB_F32 && prm.o_sc == 1 && prm.cstore % 4 == 0 && prm.o_sn % 4 == 0 &&
                 prm.o_sh % 4 == 0 && prm.o_sw % 4 == 0 && (reinterpret_cast<uintptr_t>(prm.out) & 15) == 0)
                    ? 1
                    : 0;
  if (prm.fast_out) {
    const int bw = prm.tile_w < 32 ? prm.tile_w : 32;
    const int bh = prm.tile_h < 32 / bw ? prm.tile_h : 32 / bw;
    const int bnn = 32 / (bw * bh);
    uint64_t dims[4] = {(uint64_t)prm.cstore, (uint64_t)prm.dom_w, (uint64_t)prm.dom_h, (uint64_t)prm.dom_n};
    uint64_t str[3] = {(uint64_t)prm.o_sw * 2, (uint64_t)prm.o_sh * 2, (uint64_t)prm.o_sn * 2};
    uint32_t box[4] = {64u, (uint32_t)bw, (uint32_t)bh, (uint32_t)bnn};
    int rc = make_tmap(&maps.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, prm.out, dims, str, box);
    if (rc != CDB_OK) return rc;
  }
  const int kblock_bytes = kABytes + (pair ? prm.bn / 2 : prm.bn) * 128;
  // of the 227 KB per CTA: 32 KB epilogue staging + alignment stay free
  const int budget = (getenv("CDB_IGEMM_SMEM_KB") ? atoi(getenv("CDB_IGEMM_SMEM_KB")) : 194) * 1024;
  int kgroup = budget / (3 * kblock_bytes);
  if (getenv("CDB_IGEMM_KGROUP")) kgroup = atoi(getenv("CDB_IGEMM_KGROUP"));
  if (kgroup > 4) kgroup = 4;
  if (kgroup > prm.n_taps * prm.k_chunks) kgroup = prm.n_taps * prm.k_chunks;
  if (kgroup < 1) kgroup = 1;
  prm.kgroup = kgroup;
  const int stage_bytes = kgroup * kblock_bytes;
  int stages = budget / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  prm.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 32768 + 1024;
  static size_t smem_attr[2] = {0, 0};
  if (smem > smem_attr[pair ? 1 : 0]) {
    if (pair) CDB_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CDB_CUDA_OK(cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr[pair ? 1 : 0] = smem;
  }
  prm.abort_flag = device_abort_flag_ptr();
  static long long* dbg_buf = nullptr;
  if (getenv("CDB_IGEMM_DEBUG")) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 128);
    prm.dbg = dbg_buf;
  }
  if (pair) {
    const int total = ((m_tiles + 1) / 2) * prm.n_tiles_n;
    if (total < 1) return CDB_OK;
    const int clusters = total < sm_count() / 2 ? total : sm_count() / 2;
    CDB_CUDA_OK(launch_ex(igemm_kernel<true>, dim3(2 * clusters, 1, 1), dim3(256, 1, 1), smem, stream, 2, true, maps, prm));
  } else {
    const int total = m_tiles * prm.n_tiles_n;
    int grid = total < sm_count() ? total : sm_count();
    if (grid < 1) return CDB_OK;
    CDB_CUDA_OK(launch_ex(igemm_kernel<false>, dim3(grid, 1, 1), dim3(256, 1, 1), smem, stream, 1, true, maps, prm));
  }
  CDB_LAUNCH_OK();
  if (prm.dbg) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, 128, cudaMemcpyDeviceToHost);
    fprintf(stderr,
            "[igemm dbg] pair=%d m_tiles=%d bn=%d taps=%d chunks=%d stages=%d x %d K blocks | CTA0 tiles %lld | producer: total %lld, waiting for "
            "empty %lld | mma: total %lld, waiting tempty %lld, full %lld | epilogue warp 0: total %lld, waiting tfull %lld, store "
            "%lld\n",
            (int)pair, m_tiles, prm.bn, prm.n_taps, prm.k_chunks, prm.stages, prm.kgroup, h[9], h[2], h[1], h[5], h[3], h[4], h[8], h[6],
            h[7]);
  }
  return CDB_OK;
}

static int check_act(const CdbAct* x, const char* name) {
  CDB_REQUIRE(x && x->ptr, CDB_ERR_BAD_DESC, "%s: null tensor", name);
  CDB_REQUIRE(x->dtype == CDB_BF16 || x->dtype == CDB_F32, CDB_ERR_UNSUPPORTED, "%s: bf16 or fp32 (TF32) activations",
              name);
  const int q = x->dtype == CDB_BF16 ? 8 : 4;  // elements per 16 bytes
  CDB_REQUIRE(x->c % q == 0 && x->sw % q == 0 && x->sh % q == 0 && x->sn % q == 0 &&
                  (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0,
              CDB_ERR_ALIGNMENT, "%s: channels/strides must be multiples of 16 bytes and ptr 16B aligned", name);
  return CDB_OK;
}

int launch_flat_conv(const CdbConvGeom* g, const CdbAct* x, const void* wpacked, int w_rows_pad, int w_kpad,
                     const CdbOut* y, const CdbEpilogue* ep, int flip, int use_base_offset, cudaStream_t stream);

static int env_flag(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_conv2d_fwd(const CdbConvGeom* g, const CdbAct* x, const void* wpacked,
                              int32_t w_rows_pad, int32_t w_kpad, const CdbOut* y,
                              const CdbEpilogue* ep, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(g && x && y && wpacked, CDB_ERR_BAD_DESC, "conv2d_fwd: null argument");
  int rc = check_act(x, "conv2d_fwd x");
  if (rc) return rc;
  CDB_REQUIRE(g->stride == 1 || g->stride == 2, CDB_ERR_UNSUPPORTED, "conv2d_fwd: stride %d", g->stride);
  CDB_REQUIRE(g->dil >= 1 && (g->dil == 1 || g->stride == 1), CDB_ERR_UNSUPPORTED, "conv2d_fwd: dilation with stride");
  CDB_REQUIRE(y->c >= 1 && y->cstore >= y->c, CDB_ERR_BAD_DESC, "conv2d_fwd: bad output channels");
  const bool tf32 = x->dtype == CDB_F32;
  const size_t xesz = tf32 ? 4 : 2;
  CDB_REQUIRE(w_rows_pad % 16 == 0 && w_rows_pad >= y->c && w_kpad % (tf32 ? 32 : 64) == 0, CDB_ERR_BAD_DESC,
              "conv2d_fwd: packed weight geometry");
  CDB_REQUIRE(!(tf32 && g->rowpack), CDB_ERR_UNSUPPORTED, "conv2d_fwd: rowpack is bf16 only");
  if (y->dtype == CDB_BF16 && y->sc == 1)
    CDB_REQUIRE(y->sn % 8 == 0 && y->sh % 8 == 0 && y->sw % 8 == 0 && y->cstore % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0,
                CDB_ERR_ALIGNMENT, "conv2d_fwd: bf16 NHWC output must be 16B aligned per pixel");
  const int st = g->stride;
  const int n_real_taps = g->rowpack ? g->r : g->r * g->s;
  CDB_REQUIRE(n_real_taps <= kMaxTaps, CDB_ERR_UNSUPPORTED, "conv2d_fwd: too many taps");

  IgemmParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.cout = y->c;
  prm.cstore = y->cstore;
  prm.bn = w_rows_pad < 256 ? w_rows_pad : 256;
  {
    // small problems (inference at batch 1): narrower channel tiles until the tile count covers half of the SMs
    const int px = y->n * (g->transposed ? ceil_div(y->h, g->stride) * ceil_div(y->w, g->stride) : y->h * y->w);
    const int m_tiles = ceil_div(px, 128);
    while (prm.bn > 64 && prm.bn % 32 == 0 && m_tiles * ceil_div(w_rows_pad, prm.bn) * 2 < sm_count()) prm.bn /= 2;
  }
  prm.n_tiles_n = ceil_div(w_rows_pad, prm.bn);
  prm.out_dtype = y->dtype;
  prm.act = ep ? ep->act : CDB_ACT_NONE;
  prm.slope = ep ? ep->slope : 0.f;
  prm.bias = ep ? ep->bias : nullptr;
  prm.stats = ep ? ep->stats : nullptr;
  prm.stats_on = prm.stats != nullptr;
  prm.stats_batch = (ep && (ep->flags & CDB_EP_STATS_BATCH)) ? 1 : 0;
  prm.tf32 = tf32 ? 1 : 0;
  prm.round_out = (ep && (ep->flags & CDB_EP_ROUND_TF32)) ? 1 : 0;
  prm.k_chunks = w_kpad / (tf32 ? 32 : 64);
  prm.dom_n = y->n;

  CDB_REQUIRE(!(g->flip && (g->transposed || g->rowpack)), CDB_ERR_UNSUPPORTED, "conv2d_fwd: flip with transposed/rowpack");
  // Stride-1 convolutions over a contiguous buffer with materialised padding take the flat kernel
  // (A rows shared by the S taps of a filter row, two accumulators per weight tile).
  // (the buffer may be a channel prefix of a wider concatenation buffer: pixel pitch sw >= c, regular rows)
  if (!(tf32 && env_flag("CDB_TF32_DISABLE_FLAT", 0)) && !g->transposed && !g->rowpack && st == 1 && g->pad_h == 0 && g->pad_w == 0 && x->sw >= x->c &&
      x->sh == (int64_t)x->w * x->sw && x->sn == (int64_t)x->h * x->w * x->sw &&
      y->h <= x->h - (g->r - 1) * g->dil && y->w <= x->w - (g->s - 1) * g->dil && y->n == x->n &&
      (g->s - 1) * g->dil <= 64 && (int64_t)y->h * x->w >= 256 && !env_flag("CDB_DISABLE_FLAT", 0)) {
    return launch_flat_conv(g, x, wpacked, w_rows_pad, w_kpad, y, ep, g->flip, env_flag("CDB_FLAT_BASE_OFFSET", 0),
                            stream);
  }
  if (!g->transposed) {
    // ---------------- direct convolution: parity views of the input for stride 2
    SrcView views[4];
    int n_views = st * st;
    if (g->rowpack) {
      CDB_REQUIRE(g->pad_w == 0 && g->s * g->rowpack <= 64 && (g->rowpack == 8 || g->rowpack == 16) &&
                      x->c == g->rowpack && w_kpad == 64,
                  CDB_ERR_BAD_DESC, "conv2d_fwd: rowpack geometry");
      const int span = 64 / g->rowpack;  // pixels covered by one K block
      CDB_REQUIRE(x->w >= st * (y->w - 1) + span, CDB_ERR_BAD_DESC,
                  "conv2d_fwd: rowpack input needs w >= %d (has %d)", st * (y->w - 1) + span, x->w);
      n_views = st;
      for (int a = 0; a < st; ++a) {
        views[a].ptr = static_cast<__nv_bfloat16*>(x->ptr) + a * x->sh;
        views[a].dims[0] = 64;
        views[a].dims[1] = (x->w - span) / st + 1;
        views[a].dims[2] = (x->h - a + st - 1) / st;
        views[a].dims[3] = x->n;
        views[a].strides[0] = x->sw * st;
        views[a].strides[1] = x->sh * st;
        views[a].strides[2] = x->sn;
      }
      prm.n_taps = g->r;
      for (int r = 0; r < g->r; ++r) {
        const int ih = r * g->dil - g->pad_h;
        const int a = pos_mod(ih, st);
        prm.taps[r].map = (int16_t)a;
        prm.taps[r].dh = (int16_t)((ih - a) / st);
        prm.taps[r].dw = 0;
        prm.taps[r].wk = r * 64;
      }
    } else {
      for (int a = 0; a < st; ++a)
        for (int b = 0; b < st; ++b) {
          SrcView& v = views[a * st + b];
          v.ptr = static_cast<char*>(x->ptr) + (a * x->sh + b * x->sw) * (int64_t)xesz;
          v.dims[0] = x->c;
          v.dims[1] = (x->w - b + st - 1) / st;
          v.dims[2] = (x->h - a + st - 1) / st;
          v.dims[3] = x->n;
          v.strides[0] = x->sw * st;
          v.strides[1] = x->sh * st;
          v.strides[2] = x->sn;
        }
      prm.n_taps = g->r * g->s;
      for (int r = 0; r < g->r; ++r)
        for (int s = 0; s < g->s; ++s) {
          const int ih = r * g->dil - g->pad_h, iw = s * g->dil - g->pad_w;
          const int a = pos_mod(ih, st), b = pos_mod(iw, st);
          IgemmTap& t = prm.taps[r * g->s + s];
          t.map = (int16_t)(a * st + b);
          t.dh = (int16_t)((ih - a) / st);
          t.dw = (int16_t)((iw - b) / st);
          t.wk = (g->flip ? (g->r * g->s - 1 - (r * g->s + s)) : (r * g->s + s)) * w_kpad;
        }
    }
    prm.dom_h = y->h;
    prm.dom_w = y->w;
    choose_tile(y->n, y->h, y->w, &prm.tile_w, &prm.tile_h, &prm.tile_n);
    prm.tiles_w = ceil_div(y->w, prm.tile_w);
    prm.tiles_h = ceil_div(y->h, prm.tile_h);
    prm.tiles_n = ceil_div(y->n, prm.tile_n);
    prm.out = y->ptr;
    prm.o_sn = y->sn;
    prm.o_sh = y->sh;
    prm.o_sw = y->sw;
    prm.o_sc = y->sc;
    const int ktotal = (g->rowpack ? g->r : g->r * g->s) * w_kpad;
    return launch_igemm(views, n_views, wpacked, w_rows_pad, ktotal, prm, stream);
  }

  // ---------------- transposed convolution: one launch per output parity class
  CDB_REQUIRE(!g->rowpack, CDB_ERR_UNSUPPORTED, "conv2d_fwd: rowpack with transposed");
  SrcView view;
  view.ptr = x->ptr;
  view.dims[0] = x->c;
  view.dims[1] = x->w;
  view.dims[2] = x->h;
  view.dims[3] = x->n;
  view.strides[0] = x->sw;
  view.strides[1] = x->sh;
  view.strides[2] = x->sn;
  const int ktotal = g->r * g->s * w_kpad;
  const size_t esz = y->dtype == CDB_BF16 ? 2 : 4;
  for (int a = 0; a < st; ++a)
    for (int b = 0; b < st; ++b) {
      const int dom_h = (y->h - a + st - 1) / st, dom_w = (y->w - b + st - 1) / st;
      if (dom_h <= 0 || dom_w <= 0) continue;
      IgemmParams q = prm;
      int nt = 0;
      for (int r = 0; r < g->r; ++r) {
        const int ih = a + g->pad_h - r * g->dil;
        if (pos_mod(ih, st) != 0) continue;
        for (int s = 0; s < g->s; ++s) {
          const int iw = b + g->pad_w - s * g->dil;
          if (pos_mod(iw, st) != 0) continue;
          IgemmTap& t = q.taps[nt++];
          t.map = 0;
          t.dh = (int16_t)(ih / st);
          t.dw = (int16_t)(iw / st);
          t.wk = (r * g->s + s) * w_kpad;
        }
      }
      q.n_taps = nt;
      q.dom_h = dom_h;
      q.dom_w = dom_w;
      choose_tile(y->n, dom_h, dom_w, &q.tile_w, &q.tile_h, &q.tile_n);
      q.tiles_w = ceil_div(dom_w, q.tile_w);
      q.tiles_h = ceil_div(dom_h, q.tile_h);
      q.tiles_n = ceil_div(y->n, q.tile_n);
      q.out = static_cast<char*>(y->ptr) + (a * y->sh + b * y->sw) * (int64_t)esz;
      q.o_sn = y->sn;
      q.o_sh = y->sh * st;
      q.o_sw = y->sw * st;
      q.o_sc = y->sc;
      if (nt == 0) return fail(CDB_ERR_UNSUPPORTED, "conv2d_fwd: transposed parity class without taps");
      rc = launch_igemm(&view, 1, wpacked, w_rows_pad, ktotal, q, stream);
      if (rc) return rc;
    }
  return CDB_OK;
}
