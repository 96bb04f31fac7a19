// K1c — the IMAGE layers: stride-1 convolutions whose input has at most 8 channels (the 7x7 c7s1-64 input layer of
// the generators, models/networks.py:158, and the data gradient of the 7x7 c7s1-3 output layer, :185, whose
// "input" is the 3-channel dy).
//
// With 8-channel (16-byte) pixels a filter row is ONE K block: k = s * 8 + c, 8 taps x 8 channels = 64.  The A
// operand of that block is the Toeplitz matrix  A[m][s * 8 + c] = P[m + s][c]  over the contiguous pixel row P — and
// tcgen05 can read it straight from a PLAIN copy of the row segment through a no-swizzle K-major descriptor with
// LBO = 16 B (next K chunk = next pixel) and SBO = 128 B (next group of 8 rows = 8 pixels further), i.e. overlapping
// core matrices (verified on B200, tools/toeplitz_probe.cu).  A tile of 128 flat output positions therefore needs
// R bulk copies of (128 + 7) x 16 B — 15 KB — where the row-packed implicit-GEMM path of conv_igemm.cu fetches 128
// overlapping 128-byte rows per filter row (8x read amplification: 325 us per launch at batch 24).
//
// Geometry as in conv_flat.cu: the padded input [N, Hp, Wp, 8] is a flat pixel array, an output position is its
// flat index f = h * Wp + w in the same pitch, tap (r, s) reads pixel f + r * Wp + s; the (Wp - Q) junk positions
// per row are computed and dropped.  The R x (cout x 64) weight blocks stay resident in shared memory.
// Warps: 0 producer (cp.async.bulk), 1 MMA issuer, 2 TMEM allocator, 4-7 epilogue (bias / activation / InstanceNorm
// statistics / bf16 / TMA store); accumulators double-buffered in TMEM.
#if 1
#define CDB_WAIT_SPIN 1
#endif
#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

constexpr int kTzBM = 128;
constexpr int kTzTaps = 8;                       // K chunks (taps) per filter row
constexpr int kTzSeg = kTzBM + 2 * kTzTaps;      // 144 pixels per row segment: 135 needed, started at a multiple of 8
                                                 // pixels so that source and destination of the bulk copy are both
                                                 // 128-byte aligned (a copy with different alignments on the two sides
                                                 // runs at ~16 B per 10 ns: 772 us instead of ~100 per launch at batch 24)
constexpr int kTzRowStride = kTzSeg * 16;        // 2304 bytes per filter row inside a stage (multiple of 128)
constexpr int kTzStages = 4;
constexpr int kTzMaxR = 8;

struct TzParams {
  int32_t R, cout, cstore, rows_pad;
  int32_t wp, rows_per_img_in;        // input pitch (pixels) and pixels per image
  int64_t total_px;                   // pixels in the whole input buffer (reads are clamped to it)
  int32_t tiles_per_img, n_img, out_rows_per_img;
  int32_t dom_h, dom_w;
  int32_t act, stats_on, stats_batch, fast_out, out_dtype;
  float slope;
  const float* bias;
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;             // [R][rows_pad x 64] no-swizzle core-matrix order
  void* out;
  int64_t o_sn, o_sh, o_sw, o_sc;
  float* stats;
  int* abort_flag;
  int32_t dbg;   // CDB_TZ_DBG bisecting switches: 1 skip the output stores, 2 skip the MMAs, 4 skip the input copies
  long long* ts; // CDB_TZ_DBG & 8: clock64 stamps of CTA 0: [role 0..5][16 tiles]
};

__device__ __forceinline__ void tz_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ float tz_act(float v, int act, float slope) {
  switch (act) {
    case CDB_ACT_RELU: return v > 0.f ? v : 0.f;
    case CDB_ACT_LEAKY: return v > 0.f ? v : v * slope;
    case CDB_ACT_TANH: return tanhf(v);
    case CDB_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// kRegStats (<= 64 output channels with statistics): the per-channel sums stay in per-thread registers (one partial
// per row lane and column) while the CTA's consecutive tiles belong to the same statistics group, and are reduced
// across the warp and added with atomics only when the group changes — a tile of this kernel is only 28 MMAs, so
// per-tile shared-memory transposes and atomics would dominate it (188 -> ~50 us at batch 24).
template <bool kRegStats>
__global__ void __launch_bounds__(256, 1)
toeplitz_conv_kernel(const __grid_constant__ CUtensorMap out_map, const __grid_constant__ TzParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kTzStages], bar_empty[kTzStages], bar_tfull[2], bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t w_row_bytes = static_cast<uint32_t>(p.rows_pad) * 128u;      // one filter row of weights
  const uint32_t w_bytes = static_cast<uint32_t>(p.R) * w_row_bytes;
  const uint32_t stage_bytes = static_cast<uint32_t>(p.R) * kTzRowStride;
  const uint32_t w_base = base, st_base = base + w_bytes;
  const uint32_t slab_base = (st_base + kTzStages * stage_bytes + 1023u) & ~1023u;   // 4 warps x 2 x 4 KB, 1024-aligned
  const int total_tiles = p.n_img * p.tiles_per_img;
  // every CTA owns a CONTIGUOUS range of tiles: its statistics group (image) changes once or twice per launch instead
  // of every few tiles, and it streams one region of the input
  const int tiles_per_cta = (total_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int tile_begin = static_cast<int>(blockIdx.x) * tiles_per_cta;
  const int tile_end = min(total_tiles, tile_begin + tiles_per_cta);
  const uint32_t tmem_cols = 2u * static_cast<uint32_t>(p.rows_pad) <= 32u    ? 32u
                             : 2u * static_cast<uint32_t>(p.rows_pad) <= 64u  ? 64u
                             : 2u * static_cast<uint32_t>(p.rows_pad) <= 128u ? 128u
                             : 2u * static_cast<uint32_t>(p.rows_pad) <= 256u ? 256u
                                                                              : 512u;

  for (uint32_t i = threadIdx.x; i < w_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gen)[i] = reinterpret_cast<const uint4*>(p.w)[i];
  // the row-segment buffers start out as zeros: a segment clamped at the end of the input leaves its tail untouched,
  // and that tail must hold finite values (it only feeds dropped positions, but 0 * NaN would poison the statistics)
  for (uint32_t i = threadIdx.x; i < kTzStages * stage_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(gen + w_bytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    abort_smem = 0;
    for (int s = 0; s < kTzStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), 4);      // one elected lane per epilogue warp arrives
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_smem();   // the weight blocks were written by ordinary stores and are read by the tensor core
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  volatile int* abort_flag = &abort_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ---------------------------------------------------------------- producer: R row segments per tile
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int img = tile / p.tiles_per_img;
        const int f0 = (tile - img * p.tiles_per_img) * kTzBM;
        if (!mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u, abort_flag)) break;
        if (p.ts && blockIdx.x == 0 && tile < 16) p.ts[0 * 16 + tile] = clock64();
        const uint32_t full = smem_u32(&bar_full[stage]);
        // segments are clamped to the end of the buffer (the rows they would add belong to dropped positions)
        uint32_t bytes[kTzMaxR];
        uint32_t total = 0;
        const int64_t px0 = static_cast<int64_t>(img) * p.rows_per_img_in + f0;
        for (int r = 0; r < p.R; ++r) {
          const int64_t start = (px0 + static_cast<int64_t>(r) * p.wp) & ~static_cast<int64_t>(7);
          int64_t n_px = p.total_px - start;
          n_px = n_px > kTzSeg ? kTzSeg : (n_px < 0 ? 0 : n_px);
          bytes[r] = static_cast<uint32_t>(n_px) * 16u;
          total += bytes[r];
        }
        if (total == 0 || (p.dbg & 4)) {   // (total == 0 cannot happen for a tile with valid positions)
          mbar_arrive(full);
        } else {
          mbar_arrive_expect_tx(full, total);
          for (int r = 0; r < p.R; ++r)
            if (bytes[r])
              tz_bulk_load(st_base + stage * stage_bytes + r * kTzRowStride,
                           p.x + ((px0 + static_cast<int64_t>(r) * p.wp) & ~static_cast<int64_t>(7)) * 8, bytes[r], full);
        }
        if (++stage == kTzStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer: R filter rows x 4 instructions
      const uint32_t idesc = make_idesc(1u, 0u, 0u, 128u, static_cast<uint32_t>(p.rows_pad));
      int stage = 0, local = 0;
      uint32_t phase = 0;
      for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
        const int buf = local & 1;
        const uint32_t tphase = (local >> 1) & 1u;
        if (!mbar_wait(smem_u32(&bar_tempty[buf]), tphase ^ 1u, abort_flag)) break;
        if (!mbar_wait(smem_u32(&bar_full[stage]), phase, abort_flag)) break;
        tc_fence_after();
        if (p.ts && blockIdx.x == 0 && local < 16) p.ts[1 * 16 + local] = clock64();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.rows_pad);
        const int img = tile / p.tiles_per_img;
        const int64_t px0 = static_cast<int64_t>(img) * p.rows_per_img_in + (tile - img * p.tiles_per_img) * kTzBM;
        for (int r = 0; r < p.R; ++r) {
          // Toeplitz A: K chunk j of row m = pixel m + j  (LBO 16 B, SBO 128 B, no swizzle); the segment was copied from
          // the 8-pixel boundary below its first pixel, which sits `off` pixels into the buffer
          const uint32_t off = static_cast<uint32_t>((px0 + static_cast<int64_t>(r) * p.wp) & 7);
          const uint64_t da = make_smem_desc(st_base + stage * stage_bytes + r * kTzRowStride + off * 16u, 16, 128, 0u);
          const uint64_t db = make_smem_desc(w_base + r * w_row_bytes, 128, 1024, 0u);
          if (!(p.dbg & 2)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16(d_tmem, da + 2u * k, db + 16u * k, idesc, (r | k) ? 1u : 0u);
          }
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        umma_commit(smem_u32(&bar_tfull[buf]));
        if (p.ts && blockIdx.x == 0 && local < 16) p.ts[5 * 16 + local] = clock64();
        if (++stage == kTzStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int ew = warp - 4;
    // two staging buffers per warp: the TMA store of slab i is still reading its buffer while slab i + 1 is packed
    const uint32_t stage_addr0 = slab_base + ew * 8192;
    float* slab = reinterpret_cast<float*>(smem_raw + (stage_addr0 - smem_u32(smem_raw)));
    uint32_t stage_sel = 0;
    const bool has_bias = p.bias != nullptr;
    const int cout = p.cout, cstore = p.cstore;
    const bool stats_on = p.stats_on != 0;
    int local = 0;
    float r1[kRegStats ? 64 : 1], r2[kRegStats ? 64 : 1];
    int cur_group = -1;
    if (kRegStats) {
#pragma unroll
      for (int j = 0; j < (kRegStats ? 64 : 1); ++j) r1[j] = r2[j] = 0.f;
    }
    auto flush = [&]() {
      if (!kRegStats || cur_group < 0) return;
      if (p.fast_out) {   // the staging area (aliased by `slab`) may still be read by the last TMA store
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      }
      float* dst = p.stats + static_cast<int64_t>(cur_group) * cout * 2;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float s[2];
#pragma unroll
        for (int which = 0; which < 2; ++which) {
#pragma unroll
          for (int j = 0; j < 16; ++j) slab[lane * 17 + j] = which ? r2[kRegStats ? q * 16 + j : 0] : r1[kRegStats ? q * 16 + j : 0];
          __syncwarp();
          float acc = 0.f;
          if (lane < 16) {
#pragma unroll 8
            for (int i = 0; i < 32; ++i) acc += slab[i * 17 + lane];
          }
          s[which] = acc;
          __syncwarp();
        }
        const int ch = q * 16 + lane;
        if (lane < 16 && ch < cout) {
          atomicAdd(dst + ch * 2, s[0]);
          atomicAdd(dst + ch * 2 + 1, s[1]);
        }
      }
#pragma unroll
      for (int j = 0; j < (kRegStats ? 64 : 1); ++j) r1[j] = r2[j] = 0.f;
    };
    for (int tile = tile_begin; tile < tile_end; ++tile, ++local) {
      const int buf = local & 1;
      const uint32_t tphase = (local >> 1) & 1u;
      const int img = tile / p.tiles_per_img;
      const int f0 = (tile - img * p.tiles_per_img) * kTzBM;
      const int f = f0 + ew * 32 + lane;
      if (kRegStats) {
        const int group = p.stats_batch ? 0 : img;
        if (group != cur_group) {
          flush();
          cur_group = group;
        }
      }
      const int h = f / p.wp, w = f - h * p.wp;
      const bool valid = (h < p.dom_h) && (w < p.dom_w);
      if (!mbar_wait(smem_u32(&bar_tfull[buf]), tphase, abort_flag)) break;
      tc_fence_after();
      if (p.ts && blockIdx.x == 0 && local < 16 && threadIdx.x == 128) p.ts[2 * 16 + local] = clock64();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(buf * p.rows_pad);
      float* stats_img = stats_on ? p.stats + static_cast<int64_t>(p.stats_batch ? 0 : img) * cout * 2 : nullptr;
      const int n_chunks = (min(p.rows_pad, cstore) + 63) / 64;
      for (int c0 = 0; c0 < p.rows_pad; c0 += 64) {
        if (c0 >= cstore) break;
        uint32_t v[64];
        if (p.rows_pad - c0 >= 64) {
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, v + 32);
        } else {
          // fewer than 64 columns left (rows_pad is a multiple of 16): 16 at a time, the rest zero
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c0 + q * 16 < p.rows_pad) {
              tmem_ld16(taddr + c0 + q * 16, v + q * 16);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[q * 16 + j] = 0u;
            }
          }
        }
        tmem_ld_wait();
        if (c0 / 64 == n_chunks - 1) {   // the accumulator is in registers: the MMA warp may start the tile after next
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bar_tempty[buf]));
          if (p.ts && blockIdx.x == 0 && local < 16 && threadIdx.x == 128) p.ts[3 * 16 + local] = clock64();
        }
        // every branch below is uniform and guarded OUTSIDE its unrolled loop: with the activation switch inside, the
        // 64 copies of the tanh / sigmoid paths cost ~9000 predicated instructions per tile (measured: 14 k cycles)
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const int ch = c0 + j;
            if (ch < cout) v[j] = __float_as_uint(__uint_as_float(v[j]) + __ldg(p.bias + ch));
          }
        }
        if (p.act == CDB_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(fmaxf(__uint_as_float(v[j]), 0.f));
        } else if (p.act == CDB_ACT_LEAKY) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float t = __uint_as_float(v[j]);
            v[j] = __float_as_uint(t > 0.f ? t : t * p.slope);
          }
        } else if (p.act == CDB_ACT_TANH) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(tanhf(__uint_as_float(v[j])));
        } else if (p.act == CDB_ACT_SIGMOID) {
#pragma unroll
          for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(1.f / (1.f + __expf(-__uint_as_float(v[j]))));
        }
        if (c0 + 64 > cout) {
#pragma unroll
          for (int j = 0; j < 64; ++j)
            if (c0 + j >= cout) v[j] = 0u;
        }
        const uint32_t stage_addr = stage_addr0 + stage_sel * 4096;
        if (p.fast_out) {
          // the store issued two slabs ago has finished reading this staging buffer; the shared-memory transposes of the
          // non-register statistics path alias buffer 0 and need every store drained
          if (lane == 0) {
            if (!kRegStats && stats_on) bulk_wait_read0();
            else bulk_wait_read1();
          }
          __syncwarp();
        }
        if (kRegStats) {
          if (valid) {
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const float t = __uint_as_float(v[j]);
              r1[kRegStats ? j : 0] += t;
              r2[kRegStats ? j : 0] = fmaf(t, t, r2[kRegStats ? j : 0]);
            }
          }
        } else if (stats_on) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (c0 + q * 16 >= cout) break;
#pragma unroll
            for (int j = 0; j < 16; ++j) slab[lane * 17 + j] = valid ? __uint_as_float(v[q * 16 + j]) : 0.f;
            __syncwarp();
            if (lane < 16) {
              float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
              for (int i = 0; i < 32; ++i) {
                const float t = slab[i * 17 + lane];
                s1 += t;
                s2 = fmaf(t, t, s2);
              }
              const int ch = c0 + q * 16 + lane;
              if (ch < cout) {
                atomicAdd(stats_img + ch * 2, s1);
                atomicAdd(stats_img + ch * 2 + 1, s2);
              }
            }
            __syncwarp();
          }
        }
        if (p.dbg & 1) continue;
        if (p.fast_out) {
          // 32 rows x 64 bf16 columns -> swizzled staging -> TMA store (rows of the pitched output buffer)
          const uint32_t rbase = stage_addr + lane * 128;
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
            tma_store_2d(&out_map, stage_addr, c0, img * p.out_rows_per_img + f0 + ew * 32);
            bulk_commit();
          }
          stage_sel ^= 1u;
        } else if (valid) {
          const int64_t obase = img * p.o_sn + h * p.o_sh + w * p.o_sw;
          if (p.out_dtype == CDB_BF16) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + obase;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const int ch = c0 + j;
              if (ch < cstore) o[ch * p.o_sc] = __float2bfloat16(__uint_as_float(v[j]));
            }
          } else {
            float* o = static_cast<float*>(p.out) + obase;
#pragma unroll
            for (int j = 0; j < 64; ++j) {
              const int ch = c0 + j;
              if (ch < cstore) o[ch * p.o_sc] = __uint_as_float(v[j]);
            }
          }
        }
      }
      if (p.ts && blockIdx.x == 0 && local < 16 && threadIdx.x == 128) p.ts[4 * 16 + local] = clock64();
    }
    if (p.fast_out) {
      if (lane == 0) bulk_wait_read0();
      __syncwarp();
    }
    flush();
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && abort_smem && p.abort_flag) atomicExch(p.abort_flag, 1);
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// fp32 W4[d0][d1][R][S] -> bf16 [R][rows_pad x 64] blocks, K-major in the no-swizzle core-matrix order:
//   element (row o, k = s * 8 + c) of filter row r at  r * rows_pad * 64 + (o / 8) * 512 + (k / 8) * 64 + (o % 8) * 8 + k % 8
// rows are d0 (rows_are_dim0) or d1, the <= 8 contracted channels the other one; flip packs W[.., R-1-r, S-1-s].
__global__ void pack_toeplitz_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int d0, int d1, int R,
                                     int S, int rows_are_dim0, int flip, int rows, int kch, int rows_pad) {
  const int total = R * rows_pad * 64;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int r = idx / (rows_pad * 64);
    int t = idx - r * rows_pad * 64;
    const int og = t / 512;
    t -= og * 512;
    const int kg = t / 64;
    t -= kg * 64;
    const int o = og * 8 + t / 8, k = kg * 8 + t % 8;
    const int s = k / 8, c = k % 8;
    float v = 0.f;
    if (o < rows && s < S && c < kch) {
      const int rr = flip ? R - 1 - r : r, ss = flip ? S - 1 - s : s;
      const int i0 = rows_are_dim0 ? o : c, i1 = rows_are_dim0 ? c : o;
      v = w[((static_cast<int64_t>(i0) * d1 + i1) * R + rr) * S + ss];
    }
    out[idx] = __float2bfloat16(v);
  }
}

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_pack_toeplitz_weight(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s,
                                        int32_t rows_are_dim0, int32_t flip, void* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(w4 && out, CDB_ERR_BAD_DESC, "pack_toeplitz_weight: null argument");
  const int rows = rows_are_dim0 ? d0 : d1, kch = rows_are_dim0 ? d1 : d0;
  CDB_REQUIRE(r >= 1 && r <= kTzMaxR && s >= 1 && s <= kTzTaps && kch >= 1 && kch <= 8 && rows >= 1, CDB_ERR_BAD_DESC,
              "pack_toeplitz_weight: needs R, S <= 8 and <= 8 contracted channels");
  const int rows_pad = round_up(rows, 16);
  const int total = r * rows_pad * 64;
  pack_toeplitz_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(w4, static_cast<__nv_bfloat16*>(out), d0, d1, r, s,
                                                                  rows_are_dim0, flip, rows, kch, rows_pad);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_conv2d_toeplitz_fwd(const CdbAct* x, const void* wpacked, int32_t w_rows_pad, int32_t r, int32_t s,
                                       const CdbOut* y, const CdbEpilogue* ep, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(x && x->ptr && wpacked && y && y->ptr, CDB_ERR_BAD_DESC, "conv2d_toeplitz_fwd: null argument");
  CDB_REQUIRE(x->dtype == CDB_BF16 && x->c == 8 && x->sw == 8 && x->sh == (int64_t)x->w * 8 &&
                  x->sn == (int64_t)x->h * x->w * 8 && (reinterpret_cast<uintptr_t>(x->ptr) & 15) == 0,
              CDB_ERR_UNSUPPORTED, "conv2d_toeplitz_fwd: x must be a contiguous bf16 buffer of 8-channel pixels");
  CDB_REQUIRE(r >= 1 && r <= kTzMaxR && s >= 1 && s <= kTzTaps, CDB_ERR_UNSUPPORTED, "conv2d_toeplitz_fwd: R, S <= 8");
  CDB_REQUIRE(w_rows_pad % 16 == 0 && w_rows_pad >= y->c && w_rows_pad <= 128 && y->cstore >= y->c, CDB_ERR_UNSUPPORTED,
              "conv2d_toeplitz_fwd: <= 128 output channels");
  CDB_REQUIRE(y->n == x->n && y->h >= 1 && y->h <= x->h - (r - 1) && y->w >= 1 && y->w <= x->w - (s - 1), CDB_ERR_BAD_DESC,
              "conv2d_toeplitz_fwd: output larger than the valid convolution of x");
  TzParams p;
  memset(&p, 0, sizeof(p));
  p.R = r;
  p.cout = y->c;
  p.cstore = y->cstore;
  p.rows_pad = w_rows_pad;
  p.wp = x->w;
  p.rows_per_img_in = x->h * x->w;
  p.total_px = (int64_t)x->n * x->h * x->w;
  p.tiles_per_img = ceil_div(y->h * x->w, kTzBM);
  p.n_img = y->n;
  p.dom_h = y->h;
  p.dom_w = y->w;
  p.act = ep ? ep->act : CDB_ACT_NONE;
  p.slope = ep ? ep->slope : 0.f;
  p.bias = ep ? ep->bias : nullptr;
  p.stats = ep ? ep->stats : nullptr;
  p.stats_on = p.stats != nullptr;
  p.stats_batch = (ep && (ep->flags & CDB_EP_STATS_BATCH)) ? 1 : 0;
  p.x = static_cast<const __nv_bfloat16*>(x->ptr);
  p.w = static_cast<const __nv_bfloat16*>(wpacked);
  p.out = y->ptr;
  p.out_dtype = y->dtype;
  p.o_sn = y->sn;
  p.o_sh = y->sh;
  p.o_sw = y->sw;
  p.o_sc = y->sc;
  p.abort_flag = device_abort_flag_ptr();
  p.dbg = getenv("CDB_TZ_DBG") ? atoi(getenv("CDB_TZ_DBG")) : 0;
  static long long* ts_buf = nullptr;
  if (p.dbg & 8) {
    if (!ts_buf) cudaMalloc(&ts_buf, 96 * sizeof(long long));
    cudaMemsetAsync(ts_buf, 0, 96 * sizeof(long long), stream);
    p.ts = ts_buf;
  }
  // fast output: bf16 NHWC rows in the INPUT pitch, whole 128-row tiles per image (ops.alloc_flat_output), >= 64 stored
  // channels per 64-column slab
  p.out_rows_per_img = (int)(y->sn / (y->cstore > 0 ? y->cstore : 1));
  p.fast_out = (y->dtype == CDB_BF16 && y->sc == 1 && y->sw == y->cstore && y->sh == (int64_t)x->w * y->cstore &&
                y->sn % ((int64_t)kTzBM * y->cstore) == 0 && p.out_rows_per_img >= p.tiles_per_img * kTzBM &&
                y->cstore % 64 == 0 && (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0)
                   ? 1
                   : 0;
  CUtensorMap out_map;
  memset(&out_map, 0, sizeof(out_map));
  if (p.fast_out) {
    uint64_t dims[2] = {(uint64_t)y->cstore, (uint64_t)y->n * p.out_rows_per_img};
    uint64_t str[1] = {(uint64_t)y->cstore * 2};
    uint32_t box[2] = {64u, 32u};
    int rc = make_tmap(&out_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y->ptr, dims, str, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)r * w_rows_pad * 128 + (size_t)kTzStages * r * kTzRowStride + 4 * 8192 + 1024 + 1024;
  static size_t smem_attr = 0;
  if (smem > smem_attr) {
    CDB_CUDA_OK(cudaFuncSetAttribute(toeplitz_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CDB_CUDA_OK(cudaFuncSetAttribute(toeplitz_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr = smem;
  }
  const int total = p.n_img * p.tiles_per_img;
  const int grid = total < sm_count() ? total : sm_count();
  if (grid < 1) return CDB_OK;
  if (p.stats_on && w_rows_pad <= 64) toeplitz_conv_kernel<true><<<grid, 256, smem, stream>>>(out_map, p);
  else toeplitz_conv_kernel<false><<<grid, 256, smem, stream>>>(out_map, p);
  if (p.ts) {
    long long h[96];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, ts_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const char* names[6] = {"producer got empty", "mma got full    ", "epilogue got tfull", "epilogue released",
                            "epilogue tile end", "mma committed   "};
    for (int r = 0; r < 6; ++r) {
      fprintf(stderr, "[tz dbg] %s:", names[r]);
      for (int i = 0; i < 12; ++i) fprintf(stderr, " %lld", h[r * 16 + i] - h[0]);
      fprintf(stderr, "\n");
    }
  }
  CDB_LAUNCH_OK();
  return CDB_OK;
}

// ------------------------------------------------------------------------------------------------------------------
// Weight gradient of the image layers: MN-major Toeplitz operand.
//
//   part[r][m][s * 8 + c] = sum_{n,h,w} S[n,h,w,m] * P[n, h + r, w + s, c]
// S: the tensor with many channels whose pixels are iterated (dy of the c7s1-64 layer, the padded input x of the
// c7s1-3 layer), P: the contiguous buffer of 8-channel pixels that is shifted (the padded image, the zero-haloed dy).
// The reduction runs over pixels, so both operands reach the tensor core MN-major: a K tile is 64 consecutive pixels
// of one image row; S comes in through TMA as in conv_wgrad.cu (128B swizzle), and the shifted operand
//   B[n = s * 8 + c][k] = P[k + s][c]
// is read from ONE plain copy of the 71-pixel row segment through a no-swizzle MN-major descriptor with SBO = 16 B
// (next 8-channel chunk of N = next pixel) and LBO = 128 B (next group of 8 pixels of K) — overlapping core matrices
// again (verified on B200, tools/toeplitz_mn_probe.cu).  All R filter rows accumulate side by side in TMEM (R x 64
// columns), so a K tile costs 8-16 KB for S plus R x 1.1 KB for P; the generic path of conv_wgrad.cu fetched R x 8 KB of
// overlapping 128-byte rows for P and re-read S (690 us per launch at batch 24).  Split-K over the K tiles, fp32
// partials in the caller's workspace, deterministic reduction in the finalize kernel.
// ------------------------------------------------------------------------------------------------------------------
namespace cdb {

constexpr int kTwKT = 64;                       // pixels per K tile
constexpr int kTwSeg = kTwKT + 2 * kTzTaps;     // 80 pixels: 71 needed, started at a multiple of 8 pixels (128-byte
                                                // aligned bulk copies, see kTzSeg)
constexpr int kTwRowStride = kTwSeg * 16;       // 1280 bytes per filter row of P inside a stage
constexpr int kTwABytes = 2 * 64 * 128;         // two 64-channel chunks of S
constexpr int kTwPBytes = 8 * kTwRowStride;     // 10240: keeps every stage 1024-aligned
constexpr int kTwStageBytes = kTwABytes + kTwPBytes;
constexpr int kTwStages = 6;

struct TwParams {
  int32_t R, cS;
  int32_t Hs, Ws, n_img, w_tiles;     // S domain and 64-pixel tiles per row
  int32_t pitchP, rowsP;              // P geometry (pixels per row, rows per image)
  int64_t total_px_P;
  int32_t splits, k_per_split, k_tiles;
  const __nv_bfloat16* P;
  float* ws;                          // [splits][R][128][64]
  int* abort_flag;
};

__global__ void __launch_bounds__(256, 1)
toeplitz_wgrad_kernel(const __grid_constant__ CUtensorMap s_map, const __grid_constant__ TwParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kTwStages], bar_empty[kTwStages], bar_tfull;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t tmem_cols = p.R * 64 <= 64 ? 64u : p.R * 64 <= 128 ? 128u : p.R * 64 <= 256 ? 256u : 512u;
  if (threadIdx.x == 0) {
    abort_smem = 0;
    for (int s = 0; s < kTwStages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_tfull), 1);
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) prefetch_tmap(&s_map);
  if (warp == 2) {
    tmem_alloc(smem_u32(&tmem_base_smem), tmem_cols);
    tmem_relinquish();
  }
  {
    // clamped P segments leave their tails untouched: they must hold finite values (they meet zero rows of S)
    uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
    for (int st = 0; st < kTwStages; ++st)
      for (uint32_t i = threadIdx.x; i < kTwPBytes / 16; i += blockDim.x)
        reinterpret_cast<uint4*>(gen + st * kTwStageBytes + kTwABytes)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  volatile int* abort_flag = &abort_smem;
  const int split = blockIdx.x;
  const int kt0 = split * p.k_per_split;
  const int kt1 = min(p.k_tiles, kt0 + p.k_per_split);

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt0; kt < kt1; ++kt) {
        const int wt = kt % p.w_tiles;
        int t = kt / p.w_tiles;
        const int h = t % p.Hs;
        const int img = t / p.Hs;
        const int w0 = wt * kTwKT;
        if (!mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u, abort_flag)) break;
        const uint32_t full = smem_u32(&bar_full[stage]);
        const uint32_t sa = base + stage * kTwStageBytes;
        uint32_t bytes[kTzMaxR];
        uint32_t total = kTwABytes;
        const int64_t px0 = (static_cast<int64_t>(img) * p.rowsP + h) * p.pitchP + w0;
        for (int r = 0; r < p.R; ++r) {
          const int64_t start = (px0 + static_cast<int64_t>(r) * p.pitchP) & ~static_cast<int64_t>(7);
          int64_t n_px = p.total_px_P - start;
          n_px = n_px > kTwSeg ? kTwSeg : (n_px < 0 ? 0 : n_px);
          bytes[r] = static_cast<uint32_t>(n_px) * 16u;
          total += bytes[r];
        }
        mbar_arrive_expect_tx(full, total);
        tma_load_4d(&s_map, full, sa, 0, w0, h, img);
        tma_load_4d(&s_map, full, sa + 64 * 128, 64, w0, h, img);   // channels 64..127 (zero-filled when cS <= 64)
        for (int r = 0; r < p.R; ++r)
          if (bytes[r])
            tz_bulk_load(sa + kTwABytes + r * kTwRowStride,
                         p.P + ((px0 + static_cast<int64_t>(r) * p.pitchP) & ~static_cast<int64_t>(7)) * 8, bytes[r], full);
        if (++stage == kTwStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = make_idesc(1u, 1u, 1u, 128u, 64u);   // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int kt = kt0; kt < kt1 && ok; ++kt) {
        if (!mbar_wait(smem_u32(&bar_full[stage]), phase, abort_flag)) {
          ok = false;
          break;
        }
        tc_fence_after();
        const uint32_t sa = base + stage * kTwStageBytes;
        const uint64_t da = make_smem_desc(sa, 64 * 128, 1024, kLayoutSW128);   // LBO: next 64-channel chunk
        const int wt = kt % p.w_tiles;
        const int t2 = kt / p.w_tiles;
        const int64_t px0 = (static_cast<int64_t>(t2 / p.Hs) * p.rowsP + t2 % p.Hs) * p.pitchP + wt * kTwKT;
        for (int r = 0; r < p.R; ++r) {
          const uint32_t off = static_cast<uint32_t>((px0 + static_cast<int64_t>(r) * p.pitchP) & 7);
          const uint64_t db = make_smem_desc(sa + kTwABytes + r * kTwRowStride + off * 16u, 128, 16, 0u);   // Toeplitz, MN-major
#pragma unroll
          for (int k = 0; k < 4; ++k)   // 16 pixels per instruction: 2048 B of S, 256 B of P
            umma_f16(tmem_base + static_cast<uint32_t>(r * 64), da + 128u * k, db + 16u * k, idesc,
                     (kt > kt0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&bar_empty[stage]));
        if (++stage == kTwStages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (ok) umma_commit(smem_u32(&bar_tfull));
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int m = ew * 32 + lane;
    const bool empty = kt0 >= kt1;
    if (empty || mbar_wait(smem_u32(&bar_tfull), 0u, abort_flag)) {
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      for (int r = 0; r < p.R; ++r) {
        float* dst = p.ws + ((static_cast<int64_t>(split) * p.R + r) * 128 + m) * 64;
#pragma unroll
        for (int c0 = 0; c0 < 64; c0 += 16) {
          uint32_t v[16];
          if (!empty) {
            tmem_ld16(taddr + r * 64 + c0, v);
            tmem_ld_wait();
          }
          if (m < p.cS) {
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(dst + c0 + j) =
                  empty ? make_float4(0.f, 0.f, 0.f, 0.f)
                        : make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                      __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0 && abort_smem && p.abort_flag) atomicExch(p.abort_flag, 1);
  if (warp == 2) tmem_dealloc(tmem_base, tmem_cols);
}

// dw[d0][d1][R][S] (+)= sum over splits of part[r'][m][s' * 8 + c]; m_is_d0: (d0, d1) = (m, c) else (c, m);
// flip: (r', s') = (R-1-r, S-1-s).
__global__ void toeplitz_wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, int d0, int d1, int R,
                                               int S, int splits, int m_is_d0, int flip, int accumulate) {
  const int total = d0 * d1 * R * S;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    int t = idx;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    t /= R;
    const int i1 = t % d1, i0 = t / d1;
    const int m = m_is_d0 ? i0 : i1, c = m_is_d0 ? i1 : i0;
    const int rr = flip ? R - 1 - r : r, ss = flip ? S - 1 - s : s;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp)
      acc += ws[((static_cast<int64_t>(sp) * R + rr) * 128 + m) * 64 + ss * 8 + c];
    dw[idx] = accumulate ? dw[idx] + acc : acc;
  }
}

static void tw_plan(const CdbAct* s_act, int* w_tiles, int* k_tiles, int* splits, int* k_per_split) {
  *w_tiles = ceil_div(s_act->w, kTwKT);
  *k_tiles = s_act->n * s_act->h * *w_tiles;
  int sp = sm_count();
  if (sp > *k_tiles / 4) sp = *k_tiles / 4;
  if (sp < 1) sp = 1;
  *k_per_split = ceil_div(*k_tiles, sp);
  *splits = ceil_div(*k_tiles, *k_per_split);
}

}  // namespace cdb

extern "C" size_t cdb_conv2d_toeplitz_wgrad_workspace(const CdbAct* s_act, int32_t r) {
  if (!s_act || r < 1) return 0;
  int w_tiles, k_tiles, splits, kps;
  tw_plan(s_act, &w_tiles, &k_tiles, &splits, &kps);
  return (size_t)splits * r * 128 * 64 * sizeof(float);
}

extern "C" int cdb_conv2d_toeplitz_wgrad(const CdbAct* s_act, const CdbAct* p_act, int32_t r, int32_t s, float* dw4,
                                         int32_t d0, int32_t d1, int32_t m_is_d0, int32_t flip, int32_t accumulate,
                                         void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(s_act && s_act->ptr && p_act && p_act->ptr && dw4 && workspace, CDB_ERR_BAD_DESC,
              "conv2d_toeplitz_wgrad: null argument");
  CDB_REQUIRE(p_act->dtype == CDB_BF16 && p_act->c == 8 && p_act->sw == 8 && p_act->sh == (int64_t)p_act->w * 8 &&
                  p_act->sn == (int64_t)p_act->h * p_act->w * 8 && (reinterpret_cast<uintptr_t>(p_act->ptr) & 15) == 0,
              CDB_ERR_UNSUPPORTED, "conv2d_toeplitz_wgrad: the shifted tensor must be a contiguous bf16 buffer of 8-channel pixels");
  CDB_REQUIRE(s_act->dtype == CDB_BF16 && s_act->c % 8 == 0 && s_act->sw % 8 == 0 && s_act->sh % 8 == 0 &&
                  s_act->sn % 8 == 0 && (reinterpret_cast<uintptr_t>(s_act->ptr) & 15) == 0,
              CDB_ERR_ALIGNMENT, "conv2d_toeplitz_wgrad: the iterated tensor must be bf16 NHWC with 16-byte pixels");
  CDB_REQUIRE(r >= 1 && r <= kTzMaxR && s >= 1 && s <= kTzTaps, CDB_ERR_UNSUPPORTED, "conv2d_toeplitz_wgrad: R, S <= 8");
  const int cS = m_is_d0 ? d0 : d1, cP = m_is_d0 ? d1 : d0;
  CDB_REQUIRE(cS >= 1 && cS <= 128 && cS <= s_act->c && cP >= 1 && cP <= 8, CDB_ERR_UNSUPPORTED,
              "conv2d_toeplitz_wgrad: <= 128 channels on the iterated side, <= 8 on the shifted side");
  CDB_REQUIRE(p_act->n == s_act->n && p_act->h >= s_act->h + r - 1 && p_act->w >= s_act->w + s - 1, CDB_ERR_BAD_DESC,
              "conv2d_toeplitz_wgrad: the shifted tensor must cover the iterated one plus the filter extent");
  TwParams p;
  memset(&p, 0, sizeof(p));
  tw_plan(s_act, &p.w_tiles, &p.k_tiles, &p.splits, &p.k_per_split);
  const size_t need = (size_t)p.splits * r * 128 * 64 * sizeof(float);
  CDB_REQUIRE(ws_bytes >= need, CDB_ERR_WORKSPACE, "conv2d_toeplitz_wgrad: workspace %zu < %zu", ws_bytes, need);
  p.R = r;
  p.cS = cS;
  p.Hs = s_act->h;
  p.Ws = s_act->w;
  p.n_img = s_act->n;
  p.pitchP = p_act->w;
  p.rowsP = p_act->h;
  p.total_px_P = (int64_t)p_act->n * p_act->h * p_act->w;
  p.P = static_cast<const __nv_bfloat16*>(p_act->ptr);
  p.ws = static_cast<float*>(workspace);
  p.abort_flag = device_abort_flag_ptr();
  CUtensorMap s_map;
  {
    uint64_t dims[4] = {(uint64_t)s_act->c, (uint64_t)s_act->w, (uint64_t)s_act->h, (uint64_t)s_act->n};
    uint64_t str[3] = {(uint64_t)s_act->sw * 2, (uint64_t)s_act->sh * 2, (uint64_t)s_act->sn * 2};
    uint32_t box[4] = {64u, (uint32_t)kTwKT, 1u, 1u};
    int rc = make_tmap(&s_map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, s_act->ptr, dims, str, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)kTwStages * kTwStageBytes + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    CDB_CUDA_OK(cudaFuncSetAttribute(toeplitz_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_done = true;
  }
  toeplitz_wgrad_kernel<<<p.splits, 256, smem, stream>>>(s_map, p);
  CDB_LAUNCH_OK();
  const int total = d0 * d1 * r * s;
  toeplitz_wgrad_finalize_kernel<<<ceil_div(total, 256), 256, 0, stream>>>(p.ws, dw4, d0, d1, r, s, p.splits, m_is_d0, flip,
                                                                            accumulate);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
