// K3 — convolution weight gradient on tcgen05, plus the weight (re)packing kernels.
//
//   out[tap][cS][cG] = sum_{n,p,q} S[n,p,q,cS] * G[n, p*stride + r*dil - pad_h, q*stride + s*dil - pad_w, cG]
// S is the "small" tensor whose pixels are iterated (dy for Conv2d, x for ConvTranspose2d), G the
// shifted one (x for Conv2d, dy for ConvTranspose2d); the result maps to W4[d0 = cS][d1 = cG][r][s]
// in both cases. The reduction runs over pixels, which is the SLOW axis of NHWC, so both operands
// reach the tensor core MN-major: a TMA box {64 ch, tile_w, tile_h, tile_n} of 64 pixels lands as
// 64 rows (k) x 128 B (64 channels, mn) in 128B-swizzled shared memory.
//   M (128 rows of TMEM) = channels of one operand, N (<= 256 columns) = channels of the other,
//   K = 64 pixels per pipeline stage, split-K over pixel tiles across CTAs (fp32 partials in a
//   caller-provided workspace, reduced deterministically by the finalize kernel).
#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

constexpr int kWgMaxTaps = 64;
constexpr int kWgMaxStages = 8;
constexpr int kChunkBytes = 64 * 128;  // 64 pixels x 64 channels bf16

struct WgTap {
  int16_t map, dh, dw, pad_;
};

struct WgParams {
  int32_t n_taps;
  int32_t tile_w, tile_h, tile_n;     // 64 pixels per K block
  int32_t tiles_w, tiles_h, tiles_n;  // pixel tiles over the S domain
  int32_t m_tiles, n_tiles, bn;       // bn multiple of 64
  int32_t splits, k_tiles_per_split;
  int32_t shift_on_a;                 // 1: the tap shift applies to the M-side operand
  int32_t stages;
  int32_t mpad, npad;
  int32_t tf32;                       // fp32 operands multiplied as TF32: 32 channels per 128-byte chunk row
  int32_t cpc;                        // channels per chunk: 64 (bf16) or 32 (tf32)
  int32_t prewait;                    // MMA thread waits for the next stage before the last instruction of this one
  int32_t tg, n_groups;               // taps per item (their shifted tiles sit side by side in N) and tap groups
  float* ws;                          // [splits][taps][mpad][npad]
  int* abort_flag;
  WgTap taps[kWgMaxTaps];
};

struct WgMaps {
  CUtensorMap fixed;     // S views: single map
  CUtensorMap shift[4];  // G parity views
};

// kPair: a CTA pair (cluster of 2, cta_group::2) computes a 256-row accumulator: each CTA stages the 128 M-side
// channels of ITS row tile and HALF of the N-side channels of every K block; the even CTA issues M = 256
// instructions for both; each CTA drains its own 128 x N accumulator.  Per SM a K block of a 256 x 256-channel
// layer is 16 + 16 KB instead of 16 + 32 KB for the same 4.2 MFLOP (the single-CTA kernel is bound by L2 -> SM
// operand traffic at 94 B/clk: tensor pipe 47 % active).
template <bool kPair>
__global__ void __launch_bounds__(256, 1)
wgrad_kernel(const __grid_constant__ WgMaps maps, const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t cpc = static_cast<uint32_t>(p.cpc);
  const uint32_t a_chunks = 128u / cpc;
  const uint32_t a_bytes = a_chunks * kChunkBytes;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;
  const uint32_t b_chunks = static_cast<uint32_t>(p.bn) / cpc / (kPair ? 2u : 1u);  // N-side chunks THIS CTA stages
  const uint32_t stage_bytes = a_bytes + static_cast<uint32_t>(p.tg) * b_chunks * kChunkBytes;
  const int m_units = kPair ? p.m_tiles / 2 : p.m_tiles;  // pair: two row tiles per work item
  const int total_items = p.n_groups * m_units * p.n_tiles * p.splits;
  const int item_first = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int item_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int k_tiles_total = p.tiles_n * p.tiles_h * p.tiles_w;

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
  const uint32_t tmem_base = tmem_base_smem;
  volatile int* abort_flag = &abort_smem;

  // item -> (split, tap group, m_tile, n_tile); split fastest so that the CTAs of one wave share operands.
  // A group holds p.tg consecutive taps (tg > 1 only when the shift is on the N side and there is one N tile):
  // the M-side tile is fetched once per K block for all of them, which is what the few-channel 7x7 layers need
  // (64-channel operands, 7 filter rows: the dy tile was re-read 7 times).
  auto decode = [&](int item, int& split, int& tap, int& mt, int& nt) {
    split = item % p.splits;
    item /= p.splits;
    nt = item % p.n_tiles;
    item /= p.n_tiles;
    mt = item % m_units;
    tap = item / m_units;
    if (kPair) mt = 2 * mt + static_cast<int>(rank);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int item = item_first; item < total_items && ok; item += item_step) {
        int split, tap, mt, nt;
        decode(item, split, tap, mt, nt);
        const int tap0 = tap * p.tg;
        const int ntap = min(p.tg, p.n_taps - tap0);
        const WgTap tp = p.taps[tap0];
        const CUtensorMap* amap = p.shift_on_a ? &maps.shift[tp.map] : &maps.fixed;
        const int a_dh = p.shift_on_a ? tp.dh : 0, a_dw = p.shift_on_a ? tp.dw : 0;
        const uint32_t item_bytes = a_bytes + static_cast<uint32_t>(ntap) * b_chunks * kChunkBytes;
        const int kt0 = split * p.k_tiles_per_split;
        int kt1 = kt0 + p.k_tiles_per_split;
        if (kt1 > k_tiles_total) kt1 = k_tiles_total;
        for (int kt = kt0; kt < kt1; ++kt) {
          int t = kt;
          const int tw = t % p.tiles_w;
          t /= p.tiles_w;
          const int th = t % p.tiles_h;
          const int tn = t / p.tiles_h;
          const int q0 = tw * p.tile_w, p0 = th * p.tile_h, img0 = tn * p.tile_n;
          if (!mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u, abort_flag)) {
            ok = false;
            break;
          }
          const uint32_t full_local = smem_u32(&bar_full[stage]);
          const uint32_t sa = smem_base + stage * stage_bytes;
          if (kPair) {
            // both CTAs' loads complete on the issuing (even) CTA's barrier, armed by it with the bytes of both
            const uint32_t full = mapa_shared(full_local, 0);
            if (rank == 0) mbar_arrive_expect_tx(full_local, 2u * item_bytes);
            for (uint32_t j = 0; j < a_chunks; ++j)
              tma_load_4d_pair(amap, full, sa + j * kChunkBytes, mt * 128 + j * cpc, q0 + a_dw, p0 + a_dh, img0);
            const WgTap tq = p.taps[tap0];
            const CUtensorMap* bmap = p.shift_on_a ? &maps.fixed : &maps.shift[tq.map];
            const int b_dh = p.shift_on_a ? 0 : tq.dh, b_dw = p.shift_on_a ? 0 : tq.dw;
            for (uint32_t j = 0; j < b_chunks; ++j)
              tma_load_4d_pair(bmap, full, sa + a_bytes + j * kChunkBytes, nt * p.bn + (rank * b_chunks + j) * cpc,
                               q0 + b_dw, p0 + b_dh, img0);
          } else {
            const uint32_t full = full_local;
            mbar_arrive_expect_tx(full, item_bytes);
            for (uint32_t j = 0; j < a_chunks; ++j)
              tma_load_4d(amap, full, sa + j * kChunkBytes, mt * 128 + j * cpc, q0 + a_dw, p0 + a_dh, img0);
            for (int tl = 0; tl < ntap; ++tl) {
              const WgTap tq = p.taps[tap0 + tl];
              const CUtensorMap* bmap = p.shift_on_a ? &maps.fixed : &maps.shift[tq.map];
              const int b_dh = p.shift_on_a ? 0 : tq.dh, b_dw = p.shift_on_a ? 0 : tq.dw;
              for (uint32_t j = 0; j < b_chunks; ++j)
                tma_load_4d(bmap, full, sa + a_bytes + (tl * b_chunks + j) * kChunkBytes, nt * p.bn + j * cpc, q0 + b_dw,
                            p0 + b_dh, img0);
            }
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // pair: the whole warp of the even CTA walks the loop and an elected lane issues (a cta_group::2 instruction
    // issued from a lone thread of a diverged warp takes 185-283 cycles, tools/mma_rate2.cu)
    if (kPair ? (rank == 0) : (lane == 0)) {
      const bool tf32 = p.tf32 != 0;
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      bool ok = true;
      for (int item = item_first; item < total_items && ok; item += item_step, ++local) {
        int split, tap, mt, nt;
        decode(item, split, tap, mt, nt);
        const int kt0 = split * p.k_tiles_per_split;
        int kt1 = kt0 + p.k_tiles_per_split;
        if (kt1 > k_tiles_total) kt1 = k_tiles_total;
        const int buf = local & 1;
        const uint32_t tphase = (local >> 1) & 1u;
        ok = mbar_wait(smem_u32(&bar_tempty[buf]), tphase ^ 1u, abort_flag);
        if (kPair) ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf) * 256u;
        const int ntap = min(p.tg, p.n_taps - tap * p.tg);
        const uint32_t idesc =
            make_idesc(tf32 ? 2u : 1u, 1u, 1u, kPair ? 256u : 128u, static_cast<uint32_t>(p.bn * ntap));
        // The tensor pipe accepts an instruction only shortly before it can start it, so whatever the issuing thread does
        // between the last instruction of one stage and the first of the next (commit, barrier round trip, fence,
        // descriptors: ~300 cycles) is idle tensor time per 512-cycle stage.  The wait for stage s + 1 is therefore made
        // BEFORE the last instruction of stage s is issued, while the earlier ones execute (p.prewait).
        bool ready = false;
        for (int kt = kt0; kt < kt1; ++kt) {
          if (!ready) {
            ok = mbar_wait(smem_u32(&bar_full[stage]), phase, abort_flag);
            if (kPair) ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc_fence_after();
          }
          ready = false;
          const uint32_t sa = smem_base + stage * stage_bytes;
          const int n_instr = tf32 ? 8 : 4;
          // MN-major operands.  bf16: 128B swizzle, LBO = distance between 64-channel chunks, SBO = 8 k-rows, 16 pixels
          // (= 2048 B) per instruction.  TF32: 128B swizzle with 32-byte atoms (4 k-rows x 128 B per atom), LBO = distance
          // between 32-channel chunks, SBO = 512 B between the two 4-row groups of one K = 8 instruction, 8 pixels
          // (= 1024 B) per instruction (pinned on B200 with tools/probe_tf32_wgrad.py).
          const uint64_t da = tf32 ? make_smem_desc(sa, kChunkBytes, 512, kLayoutSW128Base32)
                                   : make_smem_desc(sa, kChunkBytes, 1024, kLayoutSW128);
          const uint64_t db = tf32 ? make_smem_desc(sa + a_bytes, kChunkBytes, 512, kLayoutSW128Base32)
                                   : make_smem_desc(sa + a_bytes, kChunkBytes, 1024, kLayoutSW128);
          const uint32_t kstep = tf32 ? 64u : 128u;
          auto issue = [&](int k) {
            const uint32_t acc = (kt > kt0 || k > 0) ? 1u : 0u;
            if (kPair) {
              if (tf32) umma2_tf32(d_tmem, da + kstep * k, db + kstep * k, idesc, acc);
              else umma2_f16(d_tmem, da + kstep * k, db + kstep * k, idesc, acc);
            } else {
              if (tf32) umma_tf32(d_tmem, da + kstep * k, db + kstep * k, idesc, acc);
              else umma_f16(d_tmem, da + kstep * k, db + kstep * k, idesc, acc);
            }
          };
          if (!kPair || elect_one()) {
            for (int k = 0; k < n_instr - 1; ++k) issue(k);
          }
          if (kPair) __syncwarp();
          if (p.prewait && kt + 1 < kt1) {
            const int ns = stage + 1 == p.stages ? 0 : stage + 1;
            const uint32_t nphase = stage + 1 == p.stages ? phase ^ 1u : phase;
            ok = mbar_wait(smem_u32(&bar_full[ns]), nphase, abort_flag);
            if (kPair) ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc_fence_after();
            ready = true;
          }
          if (!kPair || elect_one()) {
            issue(n_instr - 1);
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
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const uint32_t tempty_remote = kPair ? mapa_shared(smem_u32(&bar_tempty[0]), 0) : 0u;
    int local = 0;
    for (int item = item_first; item < total_items; item += item_step, ++local) {
      int split, tap, mt, nt;
      decode(item, split, tap, mt, nt);
      const int kt0 = split * p.k_tiles_per_split;
      const bool empty_split = kt0 >= k_tiles_total;
      const int buf = local & 1;
      const uint32_t tphase = (local >> 1) & 1u;
      if (!mbar_wait(smem_u32(&bar_tfull[buf]), tphase, abort_flag)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) +
                             static_cast<uint32_t>(buf) * 256u;
      const int ntap = min(p.tg, p.n_taps - tap * p.tg);
      for (int c0 = 0; c0 < p.bn * ntap; c0 += 16) {
        const int tl = c0 / p.bn, cn = c0 - tl * p.bn;   // bn is a multiple of 64: a 16-column slab never straddles taps
        float* dst = p.ws + ((static_cast<int64_t>(split) * p.n_taps + tap * p.tg + tl) * p.mpad + mt * 128 + row) *
                                p.npad + nt * p.bn + cn - c0;
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 o;
          o.x = empty_split ? 0.f : __uint_as_float(v[j]);
          o.y = empty_split ? 0.f : __uint_as_float(v[j + 1]);
          o.z = empty_split ? 0.f : __uint_as_float(v[j + 2]);
          o.w = empty_split ? 0.f : __uint_as_float(v[j + 3]);
          *reinterpret_cast<float4*>(dst + c0 + j) = o;
        }
      }
      tc_fence_before();
      if (kPair) mbar_arrive_cluster(tempty_remote + static_cast<uint32_t>(buf) * 8u);
      else mbar_arrive(smem_u32(&bar_tempty[buf]));
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
// Row-sharing weight gradient (stride-1 layers over a materialised-padding input whose S-side rows are exactly one
// 64-pixel K block: the 3x3 256 -> 256 residual-block layers at 64 x 64).  The S taps of a filter row read
// OVERLAPPING windows of ONE input row: the row (W + S - 1 pixels, padded to 72) is staged once per K block and the taps
// address it through MN-major descriptors whose start is moved by s rows of 128 B — the trick of the flat forward
// kernel, on the K axis.  A CTA pair (cta_group::2, M = 256 = all dy channels) holds S accumulators of 128 x-channels
// (S * 128 <= 512 TMEM columns): per K block each SM pulls 16 KB of dy + 9 KB of x for S * 256 tensor cycles
// (33 B/clk) instead of 16 + 16 KB per 512 cycles for ONE tap (64 B/clk, above what an SM gets from L2).
// Work item = (filter row r, 128-channel half of x, split of the image rows); partials go to the same workspace
// layout as wgrad_kernel's, so wgrad_finalize_kernel is shared.
// -------------------------------------------------------------------------------------------------
struct WgRowParams {
  int32_t H, n_img, R, S, n_halves, splits, rows_per_split, stages, mpad, npad, n_taps;
  int32_t x_wp, x_rows_per_img;   // pitch (pixels) and pixel rows per image of the padded input
  float* ws;
  int* abort_flag;
};
struct WgRowMaps {
  CUtensorMap dy;  // box {64 ch, 64 px, 1, 1}
  CUtensorMap x;   // the padded input as a flat pixel matrix [n * Hp * Wp, C]: box {64 ch, 72 px}
};
constexpr int kRowABytes = 2 * kChunkBytes;     // 128 dy channels x 64 pixels
constexpr int kRowBBytes = 72 * 128;            // 64 x channels x 72 pixels
constexpr int kRowStageBytes = kRowABytes + kRowBBytes;

__global__ void __launch_bounds__(256, 1)
wgrad_rowshare_kernel(const __grid_constant__ WgRowMaps maps, const __grid_constant__ WgRowParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_empty[kWgMaxStages];
  __shared__ __align__(8) uint64_t bar_tfull;
  __shared__ __align__(8) uint64_t bar_tempty;
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = cluster_ctarank();
  const int total_items = p.R * p.n_halves * p.splits;
  const int item_first = static_cast<int>(blockIdx.x >> 1), item_step = static_cast<int>(gridDim.x >> 1);
  const int k_total = p.n_img * p.H;   // K blocks = image rows

  if (threadIdx.x == 0) {
    abort_smem = 0;
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_tfull), 1);
    mbar_init(smem_u32(&bar_tempty), 256);   // the epilogue threads of both CTAs
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(smem_u32(&tmem_base_smem), 512);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  volatile int* abort_flag = &abort_smem;

  auto decode = [&](int item, int& split, int& half, int& r) {
    split = item % p.splits;
    item /= p.splits;
    half = item % p.n_halves;
    r = item / p.n_halves;
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      for (int item = item_first; item < total_items && ok; item += item_step) {
        int split, half, r;
        decode(item, split, half, r);
        const int kt0 = split * p.rows_per_split;
        const int kt1 = min(k_total, kt0 + p.rows_per_split);
        for (int kt = kt0; kt < kt1; ++kt) {
          const int n = kt / p.H, h = kt - n * p.H;
          if (!mbar_wait(smem_u32(&bar_empty[stage]), phase ^ 1u, abort_flag)) {
            ok = false;
            break;
          }
          const uint32_t full_local = smem_u32(&bar_full[stage]);
          const uint32_t full0 = mapa_shared(full_local, 0);
          const uint32_t sa = smem_base + stage * kRowStageBytes;
          if (rank == 0) mbar_arrive_expect_tx(full_local, 2u * kRowStageBytes);
          // this CTA's 128 dy channels (two 64-channel chunks) of image row (n, h)
          tma_load_4d_pair(&maps.dy, full0, sa, static_cast<int>(rank) * 128, 0, h, n);
          tma_load_4d_pair(&maps.dy, full0, sa + kChunkBytes, static_cast<int>(rank) * 128 + 64, 0, h, n);
          // this CTA's 64 of the item's 128 x channels, padded-input row h + r, pixels 0 .. 71 (beyond the row: zero fill)
          // (the 72-pixel box runs past the 66-pixel row into the next one: those rows are never addressed)
          tma_load_2d_pair(&maps.x, full0, sa + kRowABytes, half * 128 + static_cast<int>(rank) * 64,
                           n * p.x_rows_per_img + (h + r) * p.x_wp);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {   // whole warp; the elected lane issues (see wgrad_kernel)
      const uint32_t idesc = make_idesc(1u, 1u, 1u, 256u, 128u);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      bool ok = true;
      for (int item = item_first; item < total_items && ok; item += item_step, ++local) {
        int split, half, r;
        decode(item, split, half, r);
        const int kt0 = split * p.rows_per_split;
        const int kt1 = min(k_total, kt0 + p.rows_per_split);
        ok = mbar_wait(smem_u32(&bar_tempty), (local & 1u) ^ 1u, abort_flag);
        ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        tc_fence_after();
        for (int kt = kt0; kt < kt1; ++kt) {
          ok = mbar_wait(smem_u32(&bar_full[stage]), phase, abort_flag);
          ok = __all_sync(0xffffffffu, ok);
          if (!ok) break;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * kRowStageBytes;
          if (elect_one()) {
            // MN-major, 128B swizzle: A = 128 channels (LBO = chunk distance) x 16 pixels per instruction (2048 B step);
            // B = 64 channels of this CTA, start moved by s pixel rows of 128 B for tap s
            const uint64_t da = make_smem_desc(sa, kChunkBytes, 1024, kLayoutSW128);
            for (int s = 0; s < p.S; ++s) {
              const uint64_t db = make_smem_desc(sa + kRowABytes + static_cast<uint32_t>(s) * 128u, kChunkBytes, 1024,
                                                 kLayoutSW128);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma2_f16(tmem_base + static_cast<uint32_t>(s) * 128u, da + 128u * k, db + 128u * k, idesc,
                          (kt > kt0 || k > 0) ? 1u : 0u);
            }
            umma2_commit(smem_u32(&bar_empty[stage]));
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        if (ok) {
          if (elect_one()) umma2_commit(smem_u32(&bar_tfull));
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int row = ew * 32 + lane;
    const uint32_t tempty_remote = mapa_shared(smem_u32(&bar_tempty), 0);
    int local = 0;
    for (int item = item_first; item < total_items; item += item_step, ++local) {
      int split, half, r;
      decode(item, split, half, r);
      if (!mbar_wait(smem_u32(&bar_tfull), local & 1u, abort_flag)) break;
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
      for (int c0 = 0; c0 < p.S * 128; c0 += 16) {
        const int s = c0 >> 7, cn = c0 & 127;
        float* dst = p.ws + ((static_cast<int64_t>(split) * p.n_taps + r * p.S + s) * p.mpad +
                             static_cast<int>(rank) * 128 + row) * p.npad + half * 128 + cn;
        uint32_t v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
      }
      tc_fence_before();
      mbar_arrive_cluster(tempty_remote);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0 && abort_smem && p.abort_flag) atomicExch(p.abort_flag, 1);
  if (warp == 2) tmem_dealloc_pair(tmem_base, 512);
}

// Sums the split-K partials and scatters them into the fp32 filter gradient W4[d0][d1][R][S].
// m_is_d0: the M side of the GEMM holds d0 (else d1). rowpack: tap = r, the d1-side index is
// s*rowpack + ch.
__global__ void wgrad_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dw, int d0, int d1,
                                      int R, int S, int n_taps, int splits, int mpad, int npad,
                                      int m_is_d0, int rowpack, int accumulate) {
  const int64_t total = static_cast<int64_t>(d0) * d1 * R * S;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t t = idx;
    const int s = t % S;
    t /= S;
    const int r = t % R;
    t /= R;
    const int i1 = t % d1;
    const int i0 = t / d1;
    int tap, cS = i0, cG = i1;
    if (rowpack) {
      tap = r;
      cG = s * rowpack + i1;
    } else {
      tap = r * S + s;
    }
    const int m = m_is_d0 ? cS : cG;
    const int n = m_is_d0 ? cG : cS;
    float acc = 0.f;
    for (int sp = 0; sp < splits; ++sp)
      acc += ws[((static_cast<int64_t>(sp) * n_taps + tap) * mpad + m) * npad + n];
    dw[idx] = accumulate ? dw[idx] + acc : acc;
  }
}

// Tiled form of wgrad_finalize_kernel for rowpack == 0: a block owns ONE row m of the GEMM result, 32 consecutive n
// and all taps; thread (n, tap group) sums the splits of its taps.  The workspace is read along its contiguous axis n
// (the generic kernel walks dw in order and so reads the workspace at a stride of mpad * npad floats per element: one
// 32-byte sector per 4-byte value, 15.6 us for a 256 x 256 x 3 x 3 gradient of 12 splits); the sums go through shared
// memory and dw is written in runs of 32 * R * S (M side = d0) or R * S (M side = d1) contiguous floats.
constexpr int kFinN = 32;
__global__ void __launch_bounds__(256)
wgrad_finalize_tiled_kernel(const float* __restrict__ ws, float* __restrict__ dw, int d0, int d1, int RS, int splits,
                            int mpad, int npad, int cN, int m_is_d0, int accumulate) {
  extern __shared__ float fin_tile[];              // [kFinN][RSp], RSp odd
  const int RSp = RS | 1;
  const int m = blockIdx.x, n0 = blockIdx.y * kFinN;
  const int n_l = threadIdx.x % kFinN, tg = threadIdx.x / kFinN;   // 8 tap groups
  const int n = n0 + n_l;
  if (n < cN) {
    const int64_t sstride = static_cast<int64_t>(RS) * mpad * npad;
    for (int tap = tg; tap < RS; tap += 256 / kFinN) {
      const float* src = ws + (static_cast<int64_t>(tap) * mpad + m) * npad + n;
      float acc = 0.f;
#pragma unroll 4
      for (int sp = 0; sp < splits; ++sp) acc += src[sp * sstride];
      fin_tile[n_l * RSp + tap] = acc;
    }
  }
  __syncthreads();
  const int cnt = min(kFinN, cN - n0) * RS;
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const int nl = i / RS, tp = i - nl * RS;
    const int64_t idx = m_is_d0 ? (static_cast<int64_t>(m) * d1 + n0 + nl) * RS + tp
                                : (static_cast<int64_t>(n0 + nl) * d1 + m) * RS + tp;
    const float v = fin_tile[nl * RSp + tp];
    dw[idx] = accumulate ? dw[idx] + v : v;
  }
}

static void launch_wgrad_finalize(const float* ws, float* dw4, int d0, int d1, int R, int S, int n_taps, int splits,
                                  int mpad, int npad, int m_is_d0, int rowpack, int accumulate, cudaStream_t stream) {
  static const bool tiled = !(getenv("CDB_WGRAD_FINALIZE_TILED") && atoi(getenv("CDB_WGRAD_FINALIZE_TILED")) == 0);
  // M side = d1 (the strided / transposed generator layers): writes in runs of R * S floats, measured 2-5 us slower
  // than the generic kernel; M side = d0: 2-5 us faster on the PatchGAN layers, equal on the residual-block layers
  if (tiled && m_is_d0 && rowpack == 0 && R * S <= 64 && n_taps == R * S) {
    const int cM = m_is_d0 ? d0 : d1, cN = m_is_d0 ? d1 : d0;
    dim3 grid(cM, ceil_div(cN, kFinN));
    const size_t smem = sizeof(float) * kFinN * ((R * S) | 1);
    wgrad_finalize_tiled_kernel<<<grid, 256, smem, stream>>>(ws, dw4, d0, d1, R * S, splits, mpad, npad, cN, m_is_d0,
                                                           accumulate);
    return;
  }
  const int64_t total = (int64_t)d0 * d1 * R * S;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  wgrad_finalize_kernel<<<blocks, 256, 0, stream>>>(ws, dw4, d0, d1, R, S, n_taps, splits, mpad, npad, m_is_d0, rowpack,
                                                    accumulate);
}

// fp32 W4[d0][d1][R][S] -> bf16 packed[rows_pad][taps*kpad] (see cdb_pack_conv_weight).
template <typename OutT>
__global__ void pack_weight_kernel(const float* __restrict__ w, OutT* __restrict__ out, int d0,
                                   int d1, int R, int S, int rows_are_dim0, int rowpack, int rows,
                                   int kdim, int rows_pad, int kpad, int n_taps) {
  const int64_t ktotal = static_cast<int64_t>(n_taps) * kpad;
  const int64_t total = static_cast<int64_t>(rows_pad) * ktotal;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int row = idx / ktotal;
    const int kk = idx % ktotal;
    const int tap = kk / kpad;
    int k = kk % kpad;
    int r, s;
    bool ok = row < rows;
    if (rowpack) {
      r = tap;
      s = k / rowpack;
      k = k % rowpack;
      ok = ok && s < S && k < kdim;
    } else {
      r = tap / S;
      s = tap % S;
      ok = ok && k < kdim;
    }
    float v = 0.f;
    if (ok) {
      const int i0 = rows_are_dim0 ? row : k;
      const int i1 = rows_are_dim0 ? k : row;
      v = w[((static_cast<int64_t>(i0) * d1 + i1) * R + r) * S + s];
    }
    if constexpr (sizeof(OutT) == 4) out[idx] = round_tf32(v);
    else out[idx] = __float2bfloat16(v);
  }
}

__global__ void round_tf32_kernel(float* __restrict__ buf, int64_t numel) {
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < numel;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    buf[idx] = round_tf32(buf[idx]);
}

// Multi-tensor packing: every cached bf16 operand of an optimizer's filters re-packed by ONE launch (the filter
// pointers and geometries travel in the kernel parameter block; graph-safe). Same element mapping as
// pack_weight_kernel; block b finds its (entry, chunk) by binary search over cumulative chunk counts.
constexpr int kPackMaxEntries = 320;
constexpr int kPackChunk = 8192;
struct PackEntryDev {
  const float* w;
  __nv_bfloat16* out;
  int32_t d0, d1, R, S, rows_are_dim0, rowpack, rows, kdim, rows_pad, kpad, n_taps, pad_;
};
struct PackMultiParams {
  int32_t n_entries, pad_;
  PackEntryDev e[kPackMaxEntries];
  int32_t cum[kPackMaxEntries + 1];
};

__global__ void __launch_bounds__(256) pack_weight_multi_kernel(const __grid_constant__ PackMultiParams q) {
  int lo = 0, hi = q.n_entries;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (q.cum[mid] <= static_cast<int>(blockIdx.x)) lo = mid;
    else hi = mid;
  }
  const PackEntryDev& en = q.e[lo];
  const int64_t ktotal = static_cast<int64_t>(en.n_taps) * en.kpad;
  const int64_t total = static_cast<int64_t>(en.rows_pad) * ktotal;
  const int64_t begin = static_cast<int64_t>(blockIdx.x - q.cum[lo]) * kPackChunk;
  const int64_t end = begin + kPackChunk < total ? begin + kPackChunk : total;
  for (int64_t idx = begin + threadIdx.x; idx < end; idx += 256) {
    const int row = static_cast<int>(idx / ktotal);
    const int kk = static_cast<int>(idx % ktotal);
    const int tap = kk / en.kpad;
    int k = kk % en.kpad;
    int r, sc;
    bool ok = row < en.rows;
    if (en.rowpack) {
      r = tap;
      sc = k / en.rowpack;
      k = k % en.rowpack;
      ok = ok && sc < en.S && k < en.kdim;
    } else {
      r = tap / en.S;
      sc = tap % en.S;
      ok = ok && k < en.kdim;
    }
    float v = 0.f;
    if (ok) {
      const int i0 = en.rows_are_dim0 ? row : k;
      const int i1 = en.rows_are_dim0 ? k : row;
      v = en.w[((static_cast<int64_t>(i0) * en.d1 + i1) * en.R + r) * en.S + sc];
    }
    en.out[idx] = __float2bfloat16(v);
  }
}

struct WgPlan {
  int m_is_s;  // M side = S tensor
  int tg, n_groups;
  int cM, cN, mpad, npad, bn, m_tiles, n_tiles;
  int tile_w, tile_h, tile_n, tiles_w, tiles_h, tiles_n, k_tiles, splits, k_per_split, n_taps;
  int pair;  // CTA pairs (cta_group::2): two row tiles per work item, the N-side operand split between the CTAs
  int rowshare;  // wgrad_rowshare_kernel: the S taps of a filter row share one staged input row (see the kernel)
};

static int plan_wgrad(const CdbConvGeom* g, const CdbAct* s_act, const CdbAct* g_act, WgPlan* pl) {
  // TF32: fp32 operands take twice the shared memory per channel -> N tile of at most 128 channels (3 stages)
  const bool tf32 = s_act->dtype == CDB_F32;
  const int cS = s_act->c;
  const int cG = g->rowpack ? 64 : g_act->c;
  auto cost = [](int cm, int cn) { return (int64_t)round_up(cm, 128) * round_up(cn, 64); };
  pl->m_is_s = cost(cS, cG) <= cost(cG, cS) ? 1 : 0;
  pl->cM = pl->m_is_s ? cS : cG;
  pl->cN = pl->m_is_s ? cG : cS;
  pl->mpad = round_up(pl->cM, 128);
  pl->m_tiles = pl->mpad / 128;
  const int n64 = round_up(pl->cN, 64);
  const int bn_max = tf32 ? 128 : 256;
  pl->bn = n64 < bn_max ? n64 : bn_max;
  pl->n_tiles = ceil_div(n64, pl->bn);
  pl->npad = pl->n_tiles * pl->bn;
  // 64-pixel K tiles over the S domain
  int w = 1;
  while (w < s_act->w && w < 64) w <<= 1;
  int h = 1;
  while (h < s_act->h && h * w < 64) h <<= 1;
  pl->tile_w = w;
  pl->tile_h = h;
  pl->tile_n = 64 / (w * h);
  pl->tiles_w = ceil_div(s_act->w, w);
  pl->tiles_h = ceil_div(s_act->h, h);
  pl->tiles_n = ceil_div(s_act->n, pl->tile_n);
  pl->k_tiles = pl->tiles_w * pl->tiles_h * pl->tiles_n;
  pl->n_taps = g->rowpack ? g->r : g->r * g->s;
  // tap grouping: the shifted operand is on the N side, one N tile, and several of its tiles fit in 256 columns
  pl->tg = 1;
  if (pl->m_is_s && pl->n_tiles == 1 && !tf32 && !getenv("CDB_WGRAD_NO_TAP_GROUPS")) {
    int tg = 256 / pl->bn;
    if (tg > 4) tg = 4;
    if (tg > pl->n_taps) tg = pl->n_taps;
    if (tg >= 1) pl->tg = tg;
  }
  pl->n_groups = ceil_div(pl->n_taps, pl->tg);
  static const int pair_env = getenv("CDB_WGRAD_PAIR") ? atoi(getenv("CDB_WGRAD_PAIR")) : 1;
  pl->pair = (pair_env && pl->tg == 1 && pl->m_tiles % 2 == 0 && pl->bn % 128 == 0 && !g->rowpack) ? 1 : 0;
  const int items = pl->n_groups * (pl->pair ? pl->m_tiles / 2 : pl->m_tiles) * pl->n_tiles;
  static const int waves = getenv("CDB_WGRAD_WAVES") ? atoi(getenv("CDB_WGRAD_WAVES")) : 1;
  int splits = (waves * (pl->pair ? sm_count() / 2 : sm_count())) / (items > 0 ? items : 1);
  if (splits < 1) splits = 1;
  int max_splits = pl->k_tiles / 8;
  if (max_splits < 1) max_splits = 1;
  if (splits > max_splits) splits = max_splits;
  pl->k_per_split = ceil_div(pl->k_tiles, splits);
  pl->splits = ceil_div(pl->k_tiles, pl->k_per_split);
  // row-sharing kernel: bf16, stride-1 Conv2d over a materialised-padding input, the dy rows are exactly one 64-pixel K
  // block, M = the 256 dy channels (one CTA pair), x channels in halves of 128, S accumulators of 128 columns in TMEM
  static const int rowshare_env = getenv("CDB_WGRAD_ROWSHARE") ? atoi(getenv("CDB_WGRAD_ROWSHARE")) : 1;
  pl->rowshare = (rowshare_env && pl->pair && !tf32 && !g->transposed && g->stride == 1 && g->dil == 1 && g->pad_h == 0 &&
                  g->pad_w == 0 && pl->m_is_s && s_act->w == 64 && pl->tile_w == 64 && pl->tile_h == 1 && pl->cM == 256 &&
                  pl->cN % 128 == 0 && g->s * 128 <= 512 && g->s <= 8 && g_act->w >= s_act->w + g->s - 1 &&
                  g_act->h >= s_act->h + g->r - 1 && g_act->sh == (int64_t)g_act->w * g_act->sw &&
                  g_act->sn == (int64_t)g_act->h * g_act->w * g_act->sw)
                     ? 1
                     : 0;
  if (pl->rowshare) {
    const int k_rows = s_act->n * s_act->h;
    const int row_items = g->r * (pl->cN / 128);
    int rs = (sm_count() / 2) / (row_items > 0 ? row_items : 1);
    if (rs < 1) rs = 1;
    if (rs > k_rows / 8) rs = k_rows / 8 > 0 ? k_rows / 8 : 1;
    pl->k_per_split = ceil_div(k_rows, rs);
    pl->splits = ceil_div(k_rows, pl->k_per_split);
  }
  return CDB_OK;
}

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_pack_conv_weight(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s,
                                    int32_t rows_are_dim0, int32_t rowpack, void* out,
                                    cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(w4 && out, CDB_ERR_BAD_DESC, "pack_conv_weight: null argument");
  const int rows = rows_are_dim0 ? d0 : d1;
  const int kdim = rows_are_dim0 ? d1 : d0;
  const int rows_pad = round_up(rows, 16);
  int kpad, n_taps;
  if (rowpack) {
    CDB_REQUIRE(s * rowpack <= 64 && kdim <= rowpack, CDB_ERR_BAD_DESC, "pack_conv_weight: rowpack geometry");
    kpad = 64;
    n_taps = r;
  } else {
    kpad = round_up(kdim, 64);
    n_taps = r * s;
  }
  const int64_t total = (int64_t)rows_pad * n_taps * kpad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(w4, static_cast<__nv_bfloat16*>(out), d0, d1, r, s,
                                                                rows_are_dim0, rowpack, rows, kdim, rows_pad, kpad,
                                                                n_taps);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_pack_conv_weight_tf32(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s,
                                         int32_t rows_are_dim0, void* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(w4 && out, CDB_ERR_BAD_DESC, "pack_conv_weight_tf32: null argument");
  const int rows = rows_are_dim0 ? d0 : d1;
  const int kdim = rows_are_dim0 ? d1 : d0;
  const int rows_pad = round_up(rows, 16);
  const int kpad = round_up(kdim, 32), n_taps = r * s;
  const int64_t total = (int64_t)rows_pad * n_taps * kpad;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weight_kernel<float><<<blocks, 256, 0, stream>>>(w4, static_cast<float*>(out), d0, d1, r, s, rows_are_dim0, 0,
                                                        rows, kdim, rows_pad, kpad, n_taps);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_round_tf32(float* buf, int64_t numel, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(buf || numel == 0, CDB_ERR_BAD_DESC, "round_tf32: null argument");
  if (numel <= 0) return CDB_OK;
  int64_t blocks = (numel + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  round_tf32_kernel<<<(int)blocks, 256, 0, stream>>>(buf, numel);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_pack_conv_weights_multi(const CdbPackEntry* entries_host, int32_t n_entries, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(entries_host && n_entries > 0, CDB_ERR_BAD_DESC, "pack_conv_weights_multi: bad argument");
  static thread_local PackMultiParams q;
  for (int base = 0; base < n_entries; base += kPackMaxEntries) {
    const int n = n_entries - base < kPackMaxEntries ? n_entries - base : kPackMaxEntries;
    q.n_entries = n;
    int blocks = 0;
    for (int i = 0; i < n; ++i) {
      const CdbPackEntry& h = entries_host[base + i];
      CDB_REQUIRE(h.w4 && h.out, CDB_ERR_BAD_DESC, "pack_conv_weights_multi: null pointer in entry %d", base + i);
      PackEntryDev& e = q.e[i];
      e.w = h.w4;
      e.out = static_cast<__nv_bfloat16*>(h.out);
      e.d0 = h.d0;
      e.d1 = h.d1;
      e.R = h.r;
      e.S = h.s;
      e.rows_are_dim0 = h.rows_are_dim0;
      e.rowpack = h.rowpack;
      e.rows = h.rows_are_dim0 ? h.d0 : h.d1;
      e.kdim = h.rows_are_dim0 ? h.d1 : h.d0;
      e.rows_pad = round_up(e.rows, 16);
      if (h.rowpack) {
        CDB_REQUIRE(h.s * h.rowpack <= 64 && e.kdim <= h.rowpack, CDB_ERR_BAD_DESC,
                    "pack_conv_weights_multi: rowpack geometry in entry %d", base + i);
        e.kpad = 64;
        e.n_taps = h.r;
      } else {
        e.kpad = round_up(e.kdim, 64);
        e.n_taps = h.r * h.s;
      }
      q.cum[i] = blocks;
      const int64_t total = (int64_t)e.rows_pad * e.n_taps * e.kpad;
      blocks += (int)((total + kPackChunk - 1) / kPackChunk);
    }
    q.cum[n] = blocks;
    pack_weight_multi_kernel<<<blocks, 256, 0, stream>>>(q);
    CDB_LAUNCH_OK();
  }
  return CDB_OK;
}

static void wgrad_roles(const CdbConvGeom* g, const CdbAct* x, const CdbAct* dy, const CdbAct** s_act,
                        const CdbAct** g_act) {
  if (g->transposed) {
    *s_act = x;
    *g_act = dy;
  } else {
    *s_act = dy;
    *g_act = x;
  }
}

extern "C" size_t cdb_conv2d_wgrad_workspace(const CdbConvGeom* g, const CdbAct* x, const CdbAct* dy) {
  if (!g || !x || !dy) return 0;
  const CdbAct *s_act, *g_act;
  wgrad_roles(g, x, dy, &s_act, &g_act);
  WgPlan pl;
  plan_wgrad(g, s_act, g_act, &pl);
  return (size_t)pl.splits * pl.n_taps * pl.mpad * pl.npad * sizeof(float);
}

extern "C" int cdb_conv2d_wgrad(const CdbConvGeom* g, const CdbAct* x, const CdbAct* dy, float* dw4,
                                int32_t d0, int32_t d1, int32_t accumulate, void* workspace,
                                size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(g && x && dy && dw4 && workspace, CDB_ERR_BAD_DESC, "conv2d_wgrad: null argument");
  CDB_REQUIRE(x->dtype == dy->dtype && (x->dtype == CDB_BF16 || x->dtype == CDB_F32), CDB_ERR_UNSUPPORTED,
              "conv2d_wgrad: x and dy must both be bf16 or both fp32 (TF32)");
  const bool tf32 = x->dtype == CDB_F32;
  const uint64_t esz = tf32 ? 4 : 2;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const int cpc = tf32 ? 32 : 64;
  CDB_REQUIRE(!(tf32 && g->rowpack), CDB_ERR_UNSUPPORTED, "conv2d_wgrad: rowpack is bf16 only");
  {
    const int q = tf32 ? 4 : 8;
    const CdbAct* both[2] = {x, dy};
    for (const CdbAct* t : both)
      CDB_REQUIRE(t->ptr && t->c % q == 0 && t->sw % q == 0 && t->sh % q == 0 && t->sn % q == 0 &&
                      (reinterpret_cast<uintptr_t>(t->ptr) & 15) == 0,
                  CDB_ERR_ALIGNMENT, "conv2d_wgrad: channels/strides must be multiples of 16 bytes and ptr 16B aligned");
  }
  CDB_REQUIRE(g->stride == 1 || g->stride == 2, CDB_ERR_UNSUPPORTED, "conv2d_wgrad: stride %d", g->stride);
  // rowpack applies to the SHIFTED tensor: x for Conv2d, dy for the transposed form (used for layers with
  // very few output channels: dy then carries 8 channels per pixel and one K block covers a filter row).
  const CdbAct *s_act, *g_act;
  wgrad_roles(g, x, dy, &s_act, &g_act);
  WgPlan pl;
  plan_wgrad(g, s_act, g_act, &pl);
  const size_t need = (size_t)pl.splits * pl.n_taps * pl.mpad * pl.npad * sizeof(float);
  CDB_REQUIRE(ws_bytes >= need, CDB_ERR_WORKSPACE, "conv2d_wgrad: workspace %zu < %zu", ws_bytes, need);
  CDB_REQUIRE(pl.n_taps <= kWgMaxTaps, CDB_ERR_UNSUPPORTED, "conv2d_wgrad: too many taps");
  const int st = g->stride;

  WgMaps maps;
  memset(&maps, 0, sizeof(maps));
  const uint32_t box[4] = {(uint32_t)cpc, (uint32_t)pl.tile_w, (uint32_t)pl.tile_h, (uint32_t)pl.tile_n};
  {
    uint64_t dims[4] = {(uint64_t)s_act->c, (uint64_t)s_act->w, (uint64_t)s_act->h, (uint64_t)s_act->n};
    uint64_t str[3] = {(uint64_t)s_act->sw * esz, (uint64_t)s_act->sh * esz, (uint64_t)s_act->sn * esz};
    int rc = make_tmap_swz(&maps.fixed, dt, 4, s_act->ptr, dims, str, box, tf32 ? 1 : 0);
    if (rc) return rc;
  }
  WgParams prm;
  memset(&prm, 0, sizeof(prm));
  if (g->rowpack) {
    const int span = 64 / g->rowpack;
    for (int a = 0; a < 4; ++a) {
      const int aa = a < st ? a : 0;
      uint64_t dims[4] = {64, (uint64_t)((g_act->w - span) / st + 1), (uint64_t)((g_act->h - aa + st - 1) / st),
                          (uint64_t)g_act->n};
      uint64_t str[3] = {(uint64_t)g_act->sw * st * 2, (uint64_t)g_act->sh * st * 2, (uint64_t)g_act->sn * 2};
      int rc = make_tmap(&maps.shift[a], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                         static_cast<__nv_bfloat16*>(g_act->ptr) + aa * g_act->sh, dims, str, box);
      if (rc) return rc;
    }
    for (int r = 0; r < g->r; ++r) {
      const int ih = r * g->dil - g->pad_h;
      const int a = pos_mod(ih, st);
      prm.taps[r].map = (int16_t)a;
      prm.taps[r].dh = (int16_t)((ih - a) / st);
      prm.taps[r].dw = 0;
    }
  } else {
    for (int i = 0; i < 4; ++i) {
      const int a = (i < st * st) ? i / st : 0, b = (i < st * st) ? i % st : 0;
      uint64_t dims[4] = {(uint64_t)g_act->c, (uint64_t)((g_act->w - b + st - 1) / st),
                          (uint64_t)((g_act->h - a + st - 1) / st), (uint64_t)g_act->n};
      for (int d = 0; d < 4; ++d)
        if (dims[d] == 0) dims[d] = 1;
      uint64_t str[3] = {(uint64_t)g_act->sw * st * esz, (uint64_t)g_act->sh * st * esz, (uint64_t)g_act->sn * esz};
      int rc = make_tmap_swz(&maps.shift[i], dt, 4,
                             static_cast<char*>(g_act->ptr) + (a * g_act->sh + b * g_act->sw) * (int64_t)esz, dims,
                             str, box, tf32 ? 1 : 0);
      if (rc) return rc;
    }
    for (int r = 0; r < g->r; ++r)
      for (int s = 0; s < g->s; ++s) {
        const int ih = r * g->dil - g->pad_h, iw = s * g->dil - g->pad_w;
        const int a = pos_mod(ih, st), b = pos_mod(iw, st);
        WgTap& t = prm.taps[r * g->s + s];
        t.map = (int16_t)(a * st + b);
        t.dh = (int16_t)((ih - a) / st);
        t.dw = (int16_t)((iw - b) / st);
      }
  }
  prm.n_taps = pl.n_taps;
  prm.tile_w = pl.tile_w;
  prm.tile_h = pl.tile_h;
  prm.tile_n = pl.tile_n;
  prm.tiles_w = pl.tiles_w;
  prm.tiles_h = pl.tiles_h;
  prm.tiles_n = pl.tiles_n;
  prm.m_tiles = pl.m_tiles;
  prm.n_tiles = pl.n_tiles;
  prm.bn = pl.bn;
  prm.splits = pl.splits;
  prm.k_tiles_per_split = pl.k_per_split;
  prm.shift_on_a = pl.m_is_s ? 0 : 1;
  prm.mpad = pl.mpad;
  prm.npad = pl.npad;
  prm.tf32 = tf32 ? 1 : 0;
  prm.cpc = cpc;
  prm.tg = pl.tg;
  prm.n_groups = pl.n_groups;
  static const int prewait_env = getenv("CDB_MMA_PREWAIT") ? atoi(getenv("CDB_MMA_PREWAIT")) : 0;  // measured neutral (same-box A/B, tools/ab_wgrad.sh): the kernel waits for data, not for its issuing thread
  prm.prewait = prewait_env;
  prm.ws = static_cast<float*>(workspace);
  prm.abort_flag = device_abort_flag_ptr();
  if (pl.rowshare) {
    WgRowMaps rmaps;
    memset(&rmaps, 0, sizeof(rmaps));
    {
      uint64_t dims[4] = {(uint64_t)s_act->c, (uint64_t)s_act->w, (uint64_t)s_act->h, (uint64_t)s_act->n};
      uint64_t str[3] = {(uint64_t)s_act->sw * 2, (uint64_t)s_act->sh * 2, (uint64_t)s_act->sn * 2};
      const uint32_t rbox[4] = {64u, 64u, 1u, 1u};
      int rc = make_tmap(&rmaps.dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, s_act->ptr, dims, str, rbox);
      if (rc) return rc;
    }
    {
      uint64_t dims[2] = {(uint64_t)g_act->c, (uint64_t)g_act->n * g_act->h * g_act->w};
      uint64_t str[1] = {(uint64_t)g_act->sw * 2};
      const uint32_t rbox[2] = {64u, 72u};
      int rc = make_tmap(&rmaps.x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g_act->ptr, dims, str, rbox);
      if (rc) return rc;
    }
    WgRowParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.H = s_act->h;
    rp.n_img = s_act->n;
    rp.R = g->r;
    rp.S = g->s;
    rp.n_halves = pl.cN / 128;
    rp.splits = pl.splits;
    rp.rows_per_split = pl.k_per_split;
    rp.mpad = pl.mpad;
    rp.npad = pl.npad;
    rp.n_taps = pl.n_taps;
    rp.x_wp = g_act->w;
    rp.x_rows_per_img = g_act->h * g_act->w;
    rp.ws = static_cast<float*>(workspace);
    rp.abort_flag = device_abort_flag_ptr();
    int rstages = (200 * 1024) / kRowStageBytes;
    if (rstages > kWgMaxStages) rstages = kWgMaxStages;
    rp.stages = rstages;
    const size_t rsmem = (size_t)rstages * kRowStageBytes + 1024;
    static size_t rsmem_attr = 0;
    if (rsmem > rsmem_attr) {
      CDB_CUDA_OK(cudaFuncSetAttribute(wgrad_rowshare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
      rsmem_attr = rsmem;
    }
    const int items = rp.R * rp.n_halves * rp.splits;
    const int clusters = items < sm_count() / 2 ? items : sm_count() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = rsmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CDB_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_rowshare_kernel, rmaps, rp));
    CDB_LAUNCH_OK();
    launch_wgrad_finalize(static_cast<const float*>(workspace), dw4, d0, d1, g->r, g->s, pl.n_taps, pl.splits, pl.mpad,
                          pl.npad, pl.m_is_s, 0, accumulate, stream);
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  const int stage_bytes = (128 / cpc) * kChunkBytes + pl.tg * (pl.bn / cpc) * kChunkBytes / (pl.pair ? 2 : 1);
  // shared-memory budget of the operand ring.  The weight gradients run on a companion stream next to the norm /
  // activation backward kernels of the main stream (engine._SideStream): a ring that leaves room for one of their CTAs
  // (64 KB for the cluster-fused InstanceNorm backward) lets the two kinds of kernels share an SM.
  static const int budget_kb = getenv("CDB_WGRAD_SMEM_KB") ? atoi(getenv("CDB_WGRAD_SMEM_KB")) : 200;
  int stages = (budget_kb * 1024) / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > kWgMaxStages) stages = kWgMaxStages;
  prm.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + 1024;
  static size_t smem_attr[2] = {0, 0};
  if (smem > smem_attr[pl.pair]) {
    if (pl.pair) CDB_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else CDB_CUDA_OK(cudaFuncSetAttribute(wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr[pl.pair] = smem;
  }
  if (pl.pair) {
    const int items = pl.n_groups * (pl.m_tiles / 2) * pl.n_tiles * pl.splits;
    const int clusters = items < sm_count() / 2 ? items : sm_count() / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    cfg.blockDim = dim3(256, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CDB_CUDA_OK(cudaLaunchKernelEx(&cfg, wgrad_kernel<true>, maps, prm));
  } else {
    const int items = pl.n_groups * pl.m_tiles * pl.n_tiles * pl.splits;
    int grid = items < sm_count() ? items : sm_count();
    wgrad_kernel<false><<<grid, 256, smem, stream>>>(maps, prm);
  }
  CDB_LAUNCH_OK();

  // W4[d0 = cS][d1 = cG]: the M side holds d0 when M is the S tensor.
  launch_wgrad_finalize(static_cast<const float*>(workspace), dw4, d0, d1, g->r, g->s, pl.n_taps, pl.splits, pl.mpad,
                        pl.npad, pl.m_is_s, g->rowpack, accumulate, stream);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
