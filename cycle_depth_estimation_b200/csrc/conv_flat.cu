// K1b — "flat" implicit-GEMM convolution for stride-1 convolutions whose padding is materialised in
// the input buffer (the reflect-padded residual-block convolutions, their data gradients over a
// zero-haloed dy, the 7x7 output layer).
//
// The padded NHWC input [N, Hp, Wp, C] is viewed as a 2-D matrix [N*Hp*Wp pixel rows, C]; an output
// position is its flat index f = h*Wp + w in the same pitch, so tap (r, s) of output f reads input row
// f + r*dil*Wp + s*dil.  One CTA tile is BM = 256 consecutive flat positions x BN <= 256 channels.
// For a filter row r the S taps read overlapping row ranges of ONE shared-memory block of
// BM + halo rows: it is fetched once and the taps are addressed by moving the start of the UMMA
// descriptor by s*dil rows of 128 B (measured on B200: the 128B swizzle is a function of the absolute
// shared-memory address bits, so an unaligned start needs NO base-offset in the descriptor).  This
// cuts the A-operand L2 traffic by S and, with the two 128-row accumulators sharing every weight
// tile, moves the 3x3 256->256 convolution from L2-bound (87 flop/B) to tensor-bound (195 flop/B).
// The (Wp - Q) junk positions per row are computed and dropped by the epilogue.
//
// Warps: 0 TMA producer (A ring + B ring), 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue
// (two accumulators x four lane quarters).  Per-channel sums for InstanceNorm are reduced through a
// per-warp shared-memory transpose and added with one atomic per (warp, column).
#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

constexpr int kFlatBM = 256;
constexpr int kFlatMaxA = 4;
constexpr int kFlatMaxB = 10;

struct FlatParams {
  int32_t R, S, dil, k_chunks, kpad, flip;
  int32_t wp, rows_per_img;          // input pitch (pixels) and rows per image
  int32_t tiles_per_img, n_img, n_tiles_n, bn;
  int32_t dom_h, dom_w;              // valid outputs: h < dom_h, w < dom_w with (h, w) = divmod(f, wp)
  int32_t halo_rows;                 // extra rows after the BM block (multiple of 8)
  int32_t a_stages, b_stages, b_taps, prewait;
  int32_t cout, cstore, out_dtype, act, stats_on, stats_batch, use_base_offset, sleep_ns, rotate, fast_out, out_rows_per_img;
  int32_t tf32, kelems;              // TF32 variant: fp32 operands, 32 channels per 128-byte K block (else 64 bf16)
  int32_t vec_out, round_out;        // fp32 NHWC output: float4 stores; round stored values to TF32
  float slope;
  const float* bias;
  void* out;
  int64_t o_sn, o_sh, o_sw, o_sc;
  float* stats;
  int* abort_flag;
  long long* dbg;  // optional timestamps of CTA 0 (CDB_FLAT_DEBUG=1)
};

struct FlatMaps {
  CUtensorMap a_big;    // box {64, 256}
  CUtensorMap a_small;  // box {64, halo_rows}
  CUtensorMap b;        // box {64, bn}
  CUtensorMap out;      // fast output path: bf16 [cstore, n*out_rows_per_img], box {64, 32}, 128B swizzle
};

__device__ __forceinline__ float flat_act(float v, int act, float slope) {
  switch (act) {
    case CDB_ACT_RELU: return v > 0.f ? v : 0.f;
    case CDB_ACT_LEAKY: return v > 0.f ? v : v * slope;
    case CDB_ACT_TANH: return tanhf(v);
    case CDB_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// kPair: a CTA pair (cluster of 2, cta_group::2) shares one 256-position x bn tile.  Each CTA stages 128 (+halo)
// input rows and HALF of the weight tile (the S taps of a filter row as ONE pipeline stage); the even CTA issues
// M = 256 instructions for both (rows 0-127 of A / D are its own, rows 128-255 the peer's) and each CTA drains its
// own 128 x bn accumulator.  That accumulator needs bn <= 256 of the 512 TMEM columns, so it is double buffered:
// the epilogue of tile i (7.5 us with InstanceNorm statistics = 25 % of a single-CTA 256 x 256 tile, which fills the
// TMEM) runs under the main loop of tile i + 1, the work units are half as long (less tail), and the L2 -> SM
// operand bytes per flop are those of the single-CTA kernel (65 KB per 12.6 MFLOP).
template <bool kPair>
__global__ void __launch_bounds__(384, 1)
igemm_flat_kernel(const __grid_constant__ FlatMaps maps, const __grid_constant__ FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_afull[kFlatMaxA];
  __shared__ __align__(8) uint64_t bar_aempty[kFlatMaxA];
  __shared__ __align__(8) uint64_t bar_bfull[kFlatMaxB];
  __shared__ __align__(8) uint64_t bar_bempty[kFlatMaxB];
  __shared__ __align__(8) uint64_t bar_tfull[2];
  __shared__ __align__(8) uint64_t bar_tempty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int abort_smem;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t rank = kPair ? cluster_ctarank() : 0u;  // 0 = the CTA that issues the MMAs
  const int cta_rows = kPair ? kFlatBM / 2 : kFlatBM;  // input / output positions of this CTA per tile
  const uint32_t a_bytes = static_cast<uint32_t>(cta_rows + p.halo_rows) * 128u;
  const int b_taps = kPair ? p.b_taps : 1;             // filter taps per weight pipeline stage (1 or S)
  const uint32_t b_bytes = static_cast<uint32_t>(kPair ? p.bn / 2 : p.bn) * 128u;
  const int tile_first = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int tile_step = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const uint32_t a_ring = smem_base;
  const uint32_t b_ring = a_ring + p.a_stages * a_bytes;
  const uint32_t slab_base = b_ring + p.b_stages * b_taps * b_bytes;  // 8 warps x 4 KB staging (1024-aligned)
  const int total_tiles = p.n_img * p.tiles_per_img * p.n_tiles_n;
  const int acc_cols = p.bn <= 128 ? 128 : 256;
  const int n_bufs = kPair ? 2 : 512 / (2 * acc_cols);  // single CTA: 2 when bn <= 128, else 1

  if (threadIdx.x == 0) {
    abort_smem = 0;
    for (int s = 0; s < p.a_stages; ++s) {
      mbar_init(smem_u32(&bar_afull[s]), 1);
      mbar_init(smem_u32(&bar_aempty[s]), 1);
    }
    for (int s = 0; s < p.b_stages; ++s) {
      mbar_init(smem_u32(&bar_bfull[s]), 1);
      mbar_init(smem_u32(&bar_bempty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&bar_tfull[b]), 1);
      mbar_init(smem_u32(&bar_tempty[b]), kPair ? 512 : 256);  // pair: the epilogue threads of both CTAs
    }
    fence_mbar_init();
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&maps.a_big);
    prefetch_tmap(&maps.a_small);
    prefetch_tmap(&maps.b);
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
  const int n_taps = p.R * p.S;
  const bool dbg = p.dbg != nullptr && blockIdx.x == 0;
  if (dbg && threadIdx.x == 0) p.dbg[0] = clock64();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      bool ok = true;
      // pair: both CTAs fill their own rings; every load completes on the ISSUING CTA's (rank 0) full barrier, which
      // that CTA arms with the bytes of both
      const uint32_t tx_mult = kPair ? 2u : 1u;
      for (int tile = tile_first; tile < total_tiles && ok; tile += tile_step) {
        const int n_tile = tile % p.n_tiles_n;
        const int m_tile = tile / p.n_tiles_n;
        const int img = m_tile / p.tiles_per_img;
        const int f0 = (m_tile % p.tiles_per_img) * kFlatBM + static_cast<int>(rank) * cta_rows;
        const int n0 = n_tile * p.bn + (kPair ? static_cast<int>(rank) * (p.bn / 2) : 0);
        // Every CTA walks the (chunk, filter row) groups in a different rotation: at any moment the
        // CTAs then pull DIFFERENT weight tiles, instead of all 148 SMs hitting the same L2 lines.
        const int groups = p.k_chunks * p.R;
        const int rot =
            p.rotate ? static_cast<int>((static_cast<unsigned>(tile_first) * 5u + static_cast<unsigned>(m_tile)) % groups) : 0;
        for (int gi = 0; gi < groups && ok; ++gi) {
          int g = gi + rot;
          if (g >= groups) g -= groups;
          const int c = g / p.R, r = g - c * p.R;
          {
            if (!mbar_wait(smem_u32(&bar_aempty[sa]), pa ^ 1u, abort_flag)) {
              ok = false;
              break;
            }
            const uint32_t full_local = smem_u32(&bar_afull[sa]);
            const uint32_t full = kPair ? mapa_shared(full_local, 0) : full_local;
            const uint32_t dst = a_ring + sa * a_bytes;
            const int row0 = img * p.rows_per_img + f0 + r * p.dil * p.wp;
            if (rank == 0) mbar_arrive_expect_tx(full_local, a_bytes * tx_mult);
            if (kPair) {
              tma_load_2d_pair(&maps.a_big, full, dst, c * p.kelems, row0);
              if (p.halo_rows > 0)
                tma_load_2d_pair(&maps.a_small, full, dst + cta_rows * 128, c * p.kelems, row0 + cta_rows);
            } else {
              tma_load_2d(&maps.a_big, full, dst, c * p.kelems, row0);
              if (p.halo_rows > 0) tma_load_2d(&maps.a_small, full, dst + kFlatBM * 128, c * p.kelems, row0 + kFlatBM);
            }
            if (++sa == p.a_stages) {
              sa = 0;
              pa ^= 1u;
            }
            for (int s = 0; s < p.S; s += b_taps) {
              if (!mbar_wait(smem_u32(&bar_bempty[sb]), pb ^ 1u, abort_flag)) {
                ok = false;
                break;
              }
              const uint32_t bfull_local = smem_u32(&bar_bfull[sb]);
              const uint32_t bfull = kPair ? mapa_shared(bfull_local, 0) : bfull_local;
              if (rank == 0) mbar_arrive_expect_tx(bfull_local, b_bytes * tx_mult * static_cast<uint32_t>(b_taps));
              for (int j = 0; j < b_taps; ++j) {
                const int t = r * p.S + s + j;
                const int wk = (p.flip ? (n_taps - 1 - t) : t) * p.kpad + c * p.kelems;
                const uint32_t dstb = b_ring + (sb * b_taps + j) * b_bytes;
                if (kPair) tma_load_2d_pair(&maps.b, bfull, dstb, wk, n0);
                else tma_load_2d(&maps.b, bfull, dstb, wk, n0);
              }
              if (++sb == p.b_stages) {
                sb = 0;
                pb ^= 1u;
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: rank 0 only)
    // Pair mode: the WHOLE warp walks the loop and an elected lane issues.  Measured on B200 (tools/mma_rate2.cu): a
    // cta_group::2 instruction issued from a lone thread of a diverged warp takes 185-283 cycles whatever its shape;
    // from the elected lane of a converged warp it takes the 128 cycles of the M = 256, N = 256, K = 16 tile.
    if (kPair ? (rank == 0) : (lane == 0)) {
      const bool tf32 = p.tf32 != 0;
      const uint32_t idesc = make_idesc(tf32 ? 2u : 1u, 0u, 0u, kPair ? 256u : 128u, static_cast<uint32_t>(p.bn));
      const bool dbgl = dbg && lane == 0;
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int local = 0;
      bool ok = true;
      long long wait_a = 0, wait_b = 0, wait_t = 0;
      for (int tile = tile_first; tile < total_tiles && ok; tile += tile_step, ++local) {
        const int buf = n_bufs == 2 ? (local & 1) : 0;
        const uint32_t tphase = (n_bufs == 2 ? (local >> 1) : local) & 1u;
        long long tw0 = dbgl ? clock64() : 0;
        ok = mbar_wait(smem_u32(&bar_tempty[buf]), tphase ^ 1u, abort_flag);
        if (kPair) ok = __all_sync(0xffffffffu, ok);
        if (!ok) break;
        if (dbgl) wait_t += clock64() - tw0;
        tc_fence_after();
        const uint32_t d0 = tmem_base + static_cast<uint32_t>(kPair ? buf * acc_cols : buf * 2 * acc_cols);
        const uint32_t d1 = d0 + static_cast<uint32_t>(acc_cols);
        bool first = true;
        if (dbgl && local == 0) p.dbg[1] = clock64();
        if (kPair && p.prewait && b_taps == p.S) {
          // Pair mode with the S weight tiles of a filter row in one stage: the barriers of group g + 1 are waited for
          // BEFORE the last instruction of group g is issued, so the round trips (two mbarrier polls by 32 lanes, a vote,
          // a fence: ~450 cycles) run while the instructions of group g execute instead of leaving the tensor pipe idle
          // between groups (tools/flat_dbg.py: 149 cycles per instruction against the 128 of the tile shape).
          const int groups = p.k_chunks * p.R;
          bool ready = false;
          for (int gi = 0; gi < groups && ok; ++gi) {
            if (!ready) {
              tw0 = dbgl ? clock64() : 0;
              ok = mbar_wait(smem_u32(&bar_afull[sa]), pa, abort_flag);
              ok = __all_sync(0xffffffffu, ok);
              if (!ok) break;
              if (dbgl) wait_a += clock64() - tw0;
              tw0 = dbgl ? clock64() : 0;
              ok = mbar_wait(smem_u32(&bar_bfull[sb]), pb, abort_flag);
              ok = __all_sync(0xffffffffu, ok);
              if (!ok) break;
              if (dbgl) wait_b += clock64() - tw0;
              tc_fence_after();
            }
            ready = false;
            const uint32_t abase = a_ring + sa * a_bytes;
            if (dbgl && first) p.dbg[2] = clock64();
            const int n_instr = 4 * b_taps;
            auto issue = [&](int i) {
              const int j = i >> 2, k = i & 3;
              const uint32_t a0 = abase + static_cast<uint32_t>(j * p.dil) * 128u;
              const uint64_t da0 = make_smem_desc_unaligned(a0, 16, 1024, kLayoutSW128, p.use_base_offset);
              const uint64_t db = make_smem_desc(b_ring + (sb * b_taps + j) * b_bytes, 16, 1024, kLayoutSW128);
              const uint32_t acc = (first && i == 0) ? 0u : 1u;
              if (tf32) umma2_tf32(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
              else umma2_f16(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
            };
            if (elect_one()) {
              for (int i = 0; i < n_instr - 1; ++i) issue(i);
            }
            __syncwarp();
            const int nsa = sa + 1 == p.a_stages ? 0 : sa + 1, nsb = sb + 1 == p.b_stages ? 0 : sb + 1;
            const uint32_t npa = sa + 1 == p.a_stages ? pa ^ 1u : pa, npb = sb + 1 == p.b_stages ? pb ^ 1u : pb;
            if (gi + 1 < groups) {
              tw0 = dbgl ? clock64() : 0;
              ok = mbar_wait(smem_u32(&bar_afull[nsa]), npa, abort_flag);
              ok = __all_sync(0xffffffffu, ok);
              if (!ok) break;
              if (dbgl) wait_a += clock64() - tw0;
              tw0 = dbgl ? clock64() : 0;
              ok = mbar_wait(smem_u32(&bar_bfull[nsb]), npb, abort_flag);
              ok = __all_sync(0xffffffffu, ok);
              if (!ok) break;
              if (dbgl) wait_b += clock64() - tw0;
              tc_fence_after();
              ready = true;
            }
            if (elect_one()) {
              issue(n_instr - 1);
              umma2_commit(smem_u32(&bar_bempty[sb]));
              umma2_commit(smem_u32(&bar_aempty[sa]));
            }
            __syncwarp();
            first = false;
            sa = nsa;
            pa = npa;
            sb = nsb;
            pb = npb;
          }
        } else
        for (int c = 0; c < p.k_chunks && ok; ++c) {
          for (int r = 0; r < p.R && ok; ++r) {
            tw0 = dbgl ? clock64() : 0;
            ok = mbar_wait(smem_u32(&bar_afull[sa]), pa, abort_flag);
            if (kPair) ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            if (dbgl) wait_a += clock64() - tw0;
            const uint32_t abase = a_ring + sa * a_bytes;
            if (dbgl && first) p.dbg[2] = clock64();
            for (int s = 0; s < p.S; s += b_taps) {
              tw0 = dbgl ? clock64() : 0;
              ok = mbar_wait(smem_u32(&bar_bfull[sb]), pb, abort_flag);
              if (kPair) ok = __all_sync(0xffffffffu, ok);
              if (!ok) break;
              if (dbgl) wait_b += clock64() - tw0;
              tc_fence_after();
              if (kPair) {
                // of an M = 256 instruction rows 0-127 of A / D are this CTA's and rows 128-255 the peer's (same
                // shared-memory / TMEM offsets); this CTA holds weight rows [0, bn/2), the peer [bn/2, bn)
                if (elect_one()) {
                  for (int j = 0; j < b_taps; ++j) {
                    const uint32_t a0 = abase + static_cast<uint32_t>((s + j) * p.dil) * 128u;
                    const uint64_t da0 = make_smem_desc_unaligned(a0, 16, 1024, kLayoutSW128, p.use_base_offset);
                    const uint64_t db = make_smem_desc(b_ring + (sb * b_taps + j) * b_bytes, 16, 1024, kLayoutSW128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                      const uint32_t acc = (first && j == 0 && k == 0) ? 0u : 1u;
                      if (tf32) umma2_tf32(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
                      else umma2_f16(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
                    }
                  }
                  umma2_commit(smem_u32(&bar_bempty[sb]));
                }
                __syncwarp();
              } else {
                const uint32_t a0 = abase + static_cast<uint32_t>(s * p.dil) * 128u;
                const uint64_t da0 = make_smem_desc_unaligned(a0, 16, 1024, kLayoutSW128, p.use_base_offset);
                const uint64_t da1 = make_smem_desc_unaligned(a0 + 128u * 128u, 16, 1024, kLayoutSW128, p.use_base_offset);
                const uint64_t db = make_smem_desc(b_ring + sb * b_bytes, 16, 1024, kLayoutSW128);
                if (tf32) {  // same 32-byte K step per instruction: 8 tf32 instead of 16 bf16
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    umma_tf32(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
                    umma_tf32(d1, da1 + 2u * k, db + 2u * k, idesc, acc);
                  }
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const uint32_t acc = (first && k == 0) ? 0u : 1u;
                    umma_f16(d0, da0 + 2u * k, db + 2u * k, idesc, acc);
                    umma_f16(d1, da1 + 2u * k, db + 2u * k, idesc, acc);
                  }
                }
                umma_commit(smem_u32(&bar_bempty[sb]));
              }
              first = false;
              if (++sb == p.b_stages) {
                sb = 0;
                pb ^= 1u;
              }
            }
            if (ok) {
              if (kPair) {
                if (elect_one()) umma2_commit(smem_u32(&bar_aempty[sa]));
                __syncwarp();
              } else {
                umma_commit(smem_u32(&bar_aempty[sa]));
              }
            }
            if (++sa == p.a_stages) {
              sa = 0;
              pa ^= 1u;
            }
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
        if (dbgl) {
          if (local == 0) p.dbg[3] = clock64();
          p.dbg[7] = clock64();
          p.dbg[8] = wait_a;
          p.dbg[9] = wait_b;
          p.dbg[10] = wait_t;
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue (8 warps)
    const int ew = warp - 4;          // 0..7
    const int sub = ew >> 2;          // single CTA: accumulator 0 / 1 (rows 128-255); pair: column half 0 / 1
    const int quarter = ew & 3;       // == warp % 4: the TMEM lane quarter this warp may read
    const int row = (kPair ? static_cast<int>(rank) : sub) * 128 + quarter * 32 + lane;  // within the 256-position tile
    const int c_begin = kPair ? sub * (p.bn / 2) : 0;
    const int c_end = kPair ? c_begin + p.bn / 2 : p.bn;
    const uint32_t tempty_remote = kPair ? mapa_shared(smem_u32(&bar_tempty[0]), 0) : 0u;
    const uint32_t stage_addr = slab_base + ew * 4096;
    float* slab = reinterpret_cast<float*>(smem_raw + (stage_addr - smem_u32(smem_raw)));  // aliases the staging
    const bool has_bias = p.bias != nullptr;
    const int act = p.act;
    const float slope = p.slope;
    const int cout = p.cout, cstore = p.cstore;
    const bool stats_on = p.stats_on != 0;
    int local = 0;
    for (int tile = tile_first; tile < total_tiles; tile += tile_step, ++local) {
      const int buf = n_bufs == 2 ? (local & 1) : 0;
      const uint32_t tphase = (n_bufs == 2 ? (local >> 1) : local) & 1u;
      if (!mbar_wait_relaxed(smem_u32(&bar_tfull[buf]), tphase, abort_flag, p.sleep_ns)) break;
      tc_fence_after();
      if (dbg && threadIdx.x == 128 && local == 0) p.dbg[4] = clock64();
      const int n_tile = tile % p.n_tiles_n;
      const int m_tile = tile / p.n_tiles_n;
      const int img = m_tile / p.tiles_per_img;
      const int f0 = (m_tile % p.tiles_per_img) * kFlatBM;
      const int f = f0 + row;
      const int h = f / p.wp, w = f - h * p.wp;
      const bool valid = (h < p.dom_h) && (w < p.dom_w);
      const int n0 = n_tile * p.bn;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) +
                             static_cast<uint32_t>(kPair ? buf * acc_cols : buf * 2 * acc_cols + sub * acc_cols);
      float* stats_img = stats_on ? p.stats + static_cast<int64_t>(p.stats_batch ? 0 : img) * cout * 2 : nullptr;
      if (p.fast_out) {
        // ---- fast path: 64-column slabs -> bf16 -> swizzled staging -> TMA store (coalesced, async)
        const int out_row0 = img * p.out_rows_per_img + f0 + (kPair ? static_cast<int>(rank) : sub) * 128 + quarter * 32;
        for (int c0 = c_begin; c0 < c_end; c0 += 64) {
          if (n0 + c0 >= cstore) break;
          uint32_t v[64];
          tmem_ld32(taddr + c0, v);
          tmem_ld32(taddr + c0 + 32, v + 32);
          tmem_ld_wait();
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
          if (n0 + c0 + 64 > cout) {
#pragma unroll
            for (int j = 0; j < 64; ++j)
              if (n0 + c0 + j >= cout) v[j] = 0u;
          }
          if (lane == 0) bulk_wait_read0();  // the previous slab's TMA store has finished reading the staging
          __syncwarp();
          if (stats_on) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
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
                const int ch = n0 + c0 + q * 16 + lane;
                if (ch < cout) {
                  atomicAdd(stats_img + ch * 2, s1);
                  atomicAdd(stats_img + ch * 2 + 1, s2);
                }
              }
              __syncwarp();
            }
          }
          // row `lane` of the 32 x 64 slab: 8 chunks of 16 B, chunk j stored at position j ^ (lane & 7)
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
            tma_store_2d(&maps.out, stage_addr, n0 + c0, out_row0);
            bulk_commit();
          }
        }
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
      } else {
        // ---- generic path: strided / fp32 / NCHW outputs (first and last layers), 16 columns at a time
        const int64_t obase = img * p.o_sn + h * p.o_sh + w * p.o_sw;
        for (int c0 = c_begin; c0 < c_end; c0 += 16) {
          if (n0 + c0 >= cstore) break;
          uint32_t v[16];
          tmem_ld16(taddr + c0, v);
          tmem_ld_wait();
          float fv[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int ch = n0 + c0 + j;
            float x = __uint_as_float(v[j]);
            if (has_bias && ch < cout) x += __ldg(p.bias + ch);
            fv[j] = ch < cout ? x : 0.f;
          }
          // activation applied by uniform branches OUTSIDE the unrolled loop (see the fast path)
          if (act == CDB_ACT_RELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) fv[j] = fmaxf(fv[j], 0.f);
          } else if (act == CDB_ACT_LEAKY) {
#pragma unroll
            for (int j = 0; j < 16; ++j) fv[j] = fv[j] > 0.f ? fv[j] : fv[j] * slope;
          } else if (act == CDB_ACT_TANH) {
#pragma unroll
            for (int j = 0; j < 16; ++j) fv[j] = n0 + c0 + j < cout ? tanhf(fv[j]) : 0.f;
          } else if (act == CDB_ACT_SIGMOID) {
#pragma unroll
            for (int j = 0; j < 16; ++j) fv[j] = n0 + c0 + j < cout ? 1.f / (1.f + __expf(-fv[j])) : 0.f;
          }
          if (stats_on) {
#pragma unroll
            for (int j = 0; j < 16; ++j) slab[lane * 17 + j] = valid ? fv[j] : 0.f;
            __syncwarp();
            if (lane < 16) {
              float s1 = 0.f, s2 = 0.f;
#pragma unroll 8
              for (int i = 0; i < 32; ++i) {
                const float t = slab[i * 17 + lane];
                s1 += t;
                s2 = fmaf(t, t, s2);
              }
              const int ch = n0 + c0 + lane;
              if (ch < cout) {
                atomicAdd(stats_img + ch * 2, s1);
                atomicAdd(stats_img + ch * 2 + 1, s2);
              }
            }
            __syncwarp();
          }
          if (valid) {
            if (p.out_dtype == CDB_BF16 && p.o_sc == 1) {
              __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + obase + n0 + c0;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                if (n0 + c0 + hh * 8 < cstore) {
                  uint4 pk;
                  pk.x = pack_bf16x2(fv[hh * 8 + 0], fv[hh * 8 + 1]);
                  pk.y = pack_bf16x2(fv[hh * 8 + 2], fv[hh * 8 + 3]);
                  pk.z = pack_bf16x2(fv[hh * 8 + 4], fv[hh * 8 + 5]);
                  pk.w = pack_bf16x2(fv[hh * 8 + 6], fv[hh * 8 + 7]);
                  *reinterpret_cast<uint4*>(o + hh * 8) = pk;
                }
              }
            } else if (p.out_dtype == CDB_BF16) {
              __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.out) + obase;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int ch = n0 + c0 + j;
                if (ch < cstore) o[ch * p.o_sc] = __float2bfloat16(fv[j]);
              }
            } else {
              if (p.round_out) {
#pragma unroll
                for (int j = 0; j < 16; ++j) fv[j] = round_tf32(fv[j]);
              }
              float* o = static_cast<float*>(p.out) + obase;
              if (p.vec_out) {
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                  const int ch = n0 + c0 + j;
                  if (ch < cstore)
                    *reinterpret_cast<float4*>(o + ch) = make_float4(fv[j], fv[j + 1], fv[j + 2], fv[j + 3]);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int ch = n0 + c0 + j;
                  if (ch < cstore) o[ch * p.o_sc] = fv[j];
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      if (kPair) mbar_arrive_cluster(tempty_remote + static_cast<uint32_t>(buf) * 8u);
      else mbar_arrive(smem_u32(&bar_tempty[buf]));
      if (dbg && threadIdx.x == 128) p.dbg[5] = clock64();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();  // neither CTA leaves while the other may still read its shared memory / signal it
  if (dbg && threadIdx.x == 0) p.dbg[6] = clock64();
  if (threadIdx.x == 0 && abort_smem && p.abort_flag) atomicExch(p.abort_flag, 1);
  if (warp == 2) {
    if (kPair) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// Host entry used by cdb_conv2d_fwd (conv_igemm.cu) when the geometry qualifies.
int launch_flat_conv(const CdbConvGeom* g, const CdbAct* x, const void* wpacked, int w_rows_pad, int w_kpad,
                     const CdbOut* y, const CdbEpilogue* ep, int flip, int use_base_offset, cudaStream_t stream) {
  FlatParams prm;
  memset(&prm, 0, sizeof(prm));
  prm.R = g->r;
  prm.S = g->s;
  prm.dil = g->dil;
  const bool tf32 = x->dtype == CDB_F32;
  const uint64_t esz = tf32 ? 4 : 2;
  const CUtensorMapDataType dt = tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  prm.tf32 = tf32 ? 1 : 0;
  prm.kelems = tf32 ? 32 : 64;
  prm.k_chunks = w_kpad / prm.kelems;
  prm.kpad = w_kpad;
  prm.flip = flip;
  prm.wp = x->w;
  prm.rows_per_img = x->h * x->w;
  prm.dom_h = y->h;
  prm.dom_w = y->w;
  prm.n_img = y->n;
  prm.tiles_per_img = ceil_div(y->h * x->w, kFlatBM);
  {
    const char* e = getenv("CDB_FLAT_BN");
    int cap = e ? atoi(e) : 256;
    if (!e) {
      // small problems (generator inference at batch 1: 17 position tiles per image): narrower channel tiles until
      // the tile count covers at least half of the SMs
      const int m_tiles = y->n * prm.tiles_per_img;
      while (cap > 64 && w_rows_pad > cap / 2 && m_tiles * ceil_div(w_rows_pad, cap) * 2 < sm_count()) cap /= 2;
    }
    prm.bn = w_rows_pad < cap ? w_rows_pad : cap;
    e = getenv("CDB_FLAT_ROTATE");
    prm.rotate = e ? atoi(e) : 1;
    e = getenv("CDB_FLAT_SLEEP");
    prm.sleep_ns = e ? atoi(e) : 0;
  }
  prm.n_tiles_n = ceil_div(w_rows_pad, prm.bn);
  // CTA pairs (cta_group::2) when the channel tile splits into two 64-column-aligned halves and the single-CTA
  // kernel would need more than one wave (measured on the 3x3 256 -> 256 layer: 136 tiles 60.9 k vs 62.3 k cycles,
  // 272 tiles 119.1 k vs 112.7 k)
  const int pair_env = getenv("CDB_FLAT_PAIR") ? atoi(getenv("CDB_FLAT_PAIR")) : -1;
  const bool pair = prm.bn % 128 == 0 && pair_env != 0 &&
                    (pair_env > 0 || (int64_t)prm.n_img * prm.tiles_per_img * prm.n_tiles_n > sm_count());
  prm.halo_rows = round_up((g->s - 1) * g->dil, 8);
  prm.cout = y->c;
  prm.cstore = y->cstore;
  prm.out_dtype = y->dtype;
  prm.act = ep ? ep->act : CDB_ACT_NONE;
  prm.slope = ep ? ep->slope : 0.f;
  prm.bias = ep ? ep->bias : nullptr;
  prm.stats = ep ? ep->stats : nullptr;
  prm.stats_on = prm.stats != nullptr;
  prm.stats_batch = (ep && (ep->flags & CDB_EP_STATS_BATCH)) ? 1 : 0;
  prm.round_out = (ep && (ep->flags & CDB_EP_ROUND_TF32)) ? 1 : 0;
  prm.vec_out = (y->dtype == CDB_F32 && y->sc == 1 && y->cstore % 4 == 0 && y->sn % 4 == 0 && y->sh % 4 == 0 &&
                 y->sw % 4 == 0 && (reinterpret_cast<uintptr_t>(y->ptr) & 15) == 0)
                    ? 1
                    : 0;
  prm.use_base_offset = use_base_offset;
  prm.out = y->ptr;
  prm.o_sn = y->sn;
  prm.o_sh = y->sh;
  prm.o_sw = y->sw;
  prm.o_sc = y->sc;
  prm.abort_flag = device_abort_flag_ptr();

  FlatMaps maps;
  memset(&maps, 0, sizeof(maps));
  const uint64_t total_rows = (uint64_t)x->n * x->h * x->w;
  {
    uint64_t dims[2] = {(uint64_t)x->c, total_rows};
    uint64_t str[1] = {(uint64_t)x->sw * esz};
    uint32_t box[2] = {(uint32_t)prm.kelems, pair ? 128u : 256u};
    int rc = make_tmap(&maps.a_big, dt, 2, x->ptr, dims, str, box);
    if (rc) return rc;
    uint32_t box2[2] = {(uint32_t)prm.kelems, (uint32_t)(prm.halo_rows > 0 ? prm.halo_rows : 8)};
    rc = make_tmap(&maps.a_small, dt, 2, x->ptr, dims, str, box2);
    if (rc) return rc;
  }
  {
    const int ktotal = g->r * g->s * w_kpad;
    uint64_t dims[2] = {(uint64_t)ktotal, (uint64_t)w_rows_pad};
    uint64_t str[1] = {(uint64_t)ktotal * esz};
    uint32_t box[2] = {(uint32_t)prm.kelems, (uint32_t)(pair ? prm.bn / 2 : prm.bn)};
    int rc = make_tmap(&maps.b, dt, 2, const_cast<void*>(wpacked), dims, str, box);
    if (rc) return rc;
  }
  // Fast output path: bf16 NHWC whose rows follow the INPUT pitch (sh = wp * cstore, sw = cstore) with
  // an image stride of tiles_per_img * 256 rows, so that a tile is 256 contiguous rows of the buffer.
  prm.out_rows_per_img = prm.tiles_per_img * kFlatBM;
  prm.fast_out = (y->dtype == CDB_BF16 && y->sc == 1 && y->sw == y->cstore && y->sh == (int64_t)x->w * y->cstore &&
                  y->sn == (int64_t)prm.out_rows_per_img * y->cstore && y->cstore % 8 == 0 &&
                  !getenv("CDB_FLAT_SLOW_OUT"))
                     ? 1
                     : 0;
  if (prm.fast_out) {
    uint64_t dims[2] = {(uint64_t)y->cstore, (uint64_t)y->n * prm.out_rows_per_img};
    uint64_t str[1] = {(uint64_t)y->cstore * 2};
    uint32_t box[2] = {64u, 32u};
    int rc = make_tmap(&maps.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, y->ptr, dims, str, box);
    if (rc) return rc;
  }
  const int a_bytes = ((pair ? kFlatBM / 2 : kFlatBM) + prm.halo_rows) * 128;
  const int b_bytes = (pair ? prm.bn / 2 : prm.bn) * 128;
  const int slab_bytes = 8 * 4096;
  const int budget = 222 * 1024 - 1024 - slab_bytes;
  int a_stages = 3, b_stages;
  prm.b_taps = 1;
  if (pair && 3 * a_bytes + 2 * g->s * b_bytes <= budget && !getenv("CDB_FLAT_NO_MERGE")) {
    // pair: the S weight tiles of a filter row travel as one pipeline stage (one barrier round trip per 4 S instructions)
    prm.b_taps = g->s;
    b_stages = (budget - a_stages * a_bytes) / (g->s * b_bytes);
    if (b_stages > 3) b_stages = 3;
    if ((budget - b_stages * g->s * b_bytes) / a_bytes >= 4) a_stages = 4;
  } else {
    if (pair) a_stages = 4;
    b_stages = (budget - a_stages * a_bytes) / b_bytes;
    if (b_stages < 2) {
      a_stages = 2;
      b_stages = (budget - a_stages * a_bytes) / b_bytes;
    }
  }
  if (b_stages > kFlatMaxB) b_stages = kFlatMaxB;
  if (getenv("CDB_FLAT_A")) a_stages = atoi(getenv("CDB_FLAT_A"));
  if (getenv("CDB_FLAT_B")) b_stages = atoi(getenv("CDB_FLAT_B"));
  if (b_stages < 2) return fail(CDB_ERR_UNSUPPORTED, "flat conv: shared memory budget");
  prm.a_stages = a_stages;
  prm.b_stages = b_stages;
  prm.prewait = getenv("CDB_MMA_PREWAIT") ? atoi(getenv("CDB_MMA_PREWAIT")) : 1;
  const size_t smem = (size_t)a_stages * a_bytes + (size_t)b_stages * prm.b_taps * b_bytes + slab_bytes + 1024;
  static size_t smem_attr[2] = {0, 0};
  if (smem > smem_attr[pair ? 1 : 0]) {
    if (pair)
      CDB_CUDA_OK(cudaFuncSetAttribute(igemm_flat_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      CDB_CUDA_OK(cudaFuncSetAttribute(igemm_flat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr[pair ? 1 : 0] = smem;
  }
  const int total = prm.n_img * prm.tiles_per_img * prm.n_tiles_n;
  if (total < 1) return CDB_OK;
  static long long* dbg_buf = nullptr;
  if (getenv("CDB_FLAT_DEBUG")) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 128);
    prm.dbg = dbg_buf;
  }
  if (pair) {
    const int clusters = total < sm_count() / 2 ? total : sm_count() / 2;
    CDB_CUDA_OK(launch_ex(igemm_flat_kernel<true>, dim3(2 * clusters, 1, 1), dim3(384, 1, 1), smem, stream, 2, true, maps,
                          prm));
  } else {
    const int grid = total < sm_count() ? total : sm_count();
    CDB_CUDA_OK(launch_ex(igemm_flat_kernel<false>, dim3(grid, 1, 1), dim3(384, 1, 1), smem, stream, 1, true, maps, prm));
  }
  CDB_LAUNCH_OK();
  if (prm.dbg) {
    long long h[16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, 128, cudaMemcpyDeviceToHost);
    fprintf(stderr,
            "[flat dbg] pair=%d tiles=%d mma_loop start %lld, first A %lld, tile0 issued %lld, tile0 tfull seen %lld, last tile "
            "issued %lld, last epilogue end %lld, exit %lld; issuer waits: A %lld B %lld tmem %lld\n",
            (int)pair, total, h[1] - h[0], h[2] - h[0], h[3] - h[0], h[4] - h[0], h[7] - h[0], h[5] - h[0], h[6] - h[0], h[8],
            h[9], h[10]);
  }
  return CDB_OK;
}

}  // namespace cdb
