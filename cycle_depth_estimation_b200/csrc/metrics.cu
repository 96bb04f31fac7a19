// K7 — KITTI depth metrics of new_multi/my_eval.py (compute_errors :7-31 inside eval_metric :35-108)
// for a batch of uint8 ground-truth / prediction image pairs.
//
// Per image:  pred = clip(pred_u8/255*80, 1, 50);  mask = 1 < gt < 50;  over the masked pixels
//   v = (pred - min)/(max - min)*49 + 1;  thr = max(gt/v, v/gt);  a_k = mean(thr < 1.25^k);
//   rmse = sqrt(mean((gt-v)^2));  rmse_log = sqrt(mean((log(gt)-log(v))^2));
//   abs_rel = mean(|gt-v|/gt);  sq_rel = mean((gt-v)^2/gt).
// gt has 48 admissible values and pred 256, so everything is a function of the pair (gt, pred):
//   pass A  masked min / max / count of pred_u8 (integer, exact)
//   pass L  per image look-up tables: v[256], log v[256] and, per gt value, the exact pred-index
//           interval on which each threshold test is true (evaluated in IEEE double with the
//           reference's operation order and no FMA contraction, so the COUNTS are bit-exact)
//   pass B  stream the two byte planes once more (16 pixels per 128-bit load) and accumulate
//   pass F  deterministic reduction of the per-block partials
// np.log of a uint8 array is computed in float16 by numpy; kLogU8 holds
// numpy's own values.
#include "common.cuh"

namespace cdb {

__constant__ float kLogU8[256] = {
    -INFINITY, 0.0f, 0.693359375f, 1.099609375f, 1.38671875f, 1.609375f, 1.79296875f, 1.9462890625f,
    2.080078125f, 2.197265625f, 2.302734375f, 2.3984375f, 2.486328125f, 2.56640625f, 2.640625f,
    2.708984375f, 2.7734375f, 2.833984375f, 2.890625f, 2.9453125f, 2.99609375f, 3.044921875f,
    3.091796875f, 3.13671875f, 3.1796875f, 3.220703125f, 3.259765625f, 3.296875f, 3.333984375f,
    3.369140625f, 3.40234375f, 3.435546875f, 3.466796875f, 3.498046875f, 3.52734375f, 3.556640625f,
    3.583984375f, 3.611328125f, 3.638671875f, 3.6640625f, 3.689453125f, 3.71484375f, 3.73828125f,
    3.76171875f, 3.78515625f, 3.80859375f, 3.830078125f, 3.8515625f, 3.873046875f, 3.892578125f,
    3.9140625f, 3.93359375f, 3.953125f, 3.970703125f, 3.990234375f, 4.0078125f, 4.02734375f,
    4.04296875f, 4.0625f, 4.078125f, 4.09375f, 4.11328125f, 4.12890625f, 4.14453125f, 4.16015625f,
    4.17578125f, 4.19140625f, 4.20703125f, 4.22265625f, 4.234375f, 4.25f, 4.265625f, 4.27734375f,
    4.29296875f, 4.3046875f, 4.3203125f, 4.33203125f, 4.34375f, 4.359375f, 4.37109375f, 4.3828125f,
    4.39453125f, 4.40625f, 4.421875f, 4.43359375f, 4.4453125f, 4.45703125f, 4.46875f, 4.48046875f,
    4.48828125f, 4.5f, 4.51171875f, 4.5234375f, 4.53515625f, 4.54296875f, 4.5546875f, 4.5625f,
    4.57421875f, 4.5859375f, 4.59375f, 4.60546875f, 4.61328125f, 4.625f, 4.6328125f, 4.64453125f,
    4.65234375f, 4.6640625f, 4.671875f, 4.68359375f, 4.69140625f, 4.69921875f, 4.7109375f, 4.71875f,
    4.7265625f, 4.734375f, 4.74609375f, 4.75390625f, 4.76171875f, 4.76953125f, 4.77734375f, 4.78515625f,
    4.796875f, 4.8046875f, 4.8125f, 4.8203125f, 4.828125f, 4.8359375f, 4.84375f, 4.8515625f, 4.859375f,
    4.8671875f, 4.875f, 4.8828125f, 4.890625f, 4.8984375f, 4.90625f, 4.9140625f, 4.91796875f,
    4.92578125f, 4.93359375f, 4.94140625f, 4.94921875f, 4.95703125f, 4.9609375f, 4.96875f, 4.9765625f,
    4.984375f, 4.98828125f, 4.99609375f, 5.00390625f, 5.01171875f, 5.015625f, 5.0234375f, 5.03125f,
    5.03515625f, 5.04296875f, 5.05078125f, 5.0546875f, 5.0625f, 5.0703125f, 5.07421875f, 5.08203125f,
    5.0859375f, 5.09375f, 5.09765625f, 5.10546875f, 5.11328125f, 5.1171875f, 5.125f, 5.12890625f,
    5.13671875f, 5.140625f, 5.1484375f, 5.15234375f, 5.16015625f, 5.1640625f, 5.171875f, 5.17578125f,
    5.1796875f, 5.1875f, 5.19140625f, 5.19921875f, 5.203125f, 5.2109375f, 5.21484375f, 5.21875f,
    5.2265625f, 5.23046875f, 5.234375f, 5.2421875f, 5.24609375f, 5.25f, 5.2578125f, 5.265625f,
    5.26953125f, 5.2734375f, 5.28125f, 5.28515625f, 5.2890625f, 5.296875f, 5.30078125f, 5.3046875f,
    5.30859375f, 5.31640625f, 5.3203125f, 5.32421875f, 5.328125f, 5.3359375f, 5.33984375f, 5.34375f,
    5.34765625f, 5.35546875f, 5.359375f, 5.36328125f, 5.3671875f, 5.37109375f, 5.37890625f, 5.3828125f,
    5.38671875f, 5.390625f, 5.39453125f, 5.3984375f, 5.40625f, 5.41015625f, 5.4140625f, 5.41796875f,
    5.421875f, 5.42578125f, 5.4296875f, 5.43359375f, 5.44140625f, 5.4453125f, 5.44921875f, 5.453125f,
    5.45703125f, 5.4609375f, 5.46484375f, 5.46875f, 5.47265625f, 5.4765625f, 5.48046875f, 5.48828125f,
    5.4921875f, 5.49609375f, 5.5f, 5.50390625f, 5.5078125f, 5.51171875f, 5.515625f, 5.51953125f,
    5.5234375f, 5.52734375f, 5.53125f, 5.53515625f, 5.5390625f, 5.54296875f
};

struct ImgInfo {
  int pmin, pmax, count, pad_;
};

struct ImgLut {
  float v[256];
  float lv[256];
  float gf[64];    // indexed by gt-2 (48 used)
  float invg[64];
  float lg[64];
  uint8_t lo[3][64];
  uint8_t hi[3][64];
};

__global__ void metrics_init_kernel(ImgInfo* info, int n_img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_img) {
    info[i].pmin = 255;
    info[i].pmax = 0;
    info[i].count = 0;
    info[i].pad_ = 0;
  }
}

__device__ __forceinline__ bool masked(uint32_t g) { return (g - 2u) < 48u; }

// chunk geometry shared by passes A and B: [begin, end) byte range of one block within an image
__device__ __forceinline__ void chunk_range(int64_t pixels, int chunk, int chunks, int64_t* b, int64_t* e) {
  int64_t per = (pixels + chunks - 1) / chunks;
  per = (per + 15) & ~static_cast<int64_t>(15);
  *b = static_cast<int64_t>(chunk) * per;
  *e = *b + per < pixels ? *b + per : pixels;
  if (*b > pixels) *b = pixels;
}

__global__ void __launch_bounds__(256)
metrics_minmax_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t pixels,
                      ImgInfo* __restrict__ info) {
  const int img = blockIdx.y;
  const uint8_t* g = gt + img * pixels;
  const uint8_t* p = pred + img * pixels;
  int64_t b, e;
  chunk_range(pixels, blockIdx.x, gridDim.x, &b, &e);
  int mn = 255, mx = 0, cnt = 0;
  // scalar head up to 16-byte alignment of the absolute address, vector body, scalar tail
  int64_t i = b + threadIdx.x;
  const int64_t head_end = min(e, b + ((16 - ((reinterpret_cast<uintptr_t>(g) + b) & 15)) & 15));
  for (; i < head_end; i += 256) {
    if (masked(g[i])) {
      mn = min(mn, (int)p[i]);
      mx = max(mx, (int)p[i]);
      ++cnt;
    }
  }
  const bool same_align = ((reinterpret_cast<uintptr_t>(g) ^ reinterpret_cast<uintptr_t>(p)) & 15) == 0;
  int64_t body_end = head_end;
  if (same_align) {
    const int64_t nvec = (e - head_end) / 16;
    body_end = head_end + nvec * 16;
    for (int64_t vi = threadIdx.x; vi < nvec; vi += 256) {
      const uint4 gv = *reinterpret_cast<const uint4*>(g + head_end + vi * 16);
      const uint4 pv = *reinterpret_cast<const uint4*>(p + head_end + vi * 16);
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t gg = (gw[k] >> (8 * j)) & 255u;
          const int pp = (pw[k] >> (8 * j)) & 255u;
          if (masked(gg)) {
            mn = min(mn, pp);
            mx = max(mx, pp);
            ++cnt;
          }
        }
    }
  }
  for (i = body_end + threadIdx.x; i < e; i += 256) {
    if (masked(g[i])) {
      mn = min(mn, (int)p[i]);
      mx = max(mx, (int)p[i]);
      ++cnt;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicMin(&info[img].pmin, mn);
    atomicMax(&info[img].pmax, mx);
    atomicAdd(&info[img].count, cnt);
  }
}

// pred_u8 -> clip(pred/255*80, 1, 50) exactly as numpy evaluates it (two separate roundings).
__device__ __forceinline__ double pred_value(int p) {
  double v = __dmul_rn(__ddiv_rn(static_cast<double>(p), 255.0), 80.0);
  if (v < 1.0) v = 1.0;
  if (v > 50.0) v = 50.0;
  return v;
}

__global__ void __launch_bounds__(256) metrics_lut_kernel(const ImgInfo* __restrict__ info, ImgLut* __restrict__ lut) {
  __shared__ double vd[256];
  const int img = blockIdx.x;
  const int p = threadIdx.x;
  const ImgInfo inf = info[img];
  const double lo = pred_value(inf.pmin), hi = pred_value(inf.pmax);
  // (x - min) / (max - min) * 49 + 1, each operation rounded separately (no FMA)
  const double v = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(pred_value(p), lo), __dsub_rn(hi, lo)), 49.0), 1.0);
  vd[p] = v;
  lut[img].v[p] = static_cast<float>(v);
  lut[img].lv[p] = static_cast<float>(log(v));
  __syncthreads();
  if (p < 48) {
    const int g = p + 2;
    const double gd = static_cast<double>(g);
    lut[img].gf[p] = static_cast<float>(g);
    lut[img].invg[p] = static_cast<float>(1.0 / gd);
    lut[img].lg[p] = kLogU8[g];
    const double thr[3] = {1.25, 1.5625, 1.953125};  // 1.25**k, exact in binary
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      int first = 1, last = 0;
      bool any = false;
      for (int q = inf.pmin; q <= inf.pmax; ++q) {
        const double a = __ddiv_rn(gd, vd[q]), b = __ddiv_rn(vd[q], gd);
        const double m = fmax(a, b);  // np.maximum propagates NaN; NaN < thr is false either way
        const bool ok = (a != a || b != b) ? false : (m < thr[t]);
        if (ok) {
          if (!any) first = q;
          last = q;
          any = true;
        }
      }
      lut[img].lo[t][p] = static_cast<uint8_t>(first);
      lut[img].hi[t][p] = static_cast<uint8_t>(last);
    }
  }
}

struct Acc {
  double sq, lg, ar, sr;
  int n, a1, a2, a3;
};

__device__ __forceinline__ void accum_pixel(const ImgLut& L, uint32_t g, uint32_t p, float& sq, float& lg, float& ar,
                                            float& sr, int& n, int& a1, int& a2, int& a3) {
  if (!masked(g)) return;
  const int gi = g - 2;
  const float v = L.v[p];
  const float d = L.gf[gi] - v;
  const float dl = L.lg[gi] - L.lv[p];
  const float ig = L.invg[gi];
  const float d2 = d * d;
  sq += d2;
  lg = fmaf(dl, dl, lg);
  ar = fmaf(fabsf(d), ig, ar);
  sr = fmaf(d2, ig, sr);
  ++n;
  a1 += (p >= L.lo[0][gi] && p <= L.hi[0][gi]) ? 1 : 0;
  a2 += (p >= L.lo[1][gi] && p <= L.hi[1][gi]) ? 1 : 0;
  a3 += (p >= L.lo[2][gi] && p <= L.hi[2][gi]) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
metrics_accum_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t pixels,
                     const ImgLut* __restrict__ lut, double* __restrict__ partial /*[img][chunks][8]*/) {
  __shared__ ImgLut L;
  __shared__ double red[8][8];
  const int img = blockIdx.y;
  {
    const uint32_t* s = reinterpret_cast<const uint32_t*>(lut + img);
    uint32_t* d = reinterpret_cast<uint32_t*>(&L);
    for (int i = threadIdx.x; i < (int)(sizeof(ImgLut) / 4); i += 256) d[i] = s[i];
  }
  __syncthreads();
  const uint8_t* g = gt + img * pixels;
  const uint8_t* p = pred + img * pixels;
  int64_t b, e;
  chunk_range(pixels, blockIdx.x, gridDim.x, &b, &e);
  Acc acc = {0.0, 0.0, 0.0, 0.0, 0, 0, 0, 0};
  float sq = 0.f, lg = 0.f, ar = 0.f, sr = 0.f;
  const int64_t head_end = min(e, b + ((16 - ((reinterpret_cast<uintptr_t>(g) + b) & 15)) & 15));
  for (int64_t i = b + threadIdx.x; i < head_end; i += 256)
    accum_pixel(L, g[i], p[i], sq, lg, ar, sr, acc.n, acc.a1, acc.a2, acc.a3);
  const bool same_align = ((reinterpret_cast<uintptr_t>(g) ^ reinterpret_cast<uintptr_t>(p)) & 15) == 0;
  int64_t body_end = head_end;
  if (same_align) {
    const int64_t nvec = (e - head_end) / 16;
    body_end = head_end + nvec * 16;
    for (int64_t vi = threadIdx.x; vi < nvec; vi += 256) {
      const uint4 gv = *reinterpret_cast<const uint4*>(g + head_end + vi * 16);
      const uint4 pv = *reinterpret_cast<const uint4*>(p + head_end + vi * 16);
      const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
      const uint32_t pw[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j)
          accum_pixel(L, (gw[k] >> (8 * j)) & 255u, (pw[k] >> (8 * j)) & 255u, sq, lg, ar, sr, acc.n, acc.a1,
                      acc.a2, acc.a3);
      // flush the short fp32 runs into the double accumulators
      acc.sq += sq;
      acc.lg += lg;
      acc.ar += ar;
      acc.sr += sr;
      sq = lg = ar = sr = 0.f;
    }
  }
  for (int64_t i = body_end + threadIdx.x; i < e; i += 256)
    accum_pixel(L, g[i], p[i], sq, lg, ar, sr, acc.n, acc.a1, acc.a2, acc.a3);
  acc.sq += sq;
  acc.lg += lg;
  acc.ar += ar;
  acc.sr += sr;
  double vals[8] = {acc.sq, acc.lg, acc.ar, acc.sr, (double)acc.n, (double)acc.a1, (double)acc.a2, (double)acc.a3};
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][k] = vals[k];
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    partial[(static_cast<int64_t>(img) * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = s;
  }
}

// out[img] = {abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, count}
__global__ void metrics_finalize_kernel(const double* __restrict__ partial, int chunks, int n_img,
                                        double* __restrict__ out) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= n_img) return;
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c = 0; c < chunks; ++c)
    for (int k = 0; k < 8; ++k) s[k] += partial[(static_cast<int64_t>(img) * chunks + c) * 8 + k];
  const double n = s[4];
  double* o = out + static_cast<int64_t>(img) * 8;
  o[0] = s[2] / n;
  o[1] = s[3] / n;
  o[2] = sqrt(s[0] / n);
  o[3] = sqrt(s[1] / n);
  o[4] = s[5] / n;
  o[5] = s[6] / n;
  o[6] = s[7] / n;
  o[7] = n;
}

constexpr int kMetricChunks = 8;

}  // namespace cdb

using namespace cdb;

extern "C" size_t cdb_depth_metrics_workspace(int32_t n_img) {
  size_t a = (size_t)n_img * sizeof(ImgInfo);
  a = (a + 255) & ~(size_t)255;
  size_t b = (size_t)n_img * sizeof(ImgLut);
  b = (b + 255) & ~(size_t)255;
  return a + b + (size_t)n_img * kMetricChunks * 8 * sizeof(double);
}

extern "C" int cdb_depth_metrics(const uint8_t* gt, const uint8_t* pred, int32_t n_img, int32_t h, int32_t w,
                                 double* out8_per_img, void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(gt && pred && out8_per_img && workspace && n_img > 0 && h > 0 && w > 0, CDB_ERR_BAD_DESC,
              "depth_metrics: bad argument");
  CDB_REQUIRE(ws_bytes >= cdb_depth_metrics_workspace(n_img), CDB_ERR_WORKSPACE, "depth_metrics: workspace too small");
  CDB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, CDB_ERR_ALIGNMENT, "depth_metrics: workspace alignment");
  char* ws = static_cast<char*>(workspace);
  ImgInfo* info = reinterpret_cast<ImgInfo*>(ws);
  size_t a = ((size_t)n_img * sizeof(ImgInfo) + 255) & ~(size_t)255;
  ImgLut* lut = reinterpret_cast<ImgLut*>(ws + a);
  size_t b = ((size_t)n_img * sizeof(ImgLut) + 255) & ~(size_t)255;
  double* partial = reinterpret_cast<double*>(ws + a + b);
  const int64_t pixels = (int64_t)h * w;
  metrics_init_kernel<<<ceil_div(n_img, 256), 256, 0, stream>>>(info, n_img);
  CDB_LAUNCH_OK();
  dim3 grid(kMetricChunks, n_img);
  metrics_minmax_kernel<<<grid, 256, 0, stream>>>(gt, pred, pixels, info);
  CDB_LAUNCH_OK();
  metrics_lut_kernel<<<n_img, 256, 0, stream>>>(info, lut);
  CDB_LAUNCH_OK();
  metrics_accum_kernel<<<grid, 256, 0, stream>>>(gt, pred, pixels, lut, partial);
  CDB_LAUNCH_OK();
  metrics_finalize_kernel<<<ceil_div(n_img, 128), 128, 0, stream>>>(partial, kMetricChunks, n_img, out8_per_img);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
