// K4 — fused memory-bound kernels around the convolutions: per-channel statistics,
// InstanceNorm/BatchNorm + activation (+ residual add) writing straight into the (reflect-)padded
// buffer the next convolution reads, and the matching backward passes (halo fold + activation
// mask + norm backward).  All tensors are NHWC with 8 bf16 channels (16 bytes) per thread access;
// consecutive threads own consecutive channel vectors of a pixel, so every warp request is a run of
// full 128-byte lines.  Per-channel coefficients are computed once per thread and reused over the
// pixels the thread walks.
#include <cooperative_groups.h>

#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

// Storage type of the activations: bf16 (the default network mode) or fp32 (the TF32 / error-compensated TF32
// network modes, whose convolutions read fp32 operands).  All kernels of this file are templates over it; the
// arithmetic is fp32 in both cases, only the 8-channel vector load / store differs (16 B vs 2 x 16 B).
struct View {  // NHWC view, channel stride 1; ptr is T* of the kernel's storage type
  void* ptr;
  int64_t sn, sh, sw;
};

static inline View view_of(const CdbAct* a) {
  View v;
  v.ptr = a ? a->ptr : nullptr;
  v.sn = a ? a->sn : 0;
  v.sh = a ? a->sh : 0;
  v.sw = a ? a->sw : 0;
  return v;
}

struct F8 {  // eight fp32 channels
  float4 a, b;
};
template <typename T>
struct Vec8;
template <>
struct Vec8<__nv_bfloat16> {
  typedef uint4 type;
};
template <>
struct Vec8<float> {
  typedef F8 type;
};

__device__ __forceinline__ void unpack8(const uint4& r, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack8(const F8& r, float* f) {
  f[0] = r.a.x; f[1] = r.a.y; f[2] = r.a.z; f[3] = r.a.w;
  f[4] = r.b.x; f[5] = r.b.y; f[6] = r.b.z; f[7] = r.b.w;
}
template <typename T>
__device__ __forceinline__ typename Vec8<T>::type pack8(const float* f);
template <>
__device__ __forceinline__ uint4 pack8<__nv_bfloat16>(const float* f) {
  uint4 r;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
template <>
__device__ __forceinline__ F8 pack8<float>(const float* f) {
  F8 r;
  r.a = make_float4(f[0], f[1], f[2], f[3]);
  r.b = make_float4(f[4], f[5], f[6], f[7]);
  return r;
}
__device__ __forceinline__ uint4 ld16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void st16(__nv_bfloat16* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
__device__ __forceinline__ F8 ld16(const float* p) {
  F8 r;
  r.a = reinterpret_cast<const float4*>(p)[0];
  r.b = reinterpret_cast<const float4*>(p)[1];
  return r;
}
__device__ __forceinline__ void st16(float* p, const F8& v) {
  reinterpret_cast<float4*>(p)[0] = v.a;
  reinterpret_cast<float4*>(p)[1] = v.b;
}

// Walks pixels p0+lane, p0+lane+lanes, ... of a W-wide image keeping (h, w) without divisions.
struct PixelWalk {
  int h, w, W;
  __device__ __forceinline__ void init(int px, int W_) {
    W = W_;
    h = px / W_;
    w = px - h * W_;
  }
  __device__ __forceinline__ void advance(int step) {
    w += step;
    while (w >= W) {
      w -= W;
      ++h;
    }
  }
};

// Thread -> (channel vector, pixel lane) mapping shared by all kernels of this file.
struct Mapping {
  int vt;      // channel vectors handled per block (power of two <= 256)
  int lanes;   // pixel lanes per block = 256 / vt
  int cv_tiles;
};
static Mapping mapping_for(int c) {
  const int cv = c / 8;
  int vt = 1;
  while (vt < cv && vt < 256) vt <<= 1;
  Mapping m;
  m.vt = vt;
  m.lanes = 256 / vt;
  m.cv_tiles = ceil_div(cv, vt);
  return m;
}
static int chunks_for(int pixels, int lanes, int n, int cv_tiles, int per_sm_default = 2) {
  const int per_sm = getenv("CDB_NORM_BLOCKS_PER_SM") ? atoi(getenv("CDB_NORM_BLOCKS_PER_SM")) : per_sm_default;
  int want = (per_sm * sm_count()) / (n * cv_tiles > 0 ? n * cv_tiles : 1);
  if (want < 1) want = 1;
  int max_chunks = ceil_div(pixels, lanes * 4);
  if (max_chunks < 1) max_chunks = 1;
  return want < max_chunks ? want : max_chunks;
}

// ------------------------------------------------------------------------------------------------
// statistics: stats[g][c][2] += (sum, sum of squares), g = image (instance) or 0 (batch)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
channel_stats_kernel(View y, int H, int W, int C, int vt, int per_image, float* __restrict__ stats) {
  __shared__ float red[256 * 16];
  const int v = threadIdx.x % vt, lane = threadIdx.x / vt, lanes = 256 / vt;
  const int cvec = blockIdx.z * vt + v;
  const int n = blockIdx.y;
  const int pixels = H * W;
  const int per_chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per_chunk;
  const int p1 = min(pixels, p0 + per_chunk);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  const bool active = cvec * 8 < C;
  if (active) {
    const T* base = static_cast<const T*>(y.ptr) + n * y.sn + cvec * 8;
    PixelWalk pw;
    pw.init(p0 + lane, W);
    const int ysh = static_cast<int>(y.sh), ysw = static_cast<int>(y.sw);
    for (int p = p0 + lane; p < p1; p += lanes, pw.advance(lanes)) {
      float f[8];
      unpack8(ld16(base + pw.h * ysh + pw.w * ysw), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += f[j];
        s2[j] += f[j] * f[j];
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[(j * 2) * 256 + threadIdx.x] = s1[j];
    red[(j * 2 + 1) * 256 + threadIdx.x] = s2[j];
  }
  __syncthreads();
  // threads 0 .. vt*16-1 each finish one (vector, component) pair
  for (int t = threadIdx.x; t < vt * 16; t += 256) {
    const int vv = t % vt, comp = t / vt;
    float acc = 0.f;
    for (int l = 0; l < lanes; ++l) acc += red[comp * 256 + l * vt + vv];
    const int ch = (blockIdx.z * vt + vv) * 8 + comp / 2;
    if (ch < C) atomicAdd(stats + ((per_image ? n : 0) * static_cast<int64_t>(C) + ch) * 2 + (comp & 1), acc);
  }
}

// ------------------------------------------------------------------------------------------------
// forward: out = [residual +] act(norm(y)), written to the interior of `out` and mirrored into
// its reflect halo of `pad` pixels.
// ------------------------------------------------------------------------------------------------
struct NormFwdParams {
  View y, res, out;
  int H, W, C, vt;
  int rows_per_block;
  int norm, act, pad, has_res;
  float slope, eps, inv_count;
  int per_image;           // stats indexed per image
  int use_running;         // eval-mode batch norm
  const float* stats;      // sums
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
};

__device__ __forceinline__ float act_fwd(float v, int act, float slope) {
  switch (act) {
    case CDB_ACT_RELU: return fmaxf(v, 0.f);
    case CDB_ACT_LEAKY: return v > 0.f ? v : v * slope;
    case CDB_ACT_TANH: return tanhf(v);
    case CDB_ACT_SIGMOID: return 1.f / (1.f + __expf(-v));
    default: return v;
  }
}

// scale/shift such that normalised = x*scale + shift
__device__ __forceinline__ void norm_coeffs(int norm, int use_running, const float* stats, const float* gamma,
                                            const float* beta, const float* rmean, const float* rvar, int g,
                                            int C, int ch0, float inv_count, float eps, float* scale,
                                            float* shift, float* mean_out, float* rstd_out) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = ch0 + j;
    float mean = 0.f, rstd = 1.f;
    if (norm != CDB_NORM_NONE && ch < C) {
      if (use_running) {
        mean = rmean[ch];
        rstd = rsqrtf(rvar[ch] + eps);
      } else {
        const float s1 = stats[(static_cast<int64_t>(g) * C + ch) * 2];
        const float s2 = stats[(static_cast<int64_t>(g) * C + ch) * 2 + 1];
        mean = s1 * inv_count;
        const float var = fmaxf(s2 * inv_count - mean * mean, 0.f);
        rstd = rsqrtf(var + eps);
      }
    }
    const float ga = (gamma != nullptr && ch < C) ? gamma[ch] : 1.f;
    const float be = (beta != nullptr && ch < C) ? beta[ch] : 0.f;
    scale[j] = rstd * ga;
    shift[j] = be - mean * rstd * ga;
    if (mean_out) mean_out[j] = mean;
    if (rstd_out) rstd_out[j] = rstd;
  }
}

// Row-strip mapping: a block owns `rows_per_block` image rows of one image and one tile of channel
// vectors; thread (v, lane) walks the pixels w = lane, lane + lanes, ... of each row, U pixels per
// iteration with all loads issued before the first use.  ACT / HAS_RES are compile-time so the inner
// loop carries no activation switch.
template <int ACT>
__device__ __forceinline__ float act_fwd_t(float v, float slope) {
  if (ACT == CDB_ACT_RELU) return fmaxf(v, 0.f);
  if (ACT == CDB_ACT_LEAKY) return v > 0.f ? v : v * slope;
  if (ACT == CDB_ACT_TANH) return tanhf(v);
  if (ACT == CDB_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
  return v;
}

template <typename T>
__device__ __forceinline__ void write_halo(T* ob, int osh, int osw, int h, int w, int H, int W,
                                           int pad, const typename Vec8<T>::type& o) {
  // reflect halo: interior row d (1..pad) mirrors to row -d, row H-1-d to row H-1+d (scalars, no indexed arrays: those
  // went to local memory)
  constexpr int kNone = -(1 << 30);
  const int h1 = (h >= 1 && h <= pad) ? -h : kNone;
  const int h2 = (h <= H - 2 && h >= H - 1 - pad) ? 2 * (H - 1) - h : kNone;
  const int w1 = (w >= 1 && w <= pad) ? -w : kNone;
  const int w2 = (w <= W - 2 && w >= W - 1 - pad) ? 2 * (W - 1) - w : kNone;
  if (h1 != kNone) {
    st16(ob + h1 * osh + w * osw, o);
    if (w1 != kNone) st16(ob + h1 * osh + w1 * osw, o);
    if (w2 != kNone) st16(ob + h1 * osh + w2 * osw, o);
  }
  if (h2 != kNone) {
    st16(ob + h2 * osh + w * osw, o);
    if (w1 != kNone) st16(ob + h2 * osh + w1 * osw, o);
    if (w2 != kNone) st16(ob + h2 * osh + w2 * osw, o);
  }
  if (w1 != kNone) st16(ob + h * osh + w1 * osw, o);
  if (w2 != kNone) st16(ob + h * osh + w2 * osw, o);
}

template <typename T, int ACT, bool HAS_RES>
__global__ void __launch_bounds__(256, 3) norm_act_fwd_kernel(NormFwdParams p) {
  typedef typename Vec8<T>::type V8;
  const int v = threadIdx.x % p.vt, lane = threadIdx.x / p.vt, lanes = 256 / p.vt;
  const int cvec = blockIdx.z * p.vt + v;
  if (cvec * 8 >= p.C) return;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * p.rows_per_block;
  const int r1 = min(p.H, r0 + p.rows_per_block);
  float scale[8], shift[8];
  norm_coeffs(p.norm, p.use_running, p.stats, p.gamma, p.beta, p.running_mean, p.running_var,
              p.per_image ? n : 0, p.C, cvec * 8, p.inv_count, p.eps, scale, shift, nullptr, nullptr);
  const T* yb = static_cast<const T*>(p.y.ptr) + n * p.y.sn + cvec * 8;
  const T* rb = HAS_RES ? static_cast<const T*>(p.res.ptr) + n * p.res.sn + cvec * 8 : nullptr;
  T* ob = static_cast<T*>(p.out.ptr) + n * p.out.sn + cvec * 8;
  const int pad = p.pad, H = p.H, W = p.W;
  const float slope = p.slope;
  constexpr int U = sizeof(T) == 2 ? 4 : 2;
  const int ysh = static_cast<int>(p.y.sh), ysw = static_cast<int>(p.y.sw);
  const int rsh = static_cast<int>(p.res.sh), rsw = static_cast<int>(p.res.sw);
  const int osh = static_cast<int>(p.out.sh), osw = static_cast<int>(p.out.sw);
  for (int h = r0; h < r1; ++h) {
    const T* yr_ = yb + h * ysh;
    const T* rr_ = HAS_RES ? rb + h * rsh : nullptr;
    const bool hborder = pad > 0 && (h <= pad || h >= H - 1 - pad);
    for (int w0 = lane; w0 < W; w0 += lanes * U) {
      V8 yr[U], rr[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * lanes;
        if (w < W) {
          yr[u] = ld16(yr_ + w * ysw);
          if (HAS_RES) rr[u] = ld16(rr_ + w * rsw);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int w = w0 + u * lanes;
        if (w >= W) break;
        float f[8];
        unpack8(yr[u], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = act_fwd_t<ACT>(f[j] * scale[j] + shift[j], slope);
        if (HAS_RES) {
          float r[8];
          unpack8(rr[u], r);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] += r[j];
        }
        const V8 o = pack8<T>(f);
        st16(ob + h * osh + w * osw, o);
        if (pad > 0 && (hborder || w <= pad || w >= W - 1 - pad)) write_halo<T>(ob, osh, osw, h, w, H, W, pad, o);
      }
    }
  }
}

template <typename T, bool HAS_RES>
static void launch_norm_fwd(const NormFwdParams& p, dim3 grid, cudaStream_t stream) {
  switch (p.act) {
    case CDB_ACT_NONE: norm_act_fwd_kernel<T, CDB_ACT_NONE, HAS_RES><<<grid, 256, 0, stream>>>(p); break;
    case CDB_ACT_RELU: norm_act_fwd_kernel<T, CDB_ACT_RELU, HAS_RES><<<grid, 256, 0, stream>>>(p); break;
    case CDB_ACT_LEAKY: norm_act_fwd_kernel<T, CDB_ACT_LEAKY, HAS_RES><<<grid, 256, 0, stream>>>(p); break;
    case CDB_ACT_TANH: norm_act_fwd_kernel<T, CDB_ACT_TANH, HAS_RES><<<grid, 256, 0, stream>>>(p); break;
    default: norm_act_fwd_kernel<T, CDB_ACT_SIGMOID, HAS_RES><<<grid, 256, 0, stream>>>(p); break;
  }
}

// rows per block such that the grid holds about `per_sm` blocks per SM
static int rows_per_block_for(int H, int n, int cv_tiles, int per_sm) {
  const char* e = getenv("CDB_NORM_BLOCKS_PER_SM");
  if (e) per_sm = atoi(e);
  int chunks = (per_sm * sm_count()) / (n * cv_tiles > 0 ? n * cv_tiles : 1);
  if (chunks < 1) chunks = 1;
  if (chunks > H) chunks = H;
  return ceil_div(H, chunks);
}

// Batch-norm running statistics (training mode): r = (1-m) r + m * batch (unbiased variance).
__global__ void bn_running_kernel(const float* __restrict__ stats, int C, float count, float momentum,
                                  float* __restrict__ rmean, float* __restrict__ rvar,
                                  const float* __restrict__ conv_bias) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= C) return;
  float mean = stats[ch * 2] / count;
  const float var = fmaxf(stats[ch * 2 + 1] / count - mean * mean, 0.f);
  const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
  if (conv_bias != nullptr) mean += conv_bias[ch];   // the statistics were taken on the bias-free convolution output
  rmean[ch] = (1.f - momentum) * rmean[ch] + momentum * mean;
  rvar[ch] = (1.f - momentum) * rvar[ch] + momentum * unbiased;
}

// ------------------------------------------------------------------------------------------------
// backward
//   g  = fold_reflect(dout) + dskip                      gradient w.r.t. the layer output (unpadded)
//   ga = g * act'(z),  z = scale*y + shift               (ReLU / LeakyReLU / none; recomputed from y)
//   reduce: bstats[g][c] += (sum ga, sum ga*xhat)
//   apply : dy = rstd*gamma*(ga - mean(ga) - xhat*mean(ga*xhat))   (norm none: dy = ga)
// ------------------------------------------------------------------------------------------------
struct NormBwdParams {
  View y, dout, dskip, dy, gsum;
  int H, W, C, vt;
  int rows_per_block;
  int norm, act, pad, has_dout, has_dskip, write_gsum, use_running;
  int pre_act;    // ACT_FIRST: y is act(conv); the result is multiplied by act'(y)
  int accum_f32;  // dy is an fp32 view that receives +=
  float slope, eps, inv_count;
  int per_image;
  const float* stats;
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  float* bstats;  // [g][c][2]
};

// Adds to g the reflect images of (h, w) other than (h, w) itself (border pixels only).
template <typename T>
__device__ __forceinline__ void add_folded_extras(const NormBwdParams& p, const T* db, int h, int w,
                                                  float* g) {
  const int pad = p.pad, H = p.H, W = p.W;
  int hh[3], ww[3];
  int nh = 0, nw = 0;
  hh[nh++] = h;
  ww[nw++] = w;
  if (h >= 1 && h <= pad) hh[nh++] = -h;
  if (h <= H - 2 && h >= H - 1 - pad) hh[nh++] = 2 * (H - 1) - h;
  if (w >= 1 && w <= pad) ww[nw++] = -w;
  if (w <= W - 2 && w >= W - 1 - pad) ww[nw++] = 2 * (W - 1) - w;
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      if (a == 0 && b == 0) continue;
      float t[8];
      unpack8(ld16(db + hh[a] * p.dout.sh + ww[b] * p.dout.sw), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
}

// Per-channel constants: xhat = y*a1 + b1 (a1 = rstd, b1 = -mean*rstd); z = xhat*gamma + beta (AFFINE) or xhat;
// apply: dy = c1*(ga - m1 - xhat*m2) with c1 = rstd*gamma.  Row-strip mapping as in the forward kernel.
template <typename T, bool kApply, int ACT, bool AFFINE, int U>
__global__ void __launch_bounds__(256, 2) norm_act_bwd_kernel(NormBwdParams p) {
  typedef typename Vec8<T>::type V8;
  __shared__ float red[kApply ? 1 : 256 * 16];
  const int v = threadIdx.x % p.vt, lane = threadIdx.x / p.vt, lanes = 256 / p.vt;
  const int cvec = blockIdx.z * p.vt + v;
  const bool active = cvec * 8 < p.C;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * p.rows_per_block;
  const int r1 = min(p.H, r0 + p.rows_per_block);
  const int grp = p.per_image ? n : 0;
  float a1[8], b1[8], gam[AFFINE ? 8 : 1], bet[AFFINE ? 8 : 1], c1[8], m1[8], m2[8], s1[kApply ? 1 : 8], s2[kApply ? 1 : 8];
  if (!kApply) {
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[kApply ? 0 : j] = s2[kApply ? 0 : j] = 0.f;
  }
  if (active) {
    {
      float scale[8], shift[8], mean[8], rstd[8];
      norm_coeffs(p.norm, p.use_running, p.stats, p.gamma, p.beta, p.running_mean, p.running_var, grp, p.C,
                  cvec * 8, p.inv_count, p.eps, scale, shift, mean, rstd);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a1[j] = rstd[j];
        b1[j] = -mean[j] * rstd[j];
        c1[j] = scale[j];
        m1[j] = m2[j] = 0.f;
        if (AFFINE) {
          const int ch = cvec * 8 + j;
          gam[AFFINE ? j : 0] = (p.gamma != nullptr && ch < p.C) ? p.gamma[ch] : 1.f;
          bet[AFFINE ? j : 0] = (p.beta != nullptr && ch < p.C) ? p.beta[ch] : 0.f;
        }
      }
    }
    const bool batch_stats = p.norm != CDB_NORM_NONE && !p.use_running;
    if (kApply && batch_stats) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int ch = cvec * 8 + j;
        if (ch < p.C) {
          m1[j] = p.bstats[(static_cast<int64_t>(grp) * p.C + ch) * 2] * p.inv_count;
          m2[j] = p.bstats[(static_cast<int64_t>(grp) * p.C + ch) * 2 + 1] * p.inv_count;
        }
      }
    }
    const T* yb = static_cast<const T*>(p.y.ptr) + n * p.y.sn + cvec * 8;
    const T* db = p.has_dout ? static_cast<const T*>(p.dout.ptr) + n * p.dout.sn + cvec * 8 : nullptr;
    const T* sb = p.has_dskip ? static_cast<const T*>(p.dskip.ptr) + n * p.dskip.sn + cvec * 8 : nullptr;
    const int pad = p.pad, H = p.H, W = p.W;
    const int ysh = static_cast<int>(p.y.sh), ysw = static_cast<int>(p.y.sw);
    const int dsh = static_cast<int>(p.dout.sh), dsw = static_cast<int>(p.dout.sw);
    const int ssh = static_cast<int>(p.dskip.sh), ssw = static_cast<int>(p.dskip.sw);
    const int gsh = static_cast<int>(p.gsum.sh), gsw = static_cast<int>(p.gsum.sw);
    const int osh = static_cast<int>(p.dy.sh), osw = static_cast<int>(p.dy.sw);
    T* gb = p.write_gsum ? static_cast<T*>(p.gsum.ptr) + n * p.gsum.sn + cvec * 8 : nullptr;
    T* ob = kApply ? static_cast<T*>(p.dy.ptr) + n * p.dy.sn + cvec * 8 : nullptr;
    const bool has_dout = p.has_dout, has_dskip = p.has_dskip;
    const int pre_act = p.pre_act;
    const float slope = p.slope;
    const bool plain_norm = p.norm == CDB_NORM_NONE, running = p.use_running != 0;
    for (int h = r0; h < r1; ++h) {
      const bool hborder = pad > 0 && (h <= pad || h >= H - 1 - pad);
      for (int w0 = lane; w0 < W; w0 += lanes * U) {
        V8 yr[U], dr[U], sr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * lanes;
          if (w < W) {
            yr[u] = ld16(yb + h * ysh + w * ysw);
            if (has_dout) dr[u] = ld16(db + h * dsh + w * dsw);
            if (has_dskip) sr[u] = ld16(sb + h * ssh + w * ssw);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * lanes;
          if (w >= W) break;
          float g[8], f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] = 0.f;
          if (has_dout) {
            unpack8(dr[u], g);
            if (pad > 0 && (hborder || w <= pad || w >= W - 1 - pad)) add_folded_extras<T>(p, db, h, w, g);
          }
          if (has_dskip) {
            float t[8];
            unpack8(sr[u], t);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] += t[j];
          }
          unpack8(yr[u], f);
          if (kApply && gb != nullptr) st16(gb + h * gsh + w * gsw, pack8<T>(g));
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xhat = f[j] * a1[j] + b1[j];
            float ga = g[j];
            if (ACT != CDB_ACT_NONE) {
              const float z = AFFINE ? xhat * gam[AFFINE ? j : 0] + bet[AFFINE ? j : 0] : xhat;
              if (ACT == CDB_ACT_RELU) ga = z > 0.f ? ga : 0.f;
              else ga = z > 0.f ? ga : ga * slope;
            }
            if (kApply) {
              float r = plain_norm ? ga : running ? ga * c1[j] : c1[j] * (ga - m1[j] - xhat * m2[j]);
              if (pre_act == CDB_ACT_RELU) r = f[j] > 0.f ? r : 0.f;
              else if (pre_act == CDB_ACT_LEAKY) r = f[j] > 0.f ? r : r * slope;
              o[j] = r;
            } else {
              s1[kApply ? 0 : j] += ga;
              s2[kApply ? 0 : j] += ga * xhat;
            }
          }
          if (kApply) {
            if (p.accum_f32) {
              float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.dy.ptr) + n * p.dy.sn + cvec * 8 +
                                                     static_cast<int64_t>(h) * p.dy.sh + static_cast<int64_t>(w) * p.dy.sw);
              float4 qa = o4[0], qb = o4[1];
              qa.x += o[0]; qa.y += o[1]; qa.z += o[2]; qa.w += o[3];
              qb.x += o[4]; qb.y += o[5]; qb.z += o[6]; qb.w += o[7];
              o4[0] = qa;
              o4[1] = qb;
            } else {
              st16(ob + h * osh + w * osw, pack8<T>(o));
            }
          }
        }
      }
    }
  }
  if (!kApply) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[(kApply ? 0 : (j * 2) * 256) + (kApply ? 0 : threadIdx.x)] = s1[kApply ? 0 : j];
      red[(kApply ? 0 : (j * 2 + 1) * 256) + (kApply ? 0 : threadIdx.x)] = s2[kApply ? 0 : j];
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.vt * 16; t += 256) {
      const int vv = t % p.vt, comp = t / p.vt;
      float acc = 0.f;
      for (int l = 0; l < lanes; ++l) acc += red[comp * 256 + l * p.vt + vv];
      const int ch = (blockIdx.z * p.vt + vv) * 8 + comp / 2;
      if (ch < p.C) atomicAdd(p.bstats + (static_cast<int64_t>(grp) * p.C + ch) * 2 + (comp & 1), acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused InstanceNorm backward for images of <= 4096 pixels (the 64x64 residual-block layers and the
// PatchGAN layers): ONE pass over HBM instead of reduce + apply.
//   A thread-block cluster of 8 CTAs owns (image n, 32 channels); CTA r owns 1/8 of the pixels and keeps
//   its y and g = fold(dout) + dskip vectors in REGISTERS (8 x 16 B each per thread, all loads issued
//   up-front).  Per-channel sums are reduced inside the CTA through shared memory and across the cluster
//   through distributed shared memory; the apply phase then runs from registers.
// DRAM traffic = y + dout (+ dskip) read once + dy (+ gsum) written once = the algorithmic minimum (the
// two-kernel form re-reads y and dout from DRAM in its second pass, profiles/r01_ncu_full_kernels_r01c.json).
// ------------------------------------------------------------------------------------------------
constexpr int kFusedCluster = 8;
constexpr int kFusedVecs = 8;  // pixel vectors per thread

// add_folded_extras with 32-bit element offsets and without the dynamically indexed coordinate arrays (which live in
// local memory): same mirrored positions, same order of additions.
__device__ __forceinline__ void fold_extras_i32(const __nv_bfloat16* db, int dsh, int dsw, int h, int w, int H, int W,
                                                int pad, float* g) {
  constexpr int kNone = -(1 << 30);
  const int h1 = (h >= 1 && h <= pad) ? -h : kNone;
  const int h2 = (h <= H - 2 && h >= H - 1 - pad) ? 2 * (H - 1) - h : kNone;
  const int w1 = (w >= 1 && w <= pad) ? -w : kNone;
  const int w2 = (w <= W - 2 && w >= W - 1 - pad) ? 2 * (W - 1) - w : kNone;
  auto acc = [&](int hh, int ww) {
    if (hh == kNone || ww == kNone) return;
    float t[8];
    unpack8(ld16(db + hh * dsh + ww * dsw), t);
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] += t[j];
  };
  acc(h, w1);
  acc(h, w2);
  acc(h1, w);
  acc(h1, w1);
  acc(h1, w2);
  acc(h2, w);
  acc(h2, w1);
  acc(h2, w2);
}

// (h, w) of the pixels p, p + 64, p + 128, ... of a W-wide image with ONE division per thread
struct Walk64 {
  int h, w, W, qstep, rstep;
  __device__ __forceinline__ void init(int px, int W_) {
    W = W_;
    h = px / W_;
    w = px - h * W_;
    qstep = 64 / W_;
    rstep = 64 - qstep * W_;
  }
  __device__ __forceinline__ void next() {
    w += rstep;
    h += qstep;
    if (w >= W) {
      w -= W;
      ++h;
    }
  }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <int ACT, bool ASYNC, int OCC>
__global__ void __cluster_dims__(kFusedCluster, 1, 1) __launch_bounds__(256, OCC)
norm_bwd_fused_in_kernel(NormBwdParams p, int ppc) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ uint4 fused_smem[];       // y vectors [8][256] then g vectors [8][256] (thread-private slots)
  uint4* sy = fused_smem;
  uint4* sg = fused_smem + kFusedVecs * 256;
  __shared__ float red[8 * 64];  // per-warp partials [warp][vector][channel][2]
  __shared__ float part[64];     // this CTA's partial (s1, s2) for its 32 channels: [vector][channel][2]
  __shared__ float tot[64];
  const int v = threadIdx.x & 3, lane = threadIdx.x >> 2;   // 4 channel vectors x 64 pixel lanes
  const int rank = static_cast<int>(cluster.block_rank());
  const int cvec = blockIdx.y * 4 + v;
  const int n = blockIdx.z;
  const int pixels = p.H * p.W;
  const int p0 = rank * ppc, p1 = min(pixels, p0 + ppc);
  const int W = p.W, H = p.H, pad = p.pad;
  // 32-bit element offsets inside one image (fused_in_eligible bounds H * stride) and a division-free pixel walk
  const int ysh = static_cast<int>(p.y.sh), ysw = static_cast<int>(p.y.sw);
  const int dsh = static_cast<int>(p.dout.sh), dsw = static_cast<int>(p.dout.sw);
  const int ssh = static_cast<int>(p.dskip.sh), ssw = static_cast<int>(p.dskip.sw);
  const int gsh = static_cast<int>(p.gsum.sh), gsw = static_cast<int>(p.gsum.sw);
  const int osh = static_cast<int>(p.dy.sh), osw = static_cast<int>(p.dy.sw);
  Walk64 first;
  first.init(p0 + lane, W);
  const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(p.y.ptr) + n * p.y.sn + cvec * 8;
  const __nv_bfloat16* db = p.has_dout ? static_cast<const __nv_bfloat16*>(p.dout.ptr) + n * p.dout.sn + cvec * 8 : nullptr;
  const __nv_bfloat16* sb = p.has_dskip ? static_cast<const __nv_bfloat16*>(p.dskip.ptr) + n * p.dskip.sn + cvec * 8 : nullptr;
  if (ASYNC) {
    // all 16 y / dout vectors of the thread go global -> shared memory (its private slots) without passing through
    // registers: 16 copies of 16 bytes in flight per thread = 64 KB per CTA from the first instruction on (the
    // per-channel coefficients are fetched behind them)
    Walk64 pw = first;
#pragma unroll
    for (int u = 0; u < kFusedVecs; ++u, pw.next()) {
      const int px = p0 + lane + u * 64;
      if (px < p1) {
        cp_async16(&sy[u * 256 + threadIdx.x], yb + pw.h * ysh + pw.w * ysw);
        if (p.has_dout) cp_async16(&sg[u * 256 + threadIdx.x], db + pw.h * dsh + pw.w * dsw);
      }
    }
  }
  float rstd[8], b1[8];   // xhat = y * rstd + b1, b1 = -mean * rstd
  {
    float scale[8], shift[8], mean[8];
    norm_coeffs(p.norm, 0, p.stats, nullptr, nullptr, nullptr, nullptr, n, p.C, cvec * 8, p.inv_count, p.eps,
                scale, shift, mean, rstd);
#pragma unroll
    for (int j = 0; j < 8; ++j) b1[j] = -mean[j] * rstd[j];
  }
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  // two batches of four pixels: 8-12 independent 16-byte loads in flight per thread
  Walk64 pl = first, pc = first;   // load and compute positions
#pragma unroll
  for (int ub = 0; ub < kFusedVecs; ub += 4) {
    uint4 yr[4], gr[4], sr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k, pl.next()) {
      const int px = p0 + lane + (ub + k) * 64;
      if (px < p1) {
        if (!ASYNC) {
          yr[k] = ld16(yb + pl.h * ysh + pl.w * ysw);
          if (p.has_dout) gr[k] = ld16(db + pl.h * dsh + pl.w * dsw);
        }
        if (p.has_dskip) sr[k] = ld16(sb + pl.h * ssh + pl.w * ssw);
      }
    }
    if (ASYNC) {
      if (ub == 0) cp_async_wait_all();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        yr[k] = sy[(ub + k) * 256 + threadIdx.x];
        if (p.has_dout) gr[k] = sg[(ub + k) * 256 + threadIdx.x];
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k, pc.next()) {
    const int u = ub + k;
    const int px = p0 + lane + u * 64;
    if (px >= p1) continue;
    const int h = pc.h, w = pc.w;
    float g[8], f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.f;
    if (p.has_dout) {
      unpack8(gr[k], g);
      if (pad > 0 && (h <= pad || h >= H - 1 - pad || w <= pad || w >= W - 1 - pad))
        fold_extras_i32(db, dsh, dsw, h, w, H, W, pad, g);
    }
    if (p.has_dskip) {
      float t[8];
      unpack8(sr[k], t);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += t[j];
    }
    sg[u * 256 + threadIdx.x] = pack8<__nv_bfloat16>(g);   // the summed gradient (rounded to bf16 once) for the apply phase / gsum
    if (!ASYNC) sy[u * 256 + threadIdx.x] = yr[k];
    unpack8(yr[k], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = fmaf(f[j], rstd[j], b1[j]);
      float ga = g[j];
      if (ACT == CDB_ACT_RELU) ga = xhat > 0.f ? ga : 0.f;
      else if (ACT == CDB_ACT_LEAKY) ga = xhat > 0.f ? ga : ga * p.slope;
      s1[j] += ga;
      s2[j] = fmaf(ga, xhat, s2[j]);
    }
    }
  }
  // ---- CTA reduction over the 64 pixel lanes: a warp holds 8 pixel lanes x 4 channel vectors (lane bits 2..4 are
  // the pixel lane), so three shuffle steps leave the warp's sums in its first four lanes; 8 warps x 64 values then
  // go through 2 KB of shared memory
#pragma unroll
  for (int j = 0; j < 8; ++j) {
#pragma unroll
    for (int off = 4; off < 32; off <<= 1) {
      s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], off);
      s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], off);
    }
  }
  if ((threadIdx.x & 31) < 4) {
    float* dst = red + (threadIdx.x >> 5) * 64 + v * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      dst[j * 2] = s1[j];
      dst[j * 2 + 1] = s2[j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float acc = 0.f;
#pragma unroll
    for (int wq = 0; wq < 8; ++wq) acc += red[wq * 64 + threadIdx.x];
    part[threadIdx.x] = acc;
  }
  cluster.sync();
  // ---- cluster reduction through distributed shared memory
  if (threadIdx.x < 64) {
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < kFusedCluster; ++r) acc += cluster.map_shared_rank(part, r)[threadIdx.x];
    tot[threadIdx.x] = acc * p.inv_count;
  }
  __syncthreads();
  float m1[8], m2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m1[j] = tot[v * 16 + j * 2];
    m2[j] = tot[v * 16 + j * 2 + 1];
  }
  // ---- apply from registers
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(p.dy.ptr) + n * p.dy.sn + cvec * 8;
  __nv_bfloat16* gb = p.write_gsum ? static_cast<__nv_bfloat16*>(p.gsum.ptr) + n * p.gsum.sn + cvec * 8 : nullptr;
  Walk64 pa = first;
#pragma unroll
  for (int u = 0; u < kFusedVecs; ++u, pa.next()) {
    const int px = p0 + lane + u * 64;
    if (px >= p1) continue;
    const int h = pa.h, w = pa.w;
    float g[8], f[8], o[8];
    const uint4 gv = sg[u * 256 + threadIdx.x];
    unpack8(gv, g);
    unpack8(sy[u * 256 + threadIdx.x], f);
    if (gb != nullptr) st16(gb + h * gsh + w * gsw, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xhat = fmaf(f[j], rstd[j], b1[j]);
      float ga = g[j];
      if (ACT == CDB_ACT_RELU) ga = xhat > 0.f ? ga : 0.f;
      else if (ACT == CDB_ACT_LEAKY) ga = xhat > 0.f ? ga : ga * p.slope;
      o[j] = rstd[j] * (ga - m1[j] - xhat * m2[j]);
    }
    st16(ob + h * osh + w * osw, pack8<__nv_bfloat16>(o));
  }
  cluster.sync();   // keeps every CTA's `part` alive until all ranks have read it (off the critical path)
}

// ------------------------------------------------------------------------------------------------
// Streaming two-pass backward for the common case (bf16, batch statistics, no affine parameters, activation NONE /
// ReLU / LeakyReLU after the norm): the instruction-lean form of norm_act_bwd_kernel.
//   * the reduce pass accumulates  sum ga  and  sum ga * y  (not ga * xhat): no xhat evaluation per element; the
//     finishing threads convert  sum ga * xhat = rstd * (sum ga * y - mean * sum ga)  once per block (linear);
//   * the activation mask is  y > mean  (rstd > 0), so the reduce pass needs ONE coefficient per channel;
//   * the apply pass is  dy = ga * A + y * B + D  with A = rstd, B = -rstd^2 m2, D = rstd^2 m2 mean - rstd m1;
//   * fp32 arithmetic in register PAIRS (fma.rn.f32x2 / add.rn.f32x2 of sm_100: two lanes per instruction).
// ~45 instead of ~70 instructions per 16-byte vector and <= 80 registers, i.e. 3 resident blocks per SM: the old
// kernels (and the cluster-fused one) are bound by instruction latency at 16 warps per SM
// (profiles/r01_norm_bwd_experiments.txt).
// ------------------------------------------------------------------------------------------------
// bf16 pair -> fp32 pair in two integer instructions (shift, mask): the fp32 bits of a bf16 are its 16 bits followed by
// zeros
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t x) {
  return make_float2(__uint_as_float(x << 16), __uint_as_float(x & 0xffff0000u));
}
__device__ __forceinline__ void unpack8_pairs(const uint4& r, float2* f) {
  f[0] = bf16x2_to_float2(r.x);
  f[1] = bf16x2_to_float2(r.y);
  f[2] = bf16x2_to_float2(r.z);
  f[3] = bf16x2_to_float2(r.w);
}
// Bits of the smallest bf16 strictly greater than m: for a bf16 value y,  y > m  <=>  y >= bf16_above(m)  exactly, so
// the ReLU mask of four channel pairs is four packed bf16 comparisons instead of eight fp32 compare + select pairs.
__device__ __forceinline__ uint32_t bf16_above(float m) {
  const __nv_bfloat16 c = __float2bfloat16_ru(m);   // smallest bf16 >= m
  uint32_t b = __bfloat16_as_ushort(c);
  if (__bfloat162float(c) == m) {                   // m is a bf16 value itself: the next one up
    if (m == 0.f) b = 0x0001u;
    else if (m > 0.f) b += 1u;
    else b -= 1u;
  }
  return b;
}
__device__ __forceinline__ uint32_t bf16x2_ge_mask(uint32_t y, uint32_t thr) {
  return __hge2_mask(*reinterpret_cast<const __nv_bfloat162*>(&y), *reinterpret_cast<const __nv_bfloat162*>(&thr));
}
__device__ __forceinline__ uint4 pack8_pairs(const float2* f) {
  uint4 r;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __float22bfloat162_rn(f[i]);
  return r;
}

template <int ACT>
__device__ __forceinline__ float2 mask_pair(float2 g, float2 y, float2 mean, float slope) {
  if (ACT == CDB_ACT_RELU) return make_float2(y.x > mean.x ? g.x : 0.f, y.y > mean.y ? g.y : 0.f);
  if (ACT == CDB_ACT_LEAKY) return make_float2(y.x > mean.x ? g.x : g.x * slope, y.y > mean.y ? g.y : g.y * slope);
  return g;
}

// g (pairs) += the reflect images of (h, w) other than (h, w) itself
__device__ __forceinline__ void fold_extras_pairs(const __nv_bfloat16* db, int dsh, int dsw, int h, int w, int H, int W,
                                                  int pad, float2* g) {
  float t[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) t[j] = 0.f;
  fold_extras_i32(db, dsh, dsw, h, w, H, W, pad, t);
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = __fadd2_rn(g[i], make_float2(t[2 * i], t[2 * i + 1]));
}

template <bool kApply, int ACT, bool HAS_SKIP, bool WRITE_GSUM, int OCC>
__global__ void __launch_bounds__(256, OCC) norm_bwd_stream_kernel(NormBwdParams p) {
  __shared__ float red[kApply ? 1 : 256 * 16];
  const int v = threadIdx.x % p.vt, lane = threadIdx.x / p.vt, lanes = 256 / p.vt;
  // The apply pass walks the tensor in the REVERSE block order of the reduce pass: what the reduce pass read last is
  // still in L2 (y + dout of a 16-image residual layer are 69 MB; a second cyclic sweep in the same order would find
  // every line already evicted).
  const int bx = kApply ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
  const int by = kApply ? gridDim.y - 1 - blockIdx.y : blockIdx.y;
  const int bz = kApply ? gridDim.z - 1 - blockIdx.z : blockIdx.z;
  const int cvec = bz * p.vt + v;
  const bool active = cvec * 8 < p.C;
  const int n = by;
  const int r0 = bx * p.rows_per_block;
  const int r1 = min(p.H, r0 + p.rows_per_block);
  const int grp = p.per_image ? n : 0;
  float2 mean2[4], A2[kApply ? 4 : 1], B2[kApply ? 4 : 1], D2[kApply ? 4 : 1];
  float2 sg[kApply ? 1 : 4], sgy[kApply ? 1 : 4];
  if (!kApply) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[kApply ? 0 : i] = sgy[kApply ? 0 : i] = make_float2(0.f, 0.f);
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mean[2], a[2], b[2], dd[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int ch = cvec * 8 + 2 * i + e;
        mean[e] = 0.f;
        a[e] = 1.f;
        b[e] = dd[e] = 0.f;
        if (ch < p.C) {
          const float s1 = p.stats[(static_cast<int64_t>(grp) * p.C + ch) * 2];
          const float s2 = p.stats[(static_cast<int64_t>(grp) * p.C + ch) * 2 + 1];
          mean[e] = s1 * p.inv_count;
          const float var = fmaxf(s2 * p.inv_count - mean[e] * mean[e], 0.f);
          const float rstd = rsqrtf(var + p.eps);
          if (kApply) {
            const float m1 = p.bstats[(static_cast<int64_t>(grp) * p.C + ch) * 2] * p.inv_count;
            const float m2 = p.bstats[(static_cast<int64_t>(grp) * p.C + ch) * 2 + 1] * p.inv_count;
            a[e] = rstd;
            b[e] = -rstd * rstd * m2;
            dd[e] = rstd * rstd * m2 * mean[e] - rstd * m1;
          }
        }
      }
      mean2[i] = make_float2(mean[0], mean[1]);
      if (kApply) {
        A2[kApply ? i : 0] = make_float2(a[0], a[1]);
        B2[kApply ? i : 0] = make_float2(b[0], b[1]);
        D2[kApply ? i : 0] = make_float2(dd[0], dd[1]);
      }
    }
    const __nv_bfloat16* yb = static_cast<const __nv_bfloat16*>(p.y.ptr) + n * p.y.sn + cvec * 8;
    const __nv_bfloat16* db = static_cast<const __nv_bfloat16*>(p.dout.ptr) + n * p.dout.sn + cvec * 8;
    const __nv_bfloat16* sb = HAS_SKIP ? static_cast<const __nv_bfloat16*>(p.dskip.ptr) + n * p.dskip.sn + cvec * 8 : nullptr;
    __nv_bfloat16* ob = kApply ? static_cast<__nv_bfloat16*>(p.dy.ptr) + n * p.dy.sn + cvec * 8 : nullptr;
    __nv_bfloat16* gb = (kApply && WRITE_GSUM) ? static_cast<__nv_bfloat16*>(p.gsum.ptr) + n * p.gsum.sn + cvec * 8 : nullptr;
    const int pad = p.pad, H = p.H, W = p.W;
    const int ysh = static_cast<int>(p.y.sh), ysw = static_cast<int>(p.y.sw);
    const int dsh = static_cast<int>(p.dout.sh), dsw = static_cast<int>(p.dout.sw);
    const int ssh = static_cast<int>(p.dskip.sh), ssw = static_cast<int>(p.dskip.sw);
    const int gsh = static_cast<int>(p.gsum.sh), gsw = static_cast<int>(p.gsum.sw);
    const int osh = static_cast<int>(p.dy.sh), osw = static_cast<int>(p.dy.sw);
    const float slope = p.slope;
    constexpr int U = OCC >= 3 ? 2 : 4;   // pixels in flight per thread: 3 blocks x 2 or 2 blocks x 4
    for (int h = r0; h < r1; ++h) {
      const bool hborder = pad > 0 && (h <= pad || h >= H - 1 - pad);
      for (int w0 = lane; w0 < W; w0 += lanes * U) {
        uint4 yr[U], dr[U], sr[HAS_SKIP ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * lanes;
          if (w < W) {
            yr[u] = ld16(yb + h * ysh + w * ysw);
            dr[u] = ld16(db + h * dsh + w * dsw);
            if (HAS_SKIP) sr[HAS_SKIP ? u : 0] = ld16(sb + h * ssh + w * ssw);
          }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int w = w0 + u * lanes;
          if (w >= W) break;
          float2 y2[4], g2[4];
          unpack8_pairs(yr[u], y2);
          unpack8_pairs(dr[u], g2);
          if (pad > 0 && (hborder || w <= pad || w >= W - 1 - pad)) fold_extras_pairs(db, dsh, dsw, h, w, H, W, pad, g2);
          if (HAS_SKIP) {
            float2 t2[4];
            unpack8_pairs(sr[HAS_SKIP ? u : 0], t2);
#pragma unroll
            for (int i = 0; i < 4; ++i) g2[i] = __fadd2_rn(g2[i], t2[i]);
          }
          if (kApply) {
            if (WRITE_GSUM) st16(gb + h * gsh + w * gsw, pack8_pairs(g2));
            float2 o2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 ga = mask_pair<ACT>(g2[i], y2[i], mean2[i], slope);
              o2[i] = __ffma2_rn(ga, A2[kApply ? i : 0], __ffma2_rn(y2[i], B2[kApply ? i : 0], D2[kApply ? i : 0]));
            }
            st16(ob + h * osh + w * osw, pack8_pairs(o2));
          } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 ga = mask_pair<ACT>(g2[i], y2[i], mean2[i], slope);
              sg[kApply ? 0 : i] = __fadd2_rn(sg[kApply ? 0 : i], ga);
              sgy[kApply ? 0 : i] = __ffma2_rn(ga, y2[i], sgy[kApply ? 0 : i]);
            }
          }
        }
      }
    }
  }
  if (!kApply) {
    // block reduction over the pixel lanes, then  (sum ga, rstd * (sum ga y - mean sum ga))  per channel
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      red[((2 * i) * 2) * 256 + threadIdx.x] = sg[kApply ? 0 : i].x;
      red[((2 * i) * 2 + 1) * 256 + threadIdx.x] = sgy[kApply ? 0 : i].x;
      red[((2 * i + 1) * 2) * 256 + threadIdx.x] = sg[kApply ? 0 : i].y;
      red[((2 * i + 1) * 2 + 1) * 256 + threadIdx.x] = sgy[kApply ? 0 : i].y;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < p.vt * 8; t += 256) {
      const int vv = t % p.vt, j = t / p.vt;     // channel j of vector vv
      float a_g = 0.f, a_gy = 0.f;
      for (int l = 0; l < lanes; ++l) {
        a_g += red[(j * 2) * 256 + l * p.vt + vv];
        a_gy += red[(j * 2 + 1) * 256 + l * p.vt + vv];
      }
      const int ch = (blockIdx.z * p.vt + vv) * 8 + j;
      if (ch < p.C) {
        const float s1 = p.stats[(static_cast<int64_t>(grp) * p.C + ch) * 2];
        const float s2 = p.stats[(static_cast<int64_t>(grp) * p.C + ch) * 2 + 1];
        const float mean = s1 * p.inv_count;
        const float rstd = rsqrtf(fmaxf(s2 * p.inv_count - mean * mean, 0.f) + p.eps);
        atomicAdd(p.bstats + (static_cast<int64_t>(grp) * p.C + ch) * 2, a_g);
        atomicAdd(p.bstats + (static_cast<int64_t>(grp) * p.C + ch) * 2 + 1, rstd * (a_gy - mean * a_g));
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged form of the streaming backward.  The register-load kernel above exposes one DRAM round trip per loop
// iteration of every warp (8 iterations per thread on the residual-block shape) and one more for the reflect-fold
// reads of every border pixel: 2.5-3.2 TB/s of 6.5 on the padded layers against 4.8 TB/s on the unpadded ones.
// Here the operands are contiguous row segments (every view has pixel stride == C), so ONE thread streams them with
// cp.async.bulk into a ring of shared-memory stages, `stages` chunks ahead of the eight warps, which read 16-byte
// vectors from shared memory (thread t owns vectors t, t + 256 of a chunk: conflict-free, and t % cv is the thread's
// channel vector for the whole kernel) and write dy / gsum straight from registers.
//   * chunk = kTmaV * 256 / cv pixels of one image row (8 KB per operand); the dout segment is staged together with
//     its `pad` halo pixels on either side, so the column images of the reflect fold are shared-memory reads; only
//     the 2 * pad mirror rows of an image still come from global memory;
//   * a block owns a contiguous chunk range of ONE image, the grid is (blocks per image, images), ~2 blocks per SM;
//   * the apply pass walks its range backwards: what the reduce pass read last is still in L2;
//   * same arithmetic as norm_bwd_stream_kernel (sum ga / sum ga*y, y > mean mask, 3-coefficient apply, f32x2 pairs).
// ------------------------------------------------------------------------------------------------
constexpr int kTmaMaxStages = 8;

struct TmaGeom {
  int cpr;           // chunks per image row
  int stages;
  int off_d, off_m, off_s;  // byte offsets of the dout / mirror-row / dskip slots inside a stage (y at 0)
  int stage_bytes;
  int hint;
};

__device__ __forceinline__ void add_vec_pairs(uint32_t saddr, float2* g) {
  float2 t[4];
  unpack8_pairs(ld_shared_v4(saddr), t);
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = __fadd2_rn(g[i], t[i]);
}
__device__ __forceinline__ void add_vec_pairs(const __nv_bfloat16* gptr, float2* g) {
  float2 t[4];
  unpack8_pairs(ld16(gptr), t);
#pragma unroll
  for (int i = 0; i < 4; ++i) g[i] = __fadd2_rn(g[i], t[i]);
}

// V = 16-byte vectors per thread, operand and chunk (chunk = V * 4 KB per operand)
template <bool kApply, int ACT, bool HAS_SKIP, bool WRITE_GSUM, int V>
__global__ void __launch_bounds__(256, 2) norm_bwd_tma_kernel(NormBwdParams p, TmaGeom tg, int* abort_global) {
  extern __shared__ __align__(128) uint8_t tma_smem[];
  __shared__ uint64_t bars[2 * kTmaMaxStages];   // full[s] = bars[s], empty[s] = bars[kTmaMaxStages + s]
  __shared__ int abort_smem;
  const int stages = tg.stages, cpr = tg.cpr;
  float* red = reinterpret_cast<float*>(tma_smem);   // reduce pass: 16 KB over the ring once every chunk is consumed
  const int tid = threadIdx.x;
  const int cv = p.vt;                             // channel vectors per pixel (C / 8, a power of two <= 256)
  const int lanes = 256 / cv, PX = V * lanes;      // pixels per chunk
  const int n = blockIdx.y;
  const int cpi = p.H * cpr;                       // chunks per image
  const int k0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * cpi / gridDim.x);
  const int k1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * cpi / gridDim.x);
  const int nk = k1 - k0;
  const int W = p.W, C = p.C, pad = p.pad, H = p.H;
  const int ysh = static_cast<int>(p.y.sh), dsh = static_cast<int>(p.dout.sh), ssh = static_cast<int>(p.dskip.sh);
  const __nv_bfloat16* yimg = static_cast<const __nv_bfloat16*>(p.y.ptr) + n * p.y.sn;
  const __nv_bfloat16* dimg = static_cast<const __nv_bfloat16*>(p.dout.ptr) + n * p.dout.sn;
  const __nv_bfloat16* simg = HAS_SKIP ? static_cast<const __nv_bfloat16*>(p.dskip.ptr) + n * p.dskip.sn : nullptr;
  const uint32_t stage_base = smem_u32(tma_smem);
  const uint32_t bar_base = smem_u32(&bars[0]);
  uint64_t pol = 0;
  if (tg.hint) pol = kApply ? l2_policy_evict_first() : l2_policy_evict_last();

  // the block walks its chunks k0 .. k1-1 (reduce) or k1-1 .. k0 (apply); (h, c) = (row, chunk of the row)
  const int kfirst = kApply ? k1 - 1 : k0;
  int h = kfirst / cpr, c = kfirst - h * cpr;
  int ih = h, ic = c;                              // producer position (thread 0)
  auto issue = [&](int stage) {
    const int w0 = ic * PX;
    const int npx = min(PX, W - w0);
    const uint32_t bytes = static_cast<uint32_t>(npx * C * 2);
    const int right = min(2 * pad, W + pad - (w0 + npx));           // halo / neighbour pixels staged after the segment
    const uint32_t dbytes = static_cast<uint32_t>((npx + pad + right) * C * 2);
    const uint32_t full = bar_base + stage * 8;
    const uint32_t dst = stage_base + stage * tg.stage_bytes;
    // the one row whose reflect image row ih is (rows 1..pad and H-1-pad..H-2 have one; H > 2 pad + 1)
    const int mh = pad == 0 ? 0 : ((ih >= 1 && ih <= pad) ? -ih : ((ih <= H - 2 && ih >= H - 1 - pad) ? 2 * (H - 1) - ih : 0));
    mbar_arrive_expect_tx(full, bytes * (HAS_SKIP ? 2 : 1) + dbytes * (mh != 0 ? 2 : 1));
    if (tg.hint) {
      bulk_load_1d_hint(dst, yimg + ih * ysh + w0 * C, bytes, full, pol);
      bulk_load_1d_hint(dst + tg.off_d, dimg + ih * dsh + (w0 - pad) * C, dbytes, full, pol);
      if (mh != 0) bulk_load_1d_hint(dst + tg.off_m, dimg + mh * dsh + (w0 - pad) * C, dbytes, full, pol);
      if (HAS_SKIP) bulk_load_1d_hint(dst + tg.off_s, simg + ih * ssh + w0 * C, bytes, full, pol);
    } else {
      bulk_load_1d(dst, yimg + ih * ysh + w0 * C, bytes, full);
      bulk_load_1d(dst + tg.off_d, dimg + ih * dsh + (w0 - pad) * C, dbytes, full);
      if (mh != 0) bulk_load_1d(dst + tg.off_m, dimg + mh * dsh + (w0 - pad) * C, dbytes, full);
      if (HAS_SKIP) bulk_load_1d(dst + tg.off_s, simg + ih * ssh + w0 * C, bytes, full);
    }
    if (kApply) {
      if (--ic < 0) { ic = cpr - 1; --ih; }
    } else {
      if (++ic == cpr) { ic = 0; ++ih; }
    }
  };

  int issued = 0;
  pdl_trigger();
  if (tid == 0) {
    abort_smem = 0;
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (kTmaMaxStages + s) * 8, 8);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  pdl_wait();   // the previous kernel of the stream has completed: its output may be read, ours written
  if (tid == 0)
    for (; issued < stages && issued < nk; ++issued) issue(issued);   // the ring is filled before the coefficients are read
  __syncthreads();
  volatile int* abort_flag = &abort_smem;

  const int v = tid % cv, lane = tid / cv;
  const int grp = p.per_image ? n : 0;
  float2 mean2[4], A2[kApply ? 4 : 1], B2[kApply ? 4 : 1], D2[kApply ? 4 : 1];
  float2 sg[kApply ? 1 : 4], sgy[kApply ? 1 : 4];
  // ReLU of a gradient that is still the stored bf16 value (no skip term, interior pixel): packed bf16 mask
  constexpr bool kFastMask = ACT == CDB_ACT_RELU && !HAS_SKIP;
  uint32_t thr[kFastMask ? 4 : 1];
  if (!kApply) {
#pragma unroll
    for (int i = 0; i < 4; ++i) sg[kApply ? 0 : i] = sgy[kApply ? 0 : i] = make_float2(0.f, 0.f);
  }
  {
    // norm none (a convolution + activation layer): mean 0, rstd 1, no mean terms — dy = ga, the reduce pass yields the
    // bias gradient sum ga, and the activation mask is taken from the stored OUTPUT y (act(x) > 0 <=> x > 0)
    const bool norm_none = p.norm == CDB_NORM_NONE;
    const float4* st4 = reinterpret_cast<const float4*>(p.stats + (static_cast<int64_t>(grp) * C + v * 8) * 2);
    const float4* bs4 = reinterpret_cast<const float4*>(p.bstats + (static_cast<int64_t>(grp) * C + v * 8) * 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 st = make_float4(0.f, 0.f, 0.f, 0.f);   // (s1, s2) of channels 2i, 2i + 1
      float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!norm_none) st = st4[i];
      if (kApply && !norm_none) bs = bs4[i];
      const float s1[2] = {st.x, st.z}, s2[2] = {st.y, st.w}, b1[2] = {bs.x, bs.z}, b2[2] = {bs.y, bs.w};
      float mean[2], a[2], b[2], dd[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        mean[e] = s1[e] * p.inv_count;
        const float var = fmaxf(s2[e] * p.inv_count - mean[e] * mean[e], 0.f);
        const float rstd = norm_none ? 1.f : rsqrtf(var + p.eps);
        const float m1 = b1[e] * p.inv_count, m2 = b2[e] * p.inv_count;
        a[e] = rstd;
        b[e] = -rstd * rstd * m2;
        dd[e] = rstd * rstd * m2 * mean[e] - rstd * m1;
      }
      mean2[i] = make_float2(mean[0], mean[1]);
      if (kFastMask) thr[kFastMask ? i : 0] = bf16_above(mean[0]) | (bf16_above(mean[1]) << 16);
      if (kApply) {
        A2[kApply ? i : 0] = make_float2(a[0], a[1]);
        B2[kApply ? i : 0] = make_float2(b[0], b[1]);
        D2[kApply ? i : 0] = make_float2(dd[0], dd[1]);
      }
    }
  }
  __nv_bfloat16* ob = kApply ? static_cast<__nv_bfloat16*>(p.dy.ptr) + n * p.dy.sn + v * 8 : nullptr;
  __nv_bfloat16* gb = (kApply && WRITE_GSUM) ? static_cast<__nv_bfloat16*>(p.gsum.ptr) + n * p.gsum.sn + v * 8 : nullptr;
  const int gsh = static_cast<int>(p.gsum.sh), gsw = static_cast<int>(p.gsum.sw);
  const int osh = static_cast<int>(p.dy.sh), osw = static_cast<int>(p.dy.sw);
  const float slope = p.slope;
  const uint32_t pad_bytes = static_cast<uint32_t>(pad * C * 2);
  int s = 0, ps = 0;            // stage of this chunk / of the previous one
  uint32_t ph = 0, pph = 0;     // and their phase parities
  for (int i = 0; i < nk; ++i) {
    if (tid == 0 && i >= 1 && issued < nk) {
      // refill the stage the previous chunk was read from, once all eight warps have released it
      if (mbar_wait(bar_base + (kTmaMaxStages + ps) * 8, pph, abort_flag)) {
        issue(ps);
        ++issued;
      }
    }
    const int w0 = c * PX;
    const int npx = min(PX, W - w0);
    if (!mbar_wait(bar_base + s * 8, ph, abort_flag)) break;
    const uint32_t sbase = stage_base + s * tg.stage_bytes;
    const uint32_t src = sbase + tid * 16;
    const bool hmirror = pad > 0 && ((h >= 1 && h <= pad) || (h <= H - 2 && h >= H - 1 - pad));
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const int px = lane + u * lanes;
      if (px >= npx) break;
      const int w = w0 + px;
      const uint4 yv = ld_shared_v4(src + u * 4096);
      const uint4 dv = ld_shared_v4(src + tg.off_d + pad_bytes + u * 4096);
      const bool border = pad > 0 && (hmirror || w <= pad || w >= W - 1 - pad);
      float2 y2[4], ga2[4];
      uint4 gv = dv;                       // the summed gradient, bf16 (gsum)
      unpack8_pairs(yv, y2);
      if (kFastMask && !border) {
        uint4 m;
        m.x = dv.x & bf16x2_ge_mask(yv.x, thr[0]);
        m.y = dv.y & bf16x2_ge_mask(yv.y, thr[kFastMask ? 1 : 0]);
        m.z = dv.z & bf16x2_ge_mask(yv.z, thr[kFastMask ? 2 : 0]);
        m.w = dv.w & bf16x2_ge_mask(yv.w, thr[kFastMask ? 3 : 0]);
        unpack8_pairs(m, ga2);
      } else {
        float2 g2[4];
        unpack8_pairs(dv, g2);
        if (HAS_SKIP) add_vec_pairs(src + tg.off_s + u * 4096, g2);
        if (border) {
          constexpr int kNone = -(1 << 30);
          const int w1 = (w >= 1 && w <= pad) ? -w : kNone;
          const int w2 = (w <= W - 2 && w >= W - 1 - pad) ? 2 * (W - 1) - w : kNone;
          // column images: inside the staged segment (position of pixel ww: ww - (w0 - pad))
          const uint32_t dslot = sbase + tg.off_d + v * 16;
          if (w1 != kNone) add_vec_pairs(dslot + static_cast<uint32_t>((w1 - w0 + pad) * C * 2), g2);
          if (w2 != kNone) add_vec_pairs(dslot + static_cast<uint32_t>((w2 - w0 + pad) * C * 2), g2);
          if (hmirror) {   // the mirror row, staged with the same extents
            const uint32_t mslot = sbase + tg.off_m + v * 16;
            add_vec_pairs(mslot + static_cast<uint32_t>((w - w0 + pad) * C * 2), g2);
            if (w1 != kNone) add_vec_pairs(mslot + static_cast<uint32_t>((w1 - w0 + pad) * C * 2), g2);
            if (w2 != kNone) add_vec_pairs(mslot + static_cast<uint32_t>((w2 - w0 + pad) * C * 2), g2);
          }
        }
        if (kApply && WRITE_GSUM && (HAS_SKIP || border)) gv = pack8_pairs(g2);
#pragma unroll
        for (int q = 0; q < 4; ++q) ga2[q] = mask_pair<ACT>(g2[q], y2[q], mean2[q], slope);
      }
      if (kApply) {
        if (WRITE_GSUM) st16(gb + h * gsh + w * gsw, gv);
        float2 o2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o2[q] = __ffma2_rn(ga2[q], A2[kApply ? q : 0], __ffma2_rn(y2[q], B2[kApply ? q : 0], D2[kApply ? q : 0]));
        st16(ob + h * osh + w * osw, pack8_pairs(o2));
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          sg[kApply ? 0 : q] = __fadd2_rn(sg[kApply ? 0 : q], ga2[q]);
          sgy[kApply ? 0 : q] = __ffma2_rn(ga2[q], y2[q], sgy[kApply ? 0 : q]);
        }
      }
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar_base + (kTmaMaxStages + s) * 8);   // the stage may be refilled
    ps = s;
    pph = ph;
    if (++s == stages) { s = 0; ph ^= 1u; }
    if (kApply) {
      if (--c < 0) { c = cpr - 1; --h; }
    } else {
      if (++c == cpr) { c = 0; ++h; }
    }
  }
  if (!kApply) {
    __syncthreads();   // every staged chunk has been consumed: the ring is free for the block reduction
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      red[((2 * q) * 2) * 256 + tid] = sg[kApply ? 0 : q].x;
      red[((2 * q) * 2 + 1) * 256 + tid] = sgy[kApply ? 0 : q].x;
      red[((2 * q + 1) * 2) * 256 + tid] = sg[kApply ? 0 : q].y;
      red[((2 * q + 1) * 2 + 1) * 256 + tid] = sgy[kApply ? 0 : q].y;
    }
    __syncthreads();
    if (!abort_smem) {
      // block reduction over the pixel lanes, then  (sum ga, rstd * (sum ga y - mean sum ga))  per channel
      for (int t = tid; t < cv * 8; t += 256) {
        const int vv = t % cv, j = t / cv;     // channel j of vector vv
        float a_g = 0.f, a_gy = 0.f;
        for (int l = 0; l < lanes; ++l) {
          a_g += red[(j * 2) * 256 + l * cv + vv];
          a_gy += red[(j * 2 + 1) * 256 + l * cv + vv];
        }
        const int ch = vv * 8 + j;
        float mean = 0.f, rstd = 1.f;
        if (p.norm != CDB_NORM_NONE) {
          const float s1 = p.stats[(static_cast<int64_t>(grp) * C + ch) * 2];
          const float s2 = p.stats[(static_cast<int64_t>(grp) * C + ch) * 2 + 1];
          mean = s1 * p.inv_count;
          rstd = rsqrtf(fmaxf(s2 * p.inv_count - mean * mean, 0.f) + p.eps);
        }
        atomicAdd(p.bstats + (static_cast<int64_t>(grp) * C + ch) * 2, a_g);
        atomicAdd(p.bstats + (static_cast<int64_t>(grp) * C + ch) * 2 + 1, rstd * (a_gy - mean * a_g));
      }
    }
  }
  if (tid == 0 && abort_smem && abort_global) atomicExch(abort_global, 1);
}

static bool stream_eligible(const CdbNormDesc* d, const NormBwdParams& p, bool accum_f32, int dt) {
  if (getenv("CDB_NORM_BWD_IMPL") && atoi(getenv("CDB_NORM_BWD_IMPL")) == 0) return false;   // 0: old kernels
  return dt == CDB_BF16 && d->norm != CDB_NORM_NONE && !d->use_running && d->gamma == nullptr && d->beta == nullptr &&
         !accum_f32 && p.pre_act == CDB_ACT_NONE && p.has_dout &&
         (p.act == CDB_ACT_NONE || p.act == CDB_ACT_RELU || p.act == CDB_ACT_LEAKY) &&
         (int64_t)(p.H + 2 * p.pad) * p.y.sh < (1 << 30) && (int64_t)(p.H + 2 * p.pad) * p.dout.sh < (1 << 30) &&
         (int64_t)(p.H + 2 * p.pad) * p.dskip.sh < (1 << 30) && (int64_t)(p.H + 2 * p.pad) * p.dy.sh < (1 << 30) &&
         (int64_t)(p.H + 2 * p.pad) * p.gsum.sh < (1 << 30);
}

template <bool kApply, int ACT, int OCC>
static void launch_stream_occ(const NormBwdParams& p, dim3 grid, cudaStream_t stream) {
  if (p.has_dskip) {
    if (p.write_gsum) norm_bwd_stream_kernel<kApply, ACT, true, true, OCC><<<grid, 256, 0, stream>>>(p);
    else norm_bwd_stream_kernel<kApply, ACT, true, false, OCC><<<grid, 256, 0, stream>>>(p);
  } else {
    if (p.write_gsum) norm_bwd_stream_kernel<kApply, ACT, false, true, OCC><<<grid, 256, 0, stream>>>(p);
    else norm_bwd_stream_kernel<kApply, ACT, false, false, OCC><<<grid, 256, 0, stream>>>(p);
  }
}
static int stream_occ() {
  static const int occ = getenv("CDB_NORM_STREAM_OCC") ? atoi(getenv("CDB_NORM_STREAM_OCC")) : 2;
  return occ;
}
template <bool kApply, int ACT>
static void launch_stream_act(const NormBwdParams& p, dim3 grid, cudaStream_t stream) {
  if (stream_occ() >= 3) launch_stream_occ<kApply, ACT, 3>(p, grid, stream);
  else launch_stream_occ<kApply, ACT, 2>(p, grid, stream);
}
template <bool kApply>
static void launch_stream(const NormBwdParams& p, dim3 grid, cudaStream_t stream) {
  switch (p.act) {
    case CDB_ACT_RELU: launch_stream_act<kApply, CDB_ACT_RELU>(p, grid, stream); break;
    case CDB_ACT_LEAKY: launch_stream_act<kApply, CDB_ACT_LEAKY>(p, grid, stream); break;
    default: launch_stream_act<kApply, CDB_ACT_NONE>(p, grid, stream); break;
  }
}

static int norm_bwd_impl() {
  static const int mode = getenv("CDB_NORM_BWD_IMPL") ? atoi(getenv("CDB_NORM_BWD_IMPL")) : 2;   // 2: TMA-staged
  return mode;
}
// the TMA-staged kernels need whole pixels contiguous along a row in every operand they stage, the column images of
// the reflect fold inside the first / last chunk of a row, and 16-byte aligned statistics rows
static bool tma_eligible(const NormBwdParams& p) {
  if (norm_bwd_impl() < 2) return false;
  const int cv = p.C / 8;
  if (!(p.C % 8 == 0 && p.C >= 64 && p.C <= 2048 && (cv & (cv - 1)) == 0)) return false;
  const int px = 2 * (256 / cv);
  return p.y.sw == p.C && p.dout.sw == p.C && (!p.has_dskip || p.dskip.sw == p.C) &&
         p.W * p.C * 2 >= 4096 && (p.pad == 0 || (px > p.pad && p.H > 2 * p.pad + 1 && p.W > 2 * p.pad + 1)) &&
         (reinterpret_cast<uintptr_t>(p.stats) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.bstats) & 15) == 0;
}

template <bool kApply, int ACT, bool HAS_SKIP, bool WRITE_GSUM, int V>
static void launch_tma_one(const NormBwdParams& p, dim3 grid, TmaGeom tg, cudaStream_t stream) {
  size_t smem = (size_t)tg.stages * tg.stage_bytes;
  if (!kApply && smem < 16384) smem = 16384;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(norm_bwd_tma_kernel<kApply, ACT, HAS_SKIP, WRITE_GSUM, V>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(norm_bwd_tma_kernel<kApply, ACT, HAS_SKIP, WRITE_GSUM, V>,
                         cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    attr_smem = smem;
  }
  launch_ex(norm_bwd_tma_kernel<kApply, ACT, HAS_SKIP, WRITE_GSUM, V>, grid, dim3(256, 1, 1), smem, stream, 1, true, p, tg,
            device_abort_flag_ptr());
}
template <bool kApply, int ACT, int V>
static void launch_tma_act(const NormBwdParams& p, dim3 grid, const TmaGeom& tg, cudaStream_t stream) {
  if (p.has_dskip) {
    if (kApply && p.write_gsum) launch_tma_one<kApply, ACT, true, kApply, V>(p, grid, tg, stream);
    else launch_tma_one<kApply, ACT, true, false, V>(p, grid, tg, stream);
  } else {
    if (kApply && p.write_gsum) launch_tma_one<kApply, ACT, false, kApply, V>(p, grid, tg, stream);
    else launch_tma_one<kApply, ACT, false, false, V>(p, grid, tg, stream);
  }
}
template <bool kApply, int V>
static void launch_tma_v(const NormBwdParams& p, dim3 grid, const TmaGeom& tg, cudaStream_t stream) {
  switch (p.act) {
    case CDB_ACT_RELU: launch_tma_act<kApply, CDB_ACT_RELU, V>(p, grid, tg, stream); break;
    case CDB_ACT_LEAKY: launch_tma_act<kApply, CDB_ACT_LEAKY, V>(p, grid, tg, stream); break;
    default: launch_tma_act<kApply, CDB_ACT_NONE, V>(p, grid, tg, stream); break;
  }
}
static void tma_geom(const NormBwdParams& p, int V, int stages_want, int budget, TmaGeom* tg) {
  const int cb = V * 4096;
  tg->cpr = ceil_div(p.W, V * (256 / (p.C / 8)));
  const int dslot = round_up(cb + 3 * p.pad * p.C * 2, 128);
  tg->off_d = cb;
  tg->off_m = tg->off_d + dslot;
  tg->off_s = tg->off_m + (p.pad > 0 ? dslot : 0);
  tg->stage_bytes = tg->off_s + (p.has_dskip ? cb : 0);
  int stages = stages_want < 2 ? 2 : (stages_want > kTmaMaxStages ? kTmaMaxStages : stages_want);
  if (stages > budget / tg->stage_bytes) stages = budget / tg->stage_bytes;
  tg->stages = stages;
}
template <bool kApply>
static void launch_tma(NormBwdParams p, int n, cudaStream_t stream) {
  static const int stages_env = getenv("CDB_NORM_TMA_STAGES") ? atoi(getenv("CDB_NORM_TMA_STAGES")) : 4;
  static const int per_sm = getenv("CDB_NORM_TMA_BLOCKS_PER_SM") ? atoi(getenv("CDB_NORM_TMA_BLOCKS_PER_SM")) : 2;
  static const int hint = getenv("CDB_NORM_TMA_HINT") ? atoi(getenv("CDB_NORM_TMA_HINT")) : 0;
  static const int v_env = getenv("CDB_NORM_TMA_V") ? atoi(getenv("CDB_NORM_TMA_V")) : 4;
  p.vt = p.C / 8;
  // two resident blocks per SM: 227 KB less 1 KB per block of system use and the static part
  const int budget = (per_sm >= 2 ? 110 : 220) * 1024;
  TmaGeom tg;
  int V = v_env == 4 ? 4 : 2;
  tma_geom(p, V, stages_env, budget, &tg);
  if (V == 4 && (tg.stages < 2 || p.W * p.C * 2 < 16384)) {   // 16 KB chunks do not fit twice (or rows are shorter)
    V = 2;
    tma_geom(p, V, stages_env, budget, &tg);
  }
  if (tg.stages < 2) tg.stages = 2;
  tg.hint = hint;
  const int cpi = p.H * tg.cpr;
  int bpi = (per_sm * sm_count()) / (n > 0 ? n : 1);
  if (bpi > cpi / 4) bpi = cpi / 4;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, n, 1);
  if (V == 4) launch_tma_v<kApply, 4>(p, grid, tg, stream);
  else launch_tma_v<kApply, 2>(p, grid, tg, stream);
}

// ------------------------------------------------------------------------------------------------
// TMA-staged forward (same ring as norm_bwd_tma_kernel): y (and the residual) arrive as cp.async.bulk row segments,
// the eight warps normalise from shared memory and store the output and its reflect halo from registers.  Any norm
// (instance / batch / running statistics / affine) and any activation: the coefficients are scale / shift per channel.
// ------------------------------------------------------------------------------------------------
template <int ACT, bool HAS_RES, int V>
__global__ void __launch_bounds__(256, 2) norm_fwd_tma_kernel(NormFwdParams p, TmaGeom tg, int* abort_global) {
  extern __shared__ __align__(128) uint8_t tma_smem[];
  __shared__ uint64_t bars[2 * kTmaMaxStages];
  __shared__ int abort_smem;
  const int stages = tg.stages, cpr = tg.cpr;
  const int tid = threadIdx.x;
  const int cv = p.vt;
  const int lanes = 256 / cv, PX = V * lanes;
  const int n = blockIdx.y;
  const int cpi = p.H * cpr;
  const int k0 = static_cast<int>(static_cast<int64_t>(blockIdx.x) * cpi / gridDim.x);
  const int k1 = static_cast<int>(static_cast<int64_t>(blockIdx.x + 1) * cpi / gridDim.x);
  const int nk = k1 - k0;
  const int W = p.W, C = p.C, pad = p.pad, H = p.H;
  const int ysh = static_cast<int>(p.y.sh), rsh = static_cast<int>(p.res.sh);
  const __nv_bfloat16* yimg = static_cast<const __nv_bfloat16*>(p.y.ptr) + n * p.y.sn;
  const __nv_bfloat16* rimg = HAS_RES ? static_cast<const __nv_bfloat16*>(p.res.ptr) + n * p.res.sn : nullptr;
  const uint32_t stage_base = smem_u32(tma_smem);
  const uint32_t bar_base = smem_u32(&bars[0]);
  int h = k0 / cpr, c = k0 - h * cpr;
  int ih = h, ic = c;
  auto issue = [&](int stage) {
    const int w0 = ic * PX;
    const uint32_t bytes = static_cast<uint32_t>(min(PX, W - w0) * C * 2);
    const uint32_t full = bar_base + stage * 8;
    const uint32_t dst = stage_base + stage * tg.stage_bytes;
    mbar_arrive_expect_tx(full, bytes * (HAS_RES ? 2 : 1));
    bulk_load_1d(dst, yimg + ih * ysh + w0 * C, bytes, full);
    if (HAS_RES) bulk_load_1d(dst + tg.off_s, rimg + ih * rsh + w0 * C, bytes, full);
    if (++ic == cpr) { ic = 0; ++ih; }
  };
  int issued = 0;
  pdl_trigger();
  if (tid == 0) {
    abort_smem = 0;
    for (int s = 0; s < stages; ++s) {
      mbar_init(bar_base + s * 8, 1);
      mbar_init(bar_base + (kTmaMaxStages + s) * 8, 8);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
  }
  pdl_wait();
  if (tid == 0)
    for (; issued < stages && issued < nk; ++issued) issue(issued);
  __syncthreads();
  volatile int* abort_flag = &abort_smem;
  const int v = tid % cv, lane = tid / cv;
  float2 sc2[4], sh2[4];
  {
    float scale[8], shift[8];
    norm_coeffs(p.norm, p.use_running, p.stats, p.gamma, p.beta, p.running_mean, p.running_var,
                p.per_image ? n : 0, C, v * 8, p.inv_count, p.eps, scale, shift, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc2[i] = make_float2(scale[2 * i], scale[2 * i + 1]);
      sh2[i] = make_float2(shift[2 * i], shift[2 * i + 1]);
    }
  }
  __nv_bfloat16* ob = static_cast<__nv_bfloat16*>(p.out.ptr) + n * p.out.sn + v * 8;
  const int osh = static_cast<int>(p.out.sh), osw = static_cast<int>(p.out.sw);
  const float slope = p.slope;
  int s = 0, ps = 0;
  uint32_t ph = 0, pph = 0;
  for (int i = 0; i < nk; ++i) {
    if (tid == 0 && i >= 1 && issued < nk) {
      if (mbar_wait(bar_base + (kTmaMaxStages + ps) * 8, pph, abort_flag)) {
        issue(ps);
        ++issued;
      }
    }
    const int w0 = c * PX;
    const int npx = min(PX, W - w0);
    if (!mbar_wait(bar_base + s * 8, ph, abort_flag)) break;
    const uint32_t src = stage_base + s * tg.stage_bytes + tid * 16;
    const bool hborder = pad > 0 && (h <= pad || h >= H - 1 - pad);
#pragma unroll
    for (int u = 0; u < V; ++u) {
      const int px = lane + u * lanes;
      if (px >= npx) break;
      const int w = w0 + px;
      float2 f2[4];
      unpack8_pairs(ld_shared_v4(src + u * 4096), f2);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        f2[q] = __ffma2_rn(f2[q], sc2[q], sh2[q]);
        f2[q].x = act_fwd_t<ACT>(f2[q].x, slope);
        f2[q].y = act_fwd_t<ACT>(f2[q].y, slope);
      }
      if (HAS_RES) add_vec_pairs(src + tg.off_s + u * 4096, f2);
      const uint4 o = pack8_pairs(f2);
      st16(ob + h * osh + w * osw, o);
      if (pad > 0 && (hborder || w <= pad || w >= W - 1 - pad)) write_halo<__nv_bfloat16>(ob, osh, osw, h, w, H, W, pad, o);
    }
    __syncwarp();
    if ((tid & 31) == 0) mbar_arrive(bar_base + (kTmaMaxStages + s) * 8);
    ps = s;
    pph = ph;
    if (++s == stages) { s = 0; ph ^= 1u; }
    if (++c == cpr) { c = 0; ++h; }
  }
  __syncthreads();
  if (tid == 0 && abort_smem && abort_global) atomicExch(abort_global, 1);
}

static bool fwd_tma_eligible(const NormFwdParams& p) {
  static const int mode = getenv("CDB_NORM_FWD_IMPL") ? atoi(getenv("CDB_NORM_FWD_IMPL")) : 2;   // 1: register loads
  if (mode < 2) return false;
  const int cv = p.C / 8;
  // pad > 1 (the 7x7 image layers): the halo stores of the border warps are the critical path and the register-load
  // kernel is faster (256x256x64, pad 3, batch 16: 53 vs 58 us); with a residual the staged form wins 30 -> 20 us
  return p.C % 8 == 0 && p.C >= 64 && p.C <= 2048 && (cv & (cv - 1)) == 0 && p.y.sw == p.C &&
         (!p.has_res || p.res.sw == p.C) && p.W * p.C * 2 >= 4096 && (p.pad <= 1 || mode >= 3);
}
template <int ACT, bool HAS_RES, int V>
static void launch_fwd_tma_one(const NormFwdParams& p, dim3 grid, const TmaGeom& tg, cudaStream_t stream) {
  const size_t smem = (size_t)tg.stages * tg.stage_bytes;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaFuncSetAttribute(norm_fwd_tma_kernel<ACT, HAS_RES, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(norm_fwd_tma_kernel<ACT, HAS_RES, V>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    attr_smem = smem;
  }
  launch_ex(norm_fwd_tma_kernel<ACT, HAS_RES, V>, grid, dim3(256, 1, 1), smem, stream, 1, true, p, tg,
            device_abort_flag_ptr());
}
template <bool HAS_RES, int V>
static void launch_fwd_tma_v(const NormFwdParams& p, dim3 grid, const TmaGeom& tg, cudaStream_t stream) {
  switch (p.act) {
    case CDB_ACT_NONE: launch_fwd_tma_one<CDB_ACT_NONE, HAS_RES, V>(p, grid, tg, stream); break;
    case CDB_ACT_RELU: launch_fwd_tma_one<CDB_ACT_RELU, HAS_RES, V>(p, grid, tg, stream); break;
    case CDB_ACT_LEAKY: launch_fwd_tma_one<CDB_ACT_LEAKY, HAS_RES, V>(p, grid, tg, stream); break;
    case CDB_ACT_TANH: launch_fwd_tma_one<CDB_ACT_TANH, HAS_RES, V>(p, grid, tg, stream); break;
    default: launch_fwd_tma_one<CDB_ACT_SIGMOID, HAS_RES, V>(p, grid, tg, stream); break;
  }
}
static void launch_fwd_tma(NormFwdParams p, int n, cudaStream_t stream) {
  static const int stages_env = getenv("CDB_NORM_TMA_STAGES") ? atoi(getenv("CDB_NORM_TMA_STAGES")) : 4;
  static const int per_sm_env = getenv("CDB_NORM_FWD_BLOCKS_PER_SM") ? atoi(getenv("CDB_NORM_FWD_BLOCKS_PER_SM")) : 0;
  const int per_sm = per_sm_env > 0 ? per_sm_env : (p.has_res ? 3 : 2);   // measured: 19.6 vs 20.8 us with a residual
  static const int v_env = getenv("CDB_NORM_TMA_V") ? atoi(getenv("CDB_NORM_TMA_V")) : 4;
  p.vt = p.C / 8;
  const int V = (v_env == 4 && p.W * p.C * 2 >= 16384) ? 4 : 2;
  const int cb = V * 4096;
  TmaGeom tg;
  tg.cpr = ceil_div(p.W, V * (256 / p.vt));
  tg.off_d = tg.off_m = 0;
  tg.off_s = cb;
  tg.stage_bytes = cb * (p.has_res ? 2 : 1);
  tg.hint = 0;
  int stages = stages_env < 2 ? 2 : (stages_env > kTmaMaxStages ? kTmaMaxStages : stages_env);
  const int budget = (per_sm >= 3 ? 72 : (per_sm == 2 ? 110 : 220)) * 1024;   // the kernels need <= 84 registers
  if (stages > budget / tg.stage_bytes) stages = budget / tg.stage_bytes;
  tg.stages = stages;
  const int cpi = p.H * tg.cpr;
  int bpi = (per_sm * sm_count()) / (n > 0 ? n : 1);
  if (bpi > cpi / 4) bpi = cpi / 4;
  if (bpi < 1) bpi = 1;
  dim3 grid(bpi, n, 1);
  if (p.has_res) {
    if (V == 4) launch_fwd_tma_v<true, 4>(p, grid, tg, stream);
    else launch_fwd_tma_v<true, 2>(p, grid, tg, stream);
  } else {
    if (V == 4) launch_fwd_tma_v<false, 4>(p, grid, tg, stream);
    else launch_fwd_tma_v<false, 2>(p, grid, tg, stream);
  }
}

static bool fused_in_eligible(const CdbNormDesc* d, const NormBwdParams& p, bool accum_f32) {
  if (getenv("CDB_NORM_NO_FUSED")) return false;
  const int pixels = p.H * p.W;
  return d->norm == CDB_NORM_INSTANCE && !d->use_running && d->gamma == nullptr && d->beta == nullptr && !accum_f32 &&
         p.pre_act == CDB_ACT_NONE && p.C % 32 == 0 && pixels <= kFusedCluster * kFusedVecs * 64 && pixels >= 64 &&
         // 32-bit element offsets inside one image (mirrored halo rows included)
         (int64_t)(p.H + 2 * p.pad) * p.y.sh < (1 << 30) && (int64_t)(p.H + 2 * p.pad) * p.dout.sh < (1 << 30) &&
         (int64_t)(p.H + 2 * p.pad) * p.dskip.sh < (1 << 30) && (int64_t)(p.H + 2 * p.pad) * p.dy.sh < (1 << 30) &&
         (int64_t)(p.H + 2 * p.pad) * p.gsum.sh < (1 << 30) && p.W <= 4096;
}

static void launch_norm_bwd_fused(const NormBwdParams& p, int n, cudaStream_t stream) {
  const int pixels = p.H * p.W;
  const int ppc = ceil_div(pixels, kFusedCluster);
  dim3 grid(kFusedCluster, p.C / 32, n);
  const size_t smem = 2 * kFusedVecs * 256 * sizeof(uint4);   // 64 KB of thread-private y / g slots (3 CTAs per SM)
  // OCC = resident CTAs per SM the kernel is compiled for: 2 (<= 128 registers) or 3 (<= 85 registers: measured
  // 2.3x SLOWER on B200, the kernel then spills 1 KB per thread — profiles/r01_norm_bwd_experiments.txt)
  static const int occ = getenv("CDB_NORM_FUSED_OCC") ? atoi(getenv("CDB_NORM_FUSED_OCC")) : 2;
  static const bool async = getenv("CDB_NORM_FUSED_ASYNC") ? atoi(getenv("CDB_NORM_FUSED_ASYNC")) != 0 : true;
#define CDB_FUSED_LAUNCH(ACT_, ASYNC_, OCC_)                                                                       \
  do {                                                                                                             \
    static bool attr_done = false;                                                                                 \
    if (!attr_done) {                                                                                              \
      cudaFuncSetAttribute(norm_bwd_fused_in_kernel<ACT_, ASYNC_, OCC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                           (int)smem);                                                                             \
      attr_done = true;                                                                                            \
    }                                                                                                              \
    norm_bwd_fused_in_kernel<ACT_, ASYNC_, OCC_><<<grid, 256, smem, stream>>>(p, ppc);                             \
  } while (0)
#define CDB_FUSED_ACT(ASYNC_, OCC_)                                          \
  do {                                                                       \
    switch (p.act) {                                                         \
      case CDB_ACT_RELU: CDB_FUSED_LAUNCH(CDB_ACT_RELU, ASYNC_, OCC_); break; \
      case CDB_ACT_LEAKY: CDB_FUSED_LAUNCH(CDB_ACT_LEAKY, ASYNC_, OCC_); break; \
      default: CDB_FUSED_LAUNCH(CDB_ACT_NONE, ASYNC_, OCC_); break;          \
    }                                                                        \
  } while (0)
  if (async && occ == 3) CDB_FUSED_ACT(true, 3);
  else if (async) CDB_FUSED_ACT(true, 2);
  else CDB_FUSED_ACT(false, 2);
#undef CDB_FUSED_ACT
#undef CDB_FUSED_LAUNCH
}

// Pixels per thread and iteration (all their loads are issued before the first use). 4 measured no faster than 2
// on B200 (the kernels are bound by instruction latency at 2 blocks per SM, not by loads in flight:
// profiles/r01_norm_bwd_experiments.txt), so 2 is the default and 4 stays selectable for experiments.
static int bwd_unroll(const NormBwdParams& p) {
  static const int forced = getenv("CDB_NORM_BWD_U") ? atoi(getenv("CDB_NORM_BWD_U")) : 0;
  (void)p;
  return forced == 4 ? 4 : 2;
}
template <typename T, bool kApply, bool AFFINE>
static void launch_norm_bwd_act(const NormBwdParams& p, dim3 grid, cudaStream_t stream) {
  if (sizeof(T) == 2 && bwd_unroll(p) == 4) {
    switch (p.act) {
      case CDB_ACT_RELU: norm_act_bwd_kernel<T, kApply, CDB_ACT_RELU, AFFINE, 4><<<grid, 256, 0, stream>>>(p); break;
      case CDB_ACT_LEAKY: norm_act_bwd_kernel<T, kApply, CDB_ACT_LEAKY, AFFINE, 4><<<grid, 256, 0, stream>>>(p); break;
      default: norm_act_bwd_kernel<T, kApply, CDB_ACT_NONE, AFFINE, 4><<<grid, 256, 0, stream>>>(p); break;
    }
    return;
  }
  constexpr int U = sizeof(T) == 2 ? 2 : 1;
  switch (p.act) {
    case CDB_ACT_RELU: norm_act_bwd_kernel<T, kApply, CDB_ACT_RELU, AFFINE, U><<<grid, 256, 0, stream>>>(p); break;
    case CDB_ACT_LEAKY: norm_act_bwd_kernel<T, kApply, CDB_ACT_LEAKY, AFFINE, U><<<grid, 256, 0, stream>>>(p); break;
    default: norm_act_bwd_kernel<T, kApply, CDB_ACT_NONE, AFFINE, U><<<grid, 256, 0, stream>>>(p); break;
  }
}
template <typename T, bool kApply>
static void launch_norm_bwd(const NormBwdParams& p, dim3 grid, cudaStream_t stream) {
  if (p.gamma != nullptr || p.beta != nullptr) launch_norm_bwd_act<T, kApply, true>(p, grid, stream);
  else launch_norm_bwd_act<T, kApply, false>(p, grid, stream);
}

}  // namespace cdb

using namespace cdb;

// dtype: the storage type every view of the call must have (bf16, or fp32 for the TF32 network modes)
static int check_view(const CdbAct* a, const char* what, int dtype = CDB_BF16) {
  CDB_REQUIRE(a && a->ptr, CDB_ERR_BAD_DESC, "%s: null tensor", what);
  CDB_REQUIRE(a->dtype == dtype, CDB_ERR_UNSUPPORTED, "%s: all views of a call must share one storage type (%s)", what,
              dtype == CDB_BF16 ? "bf16" : "fp32");
  CDB_REQUIRE(a->sn % 8 == 0 && a->sh % 8 == 0 && a->sw % 8 == 0 && (reinterpret_cast<uintptr_t>(a->ptr) & 15) == 0,
              CDB_ERR_ALIGNMENT, "%s: 16-byte alignment of pixels required", what);
  return CDB_OK;
}
static int storage_type(const CdbAct* y, const char* what, int* dtype) {
  CDB_REQUIRE(y && (y->dtype == CDB_BF16 || y->dtype == CDB_F32), CDB_ERR_UNSUPPORTED, "%s: bf16 or fp32 storage", what);
  *dtype = y->dtype;
  return CDB_OK;
}

extern "C" int cdb_channel_stats(const CdbAct* y, int32_t c_real, int32_t per_image, float* stats,
                                 cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int dt;
  int rc = storage_type(y, "channel_stats y", &dt);
  if (rc) return rc;
  rc = check_view(y, "channel_stats y", dt);
  if (rc) return rc;
  CDB_REQUIRE(stats && c_real >= 1 && c_real <= y->c, CDB_ERR_BAD_DESC, "channel_stats: bad arguments");
  const Mapping m = mapping_for(round_up(c_real, 8));
  const int chunks = chunks_for(y->h * y->w, m.lanes, y->n, m.cv_tiles);
  dim3 grid(chunks, y->n, m.cv_tiles);
  if (dt == CDB_F32)
    channel_stats_kernel<float><<<grid, 256, 0, stream>>>(view_of(y), y->h, y->w, c_real, m.vt, per_image, stats);
  else
    channel_stats_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(view_of(y), y->h, y->w, c_real, m.vt, per_image, stats);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

static int fill_common(const CdbNormDesc* d, const CdbAct* y, int* C) {
  CDB_REQUIRE(d->channels >= 1 && round_up(d->channels, 8) <= y->c, CDB_ERR_BAD_DESC,
              "norm: channels %d do not fit the stored %d", d->channels, y->c);
  CDB_REQUIRE(d->pad >= 0 && d->pad < y->h && d->pad < y->w, CDB_ERR_BAD_DESC, "norm: pad too large");
  CDB_REQUIRE(d->norm == CDB_NORM_NONE || d->use_running || d->stats, CDB_ERR_BAD_DESC, "norm: stats missing");
  CDB_REQUIRE(!d->use_running || (d->running_mean && d->running_var), CDB_ERR_BAD_DESC, "norm: running stats missing");
  *C = d->channels;
  return CDB_OK;
}

static float count_scale_of(const CdbNormDesc* d) {
  return (d->norm == CDB_NORM_BATCH && d->count_scale > 1.f) ? d->count_scale : 1.f;
}

static float inv_count_of(const CdbNormDesc* d, const CdbAct* y) {
  return 1.f / (d->norm == CDB_NORM_INSTANCE ? (float)(y->h * y->w)
                                             : (float)((int64_t)y->n * y->h * y->w) * count_scale_of(d));
}

extern "C" int cdb_norm_act_fwd(const CdbNormDesc* d, const CdbAct* y, const CdbAct* residual,
                                const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(d, CDB_ERR_BAD_DESC, "norm_act_fwd: null desc");
  int dt;
  int rc = storage_type(y, "norm_act_fwd y", &dt);
  if (rc) return rc;
  rc = check_view(y, "norm_act_fwd y", dt);
  if (rc) return rc;
  rc = check_view(out, "norm_act_fwd out", dt);
  if (rc) return rc;
  const bool has_res = residual && residual->ptr;
  if (has_res) {
    rc = check_view(residual, "norm_act_fwd residual", dt);
    if (rc) return rc;
  }
  CDB_REQUIRE(out->n == y->n && out->h == y->h && out->w == y->w, CDB_ERR_BAD_DESC,
              "norm_act_fwd: out must be the interior view with the shape of y");
  NormFwdParams p;
  memset(&p, 0, sizeof(p));
  rc = fill_common(d, y, &p.C);
  if (rc) return rc;
  p.y = view_of(y);
  p.out = view_of(out);
  p.has_res = has_res;
  if (has_res) p.res = view_of(residual);
  p.H = y->h;
  p.W = y->w;
  const Mapping m = mapping_for(round_up(p.C, 8));
  p.vt = m.vt;
  p.norm = d->norm;
  p.act = (d->flags & CDB_NORM_FLAG_ACT_FIRST) ? CDB_ACT_NONE : d->act;
  p.slope = d->slope;
  p.eps = d->eps;
  p.pad = d->pad;
  p.per_image = d->norm == CDB_NORM_INSTANCE;
  p.inv_count = inv_count_of(d, y);
  p.use_running = d->use_running;
  p.stats = d->stats;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.running_mean = d->running_mean;
  p.running_var = d->running_var;
  p.rows_per_block = rows_per_block_for(y->h, y->n, m.cv_tiles, 3);
  dim3 grid(ceil_div(y->h, p.rows_per_block), y->n, m.cv_tiles);
  if (dt == CDB_F32) {
    if (has_res) launch_norm_fwd<float, true>(p, grid, stream);
    else launch_norm_fwd<float, false>(p, grid, stream);
  } else if (fwd_tma_eligible(p)) {
    launch_fwd_tma(p, y->n, stream);
  } else {
    if (has_res) launch_norm_fwd<__nv_bfloat16, true>(p, grid, stream);
    else launch_norm_fwd<__nv_bfloat16, false>(p, grid, stream);
  }
  CDB_LAUNCH_OK();
  if (d->norm == CDB_NORM_BATCH && !d->use_running && d->update_running && d->running_mean && d->running_var) {
    const float count = (float)((int64_t)y->n * y->h * y->w) * count_scale_of(d);
    bn_running_kernel<<<ceil_div(d->channels, 128), 128, 0, stream>>>(d->stats, d->channels, count, d->momentum,
                                                                      d->running_mean, d->running_var, d->conv_bias);
    CDB_LAUNCH_OK();
  }
  return CDB_OK;
}

extern "C" int cdb_norm_act_bwd(const CdbNormDesc* d, const CdbAct* y, const CdbAct* dout, const CdbAct* dskip,
                                float* bstats, const CdbAct* dy, const CdbAct* gsum, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(d, CDB_ERR_BAD_DESC, "norm_act_bwd: null desc");
  int dt;
  int rc = storage_type(y, "norm_act_bwd y", &dt);
  if (rc) return rc;
  rc = check_view(y, "norm_act_bwd y", dt);
  if (rc) return rc;
  const bool accum_f32 = (d->flags & CDB_NORM_FLAG_ACCUM_F32) != 0;
  if (accum_f32) {
    CDB_REQUIRE(dy && dy->ptr && dy->dtype == CDB_F32 && dy->sn % 8 == 0 && dy->sh % 8 == 0 && dy->sw % 8 == 0 &&
                    (reinterpret_cast<uintptr_t>(dy->ptr) & 31) == 0,
                CDB_ERR_ALIGNMENT, "norm_act_bwd: ACCUM_F32 needs an fp32 dy view with 32-byte aligned pixels");
  } else {
    rc = check_view(dy, "norm_act_bwd dy", dt);
    if (rc) return rc;
  }
  const bool has_dout = dout && dout->ptr, has_dskip = dskip && dskip->ptr, has_gsum = gsum && gsum->ptr;
  CDB_REQUIRE(has_dout || has_dskip, CDB_ERR_BAD_DESC, "norm_act_bwd: no incoming gradient");
  if (has_dout && (rc = check_view(dout, "norm_act_bwd dout", dt))) return rc;
  if (has_dskip && (rc = check_view(dskip, "norm_act_bwd dskip", dt))) return rc;
  if (has_gsum && (rc = check_view(gsum, "norm_act_bwd gsum", dt))) return rc;
  CDB_REQUIRE(d->act == CDB_ACT_NONE || d->act == CDB_ACT_RELU || d->act == CDB_ACT_LEAKY, CDB_ERR_UNSUPPORTED,
              "norm_act_bwd: activation %d", d->act);
  NormBwdParams p;
  memset(&p, 0, sizeof(p));
  rc = fill_common(d, y, &p.C);
  if (rc) return rc;
  const bool need_reduce = d->norm != CDB_NORM_NONE && !d->use_running;
  CDB_REQUIRE(!need_reduce || bstats, CDB_ERR_BAD_DESC, "norm_act_bwd: bstats missing");
  p.y = view_of(y);
  p.dy = view_of(dy);
  p.has_dout = has_dout;
  p.has_dskip = has_dskip;
  p.write_gsum = has_gsum;
  if (has_dout) p.dout = view_of(dout);
  if (has_dskip) p.dskip = view_of(dskip);
  if (has_gsum) p.gsum = view_of(gsum);
  p.H = y->h;
  p.W = y->w;
  const Mapping m = mapping_for(round_up(p.C, 8));
  p.vt = m.vt;
  p.norm = d->norm;
  const bool act_first = (d->flags & CDB_NORM_FLAG_ACT_FIRST) != 0;
  p.act = act_first ? CDB_ACT_NONE : d->act;
  p.pre_act = act_first ? d->act : CDB_ACT_NONE;
  p.accum_f32 = accum_f32 ? 1 : 0;
  p.slope = d->slope;
  p.eps = d->eps;
  p.pad = d->pad;
  p.per_image = d->norm == CDB_NORM_INSTANCE;
  p.inv_count = inv_count_of(d, y);
  p.use_running = d->use_running;
  p.stats = d->stats;
  p.gamma = d->gamma;
  p.beta = d->beta;
  p.running_mean = d->running_mean;
  p.running_var = d->running_var;
  p.bstats = bstats;
  p.rows_per_block = rows_per_block_for(y->h, y->n, m.cv_tiles, 2);
  dim3 grid(ceil_div(y->h, p.rows_per_block), y->n, m.cv_tiles);
  const bool reduce_only = (d->flags & CDB_NORM_FLAG_BWD_REDUCE_ONLY) != 0;
  const bool apply_only = (d->flags & CDB_NORM_FLAG_BWD_APPLY_ONLY) != 0;
  if (reduce_only || apply_only) {
    // data-parallel BatchNorm: the caller all-reduces bstats between the two passes (generic two-pass kernels)
    CDB_REQUIRE(!(reduce_only && apply_only) && need_reduce, CDB_ERR_BAD_DESC,
                "norm_act_bwd: REDUCE_ONLY / APPLY_ONLY need a batch-statistics normalisation and exclude each other");
    if (reduce_only) {
      if (dt == CDB_F32) launch_norm_bwd<float, false>(p, grid, stream);
      else launch_norm_bwd<__nv_bfloat16, false>(p, grid, stream);
    } else {
      if (dt == CDB_F32) launch_norm_bwd<float, true>(p, grid, stream);
      else launch_norm_bwd<__nv_bfloat16, true>(p, grid, stream);
    }
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  // a convolution + activation layer without normalisation (the PatchGAN's first layer): same staged kernels, the
  // reduce pass only when the bias gradient is wanted
  const bool none_staged = dt == CDB_BF16 && d->norm == CDB_NORM_NONE && !accum_f32 && p.pre_act == CDB_ACT_NONE &&
                           p.has_dout && d->gamma == nullptr && d->beta == nullptr &&
                           (p.act == CDB_ACT_NONE || p.act == CDB_ACT_RELU || p.act == CDB_ACT_LEAKY) && tma_eligible(p) &&
                           !(getenv("CDB_NORM_NONE_STAGED") && atoi(getenv("CDB_NORM_NONE_STAGED")) == 0);
  if (none_staged) {
    p.per_image = 0;   // one group: the bias gradient of the whole batch
    if (bstats) {
      launch_tma<false>(p, y->n, stream);
      CDB_LAUNCH_OK();
    }
    launch_tma<true>(p, y->n, stream);
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  if (stream_eligible(d, p, accum_f32, dt) && tma_eligible(p)) {
    // (walking the batch in groups of images whose operands stay in L2 between the two passes was measured: 40 MB
    // groups 30.2 -> 31.9 ms per step, 24 MB groups 34.9 ms — the ramp and tail of the extra launches cost more than the
    // HBM reads they save)
    launch_tma<false>(p, y->n, stream);
    CDB_LAUNCH_OK();
    launch_tma<true>(p, y->n, stream);
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  if (stream_eligible(d, p, accum_f32, dt)) {
    // three resident blocks per SM (the kernels are compiled for <= 85 registers)
    p.rows_per_block = rows_per_block_for(y->h, y->n, m.cv_tiles, stream_occ() >= 3 ? 3 : 2);
    dim3 sgrid(ceil_div(y->h, p.rows_per_block), y->n, m.cv_tiles);
    launch_stream<false>(p, sgrid, stream);
    CDB_LAUNCH_OK();
    launch_stream<true>(p, sgrid, stream);
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  if (dt == CDB_BF16 && fused_in_eligible(d, p, accum_f32)) {
    launch_norm_bwd_fused(p, y->n, stream);
    CDB_LAUNCH_OK();
    return CDB_OK;
  }
  if (need_reduce || (bstats && d->norm == CDB_NORM_NONE)) {
    // norm none + bstats: the reduction yields the bias gradient (sum of ga) in component 0
    if (dt == CDB_F32) launch_norm_bwd<float, false>(p, grid, stream);
    else launch_norm_bwd<__nv_bfloat16, false>(p, grid, stream);
    CDB_LAUNCH_OK();
  }
  if (dt == CDB_F32) launch_norm_bwd<float, true>(p, grid, stream);
  else launch_norm_bwd<__nv_bfloat16, true>(p, grid, stream);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
