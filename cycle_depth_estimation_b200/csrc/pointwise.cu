// K5 — small memory-bound NHWC kernels of the seg/depth networks (new_multi/networks5_ds.py) and of the
// U-Net skip topology (models/networks.py:266-316): element-wise add, casts between the bf16 activation
// views and fp32 gradient accumulators, 2x2 average pooling, the channel-attention gate
// out = base + sigmoid(mean_hw(t)) * s, bilinear x2 up-sampling (align_corners = True), PReLU with a
// learnable slope, dropout, and the NHWC bf16 -> NCHW fp32 conversion at module outputs.
// One thread handles 8 channels (16 bytes of bf16) of one pixel; consecutive threads take consecutive
// channel vectors, so a warp touches whole 128-byte lines.  Every view carries its own storage type (bf16, or
// fp32 in the TF32 network modes); the arithmetic is fp32 either way.
#include "common.cuh"
#include "ptx.cuh"

namespace cdb {

struct PView {
  char* ptr;
  int64_t sn, sh, sw;  // elements
  int n, h, w, c;
  int f32;             // storage type: 0 = bf16, 1 = fp32
};

static inline PView pview(const CdbAct* a) {
  PView v;
  v.ptr = static_cast<char*>(a->ptr);
  v.sn = a->sn;
  v.sh = a->sh;
  v.sw = a->sw;
  v.n = a->n;
  v.h = a->h;
  v.w = a->w;
  v.c = a->c;
  v.f32 = a->dtype == CDB_F32 ? 1 : 0;
  return v;
}

__device__ __forceinline__ void pw_unpack8(const uint4& r, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pw_pack8(const float* f) {
  uint4 r;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return r;
}
__device__ __forceinline__ const __nv_bfloat16* bf(const PView& v, int n, int h, int w, int c) {
  return reinterpret_cast<const __nv_bfloat16*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}
__device__ __forceinline__ __nv_bfloat16* bfw(const PView& v, int n, int h, int w, int c) {
  return reinterpret_cast<__nv_bfloat16*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}
__device__ __forceinline__ float* f32w(const PView& v, int n, int h, int w, int c) {
  return reinterpret_cast<float*>(v.ptr) + n * v.sn + h * v.sh + w * v.sw + c;
}
__device__ __forceinline__ void ld8(const PView& v, int n, int h, int w, int c, float* f) {
  if (v.f32) {
    const float4* p = reinterpret_cast<const float4*>(f32w(v, n, h, w, c));
    const float4 a = p[0], b = p[1];
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  } else {
    pw_unpack8(*reinterpret_cast<const uint4*>(bf(v, n, h, w, c)), f);
  }
}
__device__ __forceinline__ void st8(const PView& v, int n, int h, int w, int c, const float* f) {
  if (v.f32) {
    float4* p = reinterpret_cast<float4*>(f32w(v, n, h, w, c));
    p[0] = make_float4(f[0], f[1], f[2], f[3]);
    p[1] = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    *reinterpret_cast<uint4*>(bfw(v, n, h, w, c)) = pw_pack8(f);
  }
}

// idx -> (cv, w, h, n) over a view of shape (n, h, w, cvn*8)
#define PW_DECODE(idx, cvn, W, H)            \
  const int cv = (int)((idx) % (cvn));       \
  int64_t _t = (idx) / (cvn);                \
  const int w = (int)(_t % (W));             \
  _t /= (W);                                 \
  const int h = (int)(_t % (H));             \
  const int n = (int)(_t / (H));             \
  const int c = cv * 8;

#define PW_LOOP(total)                                                                   \
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (total);      \
       idx += (int64_t)gridDim.x * blockDim.x)

static int pw_grid(int64_t total) {
  int64_t b = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

// ---- out = a + b ; cast: dst = src or dst += src between any pair of storage types
__global__ void __launch_bounds__(256) pw_add_kernel(PView a, PView b, PView o, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float x[8], y[8];
    ld8(a, n, h, w, c, x);
    ld8(b, n, h, w, c, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] += y[j];
    st8(o, n, h, w, c, x);
  }
}

__global__ void __launch_bounds__(256) pw_cast_kernel(PView s, PView d, int64_t total, int cvn, int accumulate) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, d.w, d.h)
    float x[8];
    ld8(s, n, h, w, c, x);
    if (accumulate) {
      float y[8];
      ld8(d, n, h, w, c, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] += y[j];
    }
    st8(d, n, h, w, c, x);
  }
}

// ---- 2x2 average pooling, stride 2 (nn.AvgPool2d(2, 2), new_multi/networks5_ds.py:355)
__global__ void __launch_bounds__(256) pw_avgpool_fwd_kernel(PView x, PView o, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float acc[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        ld8(x, n, 2 * h + a, 2 * w + b, c, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= 0.25f;
    st8(o, n, h, w, c, acc);
  }
}
// dx (shape of the pooling INPUT) = 0.25 * g[h/2, w/2] (0 outside the pooled area for odd sizes)
__global__ void __launch_bounds__(256) pw_avgpool_bwd_kernel(PView g, PView dx, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dx.w, dx.h)
    float t[8];
    if ((h >> 1) < g.h && (w >> 1) < g.w) {
      ld8(g, n, h >> 1, w >> 1, c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] *= 0.25f;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = 0.f;
    }
    st8(dx, n, h, w, c, t);
  }
}

// ---- out = alpha * x (skip connections scaled by a constant, models/encoder_decoder.py:198-205; its own backward)
__global__ void __launch_bounds__(256) pw_scale_kernel(PView x, PView o, float alpha, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float t[8];
    ld8(x, n, h, w, c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] *= alpha;
    st8(o, n, h, w, c, t);
  }
}

// ---- nn.Upsample(scale_factor=2, mode='nearest') (models/encoder_decoder.py:193) and its backward (2x2 sums)
__global__ void __launch_bounds__(256) pw_nearest_fwd_kernel(PView x, PView o, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float t[8];
    ld8(x, n, h >> 1, w >> 1, c, t);
    st8(o, n, h, w, c, t);
  }
}
__global__ void __launch_bounds__(256) pw_nearest_bwd_kernel(PView g, PView dx, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dx.w, dx.h)
    float acc[8], t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        ld8(g, n, 2 * h + a, 2 * w + b, c, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += t[j];
      }
    st8(dx, n, h, w, c, acc);
  }
}

// ---- nn.Tanh() inside a network (models/encoder_decoder.py:112: the output blocks feed the next decoder level):
// out = tanh(x);  dx = g * (1 - out^2)
__global__ void __launch_bounds__(256) pw_tanh_fwd_kernel(PView x, PView o, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float t[8];
    ld8(x, n, h, w, c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = tanhf(t[j]);
    st8(o, n, h, w, c, t);
  }
}
__global__ void __launch_bounds__(256) pw_tanh_bwd_kernel(PView y, PView g, PView dx, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dx.w, dx.h)
    float t[8], u[8];
    ld8(y, n, h, w, c, t);
    ld8(g, n, h, w, c, u);
#pragma unroll
    for (int j = 0; j < 8; ++j) u[j] *= 1.f - t[j] * t[j];
    st8(dx, n, h, w, c, u);
  }
}

// ---- channel attention gate: out = [base +] sigmoid(att_sum[n][c][0] * inv_hw) * s
__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }

__global__ void __launch_bounds__(256)
pw_gate_fwd_kernel(PView base, int has_base, PView s, const float* __restrict__ att, int C, float inv_hw, PView o,
                   int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float x[8], r[8];
    ld8(s, n, h, w, c, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = c + j;
      const float sg = ch < C ? sigmoidf_(att[((int64_t)n * C + ch) * 2] * inv_hw) : 0.f;
      r[j] = x[j] * sg;
    }
    if (has_base) {
      float b[8];
      ld8(base, n, h, w, c, b);
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] += b[j];
    }
    st8(o, n, h, w, c, r);
  }
}

// ds = g * sigmoid(att);  dsum[n][c] += sum_px g * s.   grid (chunks, n, cv tiles), 256 threads = vt x lanes
__global__ void __launch_bounds__(256)
pw_gate_bwd_kernel(PView g, PView s, const float* __restrict__ att, int C, float inv_hw, PView ds,
                   float* __restrict__ dsum, int vt) {
  __shared__ float red[256 * 8];
  const int v = threadIdx.x % vt, lane = threadIdx.x / vt, lanes = 256 / vt;
  const int cvec = blockIdx.z * vt + v;
  const int n = blockIdx.y;
  const int pixels = g.h * g.w;
  const int per_chunk = (pixels + gridDim.x - 1) / gridDim.x;
  const int p0 = blockIdx.x * per_chunk, p1 = min(pixels, p0 + per_chunk);
  float acc[8], sg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const bool active = cvec * 8 < C;
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = cvec * 8 + j;
      sg[j] = ch < C ? sigmoidf_(att[((int64_t)n * C + ch) * 2] * inv_hw) : 0.f;
    }
    for (int px = p0 + lane; px < p1; px += lanes) {
      const int h = px / g.w, w = px - h * g.w;
      float gv[8], sv[8], o[8];
      ld8(g, n, h, w, cvec * 8, gv);
      ld8(s, n, h, w, cvec * 8, sv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += gv[j] * sv[j];
        o[j] = gv[j] * sg[j];
      }
      st8(ds, n, h, w, cvec * 8, o);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[j * 256 + threadIdx.x] = acc[j];
  __syncthreads();
  for (int t = threadIdx.x; t < vt * 8; t += 256) {
    const int vv = t % vt, comp = t / vt;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += red[comp * 256 + l * vt + vv];
    const int ch = (blockIdx.z * vt + vv) * 8 + comp;
    if (ch < C) atomicAdd(dsum + (int64_t)n * C + ch, a);
  }
}

// dt[n,h,w,c] = dsum[n][c] * sig'(att) * inv_hw   (backward of sigmoid(mean_hw(t)))
__global__ void __launch_bounds__(256)
pw_gate_bcast_kernel(const float* __restrict__ dsum, const float* __restrict__ att, int C, float inv_hw, PView dt,
                     int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dt.w, dt.h)
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = c + j;
      float v = 0.f;
      if (ch < C) {
        const float sg = sigmoidf_(att[((int64_t)n * C + ch) * 2] * inv_hw);
        v = dsum[(int64_t)n * C + ch] * sg * (1.f - sg) * inv_hw;
      }
      r[j] = v;
    }
    st8(dt, n, h, w, c, r);
  }
}

// ---- bilinear x2, align_corners = True (nn.UpsamplingBilinear2d(scale_factor=2), networks5_ds.py:637,713)
// ATen: src = dst * (in - 1) / (out - 1) in float; i0 = (int)src; i1 = i0 + (i0 < in - 1); l1 = src - i0.
__device__ __forceinline__ void bl_src(int d, float scale, int in, int* i0, int* i1, float* l1) {
  const float src = scale * d;
  const int a = (int)src;
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *l1 = src - a;
}

__global__ void __launch_bounds__(256)
pw_bilinear_fwd_kernel(PView x, PView o, float sh, float sw, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    int h0, h1, w0, w1;
    float lh, lw;
    bl_src(h, sh, x.h, &h0, &h1, &lh);
    bl_src(w, sw, x.w, &w0, &w1, &lw);
    float a[8], b[8], cc[8], d[8], r[8];
    ld8(x, n, h0, w0, c, a);
    ld8(x, n, h0, w1, c, b);
    ld8(x, n, h1, w0, c, cc);
    ld8(x, n, h1, w1, c, d);
    const float h0l = 1.f - lh, w0l = 1.f - lw;
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = h0l * (w0l * a[j] + lw * b[j]) + lh * (w0l * cc[j] + lw * d[j]);
    st8(o, n, h, w, c, r);
  }
}

// gather form of the backward: every input pixel sums the output pixels that interpolate from it
__global__ void __launch_bounds__(256)
pw_bilinear_bwd_kernel(PView g, PView dx, float sh, float sw, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dx.w, dx.h)
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    // candidate output rows: src in (h-1, h+1)  =>  d in ((h-1)/sh, (h+1)/sh)
    const int dh_lo = max(0, (int)floorf((h - 1) / sh) - 1), dh_hi = min(g.h - 1, (int)ceilf((h + 1) / sh) + 1);
    const int dw_lo = max(0, (int)floorf((w - 1) / sw) - 1), dw_hi = min(g.w - 1, (int)ceilf((w + 1) / sw) + 1);
    for (int oh = dh_lo; oh <= dh_hi; ++oh) {
      int h0, h1;
      float lh;
      bl_src(oh, sh, dx.h, &h0, &h1, &lh);
      float wh = 0.f;
      if (h0 == h) wh += 1.f - lh;
      if (h1 == h) wh += lh;
      if (wh == 0.f) continue;
      for (int ow = dw_lo; ow <= dw_hi; ++ow) {
        int w0, w1;
        float lw;
        bl_src(ow, sw, dx.w, &w0, &w1, &lw);
        float ww = 0.f;
        if (w0 == w) ww += 1.f - lw;
        if (w1 == w) ww += lw;
        if (ww == 0.f) continue;
        float t[8];
        ld8(g, n, oh, ow, c, t);
        const float k = wh * ww;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += k * t[j];
      }
    }
    st8(dx, n, h, w, c, acc);
  }
}

// ---- PReLU with one learnable slope (nn.PReLU(), networks5_ds.py:498,551); slope read on the device
__global__ void __launch_bounds__(256)
pw_prelu_fwd_kernel(PView x, const float* __restrict__ slope, PView o, int64_t total, int cvn) {
  const float a = *slope;
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float t[8];
    ld8(x, n, h, w, c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j] > 0.f ? t[j] : a * t[j];
    st8(o, n, h, w, c, t);
  }
}
// dx = g * (x > 0 ? 1 : a);  *dslope += sum_{x <= 0} g * x
__global__ void __launch_bounds__(256)
pw_prelu_bwd_kernel(PView x, PView g, const float* __restrict__ slope, PView dx, float* __restrict__ dslope,
                    int64_t total, int cvn) {
  __shared__ float red[8];
  const float a = *slope;
  float acc = 0.f;
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, dx.w, dx.h)
    float t[8], gv[8];
    ld8(x, n, h, w, c, t);
    ld8(g, n, h, w, c, gv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (t[j] > 0.f) {
        // dx = g
      } else {
        acc += gv[j] * t[j];
        gv[j] *= a;
      }
    }
    st8(dx, n, h, w, c, gv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    if (dslope != nullptr && s != 0.f) atomicAdd(dslope, s);
  }
}

// ---- dropout (nn.Dropout(p), models/networks.py:305-306): keep mask from a counter-based hash of
// (seed, element index); the backward pass regenerates the same mask.  out = x * keep / (1 - p).
__device__ __forceinline__ uint32_t pw_hash(uint64_t k) {
  k ^= k >> 33;
  k *= 0xff51afd7ed558ccdULL;
  k ^= k >> 33;
  k *= 0xc4ceb9fe1a85ec53ULL;
  k ^= k >> 33;
  return (uint32_t)k;
}
__global__ void __launch_bounds__(256)
pw_dropout_kernel(PView x, PView o, uint64_t seed, const uint64_t* __restrict__ seed_dev, float p_drop, int64_t total,
                  int cvn) {
  if (seed_dev != nullptr) seed += *seed_dev;   // device-resident seed: a replayed CUDA graph draws a new mask every step
  const float scale = 1.f / (1.f - p_drop);
  const uint32_t thr = (uint32_t)(p_drop * 4294967296.0);
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, o.w, o.h)
    float t[8];
    ld8(x, n, h, w, c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t r = pw_hash(seed * 0x9E3779B97F4A7C15ULL + (uint64_t)idx * 8 + j);
      t[j] = r >= thr ? t[j] * scale : 0.f;
    }
    st8(o, n, h, w, c, t);
  }
}

// ---- NHWC bf16 view -> NCHW fp32 (module outputs); dst strides in elements
__global__ void __launch_bounds__(256)
pw_to_nchw_kernel(PView x, float* __restrict__ dst, int C, int64_t d_n, int64_t d_c, int64_t d_h, int64_t d_w,
                  int64_t total) {
  // one thread per (n, h, w, 8-channel group); consecutive threads walk w so fp32 stores coalesce per channel
  const int cvn = (C + 7) / 8;
  PW_LOOP(total) {
    const int w = (int)(idx % x.w);
    int64_t t_ = idx / x.w;
    const int cv = (int)(t_ % cvn);
    t_ /= cvn;
    const int h = (int)(t_ % x.h);
    const int n = (int)(t_ / x.h);
    float t[8];
    ld8(x, n, h, w, cv * 8, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = cv * 8 + j;
      if (ch < C) dst[n * d_n + ch * d_c + h * d_h + w * d_w] = t[j];
    }
  }
}

// ---- error-compensated TF32 ("3xTF32"): x = hi + lo, hi = rna_tf32(x), lo = rna_tf32(x - hi).  A convolution of
// [hi | lo | hi] against weights [w_hi | w_hi | w_lo] concatenated along the contraction dimension accumulates
// x_hi w_hi + x_lo w_hi + x_hi w_lo in the fp32 TMEM accumulator: fp32-level accuracy (~1e-6) from kind::tf32 MMAs.
//   mode 0: out [n, h, w, 3c]  channels [hi | lo | hi]   (forward / data gradient: contraction over channels)
//   mode 1: out [3n, h, w, c]  images   [hi ; lo ; hi]   (weight gradient: contraction over pixels)
//   mode 2: out [3n, h, w, c]  images   [hi ; hi ; lo]   (the other weight-gradient operand)
//   mode 3: out [n, h, w, c]   = hi                      (single-pass TF32: the operand rounded to nearest)
__global__ void __launch_bounds__(256) pw_split_tf32_kernel(PView x, PView o, int mode, int64_t total, int cvn) {
  PW_LOOP(total) {
    PW_DECODE(idx, cvn, x.w, x.h)
    float t[8], hi[8], lo[8];
    ld8(x, n, h, w, c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      hi[j] = round_tf32(t[j]);
      lo[j] = round_tf32(t[j] - hi[j]);
    }
    if (mode == 3) {
      st8(o, n, h, w, c, hi);
    } else if (mode == 0) {
      st8(o, n, h, w, c, hi);
      st8(o, n, h, w, x.c + c, lo);
      st8(o, n, h, w, 2 * x.c + c, hi);
    } else {
      st8(o, n, h, w, c, hi);
      st8(o, n + x.n, h, w, c, mode == 1 ? lo : hi);
      st8(o, n + 2 * x.n, h, w, c, mode == 1 ? hi : lo);
    }
  }
}

// dtype: CDB_BF16 / CDB_F32 to require one storage type, -1 to accept either (each view carries its own)
static int pw_check(const CdbAct* a, const char* what, int dtype = -1) {
  CDB_REQUIRE(a && a->ptr, CDB_ERR_BAD_DESC, "%s: null tensor", what);
  CDB_REQUIRE((dtype == -1 && (a->dtype == CDB_BF16 || a->dtype == CDB_F32)) || a->dtype == dtype, CDB_ERR_UNSUPPORTED,
              "%s: unexpected dtype %d", what, a->dtype);
  CDB_REQUIRE(a->c % 8 == 0 && a->sn % 8 == 0 && a->sh % 8 == 0 && a->sw % 8 == 0 &&
                  (reinterpret_cast<uintptr_t>(a->ptr) & 15) == 0,
              CDB_ERR_ALIGNMENT, "%s: 8-channel alignment of pixels required", what);
  return CDB_OK;
}
static bool same_shape(const CdbAct* a, const CdbAct* b) {
  return a->n == b->n && a->h == b->h && a->w == b->w && a->c == b->c;
}
static int64_t vec_total(const CdbAct* a) { return (int64_t)a->n * a->h * a->w * (a->c / 8); }

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_add(const CdbAct* a, const CdbAct* b, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(a, "add a")) || (rc = pw_check(b, "add b")) || (rc = pw_check(out, "add out"))) return rc;
  CDB_REQUIRE(same_shape(a, b) && same_shape(a, out), CDB_ERR_BAD_DESC, "add: shapes differ");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_add_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(a), pview(b), pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_cast(const CdbAct* src, const CdbAct* dst, int32_t accumulate, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst, CDB_ERR_BAD_DESC, "cast: null argument");
  int rc;
  if ((rc = pw_check(src, "cast src", src->dtype)) || (rc = pw_check(dst, "cast dst", dst->dtype))) return rc;
  CDB_REQUIRE(same_shape(src, dst), CDB_ERR_BAD_DESC, "cast: shapes differ");
  const int64_t total = vec_total(dst);
  if (total == 0) return CDB_OK;
  pw_cast_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(src), pview(dst), total, dst->c / 8, accumulate ? 1 : 0);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_avgpool2_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "avgpool2 x")) || (rc = pw_check(out, "avgpool2 out"))) return rc;
  CDB_REQUIRE(out->n == x->n && out->c == x->c && out->h == x->h / 2 && out->w == x->w / 2, CDB_ERR_BAD_DESC,
              "avgpool2: output must be [n, h/2, w/2, c]");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_avgpool_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_avgpool2_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(dout, "avgpool2_bwd dout")) || (rc = pw_check(dx, "avgpool2_bwd dx"))) return rc;
  CDB_REQUIRE(dout->n == dx->n && dout->c == dx->c && dout->h == dx->h / 2 && dout->w == dx->w / 2,
              CDB_ERR_BAD_DESC, "avgpool2_bwd: dout must be [n, h/2, w/2, c]");
  const int64_t total = vec_total(dx);
  if (total == 0) return CDB_OK;
  pw_avgpool_bwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(dout), pview(dx), total, dx->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_scale(const CdbAct* x, float alpha, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "scale x")) || (rc = pw_check(out, "scale out"))) return rc;
  CDB_REQUIRE(same_shape(x, out), CDB_ERR_BAD_DESC, "scale: shapes differ");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_scale_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), alpha, total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_nearest2x_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "nearest2x x")) || (rc = pw_check(out, "nearest2x out"))) return rc;
  CDB_REQUIRE(out->n == x->n && out->c == x->c && out->h == 2 * x->h && out->w == 2 * x->w, CDB_ERR_BAD_DESC,
              "nearest2x: output must be [n, 2h, 2w, c]");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_nearest_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_nearest2x_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(dout, "nearest2x_bwd dout")) || (rc = pw_check(dx, "nearest2x_bwd dx"))) return rc;
  CDB_REQUIRE(dout->n == dx->n && dout->c == dx->c && dout->h == 2 * dx->h && dout->w == 2 * dx->w, CDB_ERR_BAD_DESC,
              "nearest2x_bwd: dout must be [n, 2h, 2w, c]");
  const int64_t total = vec_total(dx);
  if (total == 0) return CDB_OK;
  pw_nearest_bwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(dout), pview(dx), total, dx->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_tanh_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "tanh x")) || (rc = pw_check(out, "tanh out"))) return rc;
  CDB_REQUIRE(same_shape(x, out), CDB_ERR_BAD_DESC, "tanh: shapes differ");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_tanh_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_tanh_bwd(const CdbAct* out, const CdbAct* g, const CdbAct* dx, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(out, "tanh_bwd out")) || (rc = pw_check(g, "tanh_bwd g")) || (rc = pw_check(dx, "tanh_bwd dx")))
    return rc;
  CDB_REQUIRE(same_shape(out, g) && same_shape(out, dx), CDB_ERR_BAD_DESC, "tanh_bwd: shapes differ");
  const int64_t total = vec_total(dx);
  if (total == 0) return CDB_OK;
  pw_tanh_bwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(out), pview(g), pview(dx), total, dx->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_gate_fwd(const CdbAct* base, const CdbAct* s, const float* att_sum, int32_t c_real, float inv_hw,
                            const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(s, "gate s")) || (rc = pw_check(out, "gate out"))) return rc;
  const bool has_base = base && base->ptr;
  if (has_base && (rc = pw_check(base, "gate base"))) return rc;
  CDB_REQUIRE(att_sum && same_shape(s, out) && (!has_base || same_shape(base, out)) && c_real <= s->c,
              CDB_ERR_BAD_DESC, "gate: bad arguments");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_gate_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(has_base ? pview(base) : pview(s), has_base ? 1 : 0, pview(s),
                                                         att_sum, c_real, inv_hw, pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_gate_bwd(const CdbAct* g, const CdbAct* s, const float* att_sum, int32_t c_real, float inv_hw,
                            const CdbAct* ds, float* dsum, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(g, "gate_bwd g")) || (rc = pw_check(s, "gate_bwd s")) || (rc = pw_check(ds, "gate_bwd ds")))
    return rc;
  CDB_REQUIRE(att_sum && dsum && same_shape(g, s) && same_shape(g, ds) && c_real <= g->c, CDB_ERR_BAD_DESC,
              "gate_bwd: bad arguments");
  const int cv = round_up(c_real, 8) / 8;
  int vt = 1;
  while (vt < cv && vt < 256) vt <<= 1;
  const int lanes = 256 / vt, cv_tiles = ceil_div(cv, vt);
  int chunks = (4 * sm_count()) / (g->n * cv_tiles > 0 ? g->n * cv_tiles : 1);
  const int max_chunks = ceil_div(g->h * g->w, lanes * 4);
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, g->n, cv_tiles);
  pw_gate_bwd_kernel<<<grid, 256, 0, stream>>>(pview(g), pview(s), att_sum, c_real, inv_hw, pview(ds), dsum, vt);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_gate_bcast(const float* dsum, const float* att_sum, int32_t c_real, float inv_hw, const CdbAct* dt,
                              cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(dt, "gate_bcast dt"))) return rc;
  CDB_REQUIRE(dsum && att_sum && c_real <= dt->c, CDB_ERR_BAD_DESC, "gate_bcast: bad arguments");
  const int64_t total = vec_total(dt);
  if (total == 0) return CDB_OK;
  pw_gate_bcast_kernel<<<pw_grid(total), 256, 0, stream>>>(dsum, att_sum, c_real, inv_hw, pview(dt), total, dt->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_bilinear2x_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "bilinear2x x")) || (rc = pw_check(out, "bilinear2x out"))) return rc;
  CDB_REQUIRE(out->n == x->n && out->c == x->c && out->h == 2 * x->h && out->w == 2 * x->w, CDB_ERR_BAD_DESC,
              "bilinear2x: output must be [n, 2h, 2w, c]");
  const float sh = out->h > 1 ? (float)(x->h - 1) / (float)(out->h - 1) : 0.f;
  const float sw = out->w > 1 ? (float)(x->w - 1) / (float)(out->w - 1) : 0.f;
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_bilinear_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), sh, sw, total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_bilinear2x_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(dout, "bilinear2x_bwd dout")) || (rc = pw_check(dx, "bilinear2x_bwd dx"))) return rc;
  CDB_REQUIRE(dout->n == dx->n && dout->c == dx->c && dout->h == 2 * dx->h && dout->w == 2 * dx->w,
              CDB_ERR_BAD_DESC, "bilinear2x_bwd: dout must be [n, 2h, 2w, c]");
  const float sh = dout->h > 1 ? (float)(dx->h - 1) / (float)(dout->h - 1) : 0.f;
  const float sw = dout->w > 1 ? (float)(dx->w - 1) / (float)(dout->w - 1) : 0.f;
  CDB_REQUIRE(sh > 0.f && sw > 0.f, CDB_ERR_UNSUPPORTED, "bilinear2x_bwd: 1-pixel inputs");
  const int64_t total = vec_total(dx);
  if (total == 0) return CDB_OK;
  pw_bilinear_bwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(dout), pview(dx), sh, sw, total, dx->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_prelu_fwd(const CdbAct* x, const float* slope, const CdbAct* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "prelu x")) || (rc = pw_check(out, "prelu out"))) return rc;
  CDB_REQUIRE(slope && same_shape(x, out), CDB_ERR_BAD_DESC, "prelu: bad arguments");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_prelu_fwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), slope, pview(out), total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_prelu_bwd(const CdbAct* x, const CdbAct* g, const float* slope, const CdbAct* dx, float* dslope,
                             cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "prelu_bwd x")) || (rc = pw_check(g, "prelu_bwd g")) || (rc = pw_check(dx, "prelu_bwd dx")))
    return rc;
  CDB_REQUIRE(slope && same_shape(x, g) && same_shape(x, dx), CDB_ERR_BAD_DESC, "prelu_bwd: bad arguments");
  const int64_t total = vec_total(dx);
  if (total == 0) return CDB_OK;
  pw_prelu_bwd_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(g), slope, pview(dx), dslope, total,
                                                          dx->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_dropout(const CdbAct* x, const CdbAct* out, uint64_t seed, float p_drop, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "dropout x")) || (rc = pw_check(out, "dropout out"))) return rc;
  CDB_REQUIRE(same_shape(x, out) && p_drop >= 0.f && p_drop < 1.f, CDB_ERR_BAD_DESC, "dropout: bad arguments");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_dropout_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), seed, nullptr, p_drop, total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_dropout_dev(const CdbAct* x, const CdbAct* out, uint64_t seed, const uint64_t* seed_dev, float p_drop,
                               cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "dropout x")) || (rc = pw_check(out, "dropout out"))) return rc;
  CDB_REQUIRE(same_shape(x, out) && p_drop >= 0.f && p_drop < 1.f && seed_dev, CDB_ERR_BAD_DESC, "dropout_dev: bad arguments");
  const int64_t total = vec_total(out);
  if (total == 0) return CDB_OK;
  pw_dropout_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), seed, seed_dev, p_drop, total, out->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_nhwc_to_nchw(const CdbAct* x, int32_t c_real, float* dst, int64_t d_n, int64_t d_c, int64_t d_h,
                                int64_t d_w, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "nhwc_to_nchw x"))) return rc;
  CDB_REQUIRE(dst && c_real >= 1 && c_real <= x->c, CDB_ERR_BAD_DESC, "nhwc_to_nchw: bad arguments");
  const int64_t total = (int64_t)x->n * x->h * x->w * ((c_real + 7) / 8);
  if (total == 0) return CDB_OK;
  pw_to_nchw_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), dst, c_real, d_n, d_c, d_h, d_w, total);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

// ---- few-output-channel convolutions (c7s1-3 of the ResNet generator, models/networks.py:184-186, and the
// data gradient of its first layer c7s1-64): the S filter columns are folded into the GEMM N dimension, i.e.
// a first pass computes T[n,p,q',(s,o)] = sum_{r,i} x[n,p+r,q',i] W[o,i,r,s] as an R x 1 convolution with
// S*Cout output channels (7x fewer, 2x wider MMAs than 49 taps of N = 16), and this kernel finishes
//   out[n,o,p,q] = act(bias[o] + sum_s T[n,p,q+s,(s,o)])      (fp32 NCHW, strided)
namespace cdb {
__global__ void __launch_bounds__(256)
shift_add_kernel(const float* __restrict__ t, int N, int P, int Q, int wp, int S, int cout, int ct,
                 const float* __restrict__ bias, int act, float slope, float* __restrict__ out, int64_t o_sn,
                 int64_t o_sc, int64_t o_sh, int64_t o_sw) {
  const int64_t total = (int64_t)N * P * Q;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int q = (int)(idx % Q);
    const int64_t np = idx / Q;
    const int p = (int)(np % P);
    const int n = (int)(np / P);
    const float* row = t + (np * wp + q) * ct;
    for (int o = 0; o < cout; ++o) {
      float acc = bias != nullptr ? bias[o] : 0.f;
      for (int s = 0; s < S; ++s) acc += row[(int64_t)s * ct + s * cout + o];
      if (act == CDB_ACT_TANH) acc = tanhf(acc);
      else if (act == CDB_ACT_SIGMOID) acc = 1.f / (1.f + __expf(-acc));
      else if (act == CDB_ACT_RELU) acc = fmaxf(acc, 0.f);
      else if (act == CDB_ACT_LEAKY) acc = acc > 0.f ? acc : acc * slope;
      out[n * o_sn + o * o_sc + p * o_sh + q * o_sw] = acc;
    }
  }
}
}  // namespace cdb

extern "C" int cdb_shift_add_nchw(const float* t, int32_t n, int32_t p, int32_t q, int32_t wp, int32_t s_taps,
                                  int32_t cout, int32_t ct, const float* bias, int32_t act, float slope, float* out,
                                  int64_t o_sn, int64_t o_sc, int64_t o_sh, int64_t o_sw, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(t && out && n > 0 && p > 0 && q > 0 && s_taps > 0 && cout > 0 && ct >= s_taps * cout && wp >= q + s_taps - 1,
              CDB_ERR_BAD_DESC, "shift_add_nchw: bad argument");
  const int64_t total = (int64_t)n * p * q;
  shift_add_kernel<<<pw_grid(total), 256, 0, stream>>>(t, n, p, q, wp, s_taps, cout, ct, bias, act, slope, out, o_sn,
                                                       o_sc, o_sh, o_sw);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

// ---- zero frame: clears every pixel of a padded buffer OUTSIDE the interior rectangle (the materialised zero
// padding of the next convolution / the halo of a data-gradient operand) instead of memsetting the whole buffer
namespace cdb {
__global__ void __launch_bounds__(256)
zero_frame_kernel(PView full, int top, int left, int ih, int iw, int64_t frame_px, int64_t total, int cvn) {
  pdl_trigger();
  pdl_wait();
  const int W = full.w, H = full.h;
  const int64_t top_band = (int64_t)top * W, mid_band = (int64_t)ih * (W - iw);
  const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(idx % cvn);
    int64_t t = idx / cvn;
    const int64_t j = t % frame_px;
    const int n = (int)(t / frame_px);
    int h, w;
    if (j < top_band) {
      h = (int)(j / W);
      w = (int)(j % W);
    } else if (j < top_band + mid_band) {
      const int64_t jj = j - top_band;
      const int r = (int)(jj / (W - iw)), cidx = (int)(jj % (W - iw));
      h = top + r;
      w = cidx < left ? cidx : cidx + iw;
    } else {
      const int64_t jj = j - top_band - mid_band;
      h = top + ih + (int)(jj / W);
      w = (int)(jj % W);
    }
    (void)H;
    st8(full, n, h, w, cv * 8, z);
  }
}
}  // namespace cdb

extern "C" int cdb_zero_frame(const CdbAct* full, int32_t top, int32_t left, int32_t inner_h, int32_t inner_w,
                              cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(full, "zero_frame buffer"))) return rc;
  CDB_REQUIRE(top >= 0 && left >= 0 && inner_h >= 0 && inner_w >= 0 && top + inner_h <= full->h &&
                  left + inner_w <= full->w,
              CDB_ERR_BAD_DESC, "zero_frame: interior outside the buffer");
  const int64_t frame_px = (int64_t)full->h * full->w - (int64_t)inner_h * inner_w;
  const int64_t total = frame_px * full->n * (full->c / 8);
  if (total == 0) return CDB_OK;
  launch_ex(zero_frame_kernel, dim3(pw_grid(total), 1, 1), dim3(256, 1, 1), 0, stream, 1, true, pview(full), top, left,
            inner_h, inner_w, frame_px, total, full->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_split_tf32(const CdbAct* x, const CdbAct* out, int32_t mode, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int rc;
  if ((rc = pw_check(x, "split_tf32 x", CDB_F32)) || (rc = pw_check(out, "split_tf32 out", CDB_F32))) return rc;
  CDB_REQUIRE(mode >= 0 && mode <= 3, CDB_ERR_BAD_DESC, "split_tf32: mode %d", mode);
  if (mode == 3)
    CDB_REQUIRE(same_shape(x, out), CDB_ERR_BAD_DESC, "split_tf32: rounding mode needs out shaped like x");
  else if (mode == 0)
    CDB_REQUIRE(out->n == x->n && out->h == x->h && out->w == x->w && out->c == 3 * x->c, CDB_ERR_BAD_DESC,
                "split_tf32: channel mode needs out [n, h, w, 3c]");
  else
    CDB_REQUIRE(out->n == 3 * x->n && out->h == x->h && out->w == x->w && out->c == x->c, CDB_ERR_BAD_DESC,
                "split_tf32: batch mode needs out [3n, h, w, c]");
  const int64_t total = vec_total(x);
  if (total == 0) return CDB_OK;
  pw_split_tf32_kernel<<<pw_grid(total), 256, 0, stream>>>(pview(x), pview(out), mode, total, x->c / 8);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
