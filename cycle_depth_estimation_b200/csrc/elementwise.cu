// Layout conversion at the module boundary (NCHW fp32 <-> NHWC bf16), the fused GAN / L1 / BCE
// losses (forward scalar and gradient in one pass) and a multi-tensor-free fused Adam step.
#include "common.cuh"

namespace cdb {

__device__ __forceinline__ float block_reduce_sum(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? smem[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;  // valid in warp 0
}

// src NCHW fp32 (strided) -> interior of an NHWC bf16 buffer with `cs` stored channels; channels
// >= c are zero; a reflect halo of `pad` pixels is mirrored. With act_out != nullptr the value is
// multiplied by the derivative of `act` evaluated from the activation OUTPUT (backward of a final
// tanh / sigmoid / leaky layer).
struct ToNhwcParams {
  const float* src;
  const float* act_out;
  int64_t s_n, s_c, s_h, s_w;
  void* out;               // bf16 or fp32 (fp32-storage network modes) NHWC interior
  int64_t o_n, o_h, o_w;
  int N, C, H, W, cs, pad, act;
  float slope;
};

template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(ToNhwcParams p) {
  const int64_t total = static_cast<int64_t>(p.N) * p.H * p.W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = idx % p.W;
    const int h = (idx / p.W) % p.H;
    const int n = idx / (static_cast<int64_t>(p.W) * p.H);
    const int64_t sbase = n * p.s_n + h * p.s_h + w * p.s_w;
    T* ob = static_cast<T*>(p.out) + n * p.o_n;
    for (int c0 = 0; c0 < p.cs; c0 += 8) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        float v = 0.f;
        if (c < p.C) {
          v = p.src[sbase + c * p.s_c];
          if (p.act_out != nullptr) {
            const float o = p.act_out[sbase + c * p.s_c];
            if (p.act == CDB_ACT_TANH) v *= (1.f - o * o);
            else if (p.act == CDB_ACT_SIGMOID) v *= o * (1.f - o);
            else if (p.act == CDB_ACT_LEAKY) v *= (o > 0.f ? 1.f : p.slope);
            else if (p.act == CDB_ACT_RELU) v *= (o > 0.f ? 1.f : 0.f);
          }
        }
        f[j] = v;
      }
      uint4 pk;
      __nv_bfloat162* q = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
      for (int j = 0; j < 4; ++j) q[j] = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
      int hh[3], ww[3];
      int nh = 0, nw = 0;
      hh[nh++] = h;
      ww[nw++] = w;
      if (p.pad > 0) {
        if (h >= 1 && h <= p.pad) hh[nh++] = -h;
        if (h <= p.H - 2 && h >= p.H - 1 - p.pad) hh[nh++] = 2 * (p.H - 1) - h;
        if (w >= 1 && w <= p.pad) ww[nw++] = -w;
        if (w <= p.W - 2 && w >= p.W - 1 - p.pad) ww[nw++] = 2 * (p.W - 1) - w;
      }
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b) {
          T* dst = ob + hh[a] * p.o_h + ww[b] * p.o_w + c0;
          if constexpr (sizeof(T) == 2) {
            *reinterpret_cast<uint4*>(dst) = pk;
          } else {
            reinterpret_cast<float4*>(dst)[0] = make_float4(f[0], f[1], f[2], f[3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(f[4], f[5], f[6], f[7]);
          }
        }
    }
  }
}

// dst[n,c,h,w] (+)= sum of the reflect images of (h,w) in src [N,C,H+2p,W+2p] (fp32 NCHW).
__global__ void reflect_fold_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, int NC, int H,
                                         int W, int pad, int accumulate) {
  const int64_t total = static_cast<int64_t>(NC) * H * W;
  const int Hp = H + 2 * pad, Wp = W + 2 * pad;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int w = idx % W;
    const int h = (idx / W) % H;
    const int64_t nc = idx / (static_cast<int64_t>(W) * H);
    int hh[3], ww[3];
    int nh = 0, nw = 0;
    hh[nh++] = h;
    ww[nw++] = w;
    if (h >= 1 && h <= pad) hh[nh++] = -h;
    if (h <= H - 2 && h >= H - 1 - pad) hh[nh++] = 2 * (H - 1) - h;
    if (w >= 1 && w <= pad) ww[nw++] = -w;
    if (w <= W - 2 && w >= W - 1 - pad) ww[nw++] = 2 * (W - 1) - w;
    float acc = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) acc += src[(nc * Hp + hh[a] + pad) * Wp + ww[b] + pad];
    dst[idx] = accumulate ? dst[idx] + acc : acc;
  }
}

// db[c] += sum_{n,h,w} g[n,c,h,w] * act'(out[n,c,h,w])  — bias gradient of a final conv (+ activation)
// layer taken from the fp32 NCHW tensors; grid (chunks, n*c), one atomicAdd per block (db zeroed first).
__global__ void __launch_bounds__(256)
bias_grad_nchw_kernel(const float* __restrict__ g, const float* __restrict__ act_out, int act, float slope, int C,
                      int64_t HW, float* __restrict__ db) {
  __shared__ float red[32];
  const int nc = blockIdx.y;
  const int c = nc % C;
  const int64_t base = static_cast<int64_t>(nc) * HW;
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < HW;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float v = g[base + i];
    if (act_out != nullptr) {
      const float o = act_out[base + i];
      if (act == CDB_ACT_TANH) v *= (1.f - o * o);
      else if (act == CDB_ACT_SIGMOID) v *= o * (1.f - o);
      else if (act == CDB_ACT_LEAKY) v *= (o > 0.f ? 1.f : slope);
      else if (act == CDB_ACT_RELU) v *= (o > 0.f ? 1.f : 0.f);
    }
    acc += v;
  }
  const float r = block_reduce_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(db + c, r);
}

// ---- losses: loss_acc += weight * mean(f(x)); grad = weight * f'(x) / numel -----------------------
enum { kLossMse = 0, kLossL1 = 1, kLossBce = 2 };

template <int kKind>
__global__ void __launch_bounds__(256)
loss_kernel(const float* __restrict__ x, const float* __restrict__ y, float target, int64_t numel, float weight,
            float* __restrict__ loss_acc, float* __restrict__ grad) {
  __shared__ float red[32];
  const float inv_n = 1.f / static_cast<float>(numel);
  float acc = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float a = x[i];
    const float t = y != nullptr ? y[i] : target;
    float l, g;
    if (kKind == kLossMse) {
      const float d = a - t;
      l = d * d;
      g = 2.f * d;
    } else if (kKind == kLossL1) {
      const float d = a - t;
      l = fabsf(d);
      g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    } else {
      // torch.nn.BCELoss: log terms clamped at -100; backward divides by max((1-a)*a, 1e-12)
      const float la = fmaxf(logf(a), -100.f), l1a = fmaxf(logf(1.f - a), -100.f);
      l = -(t * la + (1.f - t) * l1a);
      g = (a - t) / fmaxf((1.f - a) * a, 1e-12f);
    }
    acc += l;
    if (grad != nullptr) grad[i] = weight * inv_n * g;
  }
  const float r = block_reduce_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_acc, weight * inv_n * r);
}

// out[i] = a[i] * (*scalar)   (chain rule with the incoming scalar gradient)
__global__ void scale_by_scalar_kernel(const float* __restrict__ a, const float* __restrict__ scalar,
                                       float* __restrict__ out, int64_t numel) {
  const float s = *scalar;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] * s;
}

// Adam (torch.optim.Adam semantics, no amsgrad, no weight decay): one launch per parameter tensor.
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t numel, float lr, float b1, float b2, float eps,
                            float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);  // lerp, as torch does
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// Adam with the step count read from device memory, so that a CUDA graph of the training step stays valid
// across replays (the host-side bias corrections of adam_kernel would be frozen at capture time).
__global__ void adam_dev_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                     float* __restrict__ v, int64_t numel, float lr, float b1, float b2, float eps,
                                     const int32_t* __restrict__ step_ptr) {
  const float step = static_cast<float>(*step_ptr);
  const float bc1 = 1.f - powf(b1, step);
  const float bc2_sqrt = sqrtf(1.f - powf(b2, step));
  const float step_size = lr / bc1;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float gi = g[i];
    const float mi = m[i] + (1.f - b1) * (gi - m[i]);
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}

// ImagePool.query (util/image_pool.py:12-32) with the host's decisions supplied as a device table:
// plan[i] = {return_from, store_to} (-1 = the incoming image / no store). Images are processed in order by
// every thread for its own pixel, which preserves the sequential semantics of the reference (an image
// stored by entry i can be returned by a later entry of the same batch).
__global__ void pool_apply_kernel(const float* __restrict__ fake, float* __restrict__ pool,
                                  const int32_t* __restrict__ plan, int b, int64_t chw, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < chw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    for (int k = 0; k < b; ++k) {
      const int ret = plan[2 * k], sto = plan[2 * k + 1];
      const float x = fake[k * chw + i];
      const float r = ret >= 0 ? pool[ret * chw + i] : x;
      if (sto >= 0) pool[sto * chw + i] = x;
      out[k * chw + i] = r;
    }
  }
}

// Multi-tensor Adam: up to kAdamMaxEntries parameter tensors per launch, their pointers passed in the (large,
// __grid_constant__) kernel parameter block — no pointer table in device memory, so the launch is valid in a
// CUDA graph as is. Block b finds its (tensor, chunk) by binary search over the cumulative chunk counts.
// an entry may name up to two packed bf16 GEMM operands of its filter (the forward and the
// data-gradient layout the convolution kernels read); the kernel then writes the updated value into them as
// well, so that no separate re-pack pass runs after the optimizer step.
constexpr int kAdamMaxEntries = 256;
constexpr int kAdamChunk = 8192;
struct AdamPackDev {
  __nv_bfloat16* out;
  int32_t rows_are_dim0, rowpack, kpad, n_taps;
};
struct AdamEntryDev {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  AdamPackDev pack[2];
  int32_t d1, R, S, pad_;
};
struct AdamMultiParams {
  int32_t n_entries, step;
  float lr, b1, b2, eps;
  const int32_t* step_dev;
  const float* lr_dev;
  AdamEntryDev e[kAdamMaxEntries];
  int32_t cum[kAdamMaxEntries + 1];
};

// Filters that emit packed operands (no rowpack, R * S <= kAdamTileRS) are walked in tiles of 32 d0 x 32 d1 filters
// with all their taps: the fp32 state is read / written in runs of 32 * R * S contiguous floats, the updated values
// go through shared memory as bf16 and each packed operand is written in runs of 32 consecutive k (64 bytes).  The
// element-order walk below stores every packed value as a lone 2-byte write at a stride of kpad (one 32-byte sector
// each, twice per parameter): 132 us per 11.4 M-parameter generator = 1.7 TB/s of algorithmic traffic.
constexpr int kAdamTile = 32;
constexpr int kAdamTileRS = 16;

__global__ void __launch_bounds__(256) adam_multi_kernel(const __grid_constant__ AdamMultiParams q) {
  __shared__ __nv_bfloat16 tile_sb[kAdamTile * kAdamTileRS * (kAdamTile + 1)];
  int lo = 0, hi = q.n_entries;            // largest t with cum[t] <= blockIdx.x
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (q.cum[mid] <= static_cast<int>(blockIdx.x)) lo = mid;
    else hi = mid;
  }
  const AdamEntryDev& en = q.e[lo];
  const int64_t begin = static_cast<int64_t>(blockIdx.x - q.cum[lo]) * kAdamChunk;
  const int64_t end = begin + kAdamChunk < en.numel ? begin + kAdamChunk : en.numel;
  float bc1, bc2_sqrt;
  if (q.step_dev != nullptr) {
    const float st = static_cast<float>(*q.step_dev);
    bc1 = 1.f - powf(q.b1, st);
    bc2_sqrt = sqrtf(1.f - powf(q.b2, st));
  } else {
    bc1 = static_cast<float>(1.0 - pow(static_cast<double>(q.b1), static_cast<double>(q.step)));
    bc2_sqrt = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(q.b2), static_cast<double>(q.step))));
  }
  const float lr = q.lr_dev != nullptr ? *q.lr_dev : q.lr;   // device-resident learning rate: schedulers act on replays
  const float step_size = lr / bc1;
  float* __restrict__ p = en.param;
  const float* __restrict__ g = en.grad;
  float* __restrict__ m = en.exp_avg;
  float* __restrict__ v = en.exp_avg_sq;
  if (en.pad_ != 0) {
    // ---- tiled walk (en.pad_ = d0)
    const int d0 = en.pad_, d1 = en.d1, RS = en.R * en.S;
    const int tiles1 = (d1 + kAdamTile - 1) / kAdamTile;
    const int t = static_cast<int>(blockIdx.x) - q.cum[lo];
    const int i0_0 = (t / tiles1) * kAdamTile, i1_0 = (t % tiles1) * kAdamTile;
    const int n0 = min(kAdamTile, d0 - i0_0), n1 = min(kAdamTile, d1 - i1_0);
    const int run = n1 * RS;
    const float b1 = q.b1, b2 = q.b2, eps = q.eps;
    for (int e = threadIdx.x; e < n0 * run; e += 256) {
      const int i0_l = e / run, rem = e - i0_l * run;
      const int i1_l = rem / RS, tap = rem - i1_l * RS;
      const int64_t j = (static_cast<int64_t>(i0_0 + i0_l) * d1 + i1_0) * RS + rem;
      const float gi = g[j];
      float mi = m[j], vi = v[j];
      mi = mi + (1.f - b1) * (gi - mi);
      vi = b2 * vi + (1.f - b2) * gi * gi;
      const float pn = p[j] - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
      m[j] = mi;
      v[j] = vi;
      p[j] = pn;
      tile_sb[(i0_l * RS + tap) * (kAdamTile + 1) + i1_l] = __float2bfloat16(pn);
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 2; ++w) {
      const AdamPackDev& pk = en.pack[w];
      if (pk.out == nullptr) continue;
      const int nrow = pk.rows_are_dim0 ? n0 : n1, nk = pk.rows_are_dim0 ? n1 : n0;
      const int row0 = pk.rows_are_dim0 ? i0_0 : i1_0, k0 = pk.rows_are_dim0 ? i1_0 : i0_0;
      for (int e = threadIdx.x; e < nrow * RS * kAdamTile; e += 256) {
        const int k_l = e % kAdamTile, rt = e / kAdamTile;
        const int tap = rt % RS, row_l = rt / RS;
        if (k_l >= nk) continue;
        const int i0_l = pk.rows_are_dim0 ? row_l : k_l, i1_l = pk.rows_are_dim0 ? k_l : row_l;
        pk.out[(static_cast<int64_t>(row0 + row_l) * pk.n_taps + tap) * pk.kpad + k0 + k_l] =
            tile_sb[(i0_l * RS + tap) * (kAdamTile + 1) + i1_l];
      }
    }
    return;
  }
  const bool packs = en.pack[0].out != nullptr || en.pack[1].out != nullptr;
  const int rs = en.R * en.S;
  const float b1 = q.b1, b2 = q.b2, eps = q.eps;
  auto update = [&](int64_t i, float pi, float gi, float& mi, float& vi) -> float {
    mi = mi + (1.f - b1) * (gi - mi);          // lerp, as torch does
    vi = b2 * vi + (1.f - b2) * gi * gi;
    const float pn = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
    if (packs) {
      // element (i0, i1, r, s) of the filter [d0][d1][R][S] -> packed[row][tap * kpad + k] (pack_weight_kernel's map)
      const int ii = static_cast<int>(i);
      const int tap_rs = ii % rs;
      const int i01 = ii / rs;
      const int i1 = i01 % en.d1, i0 = i01 / en.d1;
      const int r = tap_rs / en.S, sc = tap_rs - r * en.S;
      const __nv_bfloat16 b = __float2bfloat16(pn);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const AdamPackDev& pk = en.pack[t];
        if (pk.out == nullptr) continue;
        const int row = pk.rows_are_dim0 ? i0 : i1;
        const int k = pk.rows_are_dim0 ? i1 : i0;
        const int tap = pk.rowpack ? r : tap_rs;
        const int kk = pk.rowpack ? sc * pk.rowpack + k : k;
        pk.out[(static_cast<int64_t>(row) * pk.n_taps + tap) * pk.kpad + kk] = b;
      }
    }
    return pn;
  };
  // 16-byte vector body (tensors from the torch allocator are 16-byte aligned; chunk boundaries are multiples of 4)
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0;
  int64_t i = begin;
  if (vec) {
    const int64_t end4 = begin + ((end - begin) & ~static_cast<int64_t>(3));
    for (int64_t j = begin + 4 * threadIdx.x; j < end4; j += 4 * 256) {
      const float4 g4 = *reinterpret_cast<const float4*>(g + j);
      float4 p4 = *reinterpret_cast<const float4*>(p + j);
      float4 m4 = *reinterpret_cast<const float4*>(m + j);
      float4 v4 = *reinterpret_cast<const float4*>(v + j);
      p4.x = update(j, p4.x, g4.x, m4.x, v4.x);
      p4.y = update(j + 1, p4.y, g4.y, m4.y, v4.y);
      p4.z = update(j + 2, p4.z, g4.z, m4.z, v4.z);
      p4.w = update(j + 3, p4.w, g4.w, m4.w, v4.w);
      *reinterpret_cast<float4*>(m + j) = m4;
      *reinterpret_cast<float4*>(v + j) = v4;
      *reinterpret_cast<float4*>(p + j) = p4;
    }
    i = end4;
  }
  for (int64_t j = i + threadIdx.x; j < end; j += 256) {
    float mi = m[j], vi = v[j];
    p[j] = update(j, p[j], g[j], mi, vi);
    m[j] = mi;
    v[j] = vi;
  }
}

static int grid_for(int64_t n, int per_thread = 4) {
  int64_t b = (n + 256 * per_thread - 1) / (256 * per_thread);
  const int cap = sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_nchw_to_nhwc(const float* src, int32_t n, int32_t c, int32_t h, int32_t w, int64_t s_n,
                                int64_t s_c, int64_t s_h, int64_t s_w, const float* act_out, int32_t act,
                                float slope, const CdbAct* out, int32_t pad, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && out && out->ptr, CDB_ERR_BAD_DESC, "nchw_to_nhwc: null argument");
  CDB_REQUIRE((out->dtype == CDB_BF16 || out->dtype == CDB_F32) && out->c % 8 == 0 && out->c >= c, CDB_ERR_BAD_DESC,
              "nchw_to_nhwc: out channels");
  CDB_REQUIRE(out->n == n && out->h == h && out->w == w, CDB_ERR_BAD_DESC, "nchw_to_nhwc: out must be the interior view");
  CDB_REQUIRE(out->sn % 8 == 0 && out->sh % 8 == 0 && out->sw % 8 == 0 && (reinterpret_cast<uintptr_t>(out->ptr) & 15) == 0,
              CDB_ERR_ALIGNMENT, "nchw_to_nhwc: alignment");
  CDB_REQUIRE(pad >= 0 && pad < h && pad < w, CDB_ERR_BAD_DESC, "nchw_to_nhwc: pad");
  ToNhwcParams p;
  p.src = src;
  p.act_out = act_out;
  p.s_n = s_n;
  p.s_c = s_c;
  p.s_h = s_h;
  p.s_w = s_w;
  p.out = out->ptr;
  p.o_n = out->sn;
  p.o_h = out->sh;
  p.o_w = out->sw;
  p.N = n;
  p.C = c;
  p.H = h;
  p.W = w;
  p.cs = out->c;
  p.pad = pad;
  p.act = act;
  p.slope = slope;
  if (out->dtype == CDB_F32) nchw_to_nhwc_kernel<float><<<grid_for((int64_t)n * h * w, 1), 256, 0, stream>>>(p);
  else nchw_to_nhwc_kernel<__nv_bfloat16><<<grid_for((int64_t)n * h * w, 1), 256, 0, stream>>>(p);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_reflect_fold_nchw(const float* src, float* dst, int32_t nc, int32_t h, int32_t w,
                                     int32_t pad, int32_t accumulate, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst && pad >= 0 && pad < h && pad < w, CDB_ERR_BAD_DESC, "reflect_fold_nchw: bad argument");
  reflect_fold_nchw_kernel<<<grid_for((int64_t)nc * h * w, 1), 256, 0, stream>>>(src, dst, nc, h, w, pad, accumulate);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_bias_grad_nchw(const float* g, const float* act_out, int32_t act, float slope, int32_t n,
                                  int32_t c, int64_t hw, float* db, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(g && db && n > 0 && c > 0 && hw > 0, CDB_ERR_BAD_DESC, "bias_grad_nchw: bad argument");
  CDB_CUDA_OK(cudaMemsetAsync(db, 0, sizeof(float) * c, stream));
  int chunks = (int)((hw + 256 * 8 - 1) / (256 * 8));
  if (chunks > 64) chunks = 64;
  bias_grad_nchw_kernel<<<dim3(chunks, n * c), 256, 0, stream>>>(g, act_out, act, slope, c, hw, db);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_loss_mse_const(const float* x, int64_t numel, float target, float weight, float* loss_acc,
                                  float* grad, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(x && loss_acc && numel > 0, CDB_ERR_BAD_DESC, "loss_mse_const: bad argument");
  loss_kernel<kLossMse><<<grid_for(numel), 256, 0, stream>>>(x, nullptr, target, numel, weight, loss_acc, grad);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_loss_bce_const(const float* x, int64_t numel, float target, float weight, float* loss_acc,
                                  float* grad, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(x && loss_acc && numel > 0, CDB_ERR_BAD_DESC, "loss_bce_const: bad argument");
  loss_kernel<kLossBce><<<grid_for(numel), 256, 0, stream>>>(x, nullptr, target, numel, weight, loss_acc, grad);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_loss_l1(const float* a, const float* b, int64_t numel, float weight, float* loss_acc,
                           float* grad_a, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(a && b && loss_acc && numel > 0, CDB_ERR_BAD_DESC, "loss_l1: bad argument");
  loss_kernel<kLossL1><<<grid_for(numel), 256, 0, stream>>>(a, b, 0.f, numel, weight, loss_acc, grad_a);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_scale_by_scalar(const float* a, const float* scalar, float* out, int64_t numel,
                                   cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(a && scalar && out && numel > 0, CDB_ERR_BAD_DESC, "scale_by_scalar: bad argument");
  scale_by_scalar_kernel<<<grid_for(numel), 256, 0, stream>>>(a, scalar, out, numel);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                             float lr, float beta1, float beta2, float eps, int32_t step, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && numel > 0 && step >= 1, CDB_ERR_BAD_DESC, "adam_step: bad argument");
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<grid_for(numel), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2, eps,
                                                   (float)bc1, (float)sqrt(bc2));
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel,
                                 float lr, float beta1, float beta2, float eps, const int32_t* step_dev,
                                 cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev && numel > 0, CDB_ERR_BAD_DESC,
              "adam_step_dev: bad argument");
  adam_dev_step_kernel<<<grid_for(numel), 256, 0, stream>>>(param, grad, exp_avg, exp_avg_sq, numel, lr, beta1, beta2,
                                                            eps, step_dev);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_image_pool_apply(const float* fake, float* pool, const int32_t* plan_dev, int32_t batch, int64_t chw,
                                    float* out, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(fake && pool && plan_dev && out && batch > 0 && chw > 0, CDB_ERR_BAD_DESC, "image_pool_apply: bad argument");
  pool_apply_kernel<<<grid_for(chw, 1), 256, 0, stream>>>(fake, pool, plan_dev, batch, chw, out);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

static int adam_multi_launch(const CdbAdamEntry* plain, const CdbAdamPackEntry* packed, int32_t n_entries, float lr,
                             const float* lr_dev, float beta1, float beta2, float eps, int32_t step,
                             const int32_t* step_dev, cudaStream_t stream) {
  CDB_REQUIRE((plain || packed) && n_entries > 0 && (step >= 1 || step_dev), CDB_ERR_BAD_DESC, "adam_multi: bad argument");
  static thread_local AdamMultiParams q;
  for (int base = 0; base < n_entries; base += kAdamMaxEntries) {
    const int n = n_entries - base < kAdamMaxEntries ? n_entries - base : kAdamMaxEntries;
    q.n_entries = n;
    q.step = step;
    q.lr = lr;
    q.b1 = beta1;
    q.b2 = beta2;
    q.eps = eps;
    q.step_dev = step_dev;
    q.lr_dev = lr_dev;
    int blocks = 0;
    for (int i = 0; i < n; ++i) {
      AdamEntryDev& d = q.e[i];
      memset(&d, 0, sizeof(d));
      if (plain) {
        const CdbAdamEntry& en = plain[base + i];
        d.param = en.param, d.grad = en.grad, d.exp_avg = en.exp_avg, d.exp_avg_sq = en.exp_avg_sq, d.numel = en.numel;
      } else {
        const CdbAdamPackEntry& en = packed[base + i];
        d.param = en.param, d.grad = en.grad, d.exp_avg = en.exp_avg, d.exp_avg_sq = en.exp_avg_sq, d.numel = en.numel;
        d.d1 = en.d1, d.R = en.r, d.S = en.s;
        for (int t = 0; t < 2; ++t) {
          if (!en.pack[t]) continue;
          CDB_REQUIRE(en.d0 >= 1 && en.d1 >= 1 && en.r >= 1 && en.s >= 1 &&
                          (int64_t)en.d0 * en.d1 * en.r * en.s == en.numel && en.numel < (1ll << 31),
                      CDB_ERR_BAD_DESC, "adam_pack_multi: filter geometry of entry %d", base + i);
          const int kdim = en.rows_are_dim0[t] ? en.d1 : en.d0;
          AdamPackDev& pk = d.pack[t];
          pk.out = static_cast<__nv_bfloat16*>(en.pack[t]);
          pk.rows_are_dim0 = en.rows_are_dim0[t];
          pk.rowpack = en.rowpack[t];
          if (pk.rowpack) {
            CDB_REQUIRE(en.s * pk.rowpack <= 64 && kdim <= pk.rowpack, CDB_ERR_BAD_DESC,
                        "adam_pack_multi: rowpack geometry of entry %d", base + i);
            pk.kpad = 64;
            pk.n_taps = en.r;
          } else {
            pk.kpad = round_up(kdim, 64);
            pk.n_taps = en.r * en.s;
          }
        }
      }
      CDB_REQUIRE(d.param && d.grad && d.exp_avg && d.exp_avg_sq && d.numel > 0, CDB_ERR_BAD_DESC,
                  "adam_multi: bad entry %d", base + i);
      q.cum[i] = blocks;
      static const bool tiled_on = !(getenv("CDB_ADAM_TILED") && atoi(getenv("CDB_ADAM_TILED")) == 0);
      const bool has_pack = d.pack[0].out != nullptr || d.pack[1].out != nullptr;
      if (tiled_on && has_pack && d.pack[0].rowpack == 0 && d.pack[1].rowpack == 0 && d.R * d.S <= kAdamTileRS) {
        const int d0 = (int)(d.numel / ((int64_t)d.d1 * d.R * d.S));
        d.pad_ = d0;
        blocks += ((d0 + kAdamTile - 1) / kAdamTile) * ((d.d1 + kAdamTile - 1) / kAdamTile);
        continue;
      }
      blocks += (int)((d.numel + kAdamChunk - 1) / kAdamChunk);
    }
    q.cum[n] = blocks;
    adam_multi_kernel<<<blocks, 256, 0, stream>>>(q);
    CDB_LAUNCH_OK();
  }
  return CDB_OK;
}

extern "C" int cdb_adam_multi(const CdbAdamEntry* entries_host, int32_t n_entries, float lr, float beta1, float beta2,
                              float eps, int32_t step, const int32_t* step_dev, cdbStream_t stream_) {
  return adam_multi_launch(entries_host, nullptr, n_entries, lr, nullptr, beta1, beta2, eps, step, step_dev,
                           static_cast<cudaStream_t>(stream_));
}

extern "C" int cdb_adam_pack_multi(const CdbAdamPackEntry* entries_host, int32_t n_entries, float lr,
                                   const float* lr_dev, float beta1, float beta2, float eps, int32_t step,
                                   const int32_t* step_dev, cdbStream_t stream_) {
  return adam_multi_launch(nullptr, entries_host, n_entries, lr, lr_dev, beta1, beta2, eps, step, step_dev,
                           static_cast<cudaStream_t>(stream_));
}
