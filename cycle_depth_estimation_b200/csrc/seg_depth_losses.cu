// K6b — fused losses of the seg/depth step (new_multi/model5.py:280-285): 2-D cross entropy over the
// segmentation logits with ignore_index (torch.nn.CrossEntropyLoss(ignore_index=255), :281) and the
// multi-range depth loss BCEDepLoss (new_multi/networks5_ds.py:947-956).  Each kernel produces the
// forward sums and the (unnormalised or normalised) gradient in one pass over fp32 NCHW tensors.
#include "common.cuh"

namespace cdb {

__device__ __forceinline__ float sdl_block_sum(float v, float* smem) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? smem[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  }
  return r;
}

// logits [N][C][HW] fp32, labels [N][HW] int64.  acc[0] += sum of -log softmax[label] over pixels whose
// label != ignore, acc[1] += their count.  grad (optional) = softmax - onehot (0 for ignored pixels), to be
// scaled by 1 / count afterwards.  One thread per pixel: the C reads of a warp are 32 consecutive floats.
__global__ void __launch_bounds__(256)
ce2d_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int C, int64_t HW, int64_t total,
            int64_t ignore, float* __restrict__ acc, float* __restrict__ grad) {
  __shared__ float red[32];
  float lsum = 0.f, cnt = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / HW, px = i - n * HW;
    const float* x = logits + n * C * HW + px;
    const int64_t lab = labels[i];
    float m = -INFINITY;
    for (int c = 0; c < C; ++c) m = fmaxf(m, x[c * HW]);
    float se = 0.f;
    for (int c = 0; c < C; ++c) se += expf(x[c * HW] - m);
    const float lse = m + logf(se);
    const bool keep = lab != ignore && lab >= 0 && lab < C;
    if (keep) {
      lsum += lse - x[lab * HW];
      cnt += 1.f;
    }
    if (grad != nullptr) {
      float* g = grad + n * C * HW + px;
      for (int c = 0; c < C; ++c) {
        float v = 0.f;
        if (keep) v = expf(x[c * HW] - lse) - (c == lab ? 1.f : 0.f);
        g[c * HW] = v;
      }
    }
  }
  const float a = sdl_block_sum(lsum, red);
  const float b = sdl_block_sum(cnt, red);
  if (threadIdx.x == 0) {
    atomicAdd(acc, a);
    atomicAdd(acc + 1, b);
  }
}

// x [B][1][HW] (broadcast over the K target channels), t [B][K][HW];  o_m = (t == 1), z_m = (t == -1):
//   loss = ( sum_{t==1} -clamp(log((x+1)/2)) + sum_{t==-1} -clamp(log(1-(x+1)/2)) ) / (B K HW)
//          + 50 * mean |x - t|
// (BCELoss clamps its logs at -100; masked-out positions contribute exactly 0 as in the reference because
// both the prediction and the target are multiplied by the mask.)  grad_x [B][1][HW] sums over K.
__global__ void __launch_bounds__(256)
bcedep_kernel(const float* __restrict__ x, const float* __restrict__ t, int K, int64_t HW, int64_t total,
              float l1_weight, float* __restrict__ loss_acc, float* __restrict__ grad) {
  __shared__ float red[32];
  const float inv = 1.f / ((float)total * (float)K);
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / HW, px = i - b * HW;
    const float xv = x[i];
    const float p = (xv + 1.f) * 0.5f;
    float g = 0.f;
    for (int k = 0; k < K; ++k) {
      const float tv = t[(b * K + k) * HW + px];
      if (tv == 1.f) {
        // BCE(p, 1): -log p ; d/dp = (p - 1) / max((1-p) p, 1e-12) ; dp/dx = 1/2
        acc += -fmaxf(logf(p), -100.f);
        g += 0.5f * (p - 1.f) / fmaxf((1.f - p) * p, 1e-12f);
      } else if (tv == -1.f) {
        acc += -fmaxf(logf(1.f - p), -100.f);
        g += 0.5f * p / fmaxf((1.f - p) * p, 1e-12f);
      }
      const float d = xv - tv;
      acc += l1_weight * fabsf(d);
      g += l1_weight * (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    if (grad != nullptr) grad[i] = g * inv;
  }
  const float r = sdl_block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(loss_acc, r * inv);
}

static int sdl_grid(int64_t n) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cdb

using namespace cdb;

extern "C" int cdb_loss_ce2d(const float* logits, const int64_t* labels, int32_t n, int32_t c, int64_t hw,
                             int64_t ignore_index, float* acc2, float* grad, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(logits && labels && acc2 && n > 0 && c > 0 && hw > 0, CDB_ERR_BAD_DESC, "loss_ce2d: bad argument");
  const int64_t total = (int64_t)n * hw;
  ce2d_kernel<<<sdl_grid(total), 256, 0, stream>>>(logits, labels, c, hw, total, ignore_index, acc2, grad);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_loss_bcedep(const float* x, const float* target, int32_t b, int32_t k, int64_t hw, float l1_weight,
                               float* loss_acc, float* grad_x, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(x && target && loss_acc && b > 0 && k > 0 && hw > 0, CDB_ERR_BAD_DESC, "loss_bcedep: bad argument");
  const int64_t total = (int64_t)b * hw;
  bcedep_kernel<<<sdl_grid(total), 256, 0, stream>>>(x, target, k, hw, total, l1_weight, loss_acc, grad_x);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
