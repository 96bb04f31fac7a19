// K7 — KITTI depth metrics of new_multi/my_eval.py (compute_errors :7-31 inside eval_metric :35-108)
// for a batch of uint8 ground-truth / prediction image pairs.
//
// Per image:  pred = clip(pred_u8/255*80, 1, 50);  mask = 1 < gt < 50;  over the masked pixels
//   v = (pred - min)/(max - min)*49 + 1;  thr = max(gt/v, v/gt);  a_k = mean(thr < 1.25^k);
//   rmse = sqrt(mean((gt-v)^2));  rmse_log = sqrt(mean((log(gt)-log(v))^2));
//   abs_rel = mean(|gt-v|/gt);  sq_rel = mean((gt-v)^2/gt).
// gt has 48 admissible values and pred 256, so everything is a function of the pair (gt, pred): see
// metrics_hist_kernel below (one pass, joint histogram in shared memory).
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

// Single pass over HBM: ONE 1024-thread block per image builds the joint histogram H[gt-2][pred]
// (48 x 256 counters in shared memory, shared-memory atomics) of the masked pixels.  Everything the
// reference computes per pixel is a function of the pair (gt, pred), so the block then evaluates
//   pmin / pmax / count            from the occupied histogram columns,
//   v[p], log v[p]                 in IEEE double with the reference's operation order (no FMA contraction),
//   sums  = sum_{g,p} H[g][p] * f(g, p)   in double,
//   a_k   = sum_{g,p} H[g][p] * [max(g / v[p], v[p] / g) < 1.25^k]   (integer counts: bit-exact)
// and writes the 8 outputs.  Algorithmic traffic = 2 bytes per pixel, read exactly once.
constexpr int kHistRows = 48;
constexpr int kHistCells = kHistRows * 256;
constexpr int kMetricThreads = 1024;

__device__ __forceinline__ bool masked(uint32_t g) { return (g - 2u) < 48u; }

// pred_u8 -> clip(pred/255*80, 1, 50) exactly as numpy evaluates it (two separate roundings).
__device__ __forceinline__ double pred_value(int p) {
  double v = __dmul_rn(__ddiv_rn(static_cast<double>(p), 255.0), 80.0);
  if (v < 1.0) v = 1.0;
  if (v > 50.0) v = 50.0;
  return v;
}

__device__ __forceinline__ void hist_add(uint32_t* hist, uint32_t g, uint32_t p) {
  if (masked(g)) atomicAdd(&hist[(g - 2u) * 256u + p], 1u);
}

// T threads per image; kOcc blocks per SM: several images per SM overlap their zero / load / evaluate phases
template <int T, int kOcc>
__global__ void __launch_bounds__(T, kOcc)
metrics_hist_kernel(const uint8_t* __restrict__ gt, const uint8_t* __restrict__ pred, int64_t pixels,
                    double* __restrict__ out) {
  extern __shared__ uint32_t hist[];            // [48][256]
  __shared__ double vd[256], lvd[256];
  __shared__ double red[32][8];
  __shared__ int s_min[32], s_max[32];
  __shared__ int s_pmin, s_pmax;
  const int tid = threadIdx.x;
  const int img = blockIdx.x;
  for (int i = tid; i < kHistCells; i += T) hist[i] = 0u;
  __syncthreads();
  const uint8_t* g = gt + img * pixels;
  const uint8_t* p = pred + img * pixels;
  // scalar head up to 16-byte alignment of gt, vector body when pred shares the alignment, scalar tail
  const int64_t head_end = min(pixels, static_cast<int64_t>((16 - (reinterpret_cast<uintptr_t>(g) & 15)) & 15));
  for (int64_t i = tid; i < head_end; i += T) hist_add(hist, g[i], p[i]);
  const bool same_align = ((reinterpret_cast<uintptr_t>(g) ^ reinterpret_cast<uintptr_t>(p)) & 15) == 0;
  int64_t body_end = head_end;
  if (same_align) {
    const int64_t nvec = (pixels - head_end) / 16;
    body_end = head_end + nvec * 16;
    const uint4* gv4 = reinterpret_cast<const uint4*>(g + head_end);
    const uint4* pv4 = reinterpret_cast<const uint4*>(p + head_end);
    int64_t vi = tid;
    // two vectors per iteration: four 16-byte loads in flight per thread
    for (; vi + T < nvec; vi += 2 * T) {
      const uint4 ga = gv4[vi], pa = pv4[vi];
      const uint4 gb = gv4[vi + T], pb = pv4[vi + T];
      const uint32_t gw[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const uint32_t pw[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) hist_add(hist, (gw[k] >> (8 * j)) & 255u, (pw[k] >> (8 * j)) & 255u);
    }
    for (; vi < nvec; vi += T) {
      const uint4 ga = gv4[vi], pa = pv4[vi];
      const uint32_t gw[4] = {ga.x, ga.y, ga.z, ga.w};
      const uint32_t pw[4] = {pa.x, pa.y, pa.z, pa.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) hist_add(hist, (gw[k] >> (8 * j)) & 255u, (pw[k] >> (8 * j)) & 255u);
    }
  }
  for (int64_t i = body_end + tid; i < pixels; i += T) hist_add(hist, g[i], p[i]);
  __syncthreads();

  // ---- occupied prediction range
  int mn = 256, mx = -1;
  if (tid < 256) {
    uint32_t any = 0u;
    for (int r = 0; r < kHistRows; ++r) any |= hist[r * 256 + tid];
    if (any) mn = mx = tid;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    s_min[warp] = mn;
    s_max[warp] = mx;
  }
  __syncthreads();
  if (tid == 0) {
    int a = 256, b = -1;
    for (int w = 0; w < T / 32; ++w) {
      a = min(a, s_min[w]);
      b = max(b, s_max[w]);
    }
    s_pmin = a;
    s_pmax = b;
  }
  __syncthreads();
  const int pmin = s_pmin, pmax = s_pmax;
  if (tid < 256) {
    // (x - min) / (max - min) * 49 + 1, each operation rounded separately (no FMA)
    const double lo = pred_value(pmin < 256 ? pmin : 0), hi = pred_value(pmax >= 0 ? pmax : 0);
    const double v = __dadd_rn(__dmul_rn(__ddiv_rn(__dsub_rn(pred_value(tid), lo), __dsub_rn(hi, lo)), 49.0), 1.0);
    vd[tid] = v;
    lvd[tid] = log(v);
  }
  __syncthreads();

  // ---- sums over the histogram cells
  double sq = 0.0, lg = 0.0, ar = 0.0, sr = 0.0, cnt = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  for (int i = tid; i < kHistCells; i += T) {
    const uint32_t c = hist[i];
    if (c == 0u) continue;
    const int gi = (i >> 8) + 2, pi = i & 255;
    const double cd = static_cast<double>(c), gd = static_cast<double>(gi), v = vd[pi];
    const double d = __dsub_rn(gd, v);
    const double d2 = __dmul_rn(d, d);
    const double dl = __dsub_rn(static_cast<double>(kLogU8[gi]), lvd[pi]);
    sq += cd * d2;
    lg += cd * __dmul_rn(dl, dl);
    ar += cd * __ddiv_rn(fabs(d), gd);
    sr += cd * __ddiv_rn(d2, gd);
    cnt += cd;
    const double qa = __ddiv_rn(gd, v), qb = __ddiv_rn(v, gd);
    const double m = fmax(qa, qb);          // np.maximum propagates NaN; NaN < thr is false either way
    const bool nan = (qa != qa) || (qb != qb);
    if (!nan) {
      if (m < 1.25) a1 += cd;
      if (m < 1.5625) a2 += cd;
      if (m < 1.953125) a3 += cd;
    }
  }
  double vals[8] = {sq, lg, ar, sr, cnt, a1, a2, a3};
#pragma unroll
  for (int k = 0; k < 8; ++k)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) vals[k] += __shfl_xor_sync(0xffffffffu, vals[k], o);
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[warp][k] = vals[k];
  __syncthreads();
  if (tid == 0) {
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int w = 0; w < T / 32; ++w)
      for (int k = 0; k < 8; ++k) s[k] += red[w][k];
    const double n = s[4];
    double* o = out + static_cast<int64_t>(img) * 8;   // {abs_rel, sq_rel, rmse, rmse_log, a1, a2, a3, count}
    o[0] = s[2] / n;
    o[1] = s[3] / n;
    o[2] = sqrt(s[0] / n);
    o[3] = sqrt(s[1] / n);
    o[4] = s[5] / n;
    o[5] = s[6] / n;
    o[6] = s[7] / n;
    o[7] = n;
  }
}

}  // namespace cdb

using namespace cdb;

extern "C" size_t cdb_depth_metrics_workspace(int32_t n_img) {
  (void)n_img;
  return 256;  // the single-pass kernel keeps all intermediate state on chip
}

extern "C" int cdb_depth_metrics(const uint8_t* gt, const uint8_t* pred, int32_t n_img, int32_t h, int32_t w,
                                 double* out8_per_img, void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  (void)workspace;
  (void)ws_bytes;
  CDB_REQUIRE(gt && pred && out8_per_img && n_img > 0 && h > 0 && w > 0, CDB_ERR_BAD_DESC,
              "depth_metrics: bad argument");
  const int64_t pixels = (int64_t)h * w;
  const size_t smem = (size_t)kHistCells * sizeof(uint32_t);
  // 512 threads x 3 resident blocks (40 registers): 0.222 ms for 697 KITTI-sized pairs vs 0.263 ms with one 1024-thread block per SM
  // (tools/ab_metrics.sh; 256 x 4: 0.312 ms); 513 selects that default explicitly
  static const int threads = getenv("CDB_METRICS_THREADS") ? atoi(getenv("CDB_METRICS_THREADS")) : 513;
  static bool attr_set = false;
  if (!attr_set) {
    CDB_CUDA_OK(cudaFuncSetAttribute(metrics_hist_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CDB_CUDA_OK(cudaFuncSetAttribute(metrics_hist_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CDB_CUDA_OK(cudaFuncSetAttribute(metrics_hist_kernel<512, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CDB_CUDA_OK(cudaFuncSetAttribute(metrics_hist_kernel<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (threads == 512) metrics_hist_kernel<512, 2><<<n_img, 512, smem, stream>>>(gt, pred, pixels, out8_per_img);
  else if (threads == 513) metrics_hist_kernel<512, 3><<<n_img, 512, smem, stream>>>(gt, pred, pixels, out8_per_img);
  else if (threads == 256) metrics_hist_kernel<256, 4><<<n_img, 256, smem, stream>>>(gt, pred, pixels, out8_per_img);
  else metrics_hist_kernel<1024, 1><<<n_img, 1024, smem, stream>>>(gt, pred, pixels, out8_per_img);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

// ------------------------------------------------------------------------------------------------
// Device-side validation path: what new_multi/train5.py:97-110 + util/util.py:51-65 +
// my_eval.py:52-56 do through a PNG round trip, without leaving the GPU:
//   u   = uint8((x + 1) / 2 * 255)                      tensor2im (float32 arithmetic, C truncation / wrap)
//   p8  = cvRound(u / max_img(u) * 255)                 train5.py:100,110 + cv2.imwrite of a float64 image
//   out = cv2.resize(p8, (W_gt, H_gt))                  INTER_LINEAR on uint8 = OpenCV's 11-bit fixed point
// The resize reproduces OpenCV bit for bit (pinned against cv2 in tests/test_validation_path.py): column
// weights clamp to {1, 0} at the borders, row weights keep their fraction and clamp the two row INDICES;
// horizontal pass in int32 (x2048), vertical pass ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2.
// ------------------------------------------------------------------------------------------------
namespace cdb {

__global__ void __launch_bounds__(256)
pred_quant_kernel(const float* __restrict__ x, int64_t pixels, uint8_t* __restrict__ u8, int* __restrict__ img_max) {
  const int img = blockIdx.y;
  const float* src = x + img * pixels;
  uint8_t* dst = u8 + img * pixels;
  int mx = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    // (x + 1) / 2.0 * 255.0 evaluated in float32 with separate roundings, then numpy's float32 -> uint8 cast
    const float v = __fmul_rn(__fdiv_rn(__fadd_rn(src[i], 1.0f), 2.0f), 255.0f);
    const int q = static_cast<int>(static_cast<long long>(v)) & 255;   // truncation, wrap modulo 256
    dst[i] = static_cast<uint8_t>(q);
    mx = max(mx, q);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&img_max[img], mx);
}

__global__ void __launch_bounds__(256)
pred_normalise_kernel(uint8_t* __restrict__ u8, int64_t pixels, const int* __restrict__ img_max) {
  const int img = blockIdx.y;
  uint8_t* p = u8 + img * pixels;
  const double m = static_cast<double>(img_max[img]);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pixels; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = __dmul_rn(__ddiv_rn(static_cast<double>(p[i]), m), 255.0);
    int r = __double2int_rn(v);          // cvRound: round half to even; saturate_cast<uchar>
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    p[i] = static_cast<uint8_t>(r);
  }
}

// ofs[d], alpha[d][2] for one axis. clamp_weights: OpenCV clamps the WEIGHTS along x (fx = 0 at the borders)
// and only the INDICES along y.
__global__ void resize_table_kernel(int ssize, int dsize, int clamp_weights, int* __restrict__ ofs,
                                    short* __restrict__ alpha, int* __restrict__ dmax) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= dsize) return;
  const double scale = static_cast<double>(ssize) / static_cast<double>(dsize);
  float f = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (clamp_weights) {
    if (s < 0) {
      f = 0.f;
      s = 0;
    }
    if (s + 1 >= ssize) {
      atomicMin(dmax, d);
      if (s >= ssize - 1) {
        f = 0.f;
        s = ssize - 1;
      }
    }
  }
  ofs[d] = s;
  alpha[2 * d] = static_cast<short>(__float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f)));
  alpha[2 * d + 1] = static_cast<short>(__float2int_rn(__fmul_rn(f, 2048.f)));
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

__global__ void __launch_bounds__(256)
resize_linear_u8_kernel(const uint8_t* __restrict__ src, int sh, int sw, uint8_t* __restrict__ dst, int dh, int dw,
                        const int* __restrict__ xofs, const short* __restrict__ xa, const int* __restrict__ yofs,
                        const short* __restrict__ ya, const int* __restrict__ xmax_p) {
  const int img = blockIdx.z;
  const uint8_t* s = src + static_cast<int64_t>(img) * sh * sw;
  uint8_t* o = dst + static_cast<int64_t>(img) * dh * dw;
  const int xmax = *xmax_p;
  const int64_t total = static_cast<int64_t>(dh) * dw;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int dx = static_cast<int>(idx % dw), dy = static_cast<int>(idx / dw);
    const int sx = xofs[dx];
    const int sy = yofs[dy];
    const int y0 = min(max(sy, 0), sh - 1), y1 = min(max(sy + 1, 0), sh - 1);
    int h0, h1;
    if (dx >= xmax) {
      h0 = static_cast<int>(s[y0 * sw + sx]) * 2048;
      h1 = static_cast<int>(s[y1 * sw + sx]) * 2048;
    } else {
      const int a0 = xa[2 * dx], a1 = xa[2 * dx + 1];
      const int x1 = min(sx + 1, sw - 1);
      h0 = s[y0 * sw + sx] * a0 + s[y0 * sw + x1] * a1;
      h1 = s[y1 * sw + sx] * a0 + s[y1 * sw + x1] * a1;
    }
    const int b0 = ya[2 * dy], b1 = ya[2 * dy + 1];
    int r = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
    r = r < 0 ? 0 : (r > 255 ? 255 : r);
    o[idx] = static_cast<uint8_t>(r);
  }
}

}  // namespace cdb

extern "C" size_t cdb_validation_workspace(int32_t n_img, int32_t dh, int32_t dw) {
  // per-image max (int) | xmax (int) | xofs[dw] yofs[dh] (int) | xalpha[2 dw] ybeta[2 dh] (short)
  size_t ints = (size_t)n_img + 1 + dw + dh;
  size_t shorts = 2 * (size_t)dw + 2 * (size_t)dh;
  return ((ints * 4 + shorts * 2 + 255) & ~(size_t)255) + 256;
}

extern "C" int cdb_depth_pred_to_u8(const float* pred, int32_t n_img, int32_t h, int32_t w, uint8_t* out_u8,
                                    void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(pred && out_u8 && workspace && n_img > 0 && h > 0 && w > 0 && ws_bytes >= (size_t)n_img * 4,
              CDB_ERR_BAD_DESC, "depth_pred_to_u8: bad argument");
  int* img_max = static_cast<int*>(workspace);
  CDB_CUDA_OK(cudaMemsetAsync(img_max, 0, sizeof(int) * n_img, stream));
  const int64_t pixels = (int64_t)h * w;
  int chunks = (int)((pixels + 256 * 16 - 1) / (256 * 16));
  if (chunks > 64) chunks = 64;
  if (chunks < 1) chunks = 1;
  dim3 grid(chunks, n_img);
  pred_quant_kernel<<<grid, 256, 0, stream>>>(pred, pixels, out_u8, img_max);
  CDB_LAUNCH_OK();
  pred_normalise_kernel<<<grid, 256, 0, stream>>>(out_u8, pixels, img_max);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_resize_linear_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, uint8_t* dst,
                                    int32_t dh, int32_t dw, void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst && workspace && n_img > 0 && sh > 0 && sw > 0 && dh > 0 && dw > 0, CDB_ERR_BAD_DESC,
              "resize_linear_u8: bad argument");
  CDB_REQUIRE(ws_bytes >= cdb_validation_workspace(n_img, dh, dw), CDB_ERR_WORKSPACE, "resize_linear_u8: workspace");
  CDB_REQUIRE(!(sw == 2 * dw && sh == 2 * dh), CDB_ERR_UNSUPPORTED,
              "resize_linear_u8: exact 2x decimation (OpenCV switches to INTER_AREA there)");
  int* ints = static_cast<int*>(workspace) + n_img;   // after the per-image maxima of cdb_depth_pred_to_u8
  int* xmax = ints;
  int* xofs = ints + 1;
  int* yofs = xofs + dw;
  short* xa = reinterpret_cast<short*>(yofs + dh);
  short* ya = xa + 2 * dw;
  set_int_kernel<<<1, 1, 0, stream>>>(xmax, dw);
  CDB_LAUNCH_OK();
  resize_table_kernel<<<ceil_div(dw, 128), 128, 0, stream>>>(sw, dw, 1, xofs, xa, xmax);
  CDB_LAUNCH_OK();
  resize_table_kernel<<<ceil_div(dh, 128), 128, 0, stream>>>(sh, dh, 0, yofs, ya, xmax);
  CDB_LAUNCH_OK();
  const int64_t total = (int64_t)dh * dw;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  resize_linear_u8_kernel<<<dim3(blocks, 1, n_img), 256, 0, stream>>>(src, sh, sw, dst, dh, dw, xofs, xa, yofs, ya, xmax);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
