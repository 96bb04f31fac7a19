// SURVEY 8(f) row f4 — device side of the loaders' per-sample arithmetic (what sits between the image decode /
// resize and the training step): multi-range depth labels, label-id remapping, ToTensor + Normalize.
// All three are HBM-bound elementwise passes; results are bit-identical to the numpy / torch statements they
// replace (IEEE single ops in the reference's order, no FMA contraction, no reciprocal multiplication).
#include "common.cuh"

namespace cdb {

// order-preserving float <-> uint32 map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2ord(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

__global__ void depth_minmax_init_kernel(uint32_t* __restrict__ mm, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    mm[2 * i] = 0xffffffffu;  // min
    mm[2 * i + 1] = 0u;       // max
  }
}

// grid (chunks, n): per-image minimum / maximum of the raw depth map
__global__ void __launch_bounds__(256) depth_minmax_kernel(const float* __restrict__ d, int64_t hw, uint32_t* __restrict__ mm) {
  const float* img = d + static_cast<int64_t>(blockIdx.y) * hw;
  uint32_t lo = 0xffffffffu, hi = 0u;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint32_t o = f2ord(img[i]);
    lo = min(lo, o);
    hi = max(hi, o);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, off));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(mm + 2 * blockIdx.y, lo);
    atomicMax(mm + 2 * blockIdx.y + 1, hi);
  }
}

__device__ __forceinline__ float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
// 2 * (v - mn) / (mx - mn) - 1 with numpy's float32 evaluation order
__device__ __forceinline__ float norm_pm1(float v, float mn, float range) {
  return __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fsub_rn(v, mn)), range), 1.f);
}

// new_multi/try_data.py:240-272. With d = raw depth and c(lo, hi) = clamp(d, lo, hi):
//   dep_l      = N(min(d, 8000))
//   labels[0]  = N(c(5000, 8000)), labels[1] = N(c(3000, 6000)), labels[2] = N(c(1000, 4000)),
//   labels[3]  = 2 * (min(d, 2000) - m4) / (max5 - min5) - 1   with m4 = min of labels[2] (the reference subtracts the
//                minimum of the already NORMALISED fourth range, :266: -1, or NaN when that range is degenerate)
// where N(x) = 2 * (x - min x) / (max x - min x) - 1 per image. Clamping is monotone, so every per-image min / max
// follows from the raw minimum and maximum.
__global__ void __launch_bounds__(256)
depth_labels_kernel(const float* __restrict__ d, int64_t hw, const uint32_t* __restrict__ mm, float* __restrict__ dep_l,
                    float* __restrict__ labels) {
  const int n = blockIdx.y;
  const float omin = ord2f(mm[2 * n]), omax = ord2f(mm[2 * n + 1]);
  const float inf = __int_as_float(0x7f800000);
  const float mn0 = fminf(omin, 8000.f), mx0 = fminf(omax, 8000.f);
  const float mn2 = clampf(omin, 5000.f, 8000.f), mx2 = clampf(omax, 5000.f, 8000.f);
  const float mn3 = clampf(omin, 3000.f, 6000.f), mx3 = clampf(omax, 3000.f, 6000.f);
  const float mn4 = clampf(omin, 1000.f, 4000.f), mx4 = clampf(omax, 1000.f, 4000.f);
  const float mn5 = fminf(omin, 2000.f), mx5 = fminf(omax, 2000.f);
  const float r0 = __fsub_rn(mx0, mn0), r2 = __fsub_rn(mx2, mn2), r3 = __fsub_rn(mx3, mn3), r4 = __fsub_rn(mx4, mn4),
              r5 = __fsub_rn(mx5, mn5);
  const float m4 = r4 == 0.f ? __fsub_rn(inf, inf) : -1.f;   // min of the normalised fourth range
  const float* img = d + static_cast<int64_t>(n) * hw;
  float* o0 = dep_l + static_cast<int64_t>(n) * hw;
  float* o = labels + static_cast<int64_t>(n) * 4 * hw;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float v = img[i];
    o0[i] = norm_pm1(fminf(v, 8000.f), mn0, r0);
    o[i] = norm_pm1(clampf(v, 5000.f, 8000.f), mn2, r2);
    o[hw + i] = norm_pm1(clampf(v, 3000.f, 6000.f), mn3, r3);
    o[2 * hw + i] = norm_pm1(clampf(v, 1000.f, 4000.f), mn4, r4);
    o[3 * hw + i] = norm_pm1(fminf(v, 2000.f), m4, r5);
  }
}

// dst[i] = lut[src[i]] as int64 class ids (MaskToTensor, new_multi/try_data.py:26-28)
__global__ void __launch_bounds__(256)
label_lut_kernel(const uint8_t* __restrict__ src, int64_t numel, const uint8_t* __restrict__ lut, int64_t* __restrict__ dst) {
  __shared__ uint8_t s_lut[256];
  s_lut[threadIdx.x] = lut[threadIdx.x];
  __syncthreads();
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < numel;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    dst[i] = static_cast<int64_t>(s_lut[src[i]]);
}

// transforms.ToTensor() + transforms.Normalize(mean, std): dst[n][c][h][w] = (src[n][h][w][c] / 255 - mean) / std
__global__ void __launch_bounds__(256)
image_normalize_kernel(const uint8_t* __restrict__ src, int64_t hw, int c, float mean, float stdv, float* __restrict__ dst) {
  const int n = blockIdx.y;
  const uint8_t* s = src + static_cast<int64_t>(n) * hw * c;
  float* o = dst + static_cast<int64_t>(n) * hw * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hw;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    for (int ch = 0; ch < c; ++ch) {
      const float v = __fdiv_rn(static_cast<float>(s[i * c + ch]), 255.f);
      o[static_cast<int64_t>(ch) * hw + i] = __fdiv_rn(__fsub_rn(v, mean), stdv);
    }
  }
}


// ---- Pillow's ImagingResample, 8 bits per channel (third-party arithmetic behind Image.resize(size, BILINEAR) of
// datasets/dataset_synthia.py:154-161 / new_multi/try_data.py:164-167): a separable filter whose per-output-coordinate
// windows [xmin, xmin + xmax) and 22-bit fixed-point coefficients are built on the host in double precision exactly as
// precompute_coeffs / normalize_coeffs_8bpc do; each pass accumulates in int32 from 1 << 21 and clips (ss >> 22) to a byte.
constexpr int kPilPrecisionBits = 32 - 8 - 2;

__device__ __forceinline__ uint8_t pil_clip8(int v) {
  v >>= kPilPrecisionBits;  // arithmetic shift, as the table index of clip8_lookups
  return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [n][h][sw][c] -> dst [n][h][dw][c]; one thread per output byte
__global__ void __launch_bounds__(256)
pil_resample_h_kernel(const uint8_t* __restrict__ src, int64_t rows, int sw, int dw, int c, const int32_t* __restrict__ bounds,
                      const int32_t* __restrict__ kk, int ksize, uint8_t* __restrict__ dst) {
  const int64_t total = rows * dw * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    const int64_t t = i / c;
    const int xx = static_cast<int>(t % dw);
    const int64_t row = t / dw;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int32_t* k = kk + static_cast<int64_t>(xx) * ksize;
    const uint8_t* s = src + (row * sw + xmin) * c + ch;
    int ss = 1 << (kPilPrecisionBits - 1);
    for (int x = 0; x < xmax; ++x) ss += static_cast<int>(s[static_cast<int64_t>(x) * c]) * k[x];
    dst[i] = pil_clip8(ss);
  }
}

// vertical pass: src [n][sh][w][c] -> dst [n][dh][w][c]
__global__ void __launch_bounds__(256)
pil_resample_v_kernel(const uint8_t* __restrict__ src, int n, int sh, int dh, int64_t wc, const int32_t* __restrict__ bounds,
                      const int32_t* __restrict__ kk, int ksize, uint8_t* __restrict__ dst) {
  const int64_t total = static_cast<int64_t>(n) * dh * wc;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t col = i % wc;
    const int64_t t = i / wc;
    const int yy = static_cast<int>(t % dh);
    const int64_t img = t / dh;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int32_t* k = kk + static_cast<int64_t>(yy) * ksize;
    const uint8_t* s = src + (img * sh + ymin) * wc + col;
    int ss = 1 << (kPilPrecisionBits - 1);
    for (int y = 0; y < ymax; ++y) ss += static_cast<int>(s[static_cast<int64_t>(y) * wc]) * k[y];
    dst[i] = pil_clip8(ss);
  }
}

// dst[n][y][x][:] = src[n][ytab[y]][xtab[x]][:] (a negative table entry writes zeros): Pillow's NEAREST resize
// (ImagingScaleAffine's pretabulated positions, built on the host) and FLIP_LEFT_RIGHT / FLIP_TOP_BOTTOM.
__global__ void __launch_bounds__(256)
gather_rows_cols_kernel(const uint8_t* __restrict__ src, int n, int sh, int sw, int c, int dh, int dw,
                        const int32_t* __restrict__ ytab, const int32_t* __restrict__ xtab, uint8_t* __restrict__ dst) {
  const int64_t total = static_cast<int64_t>(n) * dh * dw * c;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % c);
    int64_t t = i / c;
    const int x = static_cast<int>(t % dw);
    t /= dw;
    const int y = static_cast<int>(t % dh);
    const int64_t img = t / dh;
    const int sy = ytab[y], sx = xtab[x];
    dst[i] = (sy < 0 || sx < 0) ? 0 : src[((img * sh + sy) * sw + sx) * c + ch];
  }
}

static int grid_for(int64_t per_image, int n) {
  int64_t b = (per_image + 256 * 8 - 1) / (256 * 8);
  const int64_t cap = (int64_t)sm_count() * 8 / (n > 0 ? n : 1) + 1;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}

}  // namespace cdb

using namespace cdb;

extern "C" size_t cdb_depth_labels_workspace(int32_t n_img) { return (size_t)(n_img > 0 ? n_img : 0) * 2 * sizeof(uint32_t); }

extern "C" int cdb_depth_labels(const float* depth, int32_t n_img, int64_t hw, float* dep_l, float* depth_l_s,
                                void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(depth && dep_l && depth_l_s && workspace, CDB_ERR_BAD_DESC, "depth_labels: null argument");
  CDB_REQUIRE(n_img >= 0 && hw >= 1, CDB_ERR_BAD_DESC, "depth_labels: bad sizes");
  CDB_REQUIRE(ws_bytes >= cdb_depth_labels_workspace(n_img), CDB_ERR_WORKSPACE, "depth_labels: workspace too small");
  CDB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 3) == 0, CDB_ERR_ALIGNMENT, "depth_labels: workspace alignment");
  if (n_img == 0) return CDB_OK;
  uint32_t* mm = static_cast<uint32_t*>(workspace);
  depth_minmax_init_kernel<<<ceil_div(n_img, 128), 128, 0, stream>>>(mm, n_img);
  CDB_LAUNCH_OK();
  dim3 grid(grid_for(hw, n_img), n_img);
  depth_minmax_kernel<<<grid, 256, 0, stream>>>(depth, hw, mm);
  CDB_LAUNCH_OK();
  depth_labels_kernel<<<grid, 256, 0, stream>>>(depth, hw, mm, dep_l, depth_l_s);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_label_lut_i64(const uint8_t* src, int64_t numel, const uint8_t* lut_dev, int64_t* dst,
                                 cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && lut_dev && dst, CDB_ERR_BAD_DESC, "label_lut_i64: null argument");
  if (numel <= 0) return CDB_OK;
  int64_t b = (numel + 256 * 8 - 1) / (256 * 8);
  if (b > (int64_t)sm_count() * 8) b = (int64_t)sm_count() * 8;
  label_lut_kernel<<<(int)b, 256, 0, stream>>>(src, numel, lut_dev, dst);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" int cdb_image_normalize_u8(const uint8_t* src, int32_t n_img, int64_t hw, int32_t channels, float mean,
                                      float stdv, float* dst, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst, CDB_ERR_BAD_DESC, "image_normalize_u8: null argument");
  CDB_REQUIRE(n_img >= 0 && hw >= 1 && channels >= 1 && channels <= 8, CDB_ERR_BAD_DESC, "image_normalize_u8: bad sizes");
  if (n_img == 0) return CDB_OK;
  dim3 grid(grid_for(hw, n_img), n_img);
  image_normalize_kernel<<<grid, 256, 0, stream>>>(src, hw, channels, mean, stdv, dst);
  CDB_LAUNCH_OK();
  return CDB_OK;
}

extern "C" size_t cdb_pil_resample_workspace(int32_t n_img, int32_t sh, int32_t dw, int32_t channels) {
  if (n_img <= 0 || sh <= 0 || dw <= 0 || channels <= 0) return 0;
  return (size_t)n_img * sh * dw * channels;
}

extern "C" int cdb_pil_resample_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, int32_t channels,
                                   uint8_t* dst, int32_t dh, int32_t dw, const int32_t* bounds_x, const int32_t* kk_x,
                                   int32_t ksize_x, const int32_t* bounds_y, const int32_t* kk_y, int32_t ksize_y,
                                   void* workspace, size_t ws_bytes, cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst, CDB_ERR_BAD_DESC, "pil_resample_u8: null argument");
  CDB_REQUIRE(n_img >= 0 && sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && channels >= 1 && channels <= 8, CDB_ERR_BAD_DESC,
              "pil_resample_u8: bad sizes");
  // Pillow runs a pass only when the size changes along it (ImagingResample: need_horizontal / need_vertical)
  const bool horiz = dw != sw, vert = dh != sh;
  CDB_REQUIRE(!horiz || (bounds_x && kk_x && ksize_x >= 1), CDB_ERR_BAD_DESC, "pil_resample_u8: horizontal tables missing");
  CDB_REQUIRE(!vert || (bounds_y && kk_y && ksize_y >= 1), CDB_ERR_BAD_DESC, "pil_resample_u8: vertical tables missing");
  if (n_img == 0) return CDB_OK;
  const int64_t out_bytes = (int64_t)n_img * dh * dw * channels;
  auto blocks = [](int64_t total) {
    int64_t b = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    return (int)(b > cap ? cap : (b < 1 ? 1 : b));
  };
  if (!horiz && !vert) {
    CDB_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)out_bytes, cudaMemcpyDeviceToDevice, stream));
    return CDB_OK;
  }
  const uint8_t* mid = src;
  if (horiz) {
    uint8_t* hout = dst;
    if (vert) {
      CDB_REQUIRE(workspace && ws_bytes >= cdb_pil_resample_workspace(n_img, sh, dw, channels), CDB_ERR_WORKSPACE,
                  "pil_resample_u8: workspace too small");
      hout = static_cast<uint8_t*>(workspace);
    }
    const int64_t total = (int64_t)n_img * sh * dw * channels;
    pil_resample_h_kernel<<<blocks(total), 256, 0, stream>>>(src, (int64_t)n_img * sh, sw, dw, channels, bounds_x, kk_x,
                                                            ksize_x, hout);
    CDB_LAUNCH_OK();
    mid = hout;
  }
  if (vert) {
    pil_resample_v_kernel<<<blocks(out_bytes), 256, 0, stream>>>(mid, n_img, sh, dh, (int64_t)dw * channels, bounds_y, kk_y,
                                                               ksize_y, dst);
    CDB_LAUNCH_OK();
  }
  return CDB_OK;
}

extern "C" int cdb_gather_rows_cols_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, int32_t channels,
                                       uint8_t* dst, int32_t dh, int32_t dw, const int32_t* ytab, const int32_t* xtab,
                                       cdbStream_t stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CDB_REQUIRE(src && dst && ytab && xtab, CDB_ERR_BAD_DESC, "gather_rows_cols_u8: null argument");
  CDB_REQUIRE(n_img >= 0 && sh >= 1 && sw >= 1 && dh >= 1 && dw >= 1 && channels >= 1, CDB_ERR_BAD_DESC,
              "gather_rows_cols_u8: bad sizes");
  if (n_img == 0) return CDB_OK;
  const int64_t total = (int64_t)n_img * dh * dw * channels;
  int64_t b = (total + 255) / 256;
  if (b > (int64_t)sm_count() * 16) b = (int64_t)sm_count() * 16;
  gather_rows_cols_kernel<<<(int)b, 256, 0, stream>>>(src, n_img, sh, sw, channels, dh, dw, ytab, xtab, dst);
  CDB_LAUNCH_OK();
  return CDB_OK;
}
