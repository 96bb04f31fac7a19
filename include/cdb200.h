/*
 * cdb200.h — C ABI of libcdb200.so, the B200 (sm_100a) kernel library behind the drop-in
 * replacements for the dense-convolution hot path of JosephineRabbit/cycle_depth_estimation.
 *
 * The reference has no native layer: every entry point below replaces the ATen/cuDNN call that a
 * torch.nn module of the reference reaches (reference file:line given per function).  All pointers
 * are DEVICE pointers owned by the caller unless stated otherwise; the library never allocates,
 * frees or keeps device memory, launches only on the stream it is given, never synchronises and is
 * CUDA-graph capturable.  Every function returns 0 on success or a negative CdbStatus and records a
 * message retrievable with cdb_last_error() (thread local).
 *
 * Activations are NHWC ("channels last"), channel stride 1, bf16 (dtype 0) or fp32 (dtype 1);
 * stored channel counts are multiples of 8 so that every pixel starts on a 16-byte boundary.
 */
#ifndef CDB200_H_
#define CDB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cdbStream_t; /* a cudaStream_t / CUstream */

typedef enum CdbStatus {
  CDB_OK = 0,
  CDB_ERR_BAD_DESC = -1,
  CDB_ERR_UNSUPPORTED = -2,
  CDB_ERR_ALIGNMENT = -3,
  CDB_ERR_WORKSPACE = -4,
  CDB_ERR_CUDA = -5,
  CDB_ERR_DEVICE_ABORT = -6
} CdbStatus;

enum { CDB_BF16 = 0, CDB_F32 = 1 };
enum { CDB_ACT_NONE = 0, CDB_ACT_RELU = 1, CDB_ACT_LEAKY = 2, CDB_ACT_TANH = 3, CDB_ACT_SIGMOID = 4 };
enum { CDB_NORM_NONE = 0, CDB_NORM_INSTANCE = 1, CDB_NORM_BATCH = 2 };
enum { CDB_PAD_ZERO = 0, CDB_PAD_REFLECT = 1 };

/* NHWC activation view. c = stored channels (multiple of 8); strides in elements. */
typedef struct CdbAct {
  void* ptr;
  int32_t n, h, w, c;
  int64_t sn, sh, sw;
  int32_t dtype;
  int32_t reserved;
} CdbAct;

/* Strided output of a convolution: element (n,p,q,ch) goes to ptr[n*sn + p*sh + q*sw + ch*sc].
 * c = number of real output channels; cstore >= c = channels written (the extra ones get 0), used
 * to keep channel-padded NHWC buffers clean.  With sc == 1 and dtype bf16 stores are 16-byte
 * vectors (ptr, sn, sh, sw must then be multiples of 8 elements). */
typedef struct CdbOut {
  void* ptr;
  int32_t n, h, w, c;
  int32_t cstore;
  int32_t dtype;
  int64_t sn, sh, sw, sc;
} CdbOut;

/* Convolution geometry.  Forward semantics on the STORED tensors (out-of-range x reads 0):
 *   transposed == 0:  y[n,p,q,o] = sum_{r,s,i} x[n, p*stride + r*dil - pad_h, q*stride + s*dil - pad_w, i] * W[o,i,r,s]
 *   transposed == 1:  y[n,p,q,o] = sum_{r,s,i : (p+pad_h-r*dil) % stride == 0 ...}
 *                                   x[n, (p+pad_h-r*dil)/stride, (q+pad_w-s*dil)/stride, i] * W[o,i,r,s]
 * where W is the PACKED operand (see cdb_pack_conv_weight); a materialised (reflect/zero) padding
 * is expressed by passing the padded buffer as x with pad_* = 0.
 * rowpack != 0 selects the small-channel image layers: x stores `rowpack` (8 or 16) channels per
 * pixel and one K block covers a whole filter row (s = 0..S-1) of one r; requires transposed == 0,
 * pad_w == 0 (materialised), S*rowpack <= 64 and x.w >= stride*(y.w-1) + 64/rowpack. */
typedef struct CdbConvGeom {
  int32_t r, s;
  int32_t stride;
  int32_t pad_h, pad_w;
  int32_t dil;
  int32_t transposed;
  int32_t rowpack;
  int32_t flip;    /* direct convolutions only: use the packed taps in reverse order, W[R-1-r][S-1-s]; the
                    * data gradient of a stride-1 convolution is then a direct convolution of the
                    * zero-haloed dy (halo (R-1)*dil) with flip = 1 and the dgrad packing of W */
  int32_t reserved;
} CdbConvGeom;

typedef struct CdbEpilogue {
  const float* bias; /* [c] or NULL */
  int32_t act;
  float slope;
  float* stats;      /* optional [n][c][2] fp32 (sum, sum of squares) accumulated with atomics, or NULL */
  int64_t flags;     /* CDB_EP_* */
} CdbEpilogue;
/* stats is ONE group [c][2] for the whole batch (BatchNorm) instead of one per image */
enum { CDB_EP_STATS_BATCH = 1,
       /* fp32 outputs only: round the stored value to TF32 (nearest), so that a following TF32 convolution
        * multiplies exactly what was stored */
       CDB_EP_ROUND_TF32 = 2 };

/* ---- library ------------------------------------------------------------------------------ */
int cdb_version(void);
const char* cdb_last_error(void);
/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
long long cdb_launch_count(void);
/* Reads-and-clears the device-side abort flag raised when a kernel's bounded mbarrier wait timed out. */
int cdb_device_abort_flag(void);

/* ---- K1/K2: tcgen05 implicit-GEMM convolution ----------------------------------------------
 * Replaces aten.convolution reached from nn.Conv2d / nn.ConvTranspose2d
 * (models/networks.py:158,166,178,185,214,229,277,285-301,331-356; new_multi/networks5_ds.py
 * passim) and, with the dgrad packing of the weights, the data-gradient half of
 * aten.convolution_backward.  wpacked: bf16 [rows_pad][taps*kpad] from cdb_pack_conv_weight. */
int cdb_conv2d_fwd(const CdbConvGeom* g, const CdbAct* x, const void* wpacked, int32_t w_rows_pad,
                   int32_t w_kpad, const CdbOut* y, const CdbEpilogue* ep, cdbStream_t stream);
/* TF32 variant (the reference's arithmetic is fp32): x->dtype == CDB_F32 selects tcgen05 kind::tf32 —
 * x is an fp32 NHWC view (channels and strides multiples of 4), wpacked the fp32 operand of
 * cdb_pack_conv_weight_tf32 (w_kpad a multiple of 32), fp32 accumulation, y fp32 or bf16.  The tensor core
 * ignores the low 13 mantissa bits of its operands; cdb_pack_conv_weight_tf32 / cdb_round_tf32 / CDB_EP_ROUND_TF32
 * round to nearest beforehand.  rowpack is not available in this precision. */

/* ---- K1c: image layers (<= 8 input channels, stride 1): Toeplitz operand -----------------------------
 * The 7x7 c7s1-64 input layer of the generators (models/networks.py:158) and the data gradient of the 7x7 c7s1-3
 * output layer (:185, whose contracted tensor is the 3-channel dy).  x: contiguous bf16 buffer [n][h][w][8] with the
 * padding materialised (channels beyond the real ones zero); y: y->h x y->w valid outputs per image,
 *   y[n,p,q,o] = sum_{r,s,c} x[n, p + r, q + s, c] * W[o,c,r,s]        (R, S <= 8, <= 128 output channels)
 * wpacked from cdb_pack_toeplitz_weight.  A filter row is ONE K block (8 taps x 8 channels) whose A operand the tensor
 * core reads from a plain copy of the pixel row through an overlapping no-swizzle descriptor: every input byte is
 * fetched once per filter row instead of 8 times (conv_toeplitz.cu).  Epilogue as cdb_conv2d_fwd (bias, activation,
 * InstanceNorm / BatchNorm sums).  Fast output (TMA stores): y laid out by the input pitch, i.e. y->sw == cstore,
 * y->sh == x->w * cstore, y->sn a multiple of 128 rows, cstore a multiple of 64. */
int cdb_conv2d_toeplitz_fwd(const CdbAct* x, const void* wpacked, int32_t w_rows_pad, int32_t r, int32_t s,
                            const CdbOut* y, const CdbEpilogue* ep, cdbStream_t stream);
/* fp32 W4[d0][d1][R][S] -> bf16 [R][rows_pad x 64] K-major blocks in the no-swizzle core-matrix order, k = s * 8 + c;
 * rows = d0 (rows_are_dim0) or d1 (data gradient), the other dimension (<= 8) is contracted; flip packs
 * W[..][R-1-r][S-1-s].  out holds R * round_up(rows, 16) * 64 bf16. */
int cdb_pack_toeplitz_weight(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s, int32_t rows_are_dim0,
                             int32_t flip, void* out, cdbStream_t stream);

/* Weight gradient of the image layers (the filter-gradient half of aten.convolution_backward for
 * models/networks.py:158,185):
 *   dw[..][r][s] (+)= sum_{n,h,w} S[n,h,w,m] * P[n, h + r', w + s', c],   (r', s') = (r, s) or (R-1-r, S-1-s) with flip
 * S = s_act: the bf16 NHWC tensor with up to 128 channels whose pixels are iterated (dy of c7s1-64; the padded
 * 64-channel input of c7s1-3); P = p_act: the contiguous buffer of 8-channel pixels that is shifted (the padded image;
 * the zero-haloed 3-channel dy, with flip); dw4[d0][d1][R][S] has (d0, d1) = (m, c) when m_is_d0, else (c, m).
 * Both operands are MN-major; the shifted one is read from one plain copy of each 71-pixel row segment through an
 * overlapping no-swizzle descriptor (conv_toeplitz.cu).  workspace: cdb_conv2d_toeplitz_wgrad_workspace bytes. */
size_t cdb_conv2d_toeplitz_wgrad_workspace(const CdbAct* s_act, int32_t r);
int cdb_conv2d_toeplitz_wgrad(const CdbAct* s_act, const CdbAct* p_act, int32_t r, int32_t s, float* dw4, int32_t d0,
                              int32_t d1, int32_t m_is_d0, int32_t flip, int32_t accumulate, void* workspace,
                              size_t ws_bytes, cdbStream_t stream);

/* Packs an fp32 4-D filter W4[d0][d1][R][S] (Conv2d: OIHW, ConvTranspose2d: IOHW) into the bf16
 * GEMM B operand [rows_pad][taps*kpad]:  packed[row][tap*kpad + k] = W4[row][k][r][s] when
 * rows_are_dim0, else W4[k][row][r][s]; tap = r*S+s (rowpack: tap = r, k = s*rowpack + ch, kpad 64).
 * rows_pad = round_up(rows,16), kpad = round_up(k,64).  out must hold rows_pad*taps*kpad bf16. */
int cdb_pack_conv_weight(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s,
                         int32_t rows_are_dim0, int32_t rowpack, void* out, cdbStream_t stream);

/* Same packing with fp32 storage rounded to TF32: out holds rows_pad*taps*kpad floats, kpad = round_up(k,32). */
int cdb_pack_conv_weight_tf32(const float* w4, int32_t d0, int32_t d1, int32_t r, int32_t s,
                              int32_t rows_are_dim0, void* out, cdbStream_t stream);
/* In-place round-to-nearest of an fp32 buffer to TF32 precision (activations entering the TF32 path). */
int cdb_round_tf32(float* buf, int64_t numel, cdbStream_t stream);

/* Multi-tensor form: re-packs many filters in one launch (after an optimizer step; entries_host is a HOST array,
 * the pointers travel in the kernel parameter block). Same layout as cdb_pack_conv_weight per entry. */
typedef struct CdbPackEntry {
  const float* w4;
  void* out;
  int32_t d0, d1, r, s, rows_are_dim0, rowpack;
} CdbPackEntry;
int cdb_pack_conv_weights_multi(const CdbPackEntry* entries_host, int32_t n_entries, cdbStream_t stream);

/* ---- K3: weight gradient -------------------------------------------------------------------
 * Replaces the filter-gradient half of aten.convolution_backward.
 *   dW4[d0][d1][r][s] (+)= sum_{n,p,q} dy[n,p,q,o] * x[n, p*stride + r*dil - pad_h, q*stride + s*dil - pad_w, i]
 * with (o,i) = (d0,d1) for Conv2d and x/dy swapped roles for ConvTranspose2d (g->transposed):
 *   transposed: dW4[d0][d1][r][s] (+)= sum_{n,p,q} x[n,p,q,d0] * dy[n, p*stride + r*dil - pad_h, q*stride + s*dil - pad_w, d1]
 * g->rowpack refers to the shifted tensor (x, or dy when transposed), see CdbConvGeom.
 * workspace: cdb_conv2d_wgrad_workspace() bytes of fp32 split-K partials. */
size_t cdb_conv2d_wgrad_workspace(const CdbConvGeom* g, const CdbAct* x, const CdbAct* dy);
/* x and dy both CDB_F32 selects the TF32 variant (kind::tf32, MN-major fp32 operands; no rowpack). */
int cdb_conv2d_wgrad(const CdbConvGeom* g, const CdbAct* x, const CdbAct* dy, float* dw4,
                     int32_t d0, int32_t d1, int32_t accumulate, void* workspace, size_t ws_bytes,
                     cdbStream_t stream);

/* ---- K4: statistics, norm + activation (+ residual, + reflect halo) and their backward --------
 * Replace aten.native_batch_norm(_backward) (nn.InstanceNorm2d / nn.BatchNorm2d),
 * aten.reflection_pad2d(_backward), relu_/leaky_relu_/threshold_backward and the residual add of
 * models/networks.py:157-188,206-236,277-310,330-356. */
typedef struct CdbNormDesc {
  int32_t norm;           /* CDB_NORM_* */
  int32_t act;            /* CDB_ACT_NONE / RELU / LEAKY (TANH, SIGMOID forward only) */
  float slope, eps;
  int32_t channels;       /* real channels; stored channels = round_up(channels, 8) */
  int32_t pad;            /* reflect halo (pixels) written around / folded from the interior */
  int32_t use_running;    /* eval-mode batch norm: normalise with running_mean / running_var */
  int32_t update_running; /* training-mode batch norm: also update the running statistics */
  float momentum;
  int32_t flags;          /* CDB_NORM_FLAG_* */
  const float* stats;     /* [groups][channels][2] sums (sum, sum of squares); groups = n (instance) or 1 */
  const float* gamma;     /* [channels] or NULL */
  const float* beta;      /* [channels] or NULL */
  float* running_mean;    /* [channels] or NULL */
  float* running_var;     /* [channels] or NULL */
  const float* conv_bias; /* [channels] or NULL: bias of the convolution in front of a training-mode BatchNorm. The
                           * convolution skips it (it cancels in the normalised output), but torch's running_mean
                           * tracks mean(conv + bias): the running update adds it back. */
  float count_scale;      /* data-parallel BatchNorm: `stats` (and, in the apply pass, `bstats`) were summed over this many
                           * equal shards of the batch (the per-layer all-reduce of sum, sum of squares
                           * forward and sum dy, sum dy*xhat backward), so the element count is count_scale * n*h*w.
                           * 0 or 1: single process. */
  int32_t reserved_;
} CdbNormDesc;
/* ACT_FIRST: the layer is conv -> act -> norm (new_multi/networks5_ds.py:636-638,661-676): y is the
 *   ACTIVATED convolution output (the conv epilogue applied `act`), forward = norm only, backward
 *   multiplies the gradient w.r.t. y by act'(y) so that dy is the gradient of the raw convolution.
 * ACCUM_F32: backward only: dy is an fp32 view and receives dy += value (dense-block concatenation
 *   gradients, new_multi/networks5_ds.py:122-146, summed over all consumers of a channel prefix). */
/* BWD_REDUCE_ONLY / BWD_APPLY_ONLY: cdb_norm_act_bwd runs only its reduction pass (fills bstats, writes nothing else) /
 *   only its apply pass (bstats as given), so that the caller can all-reduce bstats over the data-parallel ranks in
 *   between. */
enum { CDB_NORM_FLAG_ACT_FIRST = 1, CDB_NORM_FLAG_ACCUM_F32 = 2, CDB_NORM_FLAG_BWD_REDUCE_ONLY = 4,
       CDB_NORM_FLAG_BWD_APPLY_ONLY = 8 };

/* stats[g][c][2] += (sum, sum of squares) of y over pixels; g = image if per_image else 0. */
int cdb_channel_stats(const CdbAct* y, int32_t c_real, int32_t per_image, float* stats, cdbStream_t stream);
/* out(interior view, same n/h/w as y) = [residual +] act(norm(y)); the reflect halo of d->pad pixels
 * around the interior is written too (the buffer behind `out` must extend that far). */
int cdb_norm_act_fwd(const CdbNormDesc* d, const CdbAct* y, const CdbAct* residual, const CdbAct* out,
                     cdbStream_t stream);
/* g = fold_reflect(dout, d->pad) + dskip;  dy = norm_bwd(act_bwd(g));  optionally gsum = g.
 * bstats[groups][channels][2] (zeroed by the caller) receives (sum ga, sum ga*xhat): the gradients of
 * beta and gamma for an affine norm, and the bias gradient for norm none. */
int cdb_norm_act_bwd(const CdbNormDesc* d, const CdbAct* y, const CdbAct* dout, const CdbAct* dskip,
                     float* bstats, const CdbAct* dy, const CdbAct* gsum, cdbStream_t stream);

/* ---- module boundary: NCHW fp32 <-> NHWC bf16 ------------------------------------------------ */
/* src (fp32, strides in elements) -> interior view `out` (bf16 NHWC, out->c stored channels, extra
 * channels 0) with a reflect halo of `pad`; with act_out != NULL the value is multiplied by the
 * derivative of `act` evaluated at the activation output act_out (same layout as src). */
int cdb_nchw_to_nhwc(const float* src, int32_t n, int32_t c, int32_t h, int32_t w, int64_t s_n, int64_t s_c,
                     int64_t s_h, int64_t s_w, const float* act_out, int32_t act, float slope,
                     const CdbAct* out, int32_t pad, cdbStream_t stream);
/* dst[nc][h][w] (+)= reflect-fold of src[nc][h+2p][w+2p] (backward of nn.ReflectionPad2d on NCHW fp32). */
int cdb_reflect_fold_nchw(const float* src, float* dst, int32_t nc, int32_t h, int32_t w, int32_t pad,
                          int32_t accumulate, cdbStream_t stream);

/* db[c] = sum over n,h,w of g * act'(act_out) on contiguous fp32 NCHW tensors (bias gradient of a final
 * convolution + activation layer; act_out may be NULL for no activation). */
int cdb_bias_grad_nchw(const float* g, const float* act_out, int32_t act, float slope, int32_t n, int32_t c,
                       int64_t hw, float* db, cdbStream_t stream);

/* ---- K6: fused losses (forward scalar + gradient in one pass) ----------------------------------
 * GANLoss (models/networks.py:119-138: MSELoss / BCELoss against a constant label), L1Loss
 * (models/cycle_gan_model.py:63-64,119-134; models/pix2pix_model.py:49,93).
 * *loss_acc += weight * mean(l(x));  grad[i] = weight * l'(x[i]) / numel  (grad may be NULL). */
int cdb_loss_mse_const(const float* x, int64_t numel, float target, float weight, float* loss_acc, float* grad,
                       cdbStream_t stream);
int cdb_loss_bce_const(const float* x, int64_t numel, float target, float weight, float* loss_acc, float* grad,
                       cdbStream_t stream);
int cdb_loss_l1(const float* a, const float* b, int64_t numel, float weight, float* loss_acc, float* grad_a,
                cdbStream_t stream);
int cdb_scale_by_scalar(const float* a, const float* scalar, float* out, int64_t numel, cdbStream_t stream);

/* ---- K5: small NHWC kernels of the seg/depth networks and the U-Net skip topology ---------------
 * (new_multi/networks5_ds.py; models/networks.py:266-316).  All views are bf16 NHWC unless stated. */
/* out = a + b (gradient accumulation at fan-out points; residual / attention adds :337,649,700). */
int cdb_add(const CdbAct* a, const CdbAct* b, const CdbAct* out, cdbStream_t stream);
/* fp32 <-> bf16 view conversion (dense-block gradient accumulators); accumulate: dst(f32) += src(bf16). */
int cdb_cast(const CdbAct* src, const CdbAct* dst, int32_t accumulate, cdbStream_t stream);
/* nn.AvgPool2d(2, 2) (networks5_ds.py:355) and its backward (dx has the shape of the pooling input). */
int cdb_avgpool2_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream);
int cdb_avgpool2_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream);
/* Channel attention (networks5_ds.py:641-649,696-700):  out = [base +] sigmoid(att_sum[n][c][0]*inv_hw) * s
 * where att_sum are the per-image channel sums (cdb_channel_stats layout [n][c][2]) of the tensor that
 * nn.AdaptiveAvgPool2d(1) averages.  Backward: ds = g * sigmoid(.), dsum[n][c] += sum_px g*s, and
 * cdb_gate_bcast writes dt[n,h,w,c] = dsum[n][c] * sigmoid'(.) * inv_hw (gradient of the pooled tensor). */
int cdb_gate_fwd(const CdbAct* base, const CdbAct* s, const float* att_sum, int32_t c_real, float inv_hw,
                 const CdbAct* out, cdbStream_t stream);
int cdb_gate_bwd(const CdbAct* g, const CdbAct* s, const float* att_sum, int32_t c_real, float inv_hw,
                 const CdbAct* ds, float* dsum, cdbStream_t stream);
int cdb_gate_bcast(const float* dsum, const float* att_sum, int32_t c_real, float inv_hw, const CdbAct* dt,
                   cdbStream_t stream);
/* nn.UpsamplingBilinear2d(scale_factor=2) == align_corners=True (networks5_ds.py:637,713). */
int cdb_bilinear2x_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream);
int cdb_bilinear2x_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream);
/* SURVEY 8(f) row f3 (models/encoder_decoder.py): out = alpha * x (scaled skip connections :198-205);
 * nn.Upsample(scale_factor=2, mode='nearest') (:193) and its backward (dx = 2x2 sums of dout);
 * nn.Tanh() inside a network (:112, the output blocks feed the next decoder level): dx = g * (1 - out^2). */
int cdb_scale(const CdbAct* x, float alpha, const CdbAct* out, cdbStream_t stream);
int cdb_nearest2x_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream);
int cdb_nearest2x_bwd(const CdbAct* dout, const CdbAct* dx, cdbStream_t stream);
int cdb_tanh_fwd(const CdbAct* x, const CdbAct* out, cdbStream_t stream);
int cdb_tanh_bwd(const CdbAct* out, const CdbAct* g, const CdbAct* dx, cdbStream_t stream);
/* nn.PReLU() with one learnable slope read from device memory (networks5_ds.py:498,551); *dslope += .. */
int cdb_prelu_fwd(const CdbAct* x, const float* slope, const CdbAct* out, cdbStream_t stream);
int cdb_prelu_bwd(const CdbAct* x, const CdbAct* g, const float* slope, const CdbAct* dx, float* dslope,
                  cdbStream_t stream);
/* nn.Dropout(p) (models/networks.py:305-306): out = x * keep / (1-p), keep = hash(seed, index) >= p.
 * Calling it again on the gradient with the same seed is the backward pass. */
int cdb_dropout(const CdbAct* x, const CdbAct* out, uint64_t seed, float p_drop, cdbStream_t stream);
/* Same with the effective seed = seed + *seed_dev read from device memory at run time: a training step replayed as a
 * CUDA graph then draws a fresh mask every step (the host-side seed of cdb_dropout is frozen at capture time). */
int cdb_dropout_dev(const CdbAct* x, const CdbAct* out, uint64_t seed, const uint64_t* seed_dev, float p_drop,
                    cdbStream_t stream);
/* Error-compensated TF32 ("3xTF32", the `tf32x3` network precision): x = hi + lo with hi = rna_tf32(x),
 * lo = rna_tf32(x - hi).  mode 0: out [n,h,w,3c] = channels [hi | lo | hi] (operand of a forward / data-gradient
 * convolution whose packed weights are [w_hi | w_hi | w_lo] along the contraction dimension); mode 1 / 2: out
 * [3n,h,w,c] = images [hi ; lo ; hi] / [hi ; hi ; lo] (the two operands of a weight gradient); mode 3: out [n,h,w,c] =
 * hi (the operand of the single-pass `tf32` precision, rounded to nearest instead of truncated).  The fp32 TMEM
 * accumulator then holds x_hi w_hi + x_lo w_hi + x_hi w_lo, i.e. the fp32 result of the reference's ATen
 * convolution (models/networks.py:162-185) to ~1e-6 instead of TF32's ~3e-4.  x, out: fp32 views. */
int cdb_split_tf32(const CdbAct* x, const CdbAct* out, int32_t mode, cdbStream_t stream);
/* NHWC bf16 view -> NCHW fp32 tensor (module outputs), dst strides in elements. */
int cdb_nhwc_to_nchw(const CdbAct* x, int32_t c_real, float* dst, int64_t d_n, int64_t d_c, int64_t d_h,
                     int64_t d_w, cdbStream_t stream);
/* Clears every pixel of the padded bf16 NHWC buffer `full` outside the interior rectangle
 * [top, top+inner_h) x [left, left+inner_w): the materialised zero padding (nn.Conv2d(padding=p)) of the consuming
 * convolution, written once by the producer instead of a memset of the whole buffer. */
int cdb_zero_frame(const CdbAct* full, int32_t top, int32_t left, int32_t inner_h, int32_t inner_w, cdbStream_t stream);

/* Second half of a few-output-channel convolution whose S filter columns were folded into the GEMM N dimension
 * (c7s1-3, models/networks.py:184-186; data gradient of c7s1-64, :158): t is the fp32 NHWC result
 * [n][p][wp][ct] of the R x 1 convolution with channel (s*cout + o);
 *   out[n,o,p,q] = act(bias[o] + sum_s t[n,p,q+s,s*cout+o])   (fp32, strides in elements). */
int cdb_shift_add_nchw(const float* t, int32_t n, int32_t p, int32_t q, int32_t wp, int32_t s_taps, int32_t cout,
                       int32_t ct, const float* bias, int32_t act, float slope, float* out, int64_t o_sn,
                       int64_t o_sc, int64_t o_sh, int64_t o_sw, cdbStream_t stream);

/* ---- K6b: losses of the seg/depth step ---------------------------------------------------------------
 * torch.nn.CrossEntropyLoss(ignore_index) on [n][c][hw] logits (new_multi/model5.py:281): acc2[0] += sum of
 * the per-pixel losses, acc2[1] += number of non-ignored pixels; grad = softmax - onehot (unscaled). */
int cdb_loss_ce2d(const float* logits, const int64_t* labels, int32_t n, int32_t c, int64_t hw,
                  int64_t ignore_index, float* acc2, float* grad, cdbStream_t stream);
/* BCEDepLoss (new_multi/networks5_ds.py:947-956) with o_m = (target == 1), z_m = (target == -1)
 * (get_masks :973-982): x [b][1][hw] broadcast against target [b][k][hw]; *loss_acc += loss,
 * grad_x = d loss / d x. */
int cdb_loss_bcedep(const float* x, const float* target, int32_t b, int32_t k, int64_t hw, float l1_weight,
                    float* loss_acc, float* grad_x, cdbStream_t stream);

/* torch.optim.Adam step (no amsgrad / weight decay) on one fp32 tensor (models/cycle_gan_model.py:66-69). */
int cdb_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                  float beta1, float beta2, float eps, int32_t step, cdbStream_t stream);

/* Same update with the step count read from device memory (CUDA-graph replays of the training step). */
int cdb_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t numel, float lr,
                      float beta1, float beta2, float eps, const int32_t* step_dev, cdbStream_t stream);

/* Multi-tensor form: every parameter tensor of an optimizer in one launch per 384 tensors.
 * entries_host is a HOST array (the pointers travel in the kernel parameter block, so the launch is graph-safe);
 * step >= 1 gives host-side bias corrections, step_dev != NULL reads the step count from device memory. */
typedef struct CdbAdamEntry {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
} CdbAdamEntry;
int cdb_adam_multi(const CdbAdamEntry* entries_host, int32_t n_entries, float lr, float beta1, float beta2, float eps,
                   int32_t step, const int32_t* step_dev, cdbStream_t stream);

/* SURVEY 8(f) row f1, second half: the same multi-tensor step also EMITS the packed bf16 GEMM operands the
 * convolution kernels read (cdb_pack_conv_weight layouts), so no re-pack pass follows the optimizer
 * (models/cycle_gan_model.py:66-69,149,160 are the optimizer call sites).  Per entry up to two packed copies of the
 * filter [d0][d1][r][s] are refreshed (pack[t] == NULL: none); lr_dev != NULL reads the learning rate from device
 * memory, which lets a learning-rate scheduler (models/networks.py:24-38) act on a replayed CUDA graph. */
typedef struct CdbAdamPackEntry {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t numel;
  void* pack[2];
  int32_t d0, d1, r, s;
  int32_t rows_are_dim0[2];
  int32_t rowpack[2];
} CdbAdamPackEntry;
int cdb_adam_pack_multi(const CdbAdamPackEntry* entries_host, int32_t n_entries, float lr, const float* lr_dev,
                        float beta1, float beta2, float eps, int32_t step, const int32_t* step_dev, cdbStream_t stream);

/* ImagePool.query (util/image_pool.py:12-32) with the host's random decisions supplied as a device table
 * plan_dev[batch][2] = {return_from, store_to} (-1: the incoming image / nothing stored); fake, out:
 * [batch][chw] fp32, pool: [pool_size][chw] fp32. Entries are applied in order (the reference's semantics). */
int cdb_image_pool_apply(const float* fake, float* pool, const int32_t* plan_dev, int32_t batch, int64_t chw,
                         float* out, cdbStream_t stream);

/* ---- K7: depth metrics ---------------------------------------------------------------------------
 * new_multi/my_eval.py:7-31 (compute_errors) applied as in eval_metric :52-100 to n_img pairs of
 * uint8 [h][w] images stored back to back. out8_per_img[i] = {abs_rel, sq_rel, rmse, rmse_log, a1,
 * a2, a3, masked pixel count} in float64 (count 0 => NaNs; the reference raises there). */
size_t cdb_depth_metrics_workspace(int32_t n_img);
int cdb_depth_metrics(const uint8_t* gt, const uint8_t* pred, int32_t n_img, int32_t h, int32_t w,
                      double* out8_per_img, void* workspace, size_t ws_bytes, cdbStream_t stream);

/* ---- device-side validation path -------------------------------------------------------
 * What new_multi/train5.py:97-110 does through PNG files, on the device:
 * cdb_depth_pred_to_u8:  u = uint8((x+1)/2*255) (util/util.py:64-65), then round(u / max_image(u) * 255)
 *   (train5.py:100,110 + cv2.imwrite of the float64 image) for n_img predictions pred[n][h][w] (fp32).
 * cdb_resize_linear_u8:  cv2.resize(pred, (dw, dh)) of my_eval.py:55 on uint8 images, bit-exact with OpenCV's
 *   INTER_LINEAR fixed-point arithmetic (2x exact decimation, where OpenCV switches to INTER_AREA, is refused).
 * workspace: cdb_validation_workspace() bytes, 256-byte aligned; its first n_img ints hold the per-image maxima. */
size_t cdb_validation_workspace(int32_t n_img, int32_t dh, int32_t dw);
int cdb_depth_pred_to_u8(const float* pred, int32_t n_img, int32_t h, int32_t w, uint8_t* out_u8, void* workspace,
                         size_t ws_bytes, cdbStream_t stream);
int cdb_resize_linear_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, uint8_t* dst, int32_t dh,
                         int32_t dw, void* workspace, size_t ws_bytes, cdbStream_t stream);

/* ---- device side of the loaders' per-sample arithmetic ------------------------------------
 * cdb_depth_labels: new_multi/try_data.py:240-272 — from n_img raw depth maps depth[n][hw] (fp32, after the
 *   loader's resize) the normalised depth dep_l[n][hw] (clamped at 8000) and the four range-limited depth labels
 *   depth_l_s[n][4][hw] ([5000,8000], [3000,6000], [1000,4000], (-inf,2000] — the last one keeps the reference's quirk
 *   of subtracting the minimum of the already normalised third range, :266); bit-identical to the numpy float32
 *   statements (IEEE single ops in the same order; degenerate ranges give the same NaN / inf).
 *   workspace: cdb_depth_labels_workspace(n_img) bytes, 4-byte aligned.
 * cdb_label_lut_i64: dst[i] = lut[src[i]] as int64 class ids — the label-id remapping loops of
 *   datasets/dataset_synthia.py:172-183 / new_multi/try_data.py:199-211 composed into one 256-entry table by the host
 *   (input_pipeline.label_lut_*), followed by MaskToTensor (:26-28).
 * cdb_image_normalize_u8: transforms.ToTensor() + transforms.Normalize(mean, std) (new_multi/try_data.py:425):
 *   dst[n][c][hw] = (src[n][hw][c] / 255 - mean) / std in IEEE single arithmetic. */
size_t cdb_depth_labels_workspace(int32_t n_img);
int cdb_depth_labels(const float* depth, int32_t n_img, int64_t hw, float* dep_l, float* depth_l_s, void* workspace,
                     size_t ws_bytes, cdbStream_t stream);
int cdb_label_lut_i64(const uint8_t* src, int64_t numel, const uint8_t* lut_dev, int64_t* dst, cdbStream_t stream);
int cdb_image_normalize_u8(const uint8_t* src, int32_t n_img, int64_t hw, int32_t channels, float mean, float stdv,
                           float* dst, cdbStream_t stream);

/* cdb_pil_resample_u8: Image.resize(size, Image.BILINEAR) of datasets/dataset_synthia.py:154-161 and
 *   new_multi/try_data.py:164-167 on n_img uint8 images src[n][sh][sw][channels] -> dst[n][dh][dw][channels], bit-exact
 *   with Pillow's ImagingResample 8-bit path (third-party dependency of the reference; Pillow 12.2 here): a horizontal
 *   then a vertical pass, each an int32 sum of byte x 22-bit fixed-point coefficient from 1 << 21, shifted and clipped to
 *   a byte.  The windows bounds_*[2*i] = first source index, bounds_*[2*i+1] = taps and the coefficients
 *   kk_*[i*ksize_* + j] (device pointers) are Pillow's precompute_coeffs + normalize_coeffs_8bpc, evaluated in double
 *   on the host (input_pipeline.pil_coeffs — any filter Pillow offers).  A pass whose size does not change is skipped, as
 *   Pillow does; its tables may be null.  workspace: cdb_pil_resample_workspace() bytes when both passes run.
 * cdb_gather_rows_cols_u8: dst[n][y][x][:] = src[n][ytab[y]][xtab[x]][:] (negative entry -> 0): Image.resize(size,
 *   Image.NEAREST) of the label images (:166-167; tables = ImagingScaleAffine's accumulated positions,
 *   input_pipeline.pil_nearest_table) and the F.hflip of paired_transform (:228-232; reversed identity table). */
size_t cdb_pil_resample_workspace(int32_t n_img, int32_t sh, int32_t dw, int32_t channels);
int cdb_pil_resample_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, int32_t channels, uint8_t* dst,
                        int32_t dh, int32_t dw, const int32_t* bounds_x, const int32_t* kk_x, int32_t ksize_x,
                        const int32_t* bounds_y, const int32_t* kk_y, int32_t ksize_y, void* workspace, size_t ws_bytes,
                        cdbStream_t stream);
int cdb_gather_rows_cols_u8(const uint8_t* src, int32_t n_img, int32_t sh, int32_t sw, int32_t channels, uint8_t* dst,
                            int32_t dh, int32_t dw, const int32_t* ytab, const int32_t* xtab, cdbStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CDB200_H_ */
