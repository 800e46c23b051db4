/* adpst.h -- C-ABI of libadpst.so: the B200 (sm_100a) hot path of automated-deep-photo-style-transfer.
 *
 * The reference has NO native / FFI layer: its boundary is the Python classes
 *   components/loss.py:6-165            Loss
 *   components/matting_v2.py:6-251      MattingLaplacian (matrix-free)
 *   components/matting_v3.py:13-102     MattingLaplacian (explicit COO)
 *   components/VGG19/model.py:4-41      StyleContentModel
 *   style_transfer.py:321-344           Adam + train_step
 * Each entry point below names the reference lines it replaces.  The host-side mirror of those classes
 * (automated-deep-photo-style-transfer_b200/components/*.py) binds this file through ctypes.
 *
 * Conventions
 *   - every pointer marked "dev" is a CUDA device pointer owned by the caller (torch allocates);
 *   - images / activations are NHWC with N == 1, float32 unless a dtype argument says otherwise;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*); no internal threads,
 *     no global state except the per-thread last-error string;
 *   - return value: ADPST_OK or an error code; adpst_last_error() gives the message.
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns ADPST_ERR_CUDA.
 */
#ifndef ADPST_H
#define ADPST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* adpst_stream_t;

enum { ADPST_OK = 0, ADPST_ERR_INVALID = 1, ADPST_ERR_CUDA = 2, ADPST_ERR_UNSUPPORTED = 3 };
enum { ADPST_F32 = 0, ADPST_F64 = 1 };
enum { ADPST_LAP_V2 = 2, ADPST_LAP_V3 = 3 };
/* mat-vec kernel of a Laplacian handle: AUTO = DIA when the handle has radius 1 and float32 storage, else MATRIX_FREE */
enum { ADPST_LAP_KERNEL_AUTO = 0, ADPST_LAP_KERNEL_DIA = 1, ADPST_LAP_KERNEL_MATRIX_FREE = 2, ADPST_LAP_KERNEL_TILE = 3 };
enum { ADPST_VGG_NUM_CONV = 13, ADPST_VGG_NUM_POOL = 4 };

int adpst_version(void);
const char* adpst_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (bench.py reports it as gpu_launches). */
unsigned long long adpst_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Matting Laplacian (matrix-free for both variants)
 * ---------------------------------------------------------------------------------------------- */
typedef struct adpst_laplacian adpst_laplacian;

/* matting_v2.py:11-52 (mode V2: symmetric padding, one window per pixel) and
 * matting_v3.py:27-39,61-102 (mode V3: fully interior windows only).
 * image_dev: (H,W,3) of `io_dtype`; it is copied into the handle.  compute_dtype: arithmetic type of the
 * stencil (F64 reproduces the reference's float64 path, loss.py:160; F32 is the fast path). */
int adpst_laplacian_create(int mode, int H, int W, int radius, double epsilon, const void* image_dev,
                           int io_dtype, int compute_dtype, adpst_stream_t stream, adpst_laplacian** out);
void adpst_laplacian_destroy(adpst_laplacian* h);

/* matting_v2.py:147-176 _matmul / matting_v3.py:50-51 _matmul, fused with loss.py:160-161.
 * x_dev: (H*W,3) io_dtype.  y_dev (may be NULL): receives y_scale * (L x), same dtype.
 * xLx_dev (may be NULL): device double, receives x^T L x accumulated in float64. */
int adpst_laplacian_matvec(adpst_laplacian* h, const void* x_dev, void* y_dev, double y_scale,
                           double* xLx_dev, adpst_stream_t stream);

/* Which kernel evaluates L x (all of them are the same operator):
 *   DIA          (radius 1, float32 storage; the default there): the operator's 5x5 stencil coefficients -- the explicit matrix
 *                of matting_v3.py:61-102 in diagonal format, 12 float32 planes using symmetry and zero row sums -- and L I are
 *                precomputed ONCE in float64 when the handle is created; each call evaluates y = L I + L (x - I) in float32
 *                (2e-7 of max|y| measured; exact at x = I), HBM-bound at 96 B/px;
 *   MATRIX_FREE  window statistics recomputed from I in every call, arithmetic in compute_dtype (float64: 1e-9);
 *   TILE         the shared-memory tile variant of MATRIX_FREE (any radius; validation).
 * adpst_laplacian_kernel returns the kind that will run. */
int adpst_laplacian_set_kernel(adpst_laplacian* h, int kind, adpst_stream_t stream);
int adpst_laplacian_kernel(const adpst_laplacian* h);

/* Spatially tiled runs (one image split into column strips with halos): restrict the scalar x^T L x to columns
 * [col_lo, col_hi) of the local strip; y is still produced for every column.  (0,0) restores the full sum. */
int adpst_laplacian_set_quadratic_window(adpst_laplacian* h, int col_lo, int col_hi);

/* matting_v2.py:49-52: window means (H,W,3) and regularised inverse covariances (H,W,3,3), io_dtype. */
int adpst_laplacian_coefficients(adpst_laplacian* h, void* means_dev, void* delta_inv_dev, adpst_stream_t stream);

/* matting_v3.py:97-102: COO triplets in the reference's emission order, duplicates kept.
 * nnz = 81 (H-2r)(W-2r) for r = 1 (general: (2r+1)^4 per window). vals_dev has io_dtype. */
int64_t adpst_laplacian_nnz(const adpst_laplacian* h);
int adpst_laplacian_export_coo(adpst_laplacian* h, int64_t* rows_dev, int64_t* cols_dev, void* vals_dev,
                               adpst_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Optimiser: style_transfer.py:321-326,342-343  (TF/Keras Adam + clip_by_value(0,1)), fused.
 * state_dev: device int32[2] = {completed steps t, 0}; the kernel uses t+1 and the LAST block increments t, so the
 * call is CUDA-graph replayable.  In place on x, m, v.
 * ---------------------------------------------------------------------------------------------- */
int adpst_adam_clip_step(float* x_dev, const float* grad_dev, float* m_dev, float* v_dev, size_t n,
                         int32_t* state_dev, float lr, float beta1, float beta2, float epsilon,
                         adpst_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * VGG19 extractor: components/VGG19/model.py:4-41 (forward) and the tape.gradient of
 * style_transfer.py:341 (data gradient only; weights are frozen, model.py:11,25).
 * ---------------------------------------------------------------------------------------------- */
typedef struct adpst_vgg adpst_vgg;

/* kernels_dev[i]: HWIO (3,3,Cin,Cout) float32, biases_dev[i]: (Cout,), i over the 13 convolutions
 * block1_conv1 .. block5_conv1 in network order.  The handle keeps re-laid-out copies. */
int adpst_vgg_create(const float* const* kernels_dev, const float* const* biases_dev,
                     adpst_stream_t stream, adpst_vgg** out);
void adpst_vgg_destroy(adpst_vgg* h);
/* shape of conv output i (post-ReLU) for an H x W input. */
int adpst_vgg_conv_shape(int i, int H, int W, int* h, int* w, int* c);
/* shape of pool output j (j = 0..3). */
int adpst_vgg_pool_shape(int j, int H, int W, int* h, int* w, int* c);

/* model.py:27-30: x255, RGB->BGR, mean subtraction, then convs/pools up to conv index `last` (inclusive).
 * acts_dev[i]: conv i output (1,h,w,c) post-ReLU; pools_dev[j]: pool j output.  All caller-allocated. */
int adpst_vgg_forward(adpst_vgg* h, const float* image_dev, int H, int W, float* const* acts_dev,
                      float* const* pools_dev, int last, adpst_stream_t stream);

/* The same for convolutions first..last only (spatially tiled runs exchange halo columns between blocks): the input of conv
 * `first` is image_dev (first == 0), pools_dev[j] if a pool precedes it, else acts_dev[first-1].  pools_dev[j] may be NULL for
 * pools that are not wanted.
 * level_widths / pool_col_offset (host arrays of 5 and 4 ints, or both NULL): column geometry of a strip whose halo is not the
 * same multiple of 2^-l on every resolution level l.  Tensors of level l are level_widths[l] columns wide (level_widths[0] ==
 * W); the 2x2 max-pool of a level-l tensor has level_widths[l] / 2 columns and is stored from column pool_col_offset[l] of the
 * level-(l+1) tensor on (the columns around it are the caller's: they hold what the neighbouring strip sent). */
int adpst_vgg_forward_range(adpst_vgg* h, const float* image_dev, int H, int W, float* const* acts_dev,
                            float* const* pools_dev, int first, int last, const int* level_widths, const int* pool_col_offset,
                            adpst_stream_t stream);

/* Convolution kernel family used by this handle: 0 = tcgen05 3xFP16 implicit GEMM, float32-accurate (default;
 * block1_conv1 and the gradient to the image always use the CUDA-core kernels), 1 = exact-float32 CUDA-core kernels
 * everywhere (validation). */
int adpst_vgg_set_conv_path(adpst_vgg* h, int path);

/* The tensor-core kernels split float32 operands into FP16 pairs after a power-of-two scale derived from the
 * tensor's largest magnitude.  Inside adpst_vgg_forward / adpst_vgg_backward every kernel records max|output| for its
 * consumer; tensors that enter from outside either come with a device slot holding the float32 bit pattern of
 * max|x| (adpst_absmax) or the entry point measures it with one extra pass (slot argument NULL). */
int adpst_absmax(const float* x_dev, size_t n, uint32_t* slot_dev, adpst_stream_t stream);
/* *slot_dev = max(*slot_dev, max|x|): for tensors that were patched after their producer recorded the slot (halo columns
 * received from a neighbouring rank). */
int adpst_absmax_update(const float* x_dev, size_t n, uint32_t* slot_dev, adpst_stream_t stream);
/* slot of conv i's output as left by the most recent adpst_vgg_forward on this handle (device pointer). */
const uint32_t* adpst_vgg_act_absmax(const adpst_vgg* h, int i);

/* One layer in isolation (parity tests, per-kernel roofline in bench.py).
 * conv_forward: y = relu(conv_i(x) + b_i); for i == 0, x is the [0,1] RGB image (h,w,3).
 * conv_dgrad  : dx = conv_i^T(dpre) (i >= 1), no mask, no seed.
 * x_absmax_dev / dpre_absmax_dev: see above (may be NULL). */
int adpst_vgg_conv_forward(adpst_vgg* h, int i, const float* x_dev, int lh, int lw, float* y_dev,
                           const uint32_t* x_absmax_dev, adpst_stream_t stream);
int adpst_vgg_conv_dgrad(adpst_vgg* h, int i, const float* dpre_dev, int lh, int lw, float* dx_dev,
                         const uint32_t* dpre_absmax_dev, adpst_stream_t stream);

/* Backward to the image.  seeds_dev[i] (may be NULL): dLoss/d(conv i output), added where the chain passes.
 * scratch_dev: two buffers, each at least as large as the largest activation (conv 0).
 * dimage_dev: (1,H,W,3) gradient w.r.t. the [0,1] RGB image (the x255 and channel flip are folded in). */
int adpst_vgg_backward(adpst_vgg* h, int H, int W, const float* const* acts_dev, const float* const* pools_dev,
                       const float* const* seeds_dev, int last, float* scratch0_dev, float* scratch1_dev,
                       float* dimage_dev, adpst_stream_t stream);

/* Backward through convs last..first only (spatially tiled runs exchange halo columns between the calls).
 * Entry  grad_in_dev == NULL: the chain starts at conv `last` with its seed;
 *        conv `last` is followed by a pool: grad_in_dev = dLoss/d(pooled output of conv last);
 *        otherwise: grad_in_dev = dLoss/d(pre-activation of conv last) (already ReLU-masked), with its max|.| in the slot
 *        adpst_vgg_grad_absmax(h, last) (left there by the call that produced it; adpst_absmax_update after patching it).
 * Exit   first == 0: out_dev = dLoss/d(image);  conv `first` follows a pool: out_dev = dLoss/d(that pooled tensor);
 *        otherwise: out_dev = dLoss/d(pre-activation of conv first-1) (seed and ReLU mask of conv first-1 applied).
 * level_widths / pool_col_offset: as for adpst_vgg_forward_range (a gradient w.r.t. a pooled tensor has that tensor's layout). */
int adpst_vgg_backward_range(adpst_vgg* h, int H, int W, const float* const* acts_dev, const float* const* seeds_dev, int first,
                             int last, const float* grad_in_dev, float* scratch0_dev, float* scratch1_dev, float* out_dev,
                             const int* level_widths, const int* pool_col_offset, adpst_stream_t stream);
/* slot holding max|dLoss/d(pre-activation of conv i)| of the latest backward pass (device pointer). */
const uint32_t* adpst_vgg_grad_absmax(const adpst_vgg* h, int i);

/* ------------------------------------------------------------------------------------------------
 * Halo exchange of column strips over peer memory (spatially tiled runs; the reference is single-device, SURVEY.md 8e:
 * this is the exchange the strip decomposition of style_transfer.py:331-344 needs between two network segments).
 * Every rank owns a mailbox with room for `side_bytes` of slabs from each of its two neighbours; neighbours store into it
 * over NVLink (CUDA IPC mapping between processes, direct pointers between ranks of one process).
 * ---------------------------------------------------------------------------------------------- */
typedef struct adpst_halo adpst_halo;
int adpst_halo_create(size_t side_bytes, adpst_halo** out);        /* on the current device; side_bytes % 16 == 0 */
void adpst_halo_destroy(adpst_halo* h);
int adpst_halo_ipc_handle_bytes(void);                              /* size of the opaque handle adpst_halo_export writes */
int adpst_halo_export(const adpst_halo* h, void* handle_out);
/* side 0 = left neighbour, 1 = right neighbour */
int adpst_halo_connect_ipc(adpst_halo* h, int side, const void* handle, size_t peer_side_bytes);
int adpst_halo_connect_local(adpst_halo* h, int side, adpst_halo* neighbour);
/* x_dev: (rows, width, C) float32; own columns [own_lo, own_hi).  push: the hl own columns next to each interior boundary go
 * into the neighbours' mailboxes at `offset` and are published under `slot`.  pull: waits for both neighbours' slabs of `slot`,
 * writes them into the hl halo columns on either side and raises *absmax_slot_dev (float bits, may be NULL) to max|slab|.
 * Every rank must issue the same sequence of (slot, offset) pairs, at least two exchanges per step. */
int adpst_halo_push(adpst_halo* h, int slot, size_t offset, const float* x_dev, int rows, int width, int C, int hl,
                    int own_lo, int own_hi, adpst_stream_t stream);
int adpst_halo_pull(adpst_halo* h, int slot, size_t offset, float* x_dev, int rows, int width, int C, int hl, int own_lo,
                    int own_hi, uint32_t* absmax_slot_dev, adpst_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Loss terms: components/loss.py
 * ---------------------------------------------------------------------------------------------- */
/* tf.image.resize bilinear, half-pixel centres, no antialias (loss.py:112-113); single channel. */
int adpst_resize_bilinear(const float* src_dev, int Hs, int Ws, float* dst_dev, int Hd, int Wd,
                          adpst_stream_t stream);
/* the same for n planes at once: src (n,Hs,Ws) -> dst (n,Hd,Wd) (all K class masks of a layer in one launch) */
int adpst_resize_bilinear_batch(const float* src_dev, int n, int Hs, int Ws, float* dst_dev, int Hd, int Wd,
                                adpst_stream_t stream);

/* loss.py:96-102 for K masks at once: G[k] = (F*m_k)^T (F*m_k).  F: (h,w,C) feature map; masks: (K,h*w) or NULL
 * (K==1, all ones); G: (K,C,C).  workspace_dev: at least adpst_gram_workspace_bytes(h*w,C,K) bytes.
 * path 0 (tcgen05 3xFP16): needs the list of 2x16-pixel patches on which each class mask is non-zero --
 *   patch_ids_dev: int32 patch indices (row-major over ceil(h/2) x ceil(w/16) patches) grouped by class,
 *   patch_off_dev: int32[K+1] offsets into it.  The masks are constant, so the caller builds the list once.
 * path 1, or NULL lists: exact-float32 CUDA-core kernel.
 * F_absmax_dev / masks_absmax_dev: slots holding max|F| (adpst_vgg_act_absmax / adpst_absmax) and max|masks|
 * (adpst_absmax, once: the masks are constant), or NULL to have them measured here. */
size_t adpst_gram_workspace_bytes(int HW, int C, int K);
int adpst_gram_masked(const float* F_dev, int h, int w, int C, const float* masks_dev, int K, const int* patch_ids_dev,
                      const int* patch_off_dev, float* G_dev, int path, const uint32_t* F_absmax_dev,
                      const uint32_t* masks_absmax_dev, void* workspace_dev, adpst_stream_t stream);

/* loss.py:104-137 for one layer, forward value and gradient seed.  With
 *   L = sum_k mean((A_k - G_k)^2) / (2 C^2 HW^2):
 *   *loss_dev += loss_scale * L                      (float64 accumulator, may be NULL)
 *   dF (=|+=)  grad_scale * dL/dF                     (written, or added if accumulate != 0; may be NULL)
 * G_dev is the transfer Gram from adpst_gram_masked on the same F / masks; A_dev the style Gram.
 * loss_scale carries the 1/len(args) of loss.py:85, grad_scale additionally the style weight.
 * F is the (h,w,C) feature map; path: 0 = tcgen05 3xFP16 kernel (8x16-pixel tiles, classes absent from a tile skipped),
 * 1 = exact-float32 CUDA-core kernel (validation).  hw_norm > 0 replaces h*w in the normaliser (spatially tiled runs:
 * the pixel count of the whole image).  F_absmax_dev: slot holding max|F| (adpst_vgg_act_absmax / adpst_absmax), or
 * NULL to have it measured here.  tiles_dev: per-tile class sets of these (constant) masks from adpst_style_tiles, or
 * NULL to have them rebuilt on every call. */
int adpst_style_layer_backward(const float* F_dev, int h, int w, int C, const float* masks_dev, int K,
                               const float* G_dev, const float* A_dev, double loss_scale, double grad_scale,
                               double* loss_dev, float* dF_dev, int accumulate, int path, double hw_norm,
                               const uint32_t* F_absmax_dev, const void* tiles_dev, void* workspace_dev,
                               adpst_stream_t stream);

/* Set-up for the tensor-core style gradient: which classes are present in each 8x16-pixel tile of a (K,h*w) mask
 * stack (masks_dev NULL: one all-ones class).  tiles_dev: adpst_style_tiles_bytes(h*w) bytes, filled once per layer. */
size_t adpst_style_tiles_bytes(int HW);
int adpst_style_tiles(const float* masks_dev, int K, int h, int w, void* tiles_dev, adpst_stream_t stream);

/* loss.py:90-92 with L = mean((target - output)^2):
 *   *loss_dev += loss_scale * L;   dOut (=|+=) grad_scale * 2 (output - target) / n.
 * Spatially tiled runs: n_norm > 0 replaces n (element count of the whole layer) and, with w > 0, the scalar sums only
 * columns [col_lo, col_hi) of the (.., w, C) map; pass 0, 0, 0, 0, 0 otherwise. */
int adpst_content_layer(const float* target_dev, const float* output_dev, size_t n, double loss_scale,
                        double grad_scale, double* loss_dev, float* dOut_dev, int accumulate, double n_norm, int w, int C,
                        int col_lo, int col_hi, adpst_stream_t stream);

/* EXTENSION -- no counterpart in the reference (it has no total-variation term; BASELINE.json's north star and
 * configs[1] name one).  Semantics of tf.image.total_variation on the (H,W,3) image:
 *   TV = sum |x[y+1,x,c]-x[y,x,c]| + sum |x[y,x+1,c]-x[y,x,c]|;   *loss_dev += loss_scale * TV  (float64, may be NULL);
 *   dX (=|+=) grad_scale * dTV/dx  (sign differences, sgn(0) = 0; may be NULL).
 * Spatially tiled runs: the scalar counts only columns [col_lo, col_hi) of the local strip; (0,0) = all. */
int adpst_tv_loss(const float* x_dev, int H, int W, double loss_scale, double grad_scale, double* loss_dev, float* dX_dev,
                  int accumulate, int col_lo, int col_hi, adpst_stream_t stream);

/* loss.py:72-76: acc_dev = float64 {content, style, photo, tv} (unweighted);  out_dev = float32[6]
 * {content, style, nima (0), photo, total = sum w_i * loss_i, tv}; the tv term (extension) enters the total only if
 * w_tv > 0.  One tiny launch, keeps the step graph-replayable. */
int adpst_loss_finalize(const double* acc_dev, double w_content, double w_style, double w_photo, double w_tv,
                        float* out_dev, adpst_stream_t stream);

/* out[i] = alpha * a[i] + beta * b[i] (float32; b may be NULL).  Used to combine image gradients. */
int adpst_axpby(float* out_dev, const float* a_dev, float alpha, const float* b_dev, float beta, size_t n,
                adpst_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ADPST_H */
