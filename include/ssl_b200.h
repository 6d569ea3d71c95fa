/*
 * ssl_b200.h -- C ABI of libssl_b200.so, the sm_100a implementation of the Self-Similarity
 * Graph (SSG) loss path of ChrisDud0257/SSL.
 *
 * Plain pointers and sizes only; every pointer is a DEVICE pointer unless its name ends in
 * _host.  `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Every
 * entry point returns 0 on success and a non-zero cudaError_t / SSL_B200_E* code otherwise;
 * ssl_b200_last_error() describes the last failure on the calling thread.  Nothing here
 * synchronises the device or allocates device memory unless stated.
 *
 * Reference interfaces each entry point replaces (paths under the reference checkout,
 * GAN = GAN-Based-SR/basicsr/losses):
 *   ssl_b200_compute_similarity            GAN/similarity/similarity.h:2-11   (_compute_similarity)
 *   ssl_b200_compute_similarity_backward   GAN/similarity/similarity.h:13-23  (_compute_similarity_backward)
 *   ssl_b200_build_edge_list               GAN/similarity/similaritywrapper.py:64-68 (pad + nonzero),
 *                                          GAN/../models/realesrganssl_model.py:64-72,385-388 (mask_stride, empty test)
 *   ssl_b200_ssg_rows_forward              GAN/loss_util.py:231-244 (ssl_cuda) == :182-229 (ssl_pytorch)
 *   ssl_b200_ssg_rows_backward             autograd of the above + similaritywrapper.py:38-57
 *   ssl_b200_row_loss                      GAN/basic_loss.py:14-16,41-66 (L1Loss), :269-282 (KLDistanceLoss)
 *   ssl_b200_loss_forward_backward,        GAN/../models/realesrganssl_model.py:378-430 (the SSL block of train_net_g;
 *   ssl_b200_loss_step_host                same block in 7 other *ssl_model.py, train_BSGRAN/models/model_ssl.py:285-334,
 *                                          Diffusion-Based-SR/ldm/models/diffusion/ddpmssl.py:438-513)
 *   ssl_b200_laplacian_mask                GAN-Based-SR/scripts/data_preparation/generate_mask.py:22-31
 *   ssl_b200_crop                          GAN-Based-SR/basicsr/data/transforms.py:93-149 (paired_random_crop_img_mask)
 *   ssl_b200_pool_exchange                 GAN-Based-SR/basicsr/models/realesrganssl_model.py:326-367 (_dequeue_and_enqueue)
 */
#ifndef SSL_B200_H_
#define SSL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSL_B200_ABI_VERSION 3

/* element types of image tensors */
#define SSL_B200_F32 0
#define SSL_B200_BF16 1
#define SSL_B200_F16 2

/* what a row of the SSG holds */
#define SSL_B200_ROWS_RAW 0  /* q: raw squared patch distance (similarity.cu output)          */
#define SSL_B200_ROWS_EXP 1  /* e = exp(-q/(C*kw^2)/sigma)      (generalization=False)        */
#define SSL_B200_ROWS_NORM 2 /* s = e / (sum e + eps)           (generalization=True)         */

/* which kernels the whole-step entry points use */
#define SSL_B200_PATH_AUTO 0   /* plane kernels when available and the mask is dense enough */
#define SSL_B200_PATH_POINT 1  /* one CTA per edge pixel (ssg_point.cuh) */
#define SSL_B200_PATH_PLANE 2  /* per-tile displacement planes (ssg_plane_*.cuh) */

/* error codes beyond cudaError_t */
#define SSL_B200_EINVAL 10001
#define SSL_B200_ENOTSUP 10002

int ssl_b200_abi_version(void);
const char* ssl_b200_last_error(void);

/* ---- drop-in for similarity.h ---------------------------------------------------------- */

/* Raw patch distance, reference calling convention: `image` is the REFLECT-PADDED fp32 image
 * [channel, height, width] (height/width are padded sizes), `pos` is int32 [mc,2] (row, col) in
 * padded coordinates, `out` is [mc, psize, psize].  The reference accumulates into a
 * zero-initialised `out`; this implementation overwrites it (same result under that contract). */
int ssl_b200_compute_similarity(const float* image, const int32_t* pos, float* out, int mc, int psize,
                                int ksize, int height, int width, int channel, void* stream);

/* Backward of the above: `grads` is dL/d out [mc,psize,psize]; contributions are ADDED into
 * `image_grads` [channel,height,width] (caller zeroes it, as similaritywrapper.py:47 does). */
int ssl_b200_compute_similarity_backward(const float* image, const float* grads, const int32_t* pos,
                                         float* image_grads, int mc, int psize, int ksize, int height,
                                         int width, int channel, void* stream);

/* ---- batched path ---------------------------------------------------------------------- */

/* Bytes of scratch ssl_b200_build_edge_list needs for n_pixels = B*H*W mask pixels. */
size_t ssl_b200_edge_list_workspace_bytes(int64_t n_pixels);

/* Edge list of a batch, with no host round trip.
 *   mask          float [B, mask_channels, H, W]; channel 0 is read; a pixel is an edge iff its
 *                 value == 1.0f exactly and (mask_stride <= 1 or y % mask_stride == x % mask_stride)
 *   edges         int32 [capacity]: flat index b*H*W + y*W + x of each edge pixel, ascending
 *                 (= images in order, row-major inside an image: the reference's row order)
 *   counts        int32 [2 + B]: counts[0] = number of edges written (min(total, capacity)),
 *                 counts[1] = total found, counts[2+b] = edges of image b
 */
int ssl_b200_build_edge_list(const float* mask, int B, int mask_channels, int H, int W, int mask_stride,
                             int32_t* edges, int capacity, int32_t* counts, void* workspace,
                             size_t workspace_bytes, void* stream);

/* SSG rows of every listed edge pixel.  image: [B,C,H,W] of `dtype` (unpadded; the reflect pad of
 * loss_util.py:189-191 is applied by index mapping).  n_edges_dev points at the device count
 * (counts[0] above); max_edges bounds the grid and the rows buffer.  rows: fp32 [max_edges, ks*ks].
 * image2/rows2 may name a second image of the same shape processed in the same launch (NULL = none). */
int ssl_b200_ssg_rows_forward(const void* image, const void* image2, int dtype, int B, int C, int H, int W,
                              const int32_t* edges, const int32_t* n_edges_dev, int max_edges, int ks, int kw,
                              float sigma, float eps, int rows_mode, float* rows, float* rows2, void* stream);

/* In place: rows (as written by the forward in `rows_mode`) and grad_rows = dL/drows  ->
 * grad_rows = dL/dq (raw distance).  Chain of loss_util.py:234-243. */
int ssl_b200_rows_grad_to_distance_grad(const float* rows, float* grad_rows, const int32_t* n_edges_dev,
                                        int max_edges, int ks, int kw, int C, float sigma, int rows_mode,
                                        void* stream);

/* dL/dimage accumulated (fp32, [B,C,H,W], caller zeroes) from gq = dL/dq rows, including the
 * adjoint of the reflect pad. */
int ssl_b200_ssg_rows_backward(const void* image, int dtype, int B, int C, int H, int W, const int32_t* edges,
                               const int32_t* n_edges_dev, int max_edges, int ks, int kw, const float* gq,
                               float* grad_image, void* stream);

/* Row loss between SR rows s and GT rows t (both ROWS_EXP or ROWS_NORM):
 *   sums[0] += sum |s - t|                          (L1Loss numerator)
 *   sums[1] += sum t' (log t' - log s'),  x' = max(x, 1e-10)   (KLDistanceLoss numerator)
 * and, when gq != NULL, gq = dL/dq_sr for  L = w_l1 * sums[0] + w_kl * sums[1]  (the 1/N of the
 * 'mean' reduction is applied by the caller once the global count is known).
 * sums: double [2], ACCUMULATED (caller zeroes); scratch: double [2 * ssl_b200_row_loss_blocks()]. */
int ssl_b200_row_loss_blocks(void);
int ssl_b200_row_loss(const float* rows_sr, const float* rows_gt, const int32_t* n_edges_dev, int max_edges,
                      int ks, int kw, int C, float sigma, int rows_mode, float w_l1, float w_kl, float* gq,
                      double* sums, double* scratch, void* stream);

/* ---- plane (tile-sharing) path -------------------------------------------------------------
 * Same results as the entry points above, computed per image tile instead of per edge pixel: the
 * squared-difference plane of every search offset is built once per tile and shared by all of its
 * edge pixels (ssl_b200/csrc/plane_geom.cuh).  Available for C == 3 and the kernel sizes
 * ssl_b200_plane_supported() accepts; the whole-step entry points pick it automatically. */
int ssl_b200_plane_supported(int ks, int kw, int channel);

/* Raw patch distances (ROWS_RAW, == ssl_b200_ssg_rows_forward(..., SSL_B200_ROWS_RAW, ...)) through
 * the plane kernels.  workspace: ssl_b200_plane_rows_workspace_bytes() bytes of scratch. */
size_t ssl_b200_plane_rows_workspace_bytes(int B, int H, int W, int ks, int kw, int max_edges);
int ssl_b200_plane_rows_forward(const void* image, const void* image2, int dtype, int B, int C, int H, int W,
                                const int32_t* edges, const int32_t* n_edges_dev, int max_edges, int ks, int kw,
                                float* rows, float* rows2, void* workspace, size_t workspace_bytes, void* stream);

/* In place: raw distances (ROWS_RAW) -> `rows_mode` rows; the tail of loss_util.py:234-243. */
int ssl_b200_rows_from_distance(float* rows, const int32_t* n_edges_dev, int max_edges, int ks, int kw, int C,
                                float sigma, float eps, int rows_mode, void* stream);

/* == ssl_b200_ssg_rows_backward through the plane kernels: gq = dL/dq rows [max_edges, ks*ks] in the order of
 * `edges`; grad_image fp32 [B,C,H,W] is OVERWRITTEN.  n_edges_dev must be given (counts[0]). */
size_t ssl_b200_plane_rows_backward_workspace_bytes(int B, int H, int W, int ks, int kw, int max_edges);
int ssl_b200_plane_rows_backward(const void* image, int dtype, int B, int C, int H, int W, const int32_t* edges,
                                 const int32_t* n_edges_dev, int max_edges, int ks, int kw, const float* gq,
                                 float* grad_image, void* workspace, size_t workspace_bytes, void* stream);

/* ---- whole step ------------------------------------------------------------------------- */

/* The reference training-step block (realesrganssl_model.py:378-430: per-image loop, two
 * similarity_map calls, cat, L1Loss [+ KLDistanceLoss], and its backward) for a batch whose edge
 * list is already built (ssl_b200_build_edge_list; counts[0] is read on the device).
 *   sr, gt       [B,C,H,W] of `dtype`
 *   grad_sr      fp32 [B,C,H,W] or NULL: OVERWRITTEN with d(w_l1*sum|d| + w_kl*sumKL)/d sr, i.e.
 *                the gradient before the 1/N of the 'mean' reduction (N = n_rows*ks*ks is only
 *                known once the counts of all ranks are: the caller scales)
 *   terms        double [3], OVERWRITTEN: sum|S_sr - S_gt|, sum KL, n_rows.  If the edge list overflowed its
 *                capacity (counts[1] > counts[0]) or holds more than max_edges pixels, n_rows is NaN, so
 *                the loss and the gradient scale derived from it are NaN (no silent truncation)
 *   workspace    ssl_b200_loss_workspace_bytes(...) bytes of scratch (rows never leave it)
 *   path         SSL_B200_PATH_*; the same value must be given to ssl_b200_loss_workspace_bytes */
size_t ssl_b200_loss_workspace_bytes(int B, int C, int H, int W, int ks, int kw, int max_edges, int path);
int ssl_b200_loss_forward_backward(const void* sr, const void* gt, int dtype, int B, int C, int H, int W,
                                   const int32_t* edges, const int32_t* counts, int max_edges, int ks, int kw,
                                   float sigma, float eps, int rows_mode, float w_l1, float w_kl, float* grad_sr,
                                   double* terms, void* workspace, size_t workspace_bytes, int path, void* stream);

/* The same step straight from the edge MASK (float [B, mask_channels, H, W]; channel 0 is read, a pixel is an
 * edge iff its value == 1.0f exactly and (mask_stride <= 1 or y % mask_stride == x % mask_stride): the rule of
 * ssl_b200_build_edge_list).  On the plane path no flat edge list is built at all -- the per-tile lists the
 * kernels use are made from the mask directly; on the point path the list is built inside `workspace`.
 *   dtype_sr, dtype_gt  element types of the two images; they may differ on the plane path (e.g. bf16 SR under
 *                autocast against an fp32 GT: neither is rounded, all arithmetic is fp32)
 *   max_edges    capacity: the most edge pixels the caller expects.  No host round trip happens; if the mask
 *                holds more, terms[2] (and with it the loss and the gradient scale) is NaN
 *   terms        double [3], OVERWRITTEN: sum|S_sr - S_gt|, sum KL, n_rows (the number of edge pixels found)
 *   workspace    ssl_b200_loss_step_workspace_bytes(...) bytes
 * Everything else as in ssl_b200_loss_forward_backward. */
size_t ssl_b200_loss_step_workspace_bytes(int B, int C, int H, int W, int ks, int kw, int max_edges, int path);
int ssl_b200_loss_step(const void* sr, int dtype_sr, const void* gt, int dtype_gt, const float* mask,
                       int mask_channels, int mask_stride, int B, int C, int H, int W, int max_edges, int ks, int kw,
                       float sigma, float eps, int rows_mode, float w_l1, float w_kl, float* grad_sr, double* terms,
                       void* workspace, size_t workspace_bytes, int path, void* stream);

/* The 'mean' of the step: terms (double [3]: sum|d|, sum KL, n_rows -- local, or all-reduced over the ranks) ->
 * out4 (float [4]) = { w_l1*L1 + w_kl*KL, w_l1*L1, w_kl*KL, grad_scale / N } with N = max(n_rows * row_len, 1); the last
 * entry is the factor the gradient returned by the step is multiplied with.  A NaN row count (overflowed edge
 * list) makes all four NaN. */
int ssl_b200_loss_from_terms(const double* terms, int row_len, float w_l1, float w_kl, float grad_scale, float* out4,
                             void* stream);

/* Test / inspection hook: dL/dq (the gradient with respect to the raw patch distances, before the 1/N of the
 * 'mean') that the last ssl_b200_loss_forward_backward call with a non-NULL grad_sr left in `workspace`,
 * copied out as fp32 rows [n, ks*ks] in the order of `edges`.  Same shape arguments as that call. */
int ssl_b200_loss_export_distance_grad(const void* workspace, size_t workspace_bytes, int B, int C, int H, int W,
                                       const int32_t* edges, const int32_t* counts, int max_edges, int ks, int kw,
                                       int path, float* gq_rows, void* stream);

/* Same step on HOST buffers (the end-to-end call of a host-side plugin): copies sr/gt/mask to the
 * device, builds the edge list, runs the step, applies the 'mean' (single device: N is local) and
 * copies the results back.  fp32 images.  loss_host: float [3] = total, w_l1*L1, w_kl*KL;
 * grad_host: fp32 [B,C,H,W] = d total / d sr, or NULL.  n_rows_host: int64 [1] or NULL.
 * ALLOCATES (and caches per device) its device arena; SYNCHRONISES `stream` before returning.
 * Pinned host buffers make the copies asynchronous with respect to each other. */
int ssl_b200_loss_step_host(const float* sr_host, const float* gt_host, const float* mask_host, int mask_channels,
                            int B, int C, int H, int W, int mask_stride, int ks, int kw, float sigma, float eps,
                            int rows_mode, float w_l1, float w_kl, float* loss_host, float* grad_host,
                            int64_t* n_rows_host, void* stream);
/* Frees the cached arena of the current device. */
int ssl_b200_release_host_arena(void);

/* ---- the step upstream of the loss: crop + training-pair pool that carry the mask ------- */

/* dst[p][y][x] = src[p][top + y][left + x] for `planes` planes of 4-byte elements (any fp32 / int32 tensor viewed
 * as [planes, H, W]): the tensor branch of paired_random_crop_img_mask (GAN-Based-SR/basicsr/data/transforms.py:
 * 93-149), which crops LQ, GT and the edge mask with one (top, left) draw. */
int ssl_b200_crop(const void* src, void* dst, int planes, int H, int W, int top, int left, int h, int w, void* stream);

/* One tensor of the training-pair pool (GAN-Based-SR/basicsr/models/realesrganssl_model.py:326-367): for i < b
 *     out[i] = queue[slots[i]]   (out == NULL: enqueue only),   queue[slots[i]] = in[i]
 * on samples of `sample_elems` 4-byte elements.  bcast_channels > 1: `in` holds ONE channel per sample that is
 * written to all bcast_channels channels of the queue sample -- the reference allocates the mask queue with the
 * GT's channel count and assigns the 1-channel mask into it (:339-341,357).  slots: int32 [b] on the device. */
int ssl_b200_pool_exchange(void* queue, const void* in, void* out, const int32_t* slots, int b, int64_t sample_elems,
                           int bcast_channels, void* stream);

/* Optional per-stage timing of the entry points above.  While enabled, every stage (edge list,
 * plane lists, forward, row loss, backward, ...) is bracketed by CUDA events recorded on the stream
 * it is launched on; ssl_b200_profile_read() waits for them and returns, per stage, the summed
 * milliseconds and the number of bracketed launches since the last read.  Arrays hold
 * ssl_b200_profile_num_stages() entries.  Not thread safe; meant for benchmarks. */
int ssl_b200_profile_enable(int on);
int ssl_b200_profile_num_stages(void);
const char* ssl_b200_profile_stage_name(int i);
int ssl_b200_profile_read(float* ms, int* launches);

/* Number of kernels this library has launched in this process so far (all threads). */
uint64_t ssl_b200_launch_count(void);

/* Edge mask of generate_mask.py on the GT crop: luma of round(255*clamp(gt,0,1)), 4-neighbour
 * Laplacian with BORDER_REFLECT_101 saturated to [0,255], mask = lap > threshold.
 * gt: [B,3,H,W] of `dtype`; mask: float [B,1,H,W]. */
int ssl_b200_laplacian_mask(const void* gt, int dtype, int B, int H, int W, float threshold, float* mask,
                            void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSL_B200_H_ */
