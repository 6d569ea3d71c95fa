// libssl_b200.so -- C ABI over the sm_100a kernels (declared in include/ssl_b200.h).
// One translation unit; build with csrc/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
#include "ssl_b200.h"

#include "common.cuh"
#include "edge_list.cuh"
#include "row_ops.cuh"
#include "ssg_point.cuh"

using namespace sslb;

extern "C" int ssl_b200_abi_version(void) { return SSL_B200_ABI_VERSION; }
extern "C" const char* ssl_b200_last_error(void) { return err_buf(); }

namespace {

int check_sizes(int ks, int kw, int H, int W, int C) {
    SSLB_REQUIRE(ks >= 1 && (ks & 1), "kernel_size_search must be odd and positive, got %d", ks);
    SSLB_REQUIRE(kw >= 1 && (kw & 1), "kernel_size_window must be odd and positive, got %d", kw);
    SSLB_REQUIRE(kw <= ks, "kernel_size_window (%d) must not exceed kernel_size_search (%d)", kw, ks);
    SSLB_REQUIRE(C >= 1 && H >= 1 && W >= 1, "bad image shape C=%d H=%d W=%d", C, H, W);
    SSLB_REQUIRE(ks / 2 < H && ks / 2 < W, "reflect pad %d needs an image larger than %dx%d", ks / 2, H, W);
    return 0;
}

size_t point_smem_bytes(int C, int ks, int kw, int planes_extra, int pitch) {
    const int TP = ks + 2 * (kw / 2);
    return (size_t)((C + planes_extra) * TP * pitch + ks * ks) * sizeof(float);
}

template <typename K>
int set_smem(K kernel, size_t bytes, const DeviceInfo& di) {
    SSLB_REQUIRE(bytes <= (size_t)di.max_smem_optin, "search tile needs %zu B of shared memory (> %d)", bytes,
                 di.max_smem_optin);
    if (bytes > 48 * 1024) SSLB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return 0;
}

int launch_point_forward(PointParams& p, int dtype, int n_images, cudaStream_t st) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    if (p.max_edges <= 0) return 0;
    const bool tiled = p.ks == 25 && p.kw == 9;
    if (tiled) {
        constexpr int PITCH = 57;
        const size_t smem = point_smem_bytes(p.C, 25, 9, 0, PITCH);
        const int blocks = min(p.max_edges, di.sm_count * 8);
        SSLB_DISPATCH_DTYPE(dtype, T, {
            auto k = ssg_point_fwd_tiled<T, 25, 9, 5, PITCH>;
            if (int e = set_smem(k, smem, di)) return e;
            k<<<dim3(blocks, n_images), 128, smem, st>>>(p);
        });
    } else {
        const int TP = p.ks + 2 * (p.kw / 2);
        const size_t smem = point_smem_bytes(p.C, p.ks, p.kw, 0, TP | 1);
        const int blocks = min(p.max_edges, di.sm_count * 4);
        SSLB_DISPATCH_DTYPE(dtype, T, {
            auto k = ssg_point_fwd_generic<T>;
            if (int e = set_smem(k, smem, di)) return e;
            k<<<dim3(blocks, n_images), 256, smem, st>>>(p);
        });
    }
    return check_launch("ssg_point_fwd");
}

int launch_point_backward(PointParams& p, int dtype, cudaStream_t st) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    if (p.max_edges <= 0) return 0;
    const int TP = p.ks + 2 * (p.kw / 2);
    const size_t smem = (size_t)(p.C + 1) * TP * (TP | 1) * sizeof(float);
    const int blocks = min(p.max_edges, di.sm_count * 4);
    SSLB_DISPATCH_DTYPE(dtype, T, {
        auto k = ssg_point_bwd_generic<T>;
        if (int e = set_smem(k, smem, di)) return e;
        k<<<blocks, 256, smem, st>>>(p);
    });
    return check_launch("ssg_point_bwd");
}

}  // namespace

// ---- drop-in for similarity.h -------------------------------------------------------------

extern "C" int ssl_b200_compute_similarity(const float* image, const int32_t* pos, float* out, int mc, int psize,
                                           int ksize, int height, int width, int channel, void* stream) {
    SSLB_REQUIRE(image && out && (pos || mc == 0), "null pointer");
    SSLB_REQUIRE(mc >= 0, "negative mc");
    if (int e = check_sizes(psize, ksize, height, width, channel)) return e;
    PointParams p{};
    p.img[0] = image;
    p.rows[0] = out;
    p.edges = EdgeRef{nullptr, pos};
    p.n_edges_dev = nullptr;
    p.max_edges = mc;
    p.B = 1; p.C = channel; p.H = height; p.W = width;
    p.ks = psize; p.kw = ksize;
    p.sigma = 1.f; p.eps = 0.f;
    p.mode = SSL_B200_ROWS_RAW;
    return launch_point_forward(p, SSL_B200_F32, 1, (cudaStream_t)stream);
}

extern "C" int ssl_b200_compute_similarity_backward(const float* image, const float* grads, const int32_t* pos,
                                                    float* image_grads, int mc, int psize, int ksize, int height,
                                                    int width, int channel, void* stream) {
    SSLB_REQUIRE(image && image_grads && ((grads && pos) || mc == 0), "null pointer");
    SSLB_REQUIRE(mc >= 0, "negative mc");
    if (int e = check_sizes(psize, ksize, height, width, channel)) return e;
    PointParams p{};
    p.img[0] = image;
    p.edges = EdgeRef{nullptr, pos};
    p.max_edges = mc;
    p.B = 1; p.C = channel; p.H = height; p.W = width;
    p.ks = psize; p.kw = ksize;
    p.gq = grads;
    p.grad = image_grads;
    return launch_point_backward(p, SSL_B200_F32, (cudaStream_t)stream);
}

// ---- batched path -------------------------------------------------------------------------

extern "C" size_t ssl_b200_edge_list_workspace_bytes(int64_t n_pixels) {
    const int64_t chunks = (n_pixels + kElChunk - 1) / kElChunk;
    return (size_t)(2 * chunks) * sizeof(int32_t) + 256;
}

extern "C" int ssl_b200_build_edge_list(const float* mask, int B, int mask_channels, int H, int W, int mask_stride,
                                        int32_t* edges, int capacity, int32_t* counts, void* workspace,
                                        size_t workspace_bytes, void* stream) {
    SSLB_REQUIRE(mask && edges && counts && workspace, "null pointer");
    SSLB_REQUIRE(B >= 1 && mask_channels >= 1 && H >= 1 && W >= 1 && capacity >= 0, "bad shape");
    const long long n_pixels = (long long)B * H * W;
    SSLB_REQUIRE(n_pixels < (1ll << 31), "batch too large for int32 flat indices (%lld pixels)", n_pixels);
    SSLB_REQUIRE(workspace_bytes >= ssl_b200_edge_list_workspace_bytes(n_pixels), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    EdgeListParams p{};
    p.mask = mask; p.mask_channels = mask_channels; p.H = H; p.W = W; p.stride = mask_stride;
    p.n_pixels = n_pixels;
    p.edges = edges; p.capacity = capacity; p.counts = counts;
    p.n_chunks = (int)((n_pixels + kElChunk - 1) / kElChunk);
    p.chunk_counts = static_cast<int32_t*>(workspace);
    p.chunk_offsets = p.chunk_counts + p.n_chunks;
    SSLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (2 + B), st));
    edge_count_kernel<<<p.n_chunks, kElThreads, 0, st>>>(p);
    edge_scan_kernel<<<1, 1024, 0, st>>>(p);
    edge_emit_kernel<<<p.n_chunks, kElThreads, 0, st>>>(p);
    return check_launch("build_edge_list");
}

extern "C" int ssl_b200_ssg_rows_forward(const void* image, const void* image2, int dtype, int B, int C, int H, int W,
                                         const int32_t* edges, const int32_t* n_edges_dev, int max_edges, int ks,
                                         int kw, float sigma, float eps, int rows_mode, float* rows, float* rows2,
                                         void* stream) {
    SSLB_REQUIRE(image && rows && edges, "null pointer");
    SSLB_REQUIRE((image2 == nullptr) == (rows2 == nullptr), "image2 and rows2 go together");
    SSLB_REQUIRE(rows_mode >= SSL_B200_ROWS_RAW && rows_mode <= SSL_B200_ROWS_NORM, "bad rows_mode %d", rows_mode);
    SSLB_REQUIRE(rows_mode == SSL_B200_ROWS_RAW || sigma > 0.f, "sigma must be positive");
    SSLB_REQUIRE(max_edges >= 0 && B >= 1, "bad sizes");
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    PointParams p{};
    p.img[0] = image; p.img[1] = image2;
    p.rows[0] = rows; p.rows[1] = rows2;
    p.edges = EdgeRef{edges, nullptr};
    p.n_edges_dev = n_edges_dev;
    p.max_edges = max_edges;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.ks = ks; p.kw = kw; p.sigma = sigma; p.eps = eps; p.mode = rows_mode;
    return launch_point_forward(p, dtype, image2 ? 2 : 1, (cudaStream_t)stream);
}

extern "C" int ssl_b200_rows_grad_to_distance_grad(const float* rows, float* grad_rows, const int32_t* n_edges_dev,
                                                   int max_edges, int ks, int kw, int C, float sigma, int rows_mode,
                                                   void* stream) {
    SSLB_REQUIRE(rows && grad_rows, "null pointer");
    if (rows_mode == SSL_B200_ROWS_RAW || max_edges <= 0) return 0;
    SSLB_REQUIRE(sigma > 0.f, "sigma must be positive");
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const float chain = -1.0f / (sigma * (float)C * (float)(kw * kw));
    const int blocks = min(max_edges, di.sm_count * 16);
    rows_chain_kernel<<<blocks, kRowThreads, 0, (cudaStream_t)stream>>>(rows, grad_rows, n_edges_dev, max_edges,
                                                                       ks * ks, chain, rows_mode);
    return check_launch("rows_chain");
}

extern "C" int ssl_b200_ssg_rows_backward(const void* image, int dtype, int B, int C, int H, int W,
                                          const int32_t* edges, const int32_t* n_edges_dev, int max_edges, int ks,
                                          int kw, const float* gq, float* grad_image, void* stream) {
    SSLB_REQUIRE(image && edges && gq && grad_image, "null pointer");
    SSLB_REQUIRE(max_edges >= 0 && B >= 1, "bad sizes");
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    PointParams p{};
    p.img[0] = image;
    p.edges = EdgeRef{edges, nullptr};
    p.n_edges_dev = n_edges_dev;
    p.max_edges = max_edges;
    p.B = B; p.C = C; p.H = H; p.W = W;
    p.ks = ks; p.kw = kw;
    p.gq = gq;
    p.grad = grad_image;
    return launch_point_backward(p, dtype, (cudaStream_t)stream);
}

extern "C" int ssl_b200_row_loss_blocks(void) {
    DeviceInfo di;
    if (device_info(&di)) return 148 * 16;
    return di.sm_count * 16;
}

extern "C" int ssl_b200_row_loss(const float* rows_sr, const float* rows_gt, const int32_t* n_edges_dev, int max_edges,
                                 int ks, int kw, int C, float sigma, int rows_mode, float w_l1, float w_kl, float* gq,
                                 double* sums, double* scratch, void* stream) {
    SSLB_REQUIRE(rows_sr && rows_gt && sums && scratch, "null pointer");
    SSLB_REQUIRE(rows_mode == SSL_B200_ROWS_EXP || rows_mode == SSL_B200_ROWS_NORM,
                 "row loss is defined on exp / normalised rows");
    SSLB_REQUIRE(sigma > 0.f, "sigma must be positive");
    if (max_edges <= 0) return 0;
    RowLossParams p{};
    p.s = rows_sr; p.t = rows_gt; p.gq = gq;
    p.n_edges_dev = n_edges_dev; p.max_edges = max_edges; p.L = ks * ks;
    p.chain = -1.0f / (sigma * (float)C * (float)(kw * kw));
    p.mode = rows_mode; p.w_l1 = w_l1; p.w_kl = w_kl; p.scratch = scratch;
    const int blocks = min(max_edges, ssl_b200_row_loss_blocks());
    cudaStream_t st = (cudaStream_t)stream;
    row_loss_kernel<<<blocks, kRowThreads, 0, st>>>(p);
    row_loss_finalize_kernel<<<1, 32, 0, st>>>(scratch, blocks, sums);
    return check_launch("row_loss");
}

extern "C" int ssl_b200_laplacian_mask(const void* gt, int dtype, int B, int H, int W, float threshold, float* mask,
                                       void* stream) {
    SSLB_REQUIRE(gt && mask, "null pointer");
    SSLB_REQUIRE(B >= 1 && H >= 2 && W >= 2, "bad shape");
    const long long n = (long long)B * H * W;
    const int blocks = (int)((n + 255) / 256);
    SSLB_DISPATCH_DTYPE(dtype, T, {
        laplacian_mask_kernel<T><<<blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const T*>(gt), B, H, W,
                                                                            threshold, mask);
    });
    return check_launch("laplacian_mask");
}
