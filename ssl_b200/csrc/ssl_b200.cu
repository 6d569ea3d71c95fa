// libssl_b200.so -- C ABI over the sm_100a kernels (declared in include/ssl_b200.h).
// One translation unit; build with csrc/build.py (nvcc -gencode arch=compute_100a,code=sm_100a).
#include "ssl_b200.h"

#include "common.cuh"
#include "edge_list.cuh"
#include "row_ops.cuh"
#include "ssg_point.cuh"
#include "plane_host.cuh"
#include "pool_ops.cuh"

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
    StageTimer timer(kStagePointFwd, st);
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
    StageTimer timer(kStagePointBwd, st);
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
// Dense masks (>= 2 % of the pixels) with a supported (k_s, k_w, C = 3) run on the plane kernels -- 3-4x faster at
// the reference's mask densities and free of atomics (bitwise reproducible); everything else on the point kernels.
namespace {
bool ref_entry_uses_planes(int mc, int ks, int kw, int height, int width, int channel);
int ref_forward_planes(const float* image, const int32_t* pos, float* out, int mc, int ks, int kw, int Hp, int Wp,
                       cudaStream_t st);
int ref_backward_planes(const float* image, const float* grads, const int32_t* pos, float* image_grads, int mc, int ks,
                        int kw, int Hp, int Wp, cudaStream_t st);
}  // namespace


extern "C" int ssl_b200_compute_similarity(const float* image, const int32_t* pos, float* out, int mc, int psize,
                                           int ksize, int height, int width, int channel, void* stream) {
    SSLB_REQUIRE(image && out && (pos || mc == 0), "null pointer");
    SSLB_REQUIRE(mc >= 0, "negative mc");
    if (int e = check_sizes(psize, ksize, height, width, channel)) return e;
    if (ref_entry_uses_planes(mc, psize, ksize, height, width, channel))
        return ref_forward_planes(image, pos, out, mc, psize, ksize, height, width, (cudaStream_t)stream);
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
    if (ref_entry_uses_planes(mc, psize, ksize, height, width, channel))
        return ref_backward_planes(image, grads, pos, image_grads, mc, psize, ksize, height, width,
                                   (cudaStream_t)stream);
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
    StageTimer timer(kStageEdgeList, st);
    SSLB_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (2 + B), st));
    edge_count_kernel<<<p.n_chunks, kElThreads, 0, st>>>(p);
    edge_scan_kernel<<<1, 1024, 0, st>>>(p);
    edge_emit_kernel<<<p.n_chunks, kElThreads, 0, st>>>(p);
    return check_launch("build_edge_list", 3);
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
    StageTimer timer(kStageRowLoss, st);
    row_loss_kernel<<<blocks, kRowThreads, 0, st>>>(p);
    row_loss_finalize_kernel<<<1, 32, 0, st>>>(scratch, blocks, sums);
    return check_launch("row_loss", 2);
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

// ---- plane (tile-sharing) path ------------------------------------------------------------

extern "C" int ssl_b200_plane_supported(int ks, int kw, int channel) { return plane_supported(ks, kw, channel) ? 1 : 0; }

namespace {

// Workspace of a plane forward: lists | padded images | qT | qT2 | eout | eout2
struct PlaneFwdLayout {
    PlaneGeom g;
    int cap;
    PlaneListsLayout lists;
    size_t off_pad, off_q[2], off_eout[2], total;
};

template <typename Cfg>
PlaneFwdLayout plane_fwd_layout(int B, int H, int W, int max_edges) {
    PlaneFwdLayout l;
    l.g = geom_for<Cfg>(B, H, W);
    l.cap = slot_capacity(max_edges, l.g.n_units);
    l.lists = plane_lists_layout(l.g, l.cap);
    size_t o = l.lists.total;
    l.off_pad = o; o += align256(2 * pad_layout(B, H, W, Cfg::P).bytes_per_image_set);
    for (int i = 0; i < 2; ++i) { l.off_q[i] = o; o += align256((size_t)Cfg::L * l.cap * sizeof(float)); }
    for (int i = 0; i < 2; ++i) { l.off_eout[i] = o; o += align256((size_t)l.cap * Cfg::NCLS * Cfg::NCLS * sizeof(float)); }
    l.total = o;
    return l;
}

// The backward column lists pack (slot << 8 | row) into an int32: slots must stay below 2^23.
bool plane_slots_fit(int B, int H, int W, int max_edges) {
    // (upper bound of the unit count over every plane geometry: TYF >= 48, TXF = 64, 8 units per tile row)
    return (long long)max_edges + 3ll * B * (H / 48 + 1) * (W / 64 + 2) * 8 < (1ll << 23);
}

}  // namespace

namespace {

// (row, col) in padded coordinates -> flat index of the unpadded image; positions outside the image interior
// (which similaritywrapper.py:64-68 never produces) are dropped: their rows stay zero.
__global__ void __launch_bounds__(256) pos_to_edges_kernel(const int32_t* pos, int mc, int P, int H, int W,
                                                           int32_t* edges, int32_t* counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) { counts[0] = mc; counts[1] = mc; }
    if (i >= mc) return;
    const int y = pos[2 * i] - P, x = pos[2 * i + 1] - P;
    edges[i] = (y >= 0 && y < H && x >= 0 && x < W) ? y * W + x : -1;
}

// image_grads[c][Y][X] += G[c][Y][X] (the folded padded-domain gradient, HT x WT layout)
__global__ void __launch_bounds__(256) add_padded_grad_kernel(const float* G, int HT, int WT, int Hp, int Wp,
                                                              float* image_grads) {
    const long long n = 3ll * Hp * Wp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int X = (int)(i % Wp), Y = (int)((i / Wp) % Hp), c = (int)(i / ((long long)Wp * Hp));
        image_grads[i] += G[((long long)c * HT + Y) * WT + X];
    }
}

bool ref_entry_uses_planes(int mc, int ks, int kw, int height, int width, int channel) {
    const int P = ks / 2, H = height - 2 * P, W = width - 2 * P;
    if (!plane_supported(ks, kw, channel) || H <= P || W <= P || mc <= 0) return false;
    if (!plane_slots_fit(1, H, W, mc)) return false;
    return (double)mc >= 0.02 * (double)H * W;
}

struct RefScratch {   // stream-ordered scratch of one drop-in call
    char* base = nullptr;
    cudaStream_t st;
    ~RefScratch() { if (base) cudaFreeAsync(base, st); }
};

template <typename Cfg>
int ref_forward_planes_cfg(const float* image, const int32_t* pos, float* out, int mc, int Hp, int Wp, cudaStream_t st) {
    constexpr int P = Cfg::P;
    const int H = Hp - 2 * P, W = Wp - 2 * P;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PlaneFwdLayout l = plane_fwd_layout<Cfg>(1, H, W, mc);
    const size_t off_edges = l.total, off_counts = off_edges + align256((size_t)mc * sizeof(int32_t));
    RefScratch sc; sc.st = st;
    SSLB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sc.base), off_counts + 256, st));
    char* ws = sc.base;
    int32_t* edges = reinterpret_cast<int32_t*>(ws + off_edges);
    int32_t* counts = reinterpret_cast<int32_t*>(ws + off_counts);
    float* pad = reinterpret_cast<float*>(ws + l.off_pad);
    const PadLayout pl = pad_layout(1, H, W, P);
    pos_to_edges_kernel<<<(mc + 255) / 256, 256, 0, st>>>(pos, mc, P, H, W, edges, counts);
    pad_copy_kernel<<<di.sm_count * 4, 256, 0, st>>>(image, 3, Hp, Wp, pl.pitch, pad);
    if (int e = check_launch("ref_forward_prepare", 2)) return e;
    if (int e = launch_plane_lists(nullptr, 1, 0, edges, counts, mc, l.g, l.cap, ws, st, Cfg::SRP, 32 / Cfg::G)) return e;
    const PlaneLists lists = carve_lists(ws, l.lists, nullptr);
    float* q0 = reinterpret_cast<float*>(ws + l.off_q[0]);
    if (int e = launch_plane_forward_cfg<Cfg>(pad, 1, l.g, lists, l.cap, q0, nullptr,
                                              reinterpret_cast<float*>(ws + l.off_eout[0]), nullptr, st)) return e;
    plane_rows_to_reference_kernel<<<min(mc, di.sm_count * 16), 256, 0, st>>>(q0, l.cap, edges, counts, mc, lists.slot_map,
                                                                             Cfg::L, out);
    return check_launch("plane_rows_to_reference");
}

template <typename Cfg>
int ref_backward_planes_cfg(const float* image, const float* grads, const int32_t* pos, float* image_grads, int mc,
                            int Hp, int Wp, cudaStream_t st) {
    constexpr int P = Cfg::P;
    const int H = Hp - 2 * P, W = Wp - 2 * P;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PlaneStepLayout l = plane_step_layout<Cfg>(1, H, W, mc, 2 * di.sm_count, true);
    const size_t off_edges = l.total, off_counts = off_edges + align256((size_t)mc * sizeof(int32_t));
    RefScratch sc; sc.st = st;
    SSLB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&sc.base), off_counts + 256, st));
    char* ws = sc.base;
    int32_t* edges = reinterpret_cast<int32_t*>(ws + off_edges);
    int32_t* counts = reinterpret_cast<int32_t*>(ws + off_counts);
    float* pad = reinterpret_cast<float*>(ws + l.off_pad);
    pos_to_edges_kernel<<<(mc + 255) / 256, 256, 0, st>>>(pos, mc, P, H, W, edges, counts);
    pad_copy_kernel<<<di.sm_count * 4, 256, 0, st>>>(image, 3, Hp, Wp, l.pad.pitch, pad);
    if (int e = check_launch("ref_backward_prepare", 2)) return e;
    // dL/dq rows -> slots -> padded-domain gradient (everything of the rows backward but the pad adjoint, which
    // the caller's autograd applies: similaritywrapper.py:64)
    if (int e = launch_plane_rows_backward_padded_cfg<Cfg>(pad, 1, H, W, edges, counts, mc, grads, l, ws, st)) return e;
    add_padded_grad_kernel<<<di.sm_count * 4, 256, 0, st>>>(reinterpret_cast<const float*>(ws + l.off_gpart), l.HT, l.WT,
                                                           Hp, Wp, image_grads);
    return check_launch("add_padded_grad");
}

int ref_forward_planes(const float* image, const int32_t* pos, float* out, int mc, int ks, int kw, int Hp, int Wp,
                       cudaStream_t st) {
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, { return ref_forward_planes_cfg<Cfg>(image, pos, out, mc, Hp, Wp, st); });
}

int ref_backward_planes(const float* image, const float* grads, const int32_t* pos, float* image_grads, int mc, int ks,
                        int kw, int Hp, int Wp, cudaStream_t st) {
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg,
                            { return ref_backward_planes_cfg<Cfg>(image, grads, pos, image_grads, mc, Hp, Wp, st); });
}

}  // namespace

extern "C" size_t ssl_b200_plane_rows_workspace_bytes(int B, int H, int W, int ks, int kw, int max_edges) {
    if (!plane_supported(ks, kw, 3) || B < 1 || H < 1 || W < 1 || max_edges < 0) return 0;
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, { return plane_fwd_layout<Cfg>(B, H, W, max_edges).total; });
}

extern "C" int ssl_b200_plane_rows_forward(const void* image, const void* image2, int dtype, int B, int C, int H,
                                           int W, const int32_t* edges, const int32_t* n_edges_dev, int max_edges,
                                           int ks, int kw, float* rows, float* rows2, void* workspace,
                                           size_t workspace_bytes, void* stream) {
    SSLB_REQUIRE(image && rows && edges && workspace, "null pointer");
    SSLB_REQUIRE((image2 == nullptr) == (rows2 == nullptr), "image2 and rows2 go together");
    SSLB_REQUIRE(plane_supported(ks, kw, C), "no plane kernels for k_s=%d k_w=%d C=%d", ks, kw, C);
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    if (max_edges <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, {
        const PlaneFwdLayout l = plane_fwd_layout<Cfg>(B, H, W, max_edges);
        SSLB_REQUIRE(workspace_bytes >= l.total, "workspace too small (%zu < %zu)", workspace_bytes, l.total);
        char* ws = static_cast<char*>(workspace);
        float* pad = reinterpret_cast<float*>(ws + l.off_pad);
        if (int e = launch_pad(image, dtype, image2, dtype, B, H, W, Cfg::P, pad, st)) return e;
        if (int e = launch_plane_lists(nullptr, 1, 0, edges, n_edges_dev, max_edges, l.g, l.cap, ws, st, Cfg::SRP,
                                       32 / Cfg::G)) return e;
        const PlaneLists lists = carve_lists(ws, l.lists, nullptr);
        float* q0 = reinterpret_cast<float*>(ws + l.off_q[0]);
        float* q1 = image2 ? reinterpret_cast<float*>(ws + l.off_q[1]) : nullptr;
        if (int e = launch_plane_forward_cfg<Cfg>(pad, image2 ? 2 : 1, l.g, lists, l.cap, q0, q1,
                                                  reinterpret_cast<float*>(ws + l.off_eout[0]),
                                                  image2 ? reinterpret_cast<float*>(ws + l.off_eout[1]) : nullptr, st))
            return e;
        DeviceInfo di;
        if (int e = device_info(&di)) return e;
        const int blocks = min(max_edges, di.sm_count * 16);
        plane_rows_to_reference_kernel<<<blocks, 256, 0, st>>>(q0, l.cap, edges, n_edges_dev, max_edges, lists.slot_map,
                                                               Cfg::L, rows);
        if (image2)
            plane_rows_to_reference_kernel<<<blocks, 256, 0, st>>>(q1, l.cap, edges, n_edges_dev, max_edges,
                                                                   lists.slot_map, Cfg::L, rows2);
        return check_launch("plane_rows_to_reference", image2 ? 2 : 1);
    });
}

extern "C" int ssl_b200_rows_from_distance(float* rows, const int32_t* n_edges_dev, int max_edges, int ks, int kw,
                                           int C, float sigma, float eps, int rows_mode, void* stream) {
    SSLB_REQUIRE(rows, "null pointer");
    if (rows_mode == SSL_B200_ROWS_RAW || max_edges <= 0) return 0;
    SSLB_REQUIRE(sigma > 0.f, "sigma must be positive");
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const int blocks = min(max_edges, di.sm_count * 16);
    rows_finish_kernel<<<blocks, kRowThreads, 0, (cudaStream_t)stream>>>(rows, n_edges_dev, max_edges, ks * ks,
                                                                        (float)C * (float)(kw * kw), sigma, eps, rows_mode);
    return check_launch("rows_finish");
}

extern "C" size_t ssl_b200_plane_rows_backward_workspace_bytes(int B, int H, int W, int ks, int kw, int max_edges) {
    if (!plane_supported(ks, kw, 3) || B < 1 || H < 1 || W < 1 || max_edges < 0) return 0;
    DeviceInfo di;
    const int loss_blocks = 2 * (device_info(&di) ? 148 : di.sm_count);
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, { return plane_step_layout<Cfg>(B, H, W, max_edges, loss_blocks, true).total; });
}

extern "C" int ssl_b200_plane_rows_backward(const void* image, int dtype, int B, int C, int H, int W,
                                            const int32_t* edges, const int32_t* n_edges_dev, int max_edges, int ks,
                                            int kw, const float* gq, float* grad_image, void* workspace,
                                            size_t workspace_bytes, void* stream) {
    SSLB_REQUIRE(image && edges && n_edges_dev && gq && grad_image && workspace, "null pointer");
    SSLB_REQUIRE(plane_supported(ks, kw, C), "no plane kernels for k_s=%d k_w=%d C=%d", ks, kw, C);
    if (!plane_slots_fit(B, H, W, max_edges))
        return fail(SSL_B200_ENOTSUP, "batch too large for the plane path (%d edge pixels); use the point kernels",
                    max_edges);
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    if (max_edges <= 0) {
        SSLB_CUDA(cudaMemsetAsync(grad_image, 0, sizeof(float) * (size_t)B * C * H * W, st));
        return 0;
    }
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, {
        return launch_plane_rows_backward_cfg<Cfg>(image, dtype, B, H, W, edges, n_edges_dev, max_edges, gq, grad_image,
                                                   workspace, workspace_bytes, st);
    });
}

// ---- whole step ---------------------------------------------------------------------------

namespace {

__global__ void __launch_bounds__(256) scale_imm_kernel(float4* g, long long n4, float s) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = g[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        g[i] = v;
    }
}

// loss[0..2] = total, w_l1*L1, w_kl*KL from the terms of the (one or two) sub-batches and the host-known row count
__global__ void finalize_loss_parts_kernel(const double* terms, int n_parts, double n_tot, float w_l1, float w_kl,
                                           float* loss) {
    double s0 = 0.0, s1 = 0.0;
    for (int i = 0; i < n_parts; ++i) { s0 += terms[4 * i]; s1 += terms[4 * i + 1]; }
    const double l1 = (double)w_l1 * s0 / n_tot, kl = (double)w_kl * s1 / n_tot;
    loss[0] = (float)(l1 + kl);
    loss[1] = (float)l1;
    loss[2] = (float)kl;
}

// terms[2] = number of rows.  An edge list that overflowed its capacity (counts[1] = pixels found >
// counts[0] = pixels listed) would silently drop the trailing images of the batch: poison the count instead,
// so that the loss and the gradient scale come out NaN without any host round trip.
__global__ void set_terms_count_kernel(const int32_t* counts, int max_edges, double* terms) {
    const bool overflow = counts[1] > counts[0] || counts[0] > max_edges;
    terms[2] = overflow ? __longlong_as_double(0x7ff8000000000000ll) : (double)counts[0];
}

__global__ void __launch_bounds__(256) copy_rows_kernel(const float* src, const int32_t* n_dev, int max_edges, int L,
                                                        float* dst) {
    const long long n = (long long)edge_count(n_dev, max_edges) * L;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

}  // namespace

namespace {

size_t point_loss_workspace_bytes(int ks, int max_edges) {
    const size_t rows = align256((size_t)(max_edges > 0 ? max_edges : 1) * ks * ks * sizeof(float));
    return 2 * rows + align256(2 * sizeof(double) * (size_t)ssl_b200_row_loss_blocks());
}

// Plane kernels pay per image pixel, point kernels per edge pixel: below ~2 % density the point
// kernels win (measured crossover, profiles/).
bool use_plane_path(int path, int B, int C, int H, int W, int ks, int kw, int max_edges) {
    if (path == SSL_B200_PATH_POINT || !plane_supported(ks, kw, C)) return false;
    if (!plane_slots_fit(B, H, W, max_edges)) return false;
    if (path == SSL_B200_PATH_PLANE) return true;
    return (double)max_edges >= 0.02 * (double)B * H * W;
}

}  // namespace

extern "C" size_t ssl_b200_loss_workspace_bytes(int B, int C, int H, int W, int ks, int kw, int max_edges, int path) {
    if (B < 1 || C < 1 || H < 1 || W < 1) return 0;
    if (!use_plane_path(path, B, C, H, W, ks, kw, max_edges)) return point_loss_workspace_bytes(ks, max_edges);
    DeviceInfo di;
    const int loss_blocks = 2 * (device_info(&di) ? 148 : di.sm_count);
    SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, { return plane_step_layout<Cfg>(B, H, W, max_edges, loss_blocks, true).total; });
}

namespace {

// Shared body of the two whole-step entries.  `counts` = the edge list's counts (edges-based entry) or NULL
// (mask-based entry on the plane path: the unit lists carry their own counts).
int loss_step_impl(const StepInputs& in, const int32_t* counts, int B, int C, int H, int W, int max_edges, int ks,
                   int kw, float sigma, float eps, int rows_mode, float w_l1, float w_kl, float* grad_sr, double* terms,
                   void* workspace, size_t workspace_bytes, int path, cudaStream_t st) {
    SSLB_REQUIRE(rows_mode == SSL_B200_ROWS_EXP || rows_mode == SSL_B200_ROWS_NORM,
                 "the loss is defined on exp / normalised rows");
    SSLB_REQUIRE(workspace_bytes >= ssl_b200_loss_workspace_bytes(B, C, H, W, ks, kw, max_edges, path),
                 "workspace too small");
    SSLB_REQUIRE(path != SSL_B200_PATH_PLANE || plane_supported(ks, kw, C), "no plane kernels for k_s=%d k_w=%d C=%d",
                 ks, kw, C);
    SSLB_REQUIRE(path != SSL_B200_PATH_PLANE || max_edges <= 0 || use_plane_path(path, B, C, H, W, ks, kw, max_edges),
                 "batch too large for the plane path (%d edge pixels); use SSL_B200_PATH_AUTO or _POINT", max_edges);
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    SSLB_CUDA(cudaMemsetAsync(terms, 0, 3 * sizeof(double), st));
    const bool plane = max_edges > 0 && use_plane_path(path, B, C, H, W, ks, kw, max_edges);
    if (grad_sr && !plane) SSLB_CUDA(cudaMemsetAsync(grad_sr, 0, sizeof(float) * (size_t)B * C * H * W, st));
    if (counts) {
        set_terms_count_kernel<<<1, 1, 0, st>>>(counts, max_edges, terms);
        if (int e = check_launch("set_terms_count")) return e;
    }
    if (max_edges <= 0) return 0;
    if (plane) {
        SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, {
            return launch_plane_step_cfg<Cfg>(in, B, H, W, max_edges, sigma, eps, rows_mode, w_l1, w_kl, grad_sr, terms,
                                              workspace, workspace_bytes, st);
        });
    }
    SSLB_REQUIRE(in.edges && counts, "the point kernels need the flat edge list");
    SSLB_REQUIRE(in.dtype_sr == in.dtype_gt, "the point kernels take both images in one element type");
    const size_t rows_bytes = align256((size_t)max_edges * ks * ks * sizeof(float));
    char* ws = static_cast<char*>(workspace);
    float* rows_sr = reinterpret_cast<float*>(ws);
    float* rows_gt = reinterpret_cast<float*>(ws + rows_bytes);
    double* scratch = reinterpret_cast<double*>(ws + 2 * rows_bytes);
    if (int e = ssl_b200_ssg_rows_forward(in.sr, in.gt, in.dtype_sr, B, C, H, W, in.edges, counts, max_edges, ks, kw, sigma,
                                          eps, rows_mode, rows_sr, rows_gt, st)) return e;
    // rows_sr is overwritten in place by dL/dq when a gradient is wanted
    if (int e = ssl_b200_row_loss(rows_sr, rows_gt, counts, max_edges, ks, kw, C, sigma, rows_mode, w_l1, w_kl,
                                  grad_sr ? rows_sr : nullptr, terms, scratch, st)) return e;
    if (grad_sr)
        if (int e = ssl_b200_ssg_rows_backward(in.sr, in.dtype_sr, B, C, H, W, in.edges, counts, max_edges, ks, kw, rows_sr,
                                               grad_sr, st)) return e;
    return 0;
}

struct StepExtras {   // edge list carved behind the loss workspace (mask-based entry, point path)
    size_t off_edges, off_counts, off_elws, el_ws_bytes, total;
};

StepExtras step_extras(size_t loss_ws, int B, int H, int W, int max_edges) {
    StepExtras x;
    size_t o = align256(loss_ws);
    x.off_edges = o; o += align256((size_t)(max_edges > 0 ? max_edges : 1) * sizeof(int32_t));
    x.off_counts = o; o += align256((size_t)(2 + B) * sizeof(int32_t));
    x.el_ws_bytes = ssl_b200_edge_list_workspace_bytes((int64_t)B * H * W);
    x.off_elws = o; o += align256(x.el_ws_bytes);
    x.total = o;
    return x;
}

}  // namespace

extern "C" int ssl_b200_loss_forward_backward(const void* sr, const void* gt, int dtype, int B, int C, int H, int W,
                                              const int32_t* edges, const int32_t* counts, int max_edges, int ks,
                                              int kw, float sigma, float eps, int rows_mode, float w_l1, float w_kl,
                                              float* grad_sr, double* terms, void* workspace, size_t workspace_bytes,
                                              int path, void* stream) {
    SSLB_REQUIRE(sr && gt && edges && counts && terms && workspace, "null pointer");
    StepInputs in{};
    in.sr = sr; in.gt = gt; in.dtype_sr = dtype; in.dtype_gt = dtype;
    in.mask = nullptr; in.mask_channels = 1; in.mask_stride = 0;
    in.edges = edges; in.n_edges_dev = counts;
    return loss_step_impl(in, counts, B, C, H, W, max_edges, ks, kw, sigma, eps, rows_mode, w_l1, w_kl, grad_sr, terms,
                          workspace, workspace_bytes, path, (cudaStream_t)stream);
}

extern "C" size_t ssl_b200_loss_step_workspace_bytes(int B, int C, int H, int W, int ks, int kw, int max_edges,
                                                     int path) {
    const size_t base = ssl_b200_loss_workspace_bytes(B, C, H, W, ks, kw, max_edges, path);
    if (base == 0 || use_plane_path(path, B, C, H, W, ks, kw, max_edges)) return base;
    return step_extras(base, B, H, W, max_edges).total;
}

extern "C" int ssl_b200_loss_step(const void* sr, int dtype_sr, const void* gt, int dtype_gt, const float* mask,
                                  int mask_channels, int mask_stride, int B, int C, int H, int W, int max_edges, int ks,
                                  int kw, float sigma, float eps, int rows_mode, float w_l1, float w_kl, float* grad_sr,
                                  double* terms, void* workspace, size_t workspace_bytes, int path, void* stream) {
    SSLB_REQUIRE(sr && gt && mask && terms && workspace, "null pointer");
    SSLB_REQUIRE(B >= 1 && mask_channels >= 1 && max_edges >= 0, "bad shape");
    SSLB_REQUIRE(workspace_bytes >= ssl_b200_loss_step_workspace_bytes(B, C, H, W, ks, kw, max_edges, path),
                 "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    StepInputs in{};
    in.sr = sr; in.gt = gt; in.dtype_sr = dtype_sr; in.dtype_gt = dtype_gt;
    in.mask = mask; in.mask_channels = mask_channels; in.mask_stride = mask_stride;
    if (max_edges > 0 && use_plane_path(path, B, C, H, W, ks, kw, max_edges))
        return loss_step_impl(in, nullptr, B, C, H, W, max_edges, ks, kw, sigma, eps, rows_mode, w_l1, w_kl, grad_sr, terms,
                              workspace, workspace_bytes, path, st);
    // point kernels (sparse masks, other kernel sizes): they walk the flat edge list
    const size_t base = ssl_b200_loss_workspace_bytes(B, C, H, W, ks, kw, max_edges, path);
    const StepExtras x = step_extras(base, B, H, W, max_edges);
    char* ws = static_cast<char*>(workspace);
    int32_t* edges = reinterpret_cast<int32_t*>(ws + x.off_edges);
    int32_t* counts = reinterpret_cast<int32_t*>(ws + x.off_counts);
    if (int e = ssl_b200_build_edge_list(mask, B, mask_channels, H, W, mask_stride, edges, max_edges, counts,
                                         ws + x.off_elws, x.el_ws_bytes, stream)) return e;
    in.mask = nullptr;
    in.edges = edges; in.n_edges_dev = counts;
    return loss_step_impl(in, counts, B, C, H, W, max_edges, ks, kw, sigma, eps, rows_mode, w_l1, w_kl, grad_sr, terms,
                          workspace, base, path, st);
}

extern "C" int ssl_b200_loss_export_distance_grad(const void* workspace, size_t workspace_bytes, int B, int C, int H,
                                                  int W, const int32_t* edges, const int32_t* counts, int max_edges,
                                                  int ks, int kw, int path, float* gq_rows, void* stream) {
    SSLB_REQUIRE(workspace && edges && counts && gq_rows, "null pointer");
    SSLB_REQUIRE(workspace_bytes >= ssl_b200_loss_workspace_bytes(B, C, H, W, ks, kw, max_edges, path),
                 "workspace too small");
    if (max_edges <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const char* ws = static_cast<const char*>(workspace);
    if (use_plane_path(path, B, C, H, W, ks, kw, max_edges)) {
        SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, {
            const PlaneStepLayout l = plane_step_layout<Cfg>(B, H, W, max_edges, 2 * di.sm_count, true);
            const PlaneLists lists = carve_lists(const_cast<char*>(ws), l.lists, nullptr);
            plane_rows_to_reference_kernel<<<min(max_edges, di.sm_count * 16), 256, 0, st>>>(
                reinterpret_cast<const float*>(ws + l.off_q[0]), l.cap, edges, counts, max_edges, lists.slot_map, Cfg::L,
                gq_rows);
            return check_launch("export_distance_grad");
        });
    }
    copy_rows_kernel<<<di.sm_count * 8, 256, 0, st>>>(reinterpret_cast<const float*>(ws), counts, max_edges, ks * ks,
                                                      gq_rows);
    return check_launch("export_distance_grad");
}

namespace {

struct HostArena {
    int dev = -1;
    char* base = nullptr;
    size_t bytes = 0;
    int32_t* counts_pinned = nullptr;
    cudaStream_t copy_stream = nullptr;   // carries the GT upload while the SR half of the forward runs
    cudaEvent_t counted = nullptr, sr_sent = nullptr, grad_ready = nullptr, gt_ready[2] = {nullptr, nullptr},
                sr_ready[2] = {nullptr, nullptr};
};

HostArena& arena() {
    static thread_local HostArena a;
    return a;
}

int arena_reserve(size_t bytes) {
    HostArena& a = arena();
    int dev = 0;
    SSLB_CUDA(cudaGetDevice(&dev));
    if (a.dev == dev && a.bytes >= bytes) return 0;
    if (a.base) {
        SSLB_CUDA(cudaDeviceSynchronize());
        SSLB_CUDA(cudaFree(a.base));
        a.base = nullptr;
        a.bytes = 0;
    }
    if (!a.counts_pinned) SSLB_CUDA(cudaMallocHost(&a.counts_pinned, 4 * sizeof(int32_t)));
    if (!a.copy_stream) {
        SSLB_CUDA(cudaStreamCreateWithFlags(&a.copy_stream, cudaStreamNonBlocking));
        SSLB_CUDA(cudaEventCreateWithFlags(&a.counted, cudaEventDisableTiming));
        SSLB_CUDA(cudaEventCreateWithFlags(&a.sr_sent, cudaEventDisableTiming));
        SSLB_CUDA(cudaEventCreateWithFlags(&a.grad_ready, cudaEventDisableTiming));
        for (int h = 0; h < 2; ++h) {
            SSLB_CUDA(cudaEventCreateWithFlags(&a.gt_ready[h], cudaEventDisableTiming));
            SSLB_CUDA(cudaEventCreateWithFlags(&a.sr_ready[h], cudaEventDisableTiming));
        }
    }
    SSLB_CUDA(cudaMalloc(&a.base, bytes));
    a.bytes = bytes;
    a.dev = dev;
    return 0;
}

}  // namespace

extern "C" int ssl_b200_release_host_arena(void) {
    HostArena& a = arena();
    if (a.base) {
        SSLB_CUDA(cudaDeviceSynchronize());
        SSLB_CUDA(cudaFree(a.base));
    }
    if (a.counts_pinned) SSLB_CUDA(cudaFreeHost(a.counts_pinned));
    if (a.copy_stream) {
        cudaEventDestroy(a.counted);
        cudaEventDestroy(a.sr_sent);
        cudaEventDestroy(a.grad_ready);
        for (int h = 0; h < 2; ++h) { cudaEventDestroy(a.gt_ready[h]); cudaEventDestroy(a.sr_ready[h]); }
        cudaStreamDestroy(a.copy_stream);
    }
    a = HostArena{};
    return 0;
}

// Timeline of one call (st = the caller's stream, cs = the arena's copy stream).  A batch that runs on the plane
// kernels is processed as two half-batches h0, h1 so that transfers hide behind arithmetic:
//   st: mask -> device, count edge pixels per half, counts -> host | SR(h0) -> device |
//       step(h0) [waits for GT(h0) after the SR half of its forward] | scale grad(h0) |
//       step(h1) [after SR(h1); waits for GT(h1)] | scale grad(h1) | loss | loss, grad(h1) -> host
//   cs: (after SR(h0))  GT(h0), SR(h1), GT(h1) -> device  ...  (after scale grad(h0))  grad(h0) -> host
// The host reads the counts (they size the rows workspace and give the 1/N of the mean) while SR(h0) is still on
// the wire.  Both halves share one workspace: they run back to back on st.
extern "C" int ssl_b200_loss_step_host(const float* sr_host, const float* gt_host, const float* mask_host,
                                       int mask_channels, int B, int C, int H, int W, int mask_stride, int ks, int kw,
                                       float sigma, float eps, int rows_mode, float w_l1, float w_kl, float* loss_host,
                                       float* grad_host, int64_t* n_rows_host, void* stream) {
    SSLB_REQUIRE(sr_host && gt_host && mask_host && loss_host, "null pointer");
    SSLB_REQUIRE(B >= 1 && mask_channels >= 1, "bad shape");
    if (int e = check_sizes(ks, kw, H, W, C)) return e;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const size_t n_px = (size_t)B * H * W, img_bytes = align256(n_px * C * sizeof(float));
    SSLB_REQUIRE(n_px < (1ull << 31), "batch too large for int32 pixel indices");
    const size_t mask_bytes = align256(n_px * mask_channels * sizeof(float));
    const size_t small_bytes = 512;  // terms double[2][4] | loss float[3] | counts int32[2]
    const size_t fixed = 3 * img_bytes + mask_bytes + small_bytes;
    // two halves when the batch can be split and the plane kernels exist for it
    // (the halves of the gradient are scaled as float4: the split must fall on a 16-byte boundary)
    const int n_parts = (B >= 2 && plane_supported(ks, kw, C) && ((size_t)(B / 2) * H * W * C) % 4 == 0) ? 2 : 1;
    const int Bh[2] = {n_parts == 2 ? B / 2 : B, n_parts == 2 ? B - B / 2 : 0};
    // first pass with the arena we have (or a guess of one edge pixel in eight); rows need the edge count
    HostArena& a = arena();
    if (int e = arena_reserve(fixed + ssl_b200_loss_step_workspace_bytes(Bh[n_parts - 1], C, H, W, ks, kw,
                                                                        (int)((size_t)Bh[n_parts - 1] * H * W / 8 + 1), 0)))
        return e;
    auto carve = [&](char*& p, size_t b) { char* r = p; p += b; return r; };
    char* p = a.base;
    float* d_mask = (float*)carve(p, mask_bytes);
    char* d_small = carve(p, small_bytes);
    float* d_sr = (float*)carve(p, img_bytes);
    float* d_gt = (float*)carve(p, img_bytes);
    float* d_grad = (float*)carve(p, img_bytes);
    char* d_ws = p;
    double* d_terms = (double*)d_small;            // [2][4]
    float* d_loss = (float*)(d_small + 256);
    int32_t* d_count = (int32_t*)(d_small + 384);  // [2]
    const size_t px_img = (size_t)H * W;
    const size_t off_px[2] = {0, (size_t)Bh[0] * px_img};   // first pixel of each half
    // mask first, so the edge counts come back while the images are still on the wire
    SSLB_CUDA(cudaMemcpyAsync(d_mask, mask_host, n_px * mask_channels * sizeof(float), cudaMemcpyHostToDevice, st));
    SSLB_CUDA(cudaMemsetAsync(d_count, 0, 2 * sizeof(int32_t), st));
    {
        StageTimer timer(kStageEdgeList, st);
        for (int h = 0; h < n_parts; ++h) {
            EdgeListParams ep{};
            ep.mask = d_mask + off_px[h] * mask_channels; ep.mask_channels = mask_channels; ep.H = H; ep.W = W;
            ep.stride = mask_stride;
            ep.n_pixels = (long long)Bh[h] * px_img;
            mask_count_kernel<<<di.sm_count * 2, kElThreads, 0, st>>>(ep, d_count + h);
        }
        if (int e = check_launch("mask_count", n_parts)) return e;
    }
    SSLB_CUDA(cudaMemcpyAsync(a.counts_pinned, d_count, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    SSLB_CUDA(cudaEventRecord(a.counted, st));
    SSLB_CUDA(cudaMemcpyAsync(d_sr, sr_host, (size_t)Bh[0] * px_img * C * sizeof(float), cudaMemcpyHostToDevice, st));
    SSLB_CUDA(cudaEventRecord(a.sr_sent, st));
    // everything else follows on the copy stream, in the order it is needed (not beside SR(h0): that one is on the
    // critical path)
    SSLB_CUDA(cudaStreamWaitEvent(a.copy_stream, a.sr_sent, 0));
    for (int h = 0; h < n_parts; ++h) {
        const size_t o = off_px[h] * C, n = (size_t)Bh[h] * px_img * C;
        if (h > 0) {
            SSLB_CUDA(cudaMemcpyAsync(d_sr + o, sr_host + o, n * sizeof(float), cudaMemcpyHostToDevice, a.copy_stream));
            SSLB_CUDA(cudaEventRecord(a.sr_ready[h], a.copy_stream));
        }
        SSLB_CUDA(cudaMemcpyAsync(d_gt + o, gt_host + o, n * sizeof(float), cudaMemcpyHostToDevice, a.copy_stream));
        SSLB_CUDA(cudaEventRecord(a.gt_ready[h], a.copy_stream));
    }
    cudaError_t ce = cudaEventSynchronize(a.counted);
    if (ce != cudaSuccess) return fail((int)ce, "edge count readback: %s", cudaGetErrorString(ce));
    const int n_half[2] = {a.counts_pinned[0], n_parts == 2 ? a.counts_pinned[1] : 0};
    const long long n_rows = (long long)n_half[0] + n_half[1];
    size_t ws_bytes = 0;
    for (int h = 0; h < n_parts; ++h) {
        const size_t w = ssl_b200_loss_step_workspace_bytes(Bh[h], C, H, W, ks, kw, n_half[h], 0);
        ws_bytes = w > ws_bytes ? w : ws_bytes;
    }
    if (fixed + ws_bytes > a.bytes) {
        // grow (rare: first call, or a denser mask than ever seen) and redo the uploads
        SSLB_CUDA(cudaStreamSynchronize(a.copy_stream));
        SSLB_CUDA(cudaStreamSynchronize(st));
        if (int e = arena_reserve(fixed + ws_bytes + ws_bytes / 4)) return e;
        return ssl_b200_loss_step_host(sr_host, gt_host, mask_host, mask_channels, B, C, H, W, mask_stride, ks, kw,
                                       sigma, eps, rows_mode, w_l1, w_kl, loss_host, grad_host, n_rows_host, stream);
    }
    const double n_tot = n_rows > 0 ? (double)n_rows * ks * ks : 1.0;
    const float inv_n = (float)(1.0 / n_tot);
    for (int h = 0; h < n_parts; ++h) {
        const size_t o = off_px[h] * C;
        StepInputs in{};
        in.sr = d_sr + o; in.gt = d_gt + o; in.dtype_sr = SSL_B200_F32; in.dtype_gt = SSL_B200_F32;
        in.mask = d_mask + off_px[h] * mask_channels; in.mask_channels = mask_channels; in.mask_stride = mask_stride;
        float* grad_h = grad_host ? d_grad + o : nullptr;
        double* terms_h = d_terms + 4 * h;
        if (h > 0) SSLB_CUDA(cudaStreamWaitEvent(st, a.sr_ready[h], 0));
        const bool plane = n_half[h] > 0 && use_plane_path(SSL_B200_PATH_AUTO, Bh[h], C, H, W, ks, kw, n_half[h]);
        int rc;
        if (plane) {
            in.gt_ready = a.gt_ready[h];   // the step waits for GT itself, after the SR half of the forward
            rc = loss_step_impl(in, nullptr, Bh[h], C, H, W, n_half[h], ks, kw, sigma, eps, rows_mode, w_l1, w_kl, grad_h,
                                terms_h, d_ws, ws_bytes, SSL_B200_PATH_AUTO, st);
        } else {
            SSLB_CUDA(cudaStreamWaitEvent(st, a.gt_ready[h], 0));
            rc = ssl_b200_loss_step(in.sr, SSL_B200_F32, in.gt, SSL_B200_F32, in.mask, mask_channels, mask_stride, Bh[h],
                                    C, H, W, n_half[h], ks, kw, sigma, eps, rows_mode, w_l1, w_kl, grad_h, terms_h, d_ws,
                                    ws_bytes, SSL_B200_PATH_AUTO, stream);
        }
        if (rc) return rc;
        if (grad_host) {
            const size_t n = (size_t)Bh[h] * px_img * C;   // a multiple of 4 for all but degenerate shapes
            const long long n4 = (long long)((n + 3) / 4);  // the tail of the last half runs into our own padding
            scale_imm_kernel<<<di.sm_count * 4, 256, 0, st>>>((float4*)(d_grad + o), n4, inv_n);
            if (int e = check_launch("scale_grad")) return e;
            if (h + 1 < n_parts) {
                // the first half's gradient goes home on the copy stream while the second half is computed
                SSLB_CUDA(cudaEventRecord(a.grad_ready, st));
                SSLB_CUDA(cudaStreamWaitEvent(a.copy_stream, a.grad_ready, 0));
                SSLB_CUDA(cudaMemcpyAsync(grad_host + o, d_grad + o, n * sizeof(float), cudaMemcpyDeviceToHost, a.copy_stream));
            } else {
                SSLB_CUDA(cudaMemcpyAsync(grad_host + o, d_grad + o, n * sizeof(float), cudaMemcpyDeviceToHost, st));
            }
        }
    }
    finalize_loss_parts_kernel<<<1, 1, 0, st>>>(d_terms, n_parts, n_tot, w_l1, w_kl, d_loss);
    if (int e = check_launch("finalize_loss")) return e;
    SSLB_CUDA(cudaMemcpyAsync(loss_host, d_loss, 3 * sizeof(float), cudaMemcpyDeviceToHost, st));
    SSLB_CUDA(cudaStreamSynchronize(st));
    SSLB_CUDA(cudaStreamSynchronize(a.copy_stream));
    if (n_rows_host) *n_rows_host = n_rows;
    return 0;
}

namespace {
__global__ void loss_from_terms_kernel(const double* terms, int L, float w_l1, float w_kl, float grad_scale, float* out) {
    const double n_tot = fmax(terms[2] * (double)L, 1.0);   // NaN (poisoned count) propagates: fmax(NaN, 1) = 1 is avoided below
    const double n = terms[2] != terms[2] ? terms[2] : n_tot;
    const double l1 = (double)w_l1 * terms[0] / n, kl = (double)w_kl * terms[1] / n;
    out[0] = (float)(l1 + kl);
    out[1] = (float)l1;
    out[2] = (float)kl;
    out[3] = (float)((double)grad_scale / n);
}
}  // namespace

extern "C" int ssl_b200_loss_from_terms(const double* terms, int row_len, float w_l1, float w_kl, float grad_scale,
                                        float* out4, void* stream) {
    SSLB_REQUIRE(terms && out4 && row_len >= 1, "bad arguments");
    loss_from_terms_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(terms, row_len, w_l1, w_kl, grad_scale, out4);
    return check_launch("loss_from_terms");
}

// ---- crop + training-pair pool -------------------------------------------------

extern "C" int ssl_b200_crop(const void* src, void* dst, int planes, int H, int W, int top, int left, int h, int w,
                             void* stream) {
    SSLB_REQUIRE(src && dst, "null pointer");
    SSLB_REQUIRE(planes >= 1 && h >= 1 && w >= 1 && top >= 0 && left >= 0 && top + h <= H && left + w <= W,
                 "crop [%d:%d, %d:%d] does not fit %dx%d", top, top + h, left, left + w, H, W);
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const long long n = (long long)planes * h * w;
    const int blocks = (int)min((long long)di.sm_count * 8, (n + 255) / 256);
    crop_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(static_cast<const uint32_t*>(src), static_cast<uint32_t*>(dst),
                                                          planes, H, W, top, left, h, w);
    return check_launch("crop");
}

extern "C" int ssl_b200_pool_exchange(void* queue, const void* in, void* out, const int32_t* slots, int b,
                                      int64_t sample_elems, int bcast_channels, void* stream) {
    SSLB_REQUIRE(queue && in && slots, "null pointer");
    SSLB_REQUIRE(b >= 1 && sample_elems >= 1, "bad sizes");
    SSLB_REQUIRE(bcast_channels <= 1 || sample_elems % bcast_channels == 0, "sample size is not a multiple of the channels");
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const long long n = (long long)b * sample_elems;
    const int blocks = (int)min((long long)di.sm_count * 8, (n + 255) / 256);
    pool_exchange_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        static_cast<uint32_t*>(queue), static_cast<const uint32_t*>(in), static_cast<uint32_t*>(out), slots, b,
        (long long)sample_elems, bcast_channels, bcast_channels > 1 ? (long long)sample_elems / bcast_channels : 0);
    return check_launch("pool_exchange");
}

extern "C" int ssl_b200_profile_enable(int on) {
    Profiler& p = profiler();
    p.enabled = on != 0;
    p.n = 0;
    return 0;
}

extern "C" int ssl_b200_profile_num_stages(void) { return kNumStages; }
extern "C" const char* ssl_b200_profile_stage_name(int i) { return stage_name(i); }

extern "C" int ssl_b200_profile_read(float* ms, int* launches) {
    SSLB_REQUIRE(ms && launches, "null pointer");
    Profiler& p = profiler();
    for (int i = 0; i < kNumStages; ++i) { ms[i] = 0.f; launches[i] = 0; }
    for (int i = 0; i < p.n; ++i) {
        SSLB_CUDA(cudaEventSynchronize(p.stop[i]));
        float t = 0.f;
        SSLB_CUDA(cudaEventElapsedTime(&t, p.start[i], p.stop[i]));
        ms[p.stage[i]] += t;
        launches[p.stage[i]] += 1;
    }
    p.n = 0;
    return 0;
}

extern "C" uint64_t ssl_b200_launch_count(void) { return launch_counter().load(std::memory_order_relaxed); }
