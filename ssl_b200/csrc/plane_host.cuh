// Host-side launch code of the plane path (included by ssl_b200.cu only).
#pragma once

#include <stdlib.h>

#include "pad.cuh"
#include "plane_geom.cuh"
#include "row_loss_t.cuh"
#include "ssg_plane_fwd.cuh"
#include "ssg_plane_bwd.cuh"

namespace sslb {

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

// Which (k_search, k_window) pairs have plane kernels.  Everything else uses the point kernels.
inline bool plane_supported(int ks, int kw, int C) {
    return C == 3 && ((ks == 25 && kw == 9) || (ks == 11 && kw == 5) || (ks == 7 && kw == 3));
}

#define SSLB_DISPATCH_PLANE_CFG(ks, kw, Cfg, ...)                                            \
    if ((ks) == 25 && (kw) == 9) { using Cfg = PlaneCfg<25, 9>; __VA_ARGS__; }               \
    else if ((ks) == 11 && (kw) == 5) { using Cfg = PlaneCfg<11, 5>; __VA_ARGS__; }           \
    else if ((ks) == 7 && (kw) == 3) { using Cfg = PlaneCfg<7, 3>; __VA_ARGS__; }             \
    else return sslb::fail(SSL_B200_ENOTSUP, "no plane kernels for k_s=%d k_w=%d", (int)(ks), (int)(kw));

inline int slot_capacity(int max_edges, int n_units) {
    // every unit pads its edge count to a multiple of 4
    long long c = (long long)max_edges + 3ll * n_units;
    c = (c + kPanel - 1) & ~(long long)(kPanel - 1);   // whole panels of the rows buffers
    return (int)c;
}

// Carve the list arrays out of one workspace.  Layout is a function of (geometry, capacity) only.
struct PlaneListsLayout {
    size_t off_counts, off_unit_start, off_unit_count, off_slot_pix, off_slot_rc, off_slot_map, total;
};

inline PlaneListsLayout plane_lists_layout(const PlaneGeom& g, int cap) {
    PlaneListsLayout l;
    size_t o = 0;
    l.off_counts = o; o += align256(8 * sizeof(int32_t));
    l.off_unit_start = o; o += align256((size_t)(g.n_units + 1) * sizeof(int32_t));
    l.off_unit_count = o; o += align256((size_t)g.n_units * sizeof(int32_t));
    l.off_slot_pix = o; o += align256((size_t)cap * sizeof(int32_t));
    l.off_slot_rc = o; o += align256((size_t)cap * sizeof(int32_t));
    l.off_slot_map = o; o += align256((size_t)g.B * g.H * g.W * sizeof(int32_t));
    l.total = o;
    return l;
}

inline PlaneLists carve_lists(void* ws, const PlaneListsLayout& l, int32_t** unit_count) {
    char* b = static_cast<char*>(ws);
    PlaneLists out;
    out.counts = reinterpret_cast<int32_t*>(b + l.off_counts);
    out.unit_start = reinterpret_cast<int32_t*>(b + l.off_unit_start);
    out.slot_pix = reinterpret_cast<int32_t*>(b + l.off_slot_pix);
    out.slot_rc = reinterpret_cast<int32_t*>(b + l.off_slot_rc);
    out.slot_map = reinterpret_cast<int32_t*>(b + l.off_slot_map);
    if (unit_count) *unit_count = reinterpret_cast<int32_t*>(b + l.off_unit_count);
    return out;
}

template <typename Cfg>
inline PlaneGeom geom_for(int B, int H, int W) {
    return make_geom(B, H, W, Cfg::TYF, Cfg::TXF, Cfg::K, Cfg::P);
}

// mask (or flat edge list when mask == NULL) -> unit lists
inline int launch_plane_lists(const float* mask, int mask_channels, int stride, const int32_t* edges,
                              const int32_t* n_edges_dev, int max_edges, const PlaneGeom& g, int cap, void* ws,
                              cudaStream_t st, int srp = 17, int spread = 8) {
    const PlaneListsLayout lay = plane_lists_layout(g, cap);
    PlaneListParams p{};
    p.mask = mask; p.mask_channels = mask_channels; p.stride = stride;
    p.edges = edges; p.n_edges_dev = n_edges_dev; p.max_edges = max_edges;
    p.g = g; p.capacity = cap;
    p.out = carve_lists(ws, lay, &p.unit_count);
    StageTimer timer(kStagePlaneLists, st);
    SSLB_CUDA(cudaMemsetAsync(p.out.counts, 0, 8 * sizeof(int32_t), st));   // [4] = the row loss's block counter
    const long long npx = (long long)g.B * g.H * g.W;
    const int fill_blocks = (int)((npx + 255) / 256 < 4096 ? (npx + 255) / 256 : 4096);
    fill_i32_kernel<<<fill_blocks, 256, 0, st>>>(p.out.slot_map, npx, -1);
    int n_launch = 4;
    if (!mask && max_edges > 0) {
        plane_mark_edges_kernel<<<(max_edges + 255) / 256 < 2048 ? (max_edges + 255) / 256 : 2048, 256, 0, st>>>(p);
        ++n_launch;
    }
    const int blocks = (g.n_units * 32 + 255) / 256;
    plane_units_count_kernel<<<blocks, 256, 0, st>>>(p);
    plane_units_scan_kernel<<<1, 1024, 0, st>>>(p);
    plane_units_emit_kernel<<<(g.n_units + 7) / 8, 256, 0, st>>>(p, srp, spread);
    return check_launch("plane_lists", n_launch);
}

// ---- reflect-padded fp32 copy of the batch + its tensor maps ---------------------------------
struct PadLayout {
    int Hp, Wp, pitch;
    size_t bytes_per_image_set;   // B * 3 * Hp * pitch floats (a multiple of 16 bytes; two sets are contiguous)
};

inline PadLayout pad_layout(int B, int H, int W, int P) {
    PadLayout l;
    l.Hp = H + 2 * P; l.Wp = W + 2 * P; l.pitch = pad_pitch(W, P);
    l.bytes_per_image_set = (size_t)B * 3 * l.Hp * l.pitch * sizeof(float);
    return l;
}

// Pads img0 (and img1 right behind it when given) into `out`.
inline int launch_pad(const void* img0, int dtype0, const void* img1, int dtype1, int B, int H, int W, int P,
                      float* out, cudaStream_t st) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PadLayout l = pad_layout(B, H, W, P);
    PadParams pp{};
    pp.img[0] = img0; pp.img[1] = img1; pp.dtype[0] = dtype0; pp.dtype[1] = dtype1;
    pp.n_img = img1 ? 2 : 1;
    pp.B = B; pp.C = 3; pp.H = H; pp.W = W; pp.P = P; pp.Hp = l.Hp; pp.Wp = l.Wp; pp.pitch = l.pitch;
    pp.out = out;
    StageTimer timer(kStagePad, st);
    pad_images_kernel<<<di.sm_count * 8, 256, 0, st>>>(pp);
    return check_launch("pad_images");
}

// img_first / n_img: which of the (up to two) padded image sets this launch covers (0 = SR, 1 = GT).
template <typename Cfg>
inline int launch_plane_forward_cfg(const float* pad, int n_img, const PlaneGeom& g, const PlaneLists& lists, int cap,
                                    float* qT, float* qT2, float* eout, float* eout2, cudaStream_t st,
                                    int img_first = 0, int n_pad_sets = -1) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PadLayout pl = pad_layout(g.B, g.H, g.W, Cfg::P);
    PlaneFwdParams p{};
    p.pad = pad; p.Hp = pl.Hp; p.pitch = pl.pitch;
    p.qT[0] = qT; p.qT[1] = qT2;
    p.eout[0] = eout; p.eout[1] = eout2;
    p.lists = lists; p.g = g; p.cap = cap;
    const size_t smem = plane_fwd_smem_bytes<Cfg>();
    SSLB_REQUIRE(smem <= (size_t)di.max_smem_optin, "plane forward needs %zu B of shared memory", smem);
    const int tiles = g.B * g.nty * g.ntx;
    CUtensorMap tmap;
    // both image sets are one contiguous stack of n_img * B * 3 planes
    if (n_pad_sets < 0) n_pad_sets = n_img;
    p.img_first = img_first;
    if (int e = make_plane_map(&tmap, pad, n_pad_sets * g.B * 3, pl.Hp, pl.Wp, pl.pitch, Cfg::IPITCH, Cfg::IROWS, 3)) return e;
    {
        StageTimer timer(kStageEout, st);
        plane_eout_kernel<Cfg><<<dim3(g.n_units, n_img), kEoutThreads, 0, st>>>(p);
    }
    auto k = ssg_plane_fwd_kernel<Cfg>;
    SSLB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        StageTimer timer(kStagePlaneFwd, st);
        k<<<dim3(Cfg::NDXG, tiles, n_img), Cfg::THREADS, smem, st>>>(tmap, p);
    }
    return check_launch("plane_forward", 2);
}

// ---- whole plane step ---------------------------------------------------------------------
// Workspace: lists | padded SR, GT | qT[SR] (becomes dL/dq) | qT[GT] | eout[2] | gcls | wtab | scratch | bwd lists | gpart
struct PlaneStepLayout {
    PlaneGeom g;
    int cap;
    PlaneListsLayout lists;
    PadLayout pad;
    size_t off_pad, off_q[2], off_eout[2], off_gcls, off_wtab, off_scratch, off_tcols, off_tcum, off_tent, off_gpart,
        total;
    int ntyb, ntxb, HT, WT, n_btiles, loss_blocks;
};

template <typename Cfg>
inline PlaneStepLayout plane_step_layout(int B, int H, int W, int max_edges, int loss_blocks, bool want_grad) {
    using BG = PlaneBwdGeom<Cfg>;
    using BC = PlaneBwdCfg<BG>;
    PlaneStepLayout l;
    l.g = geom_for<Cfg>(B, H, W);
    l.cap = slot_capacity(max_edges, l.g.n_units);
    l.lists = plane_lists_layout(l.g, l.cap);
    l.pad = pad_layout(B, H, W, Cfg::P);
    l.loss_blocks = loss_blocks;
    const int Hp = H + 2 * Cfg::P, Wp = W + 2 * Cfg::P;
    l.ntyb = (Hp + BG::ROWS - 1) / BG::ROWS;
    l.ntxb = (Wp + BC::TXB - 1) / BC::TXB;
    l.HT = l.ntyb * BG::ROWS;
    l.WT = l.ntxb * BC::TXB;
    l.n_btiles = B * l.ntyb * l.ntxb;
    const size_t nc2 = (size_t)Cfg::NCLS * Cfg::NCLS;
    size_t o = l.lists.total;
    // the two padded image sets must be contiguous (one tensor map spans both): unrounded size, rounded end
    l.off_pad = o; o += align256(2 * l.pad.bytes_per_image_set);
    for (int i = 0; i < 2; ++i) { l.off_q[i] = o; o += align256((size_t)Cfg::L * l.cap * sizeof(float)); }
    for (int i = 0; i < 2; ++i) { l.off_eout[i] = o; o += align256((size_t)l.cap * nc2 * sizeof(float)); }
    l.off_gcls = o; o += align256((size_t)l.cap * nc2 * sizeof(float));
    l.off_wtab = o; o += align256((size_t)l.cap * Cfg::KW * Cfg::KW * sizeof(float));
    l.off_scratch = o; o += align256((size_t)2 * loss_blocks * sizeof(double));
    l.off_tcols = o; l.off_tcum = o; l.off_tent = o; l.off_gpart = o;
    if (want_grad) {
        o += align256((size_t)l.n_btiles * (BC::RCOLS + 1) * sizeof(int32_t));
        l.off_tcum = o; o += align256((size_t)l.n_btiles * BC::RCOLS * BC::CUM_PITCH);
        l.off_tent = o; o += align256((size_t)l.n_btiles * BC::LIST_STRIDE * sizeof(int32_t));
        l.off_gpart = o; o += align256((size_t)BG::NDXG * B * 3 * l.HT * l.WT * sizeof(float));
    }
    l.total = o;
    return l;
}

// dL/dq (offset-major, gqT) + per-class sums (gcls; NULL when wtab is already in the workspace) -> dL/dimage.  Shared by the fused step and by the
// rows backward of the operator API.  `pad` = reflect-padded fp32 image the gradient is taken for.
// grad_out == NULL: stop after the fold -- the padded-domain gradient stays in part 0 of gpart (ws + l.off_gpart).
template <typename Cfg>
inline PlaneBwdParams plane_bwd_params(int B, int H, int W, const PlaneLists& lists, const PlaneStepLayout& l, char* ws,
                                       const float* gqT) {
    PlaneBwdParams bp{};
    bp.gqT = gqT;
    bp.tile_cols = reinterpret_cast<int32_t*>(ws + l.off_tcols);
    bp.tile_cum = reinterpret_cast<uint8_t*>(ws + l.off_tcum);
    bp.tile_ent = reinterpret_cast<int32_t*>(ws + l.off_tent);
    bp.gpart = reinterpret_cast<float*>(ws + l.off_gpart);
    bp.slot_map = lists.slot_map;
    bp.B = B; bp.H = H; bp.W = W; bp.cap = l.cap;
    bp.ntyb = l.ntyb; bp.ntxb = l.ntxb; bp.HT = l.HT; bp.WT = l.WT;
    return bp;
}

// Per-tile column lists of the backward: a function of the slot lists only (not of dL/dq), so the fused step
// builds them on its side lane while the forward runs.
template <typename Cfg>
inline int launch_plane_bwd_lists_cfg(int B, int H, int W, const PlaneLists& lists, const PlaneStepLayout& l, char* ws,
                                      cudaStream_t st) {
    using BG = PlaneBwdGeom<Cfg>;
    const PlaneBwdParams bp = plane_bwd_params<Cfg>(B, H, W, lists, l, ws, nullptr);
    {
        StageTimer timer(kStagePlaneBwdLists, st);
        plane_bwd_lists_kernel<BG><<<l.n_btiles, 256, 0, st>>>(bp, const_cast<int32_t*>(bp.tile_cols),
                                                               const_cast<uint8_t*>(bp.tile_cum),
                                                               const_cast<int32_t*>(bp.tile_ent));
    }
    return check_launch("plane_bwd_lists");
}

template <typename Cfg>
inline int launch_plane_backward_cfg(const float* pad, int B, int H, int W, const PlaneLists& lists,
                                     const PlaneStepLayout& l, char* ws, const float* gqT, const float* gcls,
                                     float* grad_out, cudaStream_t st, bool tile_lists_built = false) {
    using BG = PlaneBwdGeom<Cfg>;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PlaneBwdParams bp = plane_bwd_params<Cfg>(B, H, W, lists, l, ws, gqT);
    if (!tile_lists_built)
        if (int e = launch_plane_bwd_lists_cfg<Cfg>(B, H, W, lists, l, ws, st)) return e;
    const size_t smem = plane_bwd_smem_bytes<BG>();
    SSLB_REQUIRE(smem <= (size_t)di.max_smem_optin, "plane backward needs %zu B of shared memory", smem);
    CUtensorMap tmap;
    if (int e = make_plane_map(&tmap, pad, B * 3, l.pad.Hp, l.pad.Wp, l.pad.pitch, BG::IPITCH, BG::IROWS, 3)) return e;
    PlaneFinishParams fp{};
    fp.pad = pad; fp.Hp = l.pad.Hp; fp.pitch = l.pad.pitch;
    fp.gpart = bp.gpart; fp.wtab = reinterpret_cast<float*>(ws + l.off_wtab);
    fp.slot_map = lists.slot_map; fp.grad = grad_out;
    fp.B = B; fp.H = H; fp.W = W; fp.HT = l.HT; fp.WT = l.WT; fp.n_parts = BG::NDXG; fp.cap = l.cap;
    const long long npx = (long long)B * H * W;
    auto k = ssg_plane_bwd_kernel<BG>;
    SSLB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    {
        StageTimer timer(kStagePlaneBwd, st);
        k<<<dim3(l.n_btiles, BG::NDXG), BG::THREADS, smem, st>>>(tmap, bp);
    }
    {
        StageTimer timer(kStageFinish, st);
        if (gcls)   // the fused step's row loss has written wtab already
            plane_wtab_kernel<Cfg><<<di.sm_count * 8, 128, 0, st>>>(gcls, lists.counts, l.cap,
                                                                   reinterpret_cast<float*>(ws + l.off_wtab));
        plane_fold_kernel<Cfg, BG::NDXG><<<dim3(l.WT / 16, l.HT / 16, B), 256, 0, st>>>(fp);
        if (grad_out) plane_finish_kernel<Cfg><<<(int)((npx + 255) / 256), 256, 0, st>>>(fp);
    }
    return check_launch("plane_backward", 5);
}

// Lists for the step: straight from the mask when there is one, from the flat edge list otherwise.
struct StepInputs {
    const void* sr; const void* gt; int dtype_sr, dtype_gt;
    const float* mask; int mask_channels, mask_stride;     // mask == NULL: use edges
    const int32_t* edges; const int32_t* n_edges_dev;
    cudaEvent_t gt_ready;   // optional: GT is still arriving (host entry); the SR half of the forward runs before it
};

template <typename Cfg>
inline int launch_plane_step_cfg(const StepInputs& in, int B, int H, int W, int max_edges, float sigma, float eps,
                                 int rows_mode, float w_l1, float w_kl, float* grad_sr, double* terms, void* workspace,
                                 size_t workspace_bytes, cudaStream_t st) {
    using BG = PlaneBwdGeom<Cfg>;
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const int loss_blocks = 2 * di.sm_count;   // two persistent blocks per SM
    const PlaneStepLayout l = plane_step_layout<Cfg>(B, H, W, max_edges, 2 * di.sm_count, grad_sr != nullptr);
    SSLB_REQUIRE(workspace_bytes >= l.total, "workspace too small (%zu < %zu)", workspace_bytes, l.total);
    char* ws = static_cast<char*>(workspace);
    float* pad = reinterpret_cast<float*>(ws + l.off_pad);
    float* pad_gt = pad + l.pad.bytes_per_image_set / sizeof(float);
    // Side lane: the slot lists (next to the padding) and then the backward's tile lists (next to the forward) --
    // both depend on the mask only.  join[0] = slot lists ready, join[1] = tile lists ready.
    SideLane* side = nullptr;
    if (int e = side_lane(&side)) return e;
    SSLB_CUDA(cudaEventRecord(side->fork, st));
    SSLB_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
    if (int e = launch_plane_lists(in.mask, in.mask_channels, in.mask_stride, in.edges, in.n_edges_dev, max_edges, l.g,
                                   l.cap, ws, side->stream, Cfg::SRP, 32 / Cfg::G)) return e;
    SSLB_CUDA(cudaEventRecord(side->join[0], side->stream));
    const PlaneLists lists = carve_lists(ws, l.lists, nullptr);
    if (grad_sr)
        if (int e = launch_plane_bwd_lists_cfg<Cfg>(B, H, W, lists, l, ws, side->stream)) return e;
    SSLB_CUDA(cudaEventRecord(side->join[1], side->stream));
    if (in.gt_ready) {
        if (int e = launch_pad(in.sr, in.dtype_sr, nullptr, in.dtype_sr, B, H, W, Cfg::P, pad, st)) return e;
    } else {
        if (int e = launch_pad(in.sr, in.dtype_sr, in.gt, in.dtype_gt, B, H, W, Cfg::P, pad, st)) return e;
    }
    SSLB_CUDA(cudaStreamWaitEvent(st, side->join[0], 0));
    if (in.mask) {
        plane_terms_count_kernel<<<1, 1, 0, st>>>(lists.counts, max_edges, l.cap, terms);
        if (int e = check_launch("plane_terms_count")) return e;
    }
    float* q_sr = reinterpret_cast<float*>(ws + l.off_q[0]);
    float* q_gt = reinterpret_cast<float*>(ws + l.off_q[1]);
    float* eout_sr = reinterpret_cast<float*>(ws + l.off_eout[0]);
    float* eout_gt = reinterpret_cast<float*>(ws + l.off_eout[1]);
    if (in.gt_ready) {
        // SR half first; the GT half as soon as the caller's copy of GT has landed
        if (int e = launch_plane_forward_cfg<Cfg>(pad, 1, l.g, lists, l.cap, q_sr, nullptr, eout_sr, nullptr, st, 0, 2)) return e;
        SSLB_CUDA(cudaStreamWaitEvent(st, in.gt_ready, 0));
        if (int e = launch_pad(in.gt, in.dtype_gt, nullptr, in.dtype_gt, B, H, W, Cfg::P, pad_gt, st)) return e;
        if (int e = launch_plane_forward_cfg<Cfg>(pad, 1, l.g, lists, l.cap, q_gt, nullptr, eout_gt, nullptr, st, 1, 2)) return e;
    } else {
        if (int e = launch_plane_forward_cfg<Cfg>(pad, 2, l.g, lists, l.cap, q_sr, q_gt, eout_sr, eout_gt, st)) return e;
    }
    // rows -> loss terms and dL/dq (in place over q_sr) + per-class sums
    RowLossTParams rp{};
    rp.qs = q_sr; rp.qg = q_gt; rp.slot_pix = lists.slot_pix; rp.counts = lists.counts;
    rp.cap = l.cap;
    rp.denom = 3.f * (float)(Cfg::KW * Cfg::KW); rp.sigma = sigma; rp.eps = eps;
    rp.chain = -1.0f / (sigma * 3.f * (float)(Cfg::KW * Cfg::KW));
    rp.mode = rows_mode; rp.want_grad = grad_sr ? 1 : 0; rp.w_l1 = w_l1; rp.w_kl = w_kl;
    rp.wtab = grad_sr ? reinterpret_cast<float*>(ws + l.off_wtab) : nullptr;
    rp.scratch = reinterpret_cast<double*>(ws + l.off_scratch);
    rp.terms = terms;
    rp.done = reinterpret_cast<unsigned int*>(lists.counts + 4);   // zeroed with the counts, reset by the kernel
    const size_t rl_smem = (size_t)Cfg::L * kRowTSlots * sizeof(float);
    auto rl_kernel = w_kl != 0.f ? row_loss_t_kernel<Cfg::KS, Cfg::KW, true> : row_loss_t_kernel<Cfg::KS, Cfg::KW, false>;
    SSLB_CUDA(cudaFuncSetAttribute(rl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rl_smem));  // + 21 KB static
    {
        StageTimer timer(kStageRowLoss, st);
        rl_kernel<<<loss_blocks, kRowTThreads, rl_smem, st>>>(rp);
    }
    if (int e = check_launch("row_loss_t", 1)) return e;
    SSLB_CUDA(cudaStreamWaitEvent(st, side->join[1], 0));   // always: the side lane rejoins the caller's stream
    if (!grad_sr) return 0;
    return launch_plane_backward_cfg<Cfg>(pad, B, H, W, lists, l, ws, q_sr, nullptr, grad_sr, st, true);
}

// dL/dq rows in the order of `edges` [n][L] -> padded-domain (grad == NULL, result in part 0 of gpart) or image-domain
// gradient, through the plane kernels.  `pad` holds the padded image already; workspace = plane_step_layout(...).
template <typename Cfg>
inline int launch_plane_rows_backward_padded_cfg(const float* pad, int B, int H, int W, const int32_t* edges,
                                                 const int32_t* counts, int max_edges, const float* gq_rows,
                                                 const PlaneStepLayout& l, char* ws, cudaStream_t st, float* grad = nullptr) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    if (int e = launch_plane_lists(nullptr, 1, 0, edges, counts, max_edges, l.g, l.cap, ws, st, Cfg::SRP, 32 / Cfg::G)) return e;
    const PlaneLists lists = carve_lists(ws, l.lists, nullptr);
    float* gqT = reinterpret_cast<float*>(ws + l.off_q[0]);
    float* gcls = reinterpret_cast<float*>(ws + l.off_gcls);
    int32_t* slot_ref = reinterpret_cast<int32_t*>(ws + l.off_q[1]);  // the GT rows buffer is free here
    {
        StageTimer timer(kStageRowLoss, st);
        PlaneListParams lp{};
        lp.edges = edges; lp.n_edges_dev = counts; lp.max_edges = max_edges; lp.capacity = l.cap; lp.out = lists;
        const int nb = (max_edges + 255) / 256 < 2048 ? (max_edges + 255) / 256 : 2048;
        plane_slot_ref_kernel<<<nb, 256, 0, st>>>(lp, slot_ref);
        plane_rows_to_slots_kernel<<<di.sm_count * 4, 256, 0, st>>>(gq_rows, lists.slot_pix, slot_ref, lists.counts, l.cap,
                                                                   Cfg::L, gqT);
        plane_class_sums_kernel<<<di.sm_count * 8, 256, 0, st>>>(gqT, lists.counts, l.cap, Cfg::KS, Cfg::P, Cfg::K, gcls);
    }
    if (int e = check_launch("plane_rows_to_slots", 3)) return e;
    return launch_plane_backward_cfg<Cfg>(pad, B, H, W, lists, l, ws, gqT, gcls, grad, st);
}

// Rows backward behind the reference's operator API (similarity_map / compute_similarity): dL/dq rows in the
// reference's order [n][L] -> dL/dimage, through the plane kernels.  Workspace = plane_step_layout(...).
template <typename Cfg>
inline int launch_plane_rows_backward_cfg(const void* img, int dtype, int B, int H, int W, const int32_t* edges,
                                          const int32_t* counts, int max_edges, const float* gq_rows, float* grad,
                                          void* workspace, size_t workspace_bytes, cudaStream_t st) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    const PlaneStepLayout l = plane_step_layout<Cfg>(B, H, W, max_edges, 2 * di.sm_count, true);
    SSLB_REQUIRE(workspace_bytes >= l.total, "workspace too small (%zu < %zu)", workspace_bytes, l.total);
    char* ws = static_cast<char*>(workspace);
    float* pad = reinterpret_cast<float*>(ws + l.off_pad);
    if (int e = launch_pad(img, dtype, nullptr, dtype, B, H, W, Cfg::P, pad, st)) return e;
    return launch_plane_rows_backward_padded_cfg<Cfg>(pad, B, H, W, edges, counts, max_edges, gq_rows, l, ws, st, grad);
}

}  // namespace sslb
