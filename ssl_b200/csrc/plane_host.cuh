// Host-side launch code of the plane path (included by ssl_b200.cu only).
#pragma once

#include "plane_geom.cuh"
#include "row_loss_t.cuh"
#include "ssg_plane_fwd.cuh"

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
    c = (c + 3) & ~3ll;
    return (int)c;
}

// Carve the list arrays out of one workspace.  Layout is a function of (geometry, capacity) only.
struct PlaneListsLayout {
    size_t off_counts, off_unit_start, off_unit_count, off_slot_pix, off_slot_rc, off_slot_map, total;
};

inline PlaneListsLayout plane_lists_layout(const PlaneGeom& g, int cap) {
    PlaneListsLayout l;
    size_t o = 0;
    l.off_counts = o; o += align256(4 * sizeof(int32_t));
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
    return make_geom(B, H, W, Cfg::TYF, Cfg::TXF, Cfg::K);
}

// mask (or flat edge list when mask == NULL) -> unit lists
inline int launch_plane_lists(const float* mask, int mask_channels, int stride, const int32_t* edges,
                              const int32_t* n_edges_dev, int max_edges, const PlaneGeom& g, int cap, void* ws,
                              cudaStream_t st) {
    const PlaneListsLayout lay = plane_lists_layout(g, cap);
    PlaneListParams p{};
    p.mask = mask; p.mask_channels = mask_channels; p.stride = stride;
    p.edges = edges; p.n_edges_dev = n_edges_dev; p.max_edges = max_edges;
    p.g = g; p.capacity = cap;
    p.out = carve_lists(ws, lay, &p.unit_count);
    SSLB_CUDA(cudaMemsetAsync(p.out.counts, 0, 4 * sizeof(int32_t), st));
    const long long npx = (long long)g.B * g.H * g.W;
    const int fill_blocks = (int)((npx + 255) / 256 < 4096 ? (npx + 255) / 256 : 4096);
    fill_i32_kernel<<<fill_blocks, 256, 0, st>>>(p.out.slot_map, npx, -1);
    if (!mask && max_edges > 0) plane_mark_edges_kernel<<<(max_edges + 255) / 256 < 2048 ? (max_edges + 255) / 256 : 2048, 256, 0, st>>>(p);
    const int blocks = (g.n_units * 32 + 255) / 256;
    plane_units_kernel<0><<<blocks, 256, 0, st>>>(p);
    plane_units_scan_kernel<<<1, 1024, 0, st>>>(p);
    plane_units_kernel<1><<<blocks, 256, 0, st>>>(p);
    return check_launch("plane_lists", 5);
}

template <typename Cfg>
inline int launch_plane_forward_cfg(const void* img, const void* img2, int dtype, const PlaneGeom& g,
                                    const PlaneLists& lists, int cap, float* qT, float* qT2, float* eout,
                                    float* eout2, cudaStream_t st) {
    DeviceInfo di;
    if (int e = device_info(&di)) return e;
    PlaneFwdParams p{};
    p.img[0] = img; p.img[1] = img2;
    p.qT[0] = qT; p.qT[1] = qT2;
    p.eout[0] = eout; p.eout[1] = eout2;
    p.lists = lists; p.g = g; p.cap = cap;
    const int n_img = img2 ? 2 : 1;
    const size_t smem = plane_fwd_smem_bytes<Cfg>();
    SSLB_REQUIRE(smem <= (size_t)di.max_smem_optin, "plane forward needs %zu B of shared memory", smem);
    const int tiles = g.B * g.nty * g.ntx;
    SSLB_DISPATCH_DTYPE(dtype, T, {
        plane_eout_kernel<T, Cfg><<<dim3(di.sm_count * 8, n_img), 128, 0, st>>>(p);
        auto k = ssg_plane_fwd_kernel<T, Cfg>;
        SSLB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k<<<dim3(tiles, Cfg::NDXG, n_img), Cfg::THREADS, smem, st>>>(p);
    });
    return check_launch("plane_forward", 2);
}

}  // namespace sslb
