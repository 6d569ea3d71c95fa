// Geometry shared by the "plane" (tile-sharing) kernels: ssg_plane_fwd.cuh / ssg_plane_bwd.cuh.
//
// The point kernels (ssg_point.cuh) spend C*kw^2 (subtract, fma) pairs per (edge pixel, search
// offset).  The plane kernels share that work between neighbouring edge pixels: for one search
// offset d = (dy,dx) the squared-difference plane
//     D_d(Y,X) = sum_c ( I(Y,X) - I(Y+dy,X+dx) )^2                      (reflect-padded coordinates)
// is computed once per image tile, box-summed separably, and read by every edge pixel of the tile:
//     q(p,d) = sum_{a in A(dy)} sum_{b in A(dx)} D_d(p + (a,b))  +  Eout(p,d)
// with A(t) = [max(-K,-P-t), min(K,P-t)] the part of the window whose neighbour stays inside the
// search area (the zero-padded unfold of loss_util.py:208-209 == the bounds test of
// similarity.cu:43) and Eout the remaining terms, in which the neighbour counts as zero
// (similarity.cu:46-47).  tests/dense_model.py is the NumPy statement of the same algebra.
#pragma once

#include "common.cuh"

namespace sslb {

// Compile-time geometry of one (k_search, k_window) configuration.
// ROWS = image rows one worker covers (its threads), G = consecutive dx per sweep thread, NWP = workers
// per CTA (worker w takes dy = w, w+NWP, ...).  The forward uses (64, 4, 5): it needs a halo of K rows
// on each side, so tall workers waste less; the backward has no row halo and uses (32, 4, 12): 12 warps are
// three per scheduler, which leaves each thread 168 registers (a thirteenth warp would cap all of them at 128).
// MERGE: a remainder of one dx (25 = 6*4 + 1) is folded into the last group instead of getting a
// one-plane group of its own, so the dx-groups of k_s = 25 are {4,4,4,4,4,5}.
template <int KS_, int KW_, int ROWS_ = 64, int G_ = 4, int NWP_ = 5, int TX_ = 64, bool MERGE_ = true>
struct PlaneCfg {
    static constexpr int KS = KS_, KW = KW_;
    static constexpr int P = KS / 2, K = KW / 2;
    static constexpr int L = KS * KS;
    static constexpr int G = G_;           // consecutive dx handled by one sweep thread (nominal group width)
    static constexpr int NWP = NWP_;       // workers per CTA
    static constexpr bool MERGE = MERGE_ && KS >= G && (KS % G) <= 1;
    static constexpr int NDXG = MERGE ? KS / G : (KS + G - 1) / G;
    static constexpr int GMAX = MERGE ? G + KS % G : G;   // widest group
    static constexpr int NPL = GMAX * NWP;  // box-summed planes resident per CTA
    static constexpr int ROWS = ROWS_;     // threads of a worker = image rows it covers (incl. halo in the forward)
    static constexpr int THREADS = NWP * ROWS;
    static constexpr int NCLS = 2 * K + 1; // clip classes per axis
    // forward tiles (unpadded image coordinates of the edge pixels they own)
    static constexpr int TYF = ROWS - 2 * K;
    static constexpr int TXF = TX_;        // tile width (a multiple of the 8-column chunk)
    static constexpr int CH = 8;           // columns per sweep chunk
    static constexpr int UNITS_X = TXF / CH;
    static constexpr int SWEEP = TXF + CH;   // one chunk of halo: 2K <= CH columns
    static constexpr int NCH = SWEEP / CH;
    // shared image tile: row 0 <-> Y = Ytile0 - K - P, column ICOL0 <-> X = Xtile0 - K
    static constexpr int IROWS = ROWS + 2 * P;
    static constexpr int ICOL0 = ((P + 3) / 4) * 4;
    static constexpr int ICOLS = ICOL0 + SWEEP + P + 4;
    static constexpr int IPITCH = (((ICOLS + 3) / 4) | 1) * 4;  // 4 * odd: conflict-free float4 rows
    static constexpr int TILE_FLOATS = 3 * IROWS * IPITCH;      // one TMA box {IPITCH, IROWS, 3}
    // shared ring of box-summed planes
    static constexpr int RING = 16;
    static constexpr int SRP = RING + 1;                 // odd row pitch
    static constexpr int SPS = (ROWS * SRP) | 1;         // odd plane stride
    static constexpr int RC_SMEM = 2048;                 // slot descriptors of a tile staged in shared memory
    static_assert(KS % 2 == 1 && KW % 2 == 1 && KW <= KS && KW <= 9, "unsupported kernel sizes");
    static_assert(ROWS == 32 || ROWS == 64, "a worker is one warp or a warp pair");
    static_assert(2 * K <= CH, "gather lags the sweep by one chunk");
    static_assert(IROWS <= 256 && IPITCH <= 256, "image tile must be one TMA box");
};

// Backward geometry that goes with a forward configuration.
template <typename Cfg>
using PlaneBwdGeom = PlaneCfg<Cfg::KS, Cfg::KW, 32, 4, 12, 96, false>;

// in-area range of window offsets for search offset t, per axis
__host__ __device__ constexpr int rng_lo(int t, int P, int K) { return -P - t > -K ? -P - t : -K; }
__host__ __device__ constexpr int rng_hi(int t, int P, int K) { return P - t < K ? P - t : K; }
// clip class of t: 0..K-1 clipped below (lo = -cls), K unclipped, K+1..2K clipped above (hi = 2K-cls)
__host__ __device__ constexpr int clip_class(int t, int P, int K) {
    return t < -(P - K) ? t + P : (t > P - K ? t - (P - K) + K : K);
}
__host__ __device__ constexpr int class_lo(int cls, int K) { return cls < K ? -cls : -K; }
__host__ __device__ constexpr int class_hi(int cls, int K) { return cls > K ? 2 * K - cls : K; }

// Layout of the rows buffers (q of SR, q of GT, dL/dq): panels of kPanel = 16 consecutive slots, each panel holding
// all L offsets of its slots contiguously:
//     element (offset d, slot s)  at  ((s / 16) * L + d) * 16 + s % 16
// A panel is the unit of work of the row-loss kernel (L * 64 bytes, one contiguous stream), the 4 results a forward
// gather thread produces for one offset are still one aligned float4, and the G planes of a dx-group (consecutive
// offsets) of one slot group are G consecutive 64-byte segments.  Capacities are multiples of 16.
constexpr int kPanel = 16;
__host__ __device__ __forceinline__ long long qt_index(int d, int slot, int L) {
    return ((long long)(slot >> 4) * L + d) * kPanel + (slot & (kPanel - 1));
}

// Edge pixels regrouped for the plane kernels.  A *unit* is an 8-column strip of a forward tile;
// its edge pixels occupy consecutive *slots* (row-major inside the unit, count padded to a multiple
// of 4 with empty slots) so that the 4 results a lane produces for one search offset are one
// aligned float4 of the rows buffer (panel layout, qt_index).
struct PlaneLists {
    int32_t* unit_start;   // [n_units + 1] first slot of each unit (exclusive scan of padded counts)
    int32_t* slot_pix;     // [capacity] flat pixel index b*H*W + y*W + x, or -1 (padding)
    int32_t* slot_rc;      // [capacity] (row lane << 8) | column in tile, or -1
    int32_t* slot_map;     // [B*H*W] slot of each pixel, or -1
    int32_t* counts;       // [4]: slots used, edges found, slots needed (uncapped), n_units
};

struct PlaneGeom {
    int B, H, W;
    int TYF, TXF, K;
    int xs;                // forward tiles start at image column -xs (see make_geom)
    int nty, ntx;          // forward tiles per image
    int n_units;
};

// Forward tile tx owns image columns [tx*TXF - xs, (tx+1)*TXF - xs).  The shift xs = (P - K) mod 4 makes the
// first column of every tile's shared image window -- padded X = P - K - ICOL0 - xs + tx*TXF -- a multiple
// of 4 floats: the copy engine wants the start of a box 16-byte aligned in global memory, and the kernels
// want the window's own columns float4-aligned in shared memory (ICOL0 is a multiple of 4).
inline PlaneGeom make_geom(int B, int H, int W, int TYF, int TXF, int K, int P) {
    PlaneGeom g;
    g.B = B; g.H = H; g.W = W; g.TYF = TYF; g.TXF = TXF; g.K = K;
    g.xs = (P - K) % 4;
    g.nty = (H + TYF - 1) / TYF;
    g.ntx = (W + g.xs + TXF - 1) / TXF;
    g.n_units = B * g.nty * g.ntx * (TXF / 8);
    return g;
}

struct PlaneListParams {
    const float* mask;         // [B,mask_channels,H,W], or NULL: take the pixels of `edges`
    int mask_channels, stride;
    const int32_t* edges;      // flat pixel indices (ssl_b200_build_edge_list order)
    const int32_t* n_edges_dev;
    int max_edges;
    PlaneGeom g;
    int capacity;
    int32_t* unit_count;   // [n_units] padded counts (scratch)
    PlaneLists out;
};

__device__ __forceinline__ bool mask_is_edge(const float* mask, int mask_channels, int stride, int H, int W, int b,
                                             int y, int x) {
    const float v = __ldg(mask + ((long long)b * mask_channels * H + y) * W + x);
    if (v != 1.0f) return false;  // exact compare, like the reference (loss_util.py:196)
    return stride <= 1 || (y % stride) == (x % stride);
}

// The pixels of a flat edge list are marked -2 in slot_map before the unit passes run.
__global__ void __launch_bounds__(256) plane_mark_edges_kernel(PlaneListParams p) {
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mc; i += gridDim.x * blockDim.x)
        if (p.edges[i] >= 0) p.out.slot_map[p.edges[i]] = -2;
}

// slot_ref[slot] = position of the slot's pixel in the flat edge list (after the unit passes)
__global__ void __launch_bounds__(256) plane_slot_ref_kernel(PlaneListParams p, int32_t* slot_ref) {
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < mc; i += gridDim.x * blockDim.x) {
        if (p.edges[i] < 0) continue;
        const int slot = p.out.slot_map[p.edges[i]];
        if (slot >= 0 && slot < p.capacity) slot_ref[slot] = i;
    }
}

__device__ __forceinline__ bool unit_is_edge(const PlaneListParams& p, int b, int y, int x) {
    if (p.mask) return mask_is_edge(p.mask, p.mask_channels, p.stride, p.g.H, p.g.W, b, y, x);
    return p.out.slot_map[(b * p.g.H + y) * p.g.W + x] != -1;
}

// unit -> (b, ty, tx, cx)
__device__ __forceinline__ void decode_unit(const PlaneGeom& g, int u, int& b, int& ty, int& tx, int& cx) {
    const int ux = g.TXF / 8;
    cx = u % ux; u /= ux;
    tx = u % g.ntx; u /= g.ntx;
    ty = u % g.nty;
    b = u / g.nty;
}

// Count pass, one warp per unit: the unit is walked 8 columns x 4 rows per warp step.
__global__ void __launch_bounds__(256) plane_units_count_kernel(PlaneListParams p) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= p.g.n_units) return;
    int b, ty, tx, cx;
    decode_unit(p.g, warp, b, ty, tx, cx);
    const int y0 = ty * p.g.TYF, x0 = tx * p.g.TXF + cx * 8 - p.g.xs;
    const int ly = lane >> 3, lx = lane & 7;
    int n = 0;
    for (int yy = 0; yy < p.g.TYF; yy += 4) {
        const int y = y0 + yy + ly, x = x0 + lx;
        const bool e = (yy + ly) < p.g.TYF && y < p.g.H && x >= 0 && x < p.g.W && unit_is_edge(p, b, y, x);
        n += __popc(__ballot_sync(0xffffffffu, e));
    }
    if (lane == 0) {
        p.unit_count[warp] = (n + 3) & ~3;   // padded to whole float4 groups
        if (n) atomicAdd(p.out.counts + 1, n);
    }
}

// Emit pass (after the scan), one warp per unit.  The order of a unit's edge pixels inside its slot range is
// free (slot_map / slot_pix carry the correspondence), so it is chosen for the forward gather: there a warp reads,
// in one shared-memory instruction, the same ring position of G planes for 32/G consecutive 4-slot groups, and
// the bank of a slot is h = (SRP * row + column) mod 32.  Slots are counting-sorted by h and dealt out so
// that consecutive groups get ranks n/(32/G) apart, i.e. banks ~G apart: their G-bank runs do not overlap
// (kSpreadRows = 32/G).
constexpr int kUnitMaxEdges = 512;

__global__ void __launch_bounds__(256) plane_units_emit_kernel(PlaneListParams p, int srp, int kSpreadRows) {
    __shared__ int32_t s_rc[8][kUnitMaxEdges], s_pix[8][kUnitMaxEdges];
    __shared__ uint8_t s_h[8][kUnitMaxEdges];
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int warp = blockIdx.x * 8 + wib;
    if (warp >= p.g.n_units) return;
    int b, ty, tx, cx;
    decode_unit(p.g, warp, b, ty, tx, cx);
    const int y0 = ty * p.g.TYF, x0 = tx * p.g.TXF + cx * 8 - p.g.xs;
    const int ly = lane >> 3, lx = lane & 7;
    // 1. collect the unit's edge pixels in row-major order
    int n = 0;
    for (int yy = 0; yy < p.g.TYF; yy += 4) {
        const int y = y0 + yy + ly, x = x0 + lx;
        const bool e = (yy + ly) < p.g.TYF && y < p.g.H && x >= 0 && x < p.g.W && unit_is_edge(p, b, y, x);
        const unsigned ball = __ballot_sync(0xffffffffu, e);
        if (e) {
            const int t = n + __popc(ball & ((1u << lane) - 1u));
            const int re = yy + ly + p.g.K, ex = cx * 8 + lx;
            s_rc[wib][t] = (re << 8) | ex;
            s_pix[wib][t] = (b * p.g.H + y) * p.g.W + x;
            // sort key: the shared-memory bank of the slot in the forward ring, or (kSpreadRows == 0) the column
            s_h[wib][t] = kSpreadRows > 0 ? (uint8_t)((srp * re + ex) & 31) : (uint8_t)lx;
        }
        n += __popc(ball);
    }
    __syncwarp();
    const int start = p.out.unit_start[warp], end = p.out.unit_start[warp + 1];
    for (int s = start + lane; s < end; s += 32)
        if (s < p.capacity) { p.out.slot_pix[s] = -1; p.out.slot_rc[s] = -1; }
    __syncwarp();
    if (n == 0) return;
    // 2. counting sort by bank: lane = bank value; stable in row-major order
    int cnt = 0;
    for (int t = 0; t < n; ++t) cnt += s_h[wib][t] == lane;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    int rank = incl - cnt;
    // 3. rank -> position: ranks laid out in kSpreadRows rows of m, read column by column, dealt to the
    //    4-slot groups round-robin (position = 4 * group + index in group)
    const int ksr = kSpreadRows > 0 ? kSpreadRows : 1;
    const int m = (n + ksr - 1) / ksr, nfull = n / m, part = n % m;
    const int gtot = ((n + 3) & ~3) / 4;
    int rr = rank / m, cc = rank - rr * m;   // kept incrementally below: one division per lane, none per edge pixel
    for (int t = 0; t < n; ++t) {
        if (s_h[wib][t] != lane) continue;
        const int f = cc * nfull + (cc < part ? cc : part) + rr;
        // f < 4 * gtot: quotient and remainder by comparisons
        const int fq = (f >= gtot) + (f >= 2 * gtot) + (f >= 3 * gtot), fr = f - fq * gtot;
        // kSpreadRows == 0: column-major order inside the unit (the entries of an image column are consecutive slots)
        const int slot = kSpreadRows > 0 ? start + 4 * fr + fq : start + rank;
        ++rank;
        if (++cc == m) { cc = 0; ++rr; }
        if (slot < p.capacity) {
            p.out.slot_pix[slot] = s_pix[wib][t];
            p.out.slot_rc[slot] = s_rc[wib][t];
            p.out.slot_map[s_pix[wib][t]] = slot;
        } else {
            p.out.slot_map[s_pix[wib][t]] = -1;
        }
    }
}

// Single block: exclusive scan of the padded unit counts.
__global__ void __launch_bounds__(1024) plane_units_scan_kernel(PlaneListParams p) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < p.g.n_units; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = i < p.g.n_units ? p.unit_count[i] : 0;
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) warp_tot[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < p.g.n_units) p.out.unit_start[i] = carry + (warp ? warp_tot[warp - 1] : 0) + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int total = carry_s;
        p.out.unit_start[p.g.n_units] = total;
        p.out.counts[0] = total < p.capacity ? total : p.capacity;
        p.out.counts[2] = total;
        p.out.counts[3] = p.g.n_units;
    }
}

// terms[2] = number of SSG rows of a step whose lists were built straight from the mask.  More edge pixels than
// the caller's max_edges, or more slots than the workspace holds, would drop edge pixels: the count is poisoned
// instead, so the loss and the gradient scale come out NaN without any host round trip.
__global__ void plane_terms_count_kernel(const int32_t* counts, int max_edges, int cap, double* terms) {
    const bool overflow = counts[1] > max_edges || counts[2] > cap;
    terms[2] = overflow ? __longlong_as_double(0x7ff8000000000000ll) : (double)counts[1];
}

__global__ void __launch_bounds__(256) fill_i32_kernel(int32_t* p, long long n, int32_t v) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}

}  // namespace sslb
