// Backward "plane" kernels: dL/dimage from dL/dq held in the rows buffer (panel layout, plane_geom.cuh: qt_index).
//
// Adjoint of ssg_plane_fwd.cuh, in gather form (no global atomics, fixed summation order).  For a
// search offset d the contributions of all edge pixels are first spread into a sparse plane
//     u_d = (gq(p,d) placed at p)  +  (gq(p,-d) placed at p-d)
// (v-direction applied while placing, h-direction by the same "sum of the last l" tree as the
// forward), which gives, for every pixel x of a tile,
//     Gs_d(x) = sum_{p: x in p+A(dy)xA(dx)} gq(p,d) + sum_{p: x+d in p+A(-dy)xA(-dx)} gq(p,-d)
// so that  dL/dIpad(x,c) = 2 sum_d ( I(x,c) - I(x+d,c) ) Gs_d(x)  + out-of-area terms
// (similarity.cu:73-131: +v into the centre pixel, -v into the neighbour pixel; the second sum is
// the neighbour role of x, re-indexed so that it is gathered at x too).  tests/dense_model.py holds
// the NumPy statement of these placement rules.
//
// One CTA = one 32 x TXB tile of the PADDED image (its image window arrives by one TMA tensor copy) and one
// dx-group; its NWP = 12 workers (single warps, lane = image row) take dy = w - P, w - P + NWP and one output chunk
// of the left-over dy, and free-run; the partial sums go into accumulator tiles held in tensor memory (TMEM, one
// per lane quarter), additions ordered by per-chunk tickets (deterministic summation order, no atomics on data).
#pragma once

#include <cuda/atomic>

#include "plane_geom.cuh"
#include "ssg_plane_fwd.cuh"

namespace sslb {

// ---- tensor memory (TMEM) as the accumulator store ---------------------------------------------
// Blackwell's 256 KB of tensor memory per SM (128 lanes x 512 32-bit columns) is reachable from ordinary warps
// with tcgen05.ld / tcgen05.st, on a datapath of its own: nothing of it goes through the shared-memory pipe that
// bounds this kernel.  A warp reaches the 32 lanes of its own quarter (warp id mod 4) -- with lane = image row,
// which is exactly how a worker holds its partial sums -- so the CTA keeps FOUR partial accumulator tiles, one per
// quarter, each shared by the three workers whose warp id falls into it; they are added up -- (q0 + q1) + (q2 + q3)
// -- once, at the end (all warps stage their share of the partials, one barrier, sum at the output store).  `#define SSLB_BWD_TMEM 0` builds the shared-memory accumulator of round 1 instead.
#ifndef SSLB_BWD_TMEM
#define SSLB_BWD_TMEM 1
#endif

__device__ __forceinline__ void tmem_alloc_512(uint32_t* smem_slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(smem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 8 consecutive columns of the calling thread's lane (lane = quarter base + lane id)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#ifdef SSLB_EXPERIMENT_NOGQ   // timing experiment only: how much of the kernel is the dL/dq gather?
#define SSLB_GQ(gq, packed) (1e-9f * (float)((packed) & 1023))
#else
#define SSLB_GQ(gq, packed) __ldg(gq_at<Cfg>(gq, packed))
#endif

template <typename Cfg>
struct PlaneBwdCfg {
    static constexpr int P = Cfg::P, K = Cfg::K, G = Cfg::G;
    static constexpr int TXB = Cfg::TXF;                 // same sweep geometry as the forward
    static constexpr int NCHB = TXB / 8 + 1;
    static constexpr int ACC_PITCH = TXB + 4;            // 4 * odd
    // Per-worker staging buffer of one chunk's sparse columns (G planes x 8 columns = 32 columns x ROWS rows),
    // used in two layouts one after the other:
    //   A[row][lane]      pitch 32: while the +val/-val events are entered, lane l owns column l and every one
    //                     of its addresses lies in bank l -- the updates cannot conflict whatever their rows are;
    //   B[row][column]    pitch 36 (4 * odd): the prefix-summed columns, read by the sweep (lane = row) as two
    //                     conflict-free float4 per plane.
    static constexpr int UB_PITCH = 36;
    static constexpr int U_WORKER = Cfg::ROWS * UB_PITCH;
    // edge-pixel region a tile needs: rows [Yb0-P, Yb0+ROWS+P), columns [Xb0-P-8, Xb0+TXB+P)
    static constexpr int RROWS = Cfg::ROWS + 2 * P;
    static constexpr int RCOLS = TXB + 2 * P + 8;
    static constexpr int LIST_STRIDE = RROWS * RCOLS;    // worst case entries per tile
    static constexpr int LIST_SMEM = 2560;               // entries staged in shared memory
    static constexpr int CUM_PITCH = (RROWS + 1 + 3) & ~3;  // per column: entries above each region row (uint8)
    // work distribution over the NWP workers (run_group_bwd)
    static constexpr int FULL_ROUNDS = Cfg::KS / Cfg::NWP;
    static constexpr bool SPLIT_LAST = Cfg::KS % Cfg::NWP == 1 && NCHB - 1 <= Cfg::NWP;
    static constexpr int N_ITEMS = SPLIT_LAST ? FULL_ROUNDS + 1 : (Cfg::KS + Cfg::NWP - 1) / Cfg::NWP;
#ifndef SSLB_NB
#define SSLB_NB 8
#endif
    static constexpr int NB = SSLB_NB;                   // dL/dq values per kind prefetched into registers
    static_assert((ACC_PITCH / 4) % 2 == 1, "accumulator rows must be float4 conflict-free");
    static_assert(G * 8 == 32 && Cfg::ROWS == 32, "one lane per (plane, u-column) when placing");
    static_assert(Cfg::NDXG <= 7, "dx-group dispatch");
    static_assert(RROWS < 256, "row counts are stored in a byte");
};

struct PlaneBwdParams {
    const float* gqT;         // L x cap, panel layout (qt_index)
    const int32_t* tile_cols; // [n_btiles][RCOLS+1] column starts inside the tile's entry list
    const uint8_t* tile_cum;  // [n_btiles][RCOLS][CUM_PITCH] entries of the column above region row rr
    const int32_t* tile_ent;  // [n_btiles][LIST_STRIDE] (slot << 8) | region row
    float* gpart;             // [NDXG][B][3][HT][WT] partial padded gradients
    const int32_t* slot_map;
    int B, H, W, cap;
    int ntyb, ntxb, HT, WT;
};

// ---- per-tile edge lists by column ---------------------------------------------------------
// One block per backward tile.  Entries of a column are in ascending row order; cum[col][rr] counts the
// entries of the column that lie above region row rr, so the entries whose run reaches the tile for a given
// (kind, dy) are the contiguous range [cum[rr_lo], cum[rr_hi + 1]).
template <typename Cfg>
__global__ void __launch_bounds__(256) plane_bwd_lists_kernel(PlaneBwdParams p, int32_t* tile_cols, uint8_t* tile_cum,
                                                              int32_t* tile_ent) {
    using BC = PlaneBwdCfg<Cfg>;
    constexpr int P = Cfg::P;
    __shared__ int cnt[BC::RCOLS + 1];
    const int t = blockIdx.x;
    const int txb = t % p.ntxb, tyb = (t / p.ntxb) % p.ntyb, b = t / (p.ntxb * p.ntyb);
    const int y0 = tyb * Cfg::ROWS - P - P;        // image row of region row 0 (padded Yb0-P, minus pad P)
    const int x0 = txb * BC::TXB - P - 8 - P;      // image column of region column 0
    for (int i = threadIdx.x; i <= BC::RCOLS; i += blockDim.x) cnt[i] = 0;
    __syncthreads();
    // one thread per region column: count, then (after the scan) emit in row order
    int my = 0;
    const int col = threadIdx.x;
    const int x = x0 + col;
    const bool colok = col < BC::RCOLS && x >= 0 && x < p.W;
    if (colok)
        for (int rr = 0; rr < BC::RROWS; ++rr) {
            const int y = y0 + rr;
            if (y >= 0 && y < p.H && p.slot_map[(b * p.H + y) * p.W + x] >= 0) ++my;
        }
    if (col < BC::RCOLS) cnt[col + 1] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int i = 1; i <= BC::RCOLS; ++i) { run += cnt[i]; cnt[i] = run; }
    }
    __syncthreads();
    int32_t* cols = tile_cols + (long long)t * (BC::RCOLS + 1);
    for (int i = threadIdx.x; i <= BC::RCOLS; i += blockDim.x) cols[i] = cnt[i];
    if (col < BC::RCOLS) {
        uint8_t* cum = tile_cum + ((long long)t * BC::RCOLS + col) * BC::CUM_PITCH;
        int32_t* ent = tile_ent + (long long)t * BC::LIST_STRIDE + cnt[col];
        int n = 0;
        for (int rr = 0; rr < BC::RROWS; ++rr) {
            cum[rr] = (uint8_t)n;
            const int y = y0 + rr;
            if (!colok || y < 0 || y >= p.H) continue;
            const int slot = p.slot_map[(b * p.H + y) * p.W + x];
            if (slot >= 0) { ent[n] = (slot << 8) | rr; ++n; }
        }
        for (int rr = BC::RROWS; rr < BC::CUM_PITCH; ++rr) cum[rr] = (uint8_t)n;
    }
}

// Building one column of one sparse plane (lane (plane pj, column c8) owns it) is split in two so that the
// loads of dL/dq can be issued one chunk ahead and fly while the sweep of the current chunk computes:
//     place_fetch (entry range + the first NB loads into registers)  ...sweep...  place_apply.
//
// Every edge pixel that lands in the column covers a run of rows [r0, r1] (the v-direction of the box);
// it is entered as +val at r0 and -val at r1+1 and the column is then prefix-summed, so the cost per edge
// pixel is two shared-memory updates instead of up to 2K+1.
template <int NB>
struct PlaceFetch {
    int packed[2][NB];
    float val[2][NB];
    int e0[2], e1[2];      // entries of the column whose run reaches the tile: [e0, e1); [e0, e0+NB) prefetched
};

// rows covered by an entry at region row rr: [rr + lo_off, rr + hi_off] in tile rows
template <typename Cfg>
__device__ __forceinline__ void place_offsets(int kind, int dy, int& lo_off, int& hi_off) {
    constexpr int P = Cfg::P, K = Cfg::K;
    const int alo = rng_lo(dy, P, K), ahi = rng_hi(dy, P, K);
    lo_off = kind == 0 ? -P + alo : -P - dy - ahi;
    hi_off = kind == 0 ? -P + ahi : -P - dy - alo;
}

// where an entry's value comes from: kind 0 reads offset d, kind 1 the mirrored offset -d.  Returns gqT advanced to
// the offset's 64-byte segment of panel 0; gq_at() adds the slot's panel and lane.
template <typename Cfg>
__device__ __forceinline__ const float* place_source(const PlaneBwdParams& p, int kind, int dy, int dx) {
    constexpr int P = Cfg::P;
    const int d = kind == 0 ? (dy + P) * Cfg::KS + dx + P : (-dy + P) * Cfg::KS + (-dx) + P;
    return p.gqT + d * kPanel;
}

template <typename Cfg>
__device__ __forceinline__ const float* gq_at(const float* gq, int packed) {
    const int slot = packed >> 8;
    return gq + (long long)(slot >> 4) * (Cfg::L * kPanel) + (slot & (kPanel - 1));
}

template <typename Cfg, int NB>
__device__ __forceinline__ void place_fetch(const PlaneBwdParams& p, const int32_t* cols, const uint8_t* cum,
                                            const int32_t* ent, int xu, int dy, int dx, PlaceFetch<NB>& f) {
    using BC = PlaneBwdCfg<Cfg>;
    constexpr int P = Cfg::P, K = Cfg::K;
    const int blo = rng_lo(dx, P, K), bhi = rng_hi(dx, P, K);
#pragma unroll
    for (int kind = 0; kind < 2; ++kind) {
        // region column holding the edge pixels that land in u-column xu
        const int pcol = kind == 0 ? xu - blo + P : xu + dx + bhi + P;
        int lo_off, hi_off;
        place_offsets<Cfg>(kind, dy, lo_off, hi_off);
        // region rows whose run meets tile rows [0, ROWS)
        const int rr_lo = max(0, -hi_off), rr_hi = min(BC::RROWS - 1, Cfg::ROWS - 1 - lo_off);
        // branch-free: the lists are read at a clamped column and the range is emptied afterwards (rr_lo and
        // rr_hi + 1 always lie inside the cum row: |offsets| <= 2P + K < RROWS - ROWS)
        const bool ok = pcol >= 0 && pcol < BC::RCOLS && rr_lo <= rr_hi;
        const int pc = min(max(pcol, 0), BC::RCOLS - 1);
        const int base = cols[pc];
        const uint8_t* cc = cum + pc * BC::CUM_PITCH;
        const int e0 = ok ? base + cc[rr_lo] : 0;
        const int e1 = ok ? base + cc[min(rr_hi, BC::RROWS - 1) + 1] : 0;
        f.e0[kind] = e0;
        f.e1[kind] = e1;
        const float* gq = place_source<Cfg>(p, kind, dy, dx);
#pragma unroll
        for (int m = 0; m < NB; ++m) {
            const bool on = e0 + m < e1;   // list entries are never negative: one predicate serves both loads
            f.packed[kind][m] = on ? ent[e0 + m] : -1;
            f.val[kind][m] = on ? SSLB_GQ(gq, f.packed[kind][m]) : 0.f;
        }
    }
}

// Enter NB edge pixels of ONE kind into the lane's column A[row * 32] (A already points at the lane's bank).
// Entries of a kind have distinct rows, so the NB "+val" addresses are distinct, and so are the NB "-val"
// addresses: each half is NB independent loads, NB adds, NB stores.  The two clamped cases stay out of shared
// memory: runs starting above the tile add into `head` (row 0, applied by the prefix pass), runs ending below
// it need no "-val" at all (nothing reads past the last row).
template <typename Cfg, int NB>
__device__ __forceinline__ void place_batch(float* A, const int (&packed)[NB], const float (&val)[NB], int lo_off,
                                            int hi_off, float& head) {
    // An empty slot (packed = -1) decodes to region row 255, far beyond every run that can meet the tile (region rows
    // are < RROWS <= 80 and the offsets lie in [-(2P + K), K]): all three range tests below fail for it by themselves.
    static_assert(PlaneBwdCfg<Cfg>::RROWS + 2 * Cfg::P + Cfg::K < 255 - Cfg::ROWS, "row 255 must stay out of range");
    int r0[NB], r1[NB];
    bool in0[NB], in1[NB];
    float cur[NB];
#pragma unroll
    for (int m = 0; m < NB; ++m) {
        const int rr = packed[m] & 255;
        r0[m] = rr + lo_off;
        r1[m] = rr + hi_off;
        in0[m] = (unsigned)(r0[m] - 1) < (unsigned)(Cfg::ROWS - 1);   // "+val" lands in rows 1 .. ROWS-1
        in1[m] = (unsigned)r1[m] < (unsigned)(Cfg::ROWS - 1);         // "-val" lands in rows 1 .. ROWS-1 (r1 + 1)
    }
#pragma unroll
    for (int m = 0; m < NB; ++m)
        if (in0[m]) cur[m] = A[r0[m] * 32];
#pragma unroll
    for (int m = 0; m < NB; ++m) {   // branch-free: val is zero for an empty slot
        if (in0[m]) A[r0[m] * 32] = cur[m] + val[m];
        head += r0[m] <= 0 ? val[m] : 0.f;
    }
#pragma unroll
    for (int m = 0; m < NB; ++m)
        if (in1[m]) cur[m] = A[(r1[m] + 1) * 32];
#pragma unroll
    for (int m = 0; m < NB; ++m)
        if (in1[m]) A[(r1[m] + 1) * 32] = cur[m] - val[m];
}

// All lanes of the worker call this (it contains warp barriers); lanes without a column (active == false:
// planes beyond the group's width) only take part in the barriers and the clearing.
//
// The staging buffer holds UB_PITCH * ROWS floats.  The A layout sits at its END (offset A_OFF = 4 * ROWS), the
// B layout at its start: B row i ends at 36 (i + 1) <= A_OFF + 32 (i + 1), the start of A row i + 1, so rows can
// be moved from A to B eight at a time (read 8 rows of the own column, warp barrier, write them as running
// sums) without ever overwriting a row that has not been read -- 8 live registers instead of 32.
template <typename Cfg, int NB>
__device__ __forceinline__ void place_apply(const PlaneBwdParams& p, const int32_t* ent, float* ubuf, int lane,
                                            bool active, int dy, int dx, const PlaceFetch<NB>& f) {
    using BC = PlaneBwdCfg<Cfg>;
    constexpr int A_OFF = BC::U_WORKER - 32 * Cfg::ROWS;
    static_assert(A_OFF % 4 == 0 && A_OFF >= 0, "A layout must stay float4 aligned");
    // 1. clear the A layout (ROWS x 32 floats): ROWS/4 float4 per lane
    float4* a4 = reinterpret_cast<float4*>(ubuf + A_OFF);
#pragma unroll
    for (int i = 0; i < Cfg::ROWS / 4; ++i) a4[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncwarp();
    // 2. events, each lane in its own bank
    float* A = ubuf + A_OFF + lane;
    float head = 0.f;
#ifdef SSLB_EXPERIMENT_NOPLACE
    if (active && f.val[0][0] == 123.456f) {
#else
    if (active) {
#endif
#pragma unroll
        for (int kind = 0; kind < 2; ++kind) {
            int lo_off, hi_off;
            place_offsets<Cfg>(kind, dy, lo_off, hi_off);
            place_batch<Cfg, NB>(A, f.packed[kind], f.val[kind], lo_off, hi_off, head);
            // columns with more than NB entries of a kind (dense masks): the rest, in batches of 4
#ifdef SSLB_EXPERIMENT_NOOVERFLOW   // timing experiment only: entries beyond the prefetched NB are dropped
            if (f.e0[kind] + NB < f.e1[kind] && f.val[kind][0] == 123.456f) {
#else
            if (f.e0[kind] + NB < f.e1[kind]) {
#endif
                const float* gq = place_source<Cfg>(p, kind, dy, dx);
                for (int e = f.e0[kind] + NB; e < f.e1[kind]; e += 4) {
                    int packed[4];
                    float val[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const bool on = e + m < f.e1[kind];
                        packed[m] = on ? ent[e + m] : -1;
                        val[m] = on ? SSLB_GQ(gq, packed[m]) : 0.f;
                    }
                    place_batch<Cfg, 4>(A, packed, val, lo_off, hi_off, head);
                }
            }
        }
    }
    // 3. A -> B, eight rows at a time, as a running sum down the column; the partial sums of a quad do not wait
    //    for the running total
    float* Bc = ubuf + lane;
    float run = head;
#pragma unroll
    for (int i0 = 0; i0 < Cfg::ROWS; i0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = A[(i0 + i) * 32];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            run += v[i];
            Bc[(i0 + i) * BC::UB_PITCH] = run;
        }
    }
    __syncwarp();
}

// One chunk of the h-direction tree + products for one sweep thread.
template <typename Cfg, int GI>
__device__ __forceinline__ void sweep_chunk_bwd(const float* tile, const float* uworker, int r, int dy, int k, bool prime,
                                                BoxCarry (&carry)[GroupConsts<Cfg, GI>::GJ], float (&acc)[3][8]) {
    using GC = GroupConsts<Cfg, GI>;
    constexpr int P = Cfg::P, GJ = GC::GJ, OFF = GC::OFF, NV4 = GC::NV4;
    float gs[GJ][8];
    BoxDispatch<0, GJ>::template run<Cfg, GC::DX0>([&](auto jc, auto lenc) {
        constexpr int j = decltype(jc)::value, len = decltype(lenc)::value;
        const float4* up = reinterpret_cast<const float4*>(uworker + r * PlaneBwdCfg<Cfg>::UB_PITCH + j * 8);
        float cur[8];
        *reinterpret_cast<float4*>(&cur[0]) = up[0];
        *reinterpret_cast<float4*>(&cur[4]) = up[1];
        box_last<len>(cur, carry[j], gs[j]);
    });
    if (prime) return;  // the first chunk of an item only primes the tree (its outputs belong to the chunk before)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float* rowb = tile + (c * Cfg::IROWS + r + P) * Cfg::IPITCH + Cfg::ICOL0 + 8 * k;
        const float* rown = tile + (c * Cfg::IROWS + r + P + dy) * Cfg::IPITCH + Cfg::ICOL0 + 8 * k + (GC::DX0 - OFF);
        float base[8], nb[4 * NV4];
        *reinterpret_cast<float4*>(&base[0]) = *reinterpret_cast<const float4*>(rowb);
        *reinterpret_cast<float4*>(&base[4]) = *reinterpret_cast<const float4*>(rowb + 4);
#pragma unroll
        for (int v = 0; v < NV4; ++v)
            *reinterpret_cast<float4*>(&nb[4 * v]) = *reinterpret_cast<const float4*>(rown + 4 * v);
#pragma unroll
        for (int j = 0; j < GJ; ++j)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[c][i] = fmaf(base[i] - nb[OFF + i + j], gs[j][i], acc[c][i]);
    }
}

// Workers free-run through their own (dy, chunk) sequence.  The only shared state is the accumulator
// tile; additions into its 8-column chunk c are serialised by a ticket per chunk, handed out in the
// fixed order (dy-round, worker), so the summation order -- and the result -- is the same every run.
// The ticket is a block-scope atomic: lane 0 acquires it, __syncwarp extends the ordering to the other
// lanes, and lane 0 releases the next ticket after the warp's updates.
template <typename Cfg, int GI>
__device__ __forceinline__ void run_group_bwd(const PlaneBwdParams& p, const float* tile, float* ubuf, float* accT,
                                              const int32_t* cols, const uint8_t* cum, const int32_t* ent, int* turn,
                                              uint32_t tmem_base) {
    using GC = GroupConsts<Cfg, GI>;
    using BC = PlaneBwdCfg<Cfg>;
    constexpr int P = Cfg::P, GJ = GC::GJ, NWP = Cfg::NWP;
    const int tid = threadIdx.x;
    const int wp = tid / Cfg::ROWS, r = tid % Cfg::ROWS;
    float* uworker = ubuf + wp * BC::U_WORKER;
    // place-side role: (u-column c8, plane pj)
    const int c8 = r & 7, pj = r >> 3;
    const bool placing = pj < GJ;
    BoxCarry carry[GJ];
    // Work items of this worker: whole rows of offsets dy = wp - P, wp - P + NWP, ... (chunks 0 .. NCHB-1, chunk 0
    // only primes the h-tree), and -- when exactly one dy is left over (k_s = 25 on 12 workers) -- one output chunk
    // of that last dy: the h-tree reaches back 2K <= 8 columns, so output chunk c needs only chunks c (priming)
    // and c + 1.  Tickets number the additions into an output chunk in the fixed order (item, worker).
    for (int item = 0; item < BC::N_ITEMS; ++item) {
        int dy, k_begin, k_end;
        if (item < BC::FULL_ROUNDS || !BC::SPLIT_LAST) {
            dy = wp - P + item * NWP;
            k_begin = 0;
            k_end = BC::NCHB;
            if (dy > P) break;
        } else {
            dy = P;
            k_begin = wp;
            k_end = wp + 2;
            if (wp >= BC::NCHB - 1) break;
        }
#if SSLB_BWD_TMEM
        // additions into (quarter, output chunk) are ordered among the workers of the quarter: (item, worker / 4)
        constexpr int WQ = (NWP + 3) / 4;
        const int ticket = item * WQ + (item < BC::FULL_ROUNDS || !BC::SPLIT_LAST ? wp / 4 : 0);
#else
        const int ticket = item * NWP + (item < BC::FULL_ROUNDS || !BC::SPLIT_LAST ? wp : 0);
#endif
#pragma unroll
        for (int j = 0; j < GJ; ++j) box_carry_reset(carry[j]);
        constexpr int NB = BC::NB;
        PlaceFetch<NB> pf;
        if (placing) place_fetch<Cfg, NB>(p, cols, cum, ent, 8 * k_begin + c8, dy, GC::DX0 + pj, pf);
        for (int k = k_begin; k < k_end; ++k) {
            __syncwarp();  // the previous chunk's sweep has finished reading the staging buffer
            // 1. build the chunk's 8 u-columns of every plane (one lane per column), then start the loads of
            //    the next chunk's columns: they fly during the sweep below
            place_apply<Cfg, NB>(p, ent, uworker, r, placing, dy, GC::DX0 + pj, pf);
            if (placing && k + 1 < k_end) place_fetch<Cfg, NB>(p, cols, cum, ent, 8 * (k + 1) + c8, dy, GC::DX0 + pj, pf);
            // 2. h-direction + products
            float acc[3][8];
#pragma unroll
            for (int c = 0; c < 3; ++c)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[c][i] = 0.f;
            sweep_chunk_bwd<Cfg, GI>(tile, uworker, r, dy, k, k == k_begin, carry, acc);
            // 3. add into the accumulator tile (output chunk k-1) when it is this worker's turn
            if (k > k_begin) {
#if SSLB_BWD_TMEM
                cuda::atomic_ref<int, cuda::thread_scope_block> tk(turn[(wp & 3) * BC::NCHB + k - 1]);
#else
                cuda::atomic_ref<int, cuda::thread_scope_block> tk(turn[k - 1]);
#endif
#ifndef SSLB_EXPERIMENT_NOTICKET
                if (r == 0)
                    while (tk.load(cuda::memory_order_acquire) != ticket) __nanosleep(32);
#endif
                __syncwarp();
#if SSLB_BWD_TMEM
                tmem_fence_after_sync();
                // lane = row: the thread's 8 columns of every channel are 8 consecutive TMEM columns of its lane
                const uint32_t ta = tmem_base + ((uint32_t)((wp & 3) * 32) << 16) + 8 * (k - 1);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float v[8];
                    tmem_ld8(ta + c * BC::TXB, v);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += acc[c][i];
                    tmem_st8(ta + c * BC::TXB, v);
                }
                tmem_wait_st();
                tmem_fence_before_sync();
#else
#ifdef SSLB_EXPERIMENT_NOACC
                if (acc[0][0] == 123.456f)
#endif
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    float4* dst = reinterpret_cast<float4*>(accT + (c * Cfg::ROWS + r) * BC::ACC_PITCH + 8 * (k - 1));
                    float4 v0 = dst[0], v1 = dst[1];
                    v0.x += acc[c][0]; v0.y += acc[c][1]; v0.z += acc[c][2]; v0.w += acc[c][3];
                    v1.x += acc[c][4]; v1.y += acc[c][5]; v1.z += acc[c][6]; v1.w += acc[c][7];
                    dst[0] = v0; dst[1] = v1;
                }
#endif
                __syncwarp();
                if (r == 0) tk.store(ticket + 1, cuda::memory_order_release);
            }
        }
    }
}

template <typename Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) ssg_plane_bwd_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                        PlaneBwdParams p) {
    using BC = PlaneBwdCfg<Cfg>;
    extern __shared__ __align__(1024) unsigned char plane_smem_raw[];
    float* tile = reinterpret_cast<float*>(plane_smem_raw);
    float* ubuf = tile + Cfg::TILE_FLOATS;
    float* accT = ubuf + Cfg::NWP * BC::U_WORKER;
    int32_t* ent_s = reinterpret_cast<int32_t*>(accT + 3 * Cfg::ROWS * BC::ACC_PITCH);
    uint8_t* cum_s = reinterpret_cast<uint8_t*>(ent_s + BC::LIST_SMEM);
    __shared__ int32_t cols_s[BC::RCOLS + 1];
    __shared__ int turn_s[4 * BC::NCHB];
    __shared__ __align__(8) uint64_t tile_bar;
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 4 * BC::NCHB) turn_s[threadIdx.x] = 0;
    const int t = blockIdx.x;
    const int txb = t % p.ntxb, tyb = (t / p.ntxb) % p.ntyb, b = t / (p.ntxb * p.ntyb);
    const int32_t* cols_g = p.tile_cols + (long long)t * (BC::RCOLS + 1);
    const int n_ent = cols_g[BC::RCOLS];
    float* out = p.gpart + ((long long)blockIdx.y * p.B + b) * 3 * p.HT * p.WT;
    const int Yb0 = tyb * Cfg::ROWS, Xb0 = txb * BC::TXB;
    if (n_ent == 0) {  // nothing lands in this tile
        for (int i = threadIdx.x; i < 3 * Cfg::ROWS * BC::TXB; i += blockDim.x) {
            const int xo = i % BC::TXB, rr = (i / BC::TXB) % Cfg::ROWS, c = i / (BC::TXB * Cfg::ROWS);
            out[((long long)c * p.HT + Yb0 + rr) * p.WT + Xb0 + xo] = 0.f;
        }
        return;
    }
    issue_tile_load<Cfg>(tile, &tmap, &tile_bar, Xb0 - 8 - Cfg::ICOL0, Yb0 - Cfg::P, b * 3);
#if SSLB_BWD_TMEM
    static_assert(3 * BC::TXB <= 512 && Cfg::ROWS == 32, "one TMEM lane per image row, 3 * TXB columns per quarter");
    if (threadIdx.x < 32) tmem_alloc_512(&tmem_slot);   // warp 0 (the whole warp) allocates all 512 columns
    tmem_fence_before_sync();
#endif
#if !SSLB_BWD_TMEM
    for (int i = threadIdx.x; i < 3 * Cfg::ROWS * BC::ACC_PITCH; i += blockDim.x) accT[i] = 0.f;
#endif
    for (int i = threadIdx.x; i <= BC::RCOLS; i += blockDim.x) cols_s[i] = cols_g[i];
    {
        const uint32_t* cg = reinterpret_cast<const uint32_t*>(p.tile_cum + (long long)t * BC::RCOLS * BC::CUM_PITCH);
        uint32_t* cs = reinterpret_cast<uint32_t*>(cum_s);
        for (int i = threadIdx.x; i < BC::RCOLS * BC::CUM_PITCH / 4; i += blockDim.x) cs[i] = cg[i];
    }
    const int32_t* ent_g = p.tile_ent + (long long)t * BC::LIST_STRIDE;
    const bool staged = n_ent <= BC::LIST_SMEM;
    if (staged)
        for (int i = threadIdx.x; i < n_ent; i += blockDim.x) ent_s[i] = ent_g[i];
    __syncthreads();
    uint32_t tmem_base = 0;
#if SSLB_BWD_TMEM
    tmem_fence_after_sync();
    tmem_base = tmem_slot;
    if (threadIdx.x < 128) {   // warps 0..3 clear their quarter's partial accumulator
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const uint32_t ta = tmem_base + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
        for (int c = 0; c < 3 * BC::TXB; c += 8) tmem_st8(ta + c, z);
        tmem_wait_st();
    }
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
#endif
    mbar_wait(&tile_bar, 0);
    // two copies of the dispatch: with the entry list staged (always, below ~35 % mask density) the compiler sees a
    // shared-memory pointer and emits LDS instead of generic loads
    auto dispatch = [&](const int32_t* ent) {
        switch (blockIdx.y) {
            case 0: run_group_bwd<Cfg, 0>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 1: if constexpr (Cfg::NDXG > 1) run_group_bwd<Cfg, 1>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 2: if constexpr (Cfg::NDXG > 2) run_group_bwd<Cfg, 2>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 3: if constexpr (Cfg::NDXG > 3) run_group_bwd<Cfg, 3>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 4: if constexpr (Cfg::NDXG > 4) run_group_bwd<Cfg, 4>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 5: if constexpr (Cfg::NDXG > 5) run_group_bwd<Cfg, 5>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            case 6: if constexpr (Cfg::NDXG > 6) run_group_bwd<Cfg, 6>(p, tile, ubuf, accT, cols_s, cum_s, ent, turn_s, tmem_base); break;
            default: break;
        }
    };
    if (staged) dispatch(ent_s);
    else dispatch(ent_g);
#if SSLB_BWD_TMEM
    // Epilogue: every warp copies a third of its quarter's partial tile from TMEM into a staging area (the image
    // tile, the placement buffers, accT and the lists are free now and contiguous), then each thread adds the four partials of
    // its output elements in the fixed order (q0 + q1) + (q2 + q3) and writes the result.  TMEM is released in between.
    constexpr int QS = 3 * Cfg::ROWS * BC::ACC_PITCH;              // floats of one staged partial tile
    static_assert(4 * QS * sizeof(float) <= (Cfg::TILE_FLOATS + Cfg::NWP * BC::U_WORKER + QS) * sizeof(float) +
                                                BC::LIST_SMEM * sizeof(int32_t) + BC::RCOLS * BC::CUM_PITCH,
                  "the staging area of the epilogue is the kernel's whole dynamic shared memory (lists included)");
    static_assert(Cfg::NWP % 4 == 0 && (3 * BC::TXB / 8) % (Cfg::NWP / 4) == 0, "columns split evenly over a quarter's warps");
    float* S = tile;
    tmem_fence_before_sync();
    __syncthreads();
    tmem_fence_after_sync();
    {
        const int w = threadIdx.x >> 5, r = threadIdx.x & 31, q = w & 3;
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16);
        for (int j = w >> 2; j < 3 * BC::TXB / 8; j += Cfg::NWP / 4) {
            float v[8];
            tmem_ld8(ta + 8 * j, v);
            const int c = (8 * j) / BC::TXB, x = (8 * j) % BC::TXB;
            float4* dst = reinterpret_cast<float4*>(S + q * QS + (c * Cfg::ROWS + r) * BC::ACC_PITCH + x);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
    }
    tmem_fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc_512(tmem_base);
    // the factor 2 of d(t^2) is applied here, once
    for (int i = threadIdx.x; i < 3 * Cfg::ROWS * BC::TXB; i += blockDim.x) {
        const int xo = i % BC::TXB, rr = (i / BC::TXB) % Cfg::ROWS, c = i / (BC::TXB * Cfg::ROWS);
        const float* sp = S + (c * Cfg::ROWS + rr) * BC::ACC_PITCH + xo;
        out[((long long)c * p.HT + Yb0 + rr) * p.WT + Xb0 + xo] = 2.f * ((sp[0] + sp[QS]) + (sp[2 * QS] + sp[3 * QS]));
    }
#else
    __syncthreads();
    // the factor 2 of d(t^2) is applied here, once
    for (int i = threadIdx.x; i < 3 * Cfg::ROWS * BC::TXB; i += blockDim.x) {
        const int xo = i % BC::TXB, rr = (i / BC::TXB) % Cfg::ROWS, c = i / (BC::TXB * Cfg::ROWS);
        out[((long long)c * p.HT + Yb0 + rr) * p.WT + Xb0 + xo] = 2.f * accT[(c * Cfg::ROWS + rr) * BC::ACC_PITCH + xo];
    }
#endif
}

template <typename Cfg>
constexpr size_t plane_bwd_smem_bytes() {
    using BC = PlaneBwdCfg<Cfg>;
    return (size_t)(Cfg::TILE_FLOATS + Cfg::NWP * BC::U_WORKER + 3 * Cfg::ROWS * BC::ACC_PITCH) * sizeof(float) +
           (size_t)BC::LIST_SMEM * sizeof(int32_t) + (size_t)BC::RCOLS * BC::CUM_PITCH;
}

// ---- out-of-area terms and the reflect-pad adjoint -------------------------------------------
// wtab[slot][(a+K)*KW + (b+K)] = sum of dL/dq over the classes for which window offset (a,b) is out
// of area (similarity.cu:123-124: the neighbour is zero there, so only the centre pixel gets 2*I*g).
template <typename Cfg>
__global__ void __launch_bounds__(128) plane_wtab_kernel(const float* gcls, const int32_t* counts, int cap, float* wtab) {
    constexpr int K = Cfg::K, KW = Cfg::KW, NC = Cfg::NCLS;
    __shared__ float sG[4][NC * NC], sR[4][NC], sT[4][NC * KW];
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slots = min(counts[0], cap);
    for (int slot = blockIdx.x * 4 + w; slot < n_slots; slot += gridDim.x * 4) {
        for (int i = lane; i < NC * NC; i += 32) sG[w][i] = gcls[(long long)i * cap + slot];
        __syncwarp();
        if (lane < NC) {
            float s = 0.f;
            for (int cb = 0; cb < NC; ++cb) s += sG[w][lane * NC + cb];
            sR[w][lane] = s;
        }
        for (int i = lane; i < NC * KW; i += 32) {
            const int ca = i / KW, b = i % KW - K;
            float s = 0.f;
            for (int cb = 0; cb < NC; ++cb)
                if (b < class_lo(cb, K) || b > class_hi(cb, K)) s += sG[w][ca * NC + cb];
            sT[w][i] = s;
        }
        __syncwarp();
        for (int i = lane; i < KW * KW; i += 32) {
            const int a = i / KW - K, b = i % KW;
            float s = 0.f;
            for (int ca = 0; ca < NC; ++ca) s += (a < class_lo(ca, K) || a > class_hi(ca, K)) ? sR[w][ca] : sT[w][ca * KW + b];
            wtab[(long long)slot * (KW * KW) + i] = s;
        }
        __syncwarp();
    }
}

// rows in the reference's order [n][L] -> offset-major gqT[L][cap] (padding slots get zeros).
// One thread per slot walks its row; consecutive threads write consecutive slots.
__global__ void __launch_bounds__(256) plane_rows_to_slots_kernel(const float* rows, const int32_t* slot_pix,
                                                                  const int32_t* slot_ref, const int32_t* counts, int cap,
                                                                  int L, float* gqT) {
    const int n_slots = min(counts[0], cap);
    for (int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots; slot += gridDim.x * blockDim.x) {
        const bool real = slot_pix[slot] >= 0;
        const float* src = rows + (long long)(real ? slot_ref[slot] : 0) * L;
        for (int d = 0; d < L; ++d) gqT[qt_index(d, slot, L)] = real ? __ldg(src + d) : 0.f;
    }
}

// gcls[class][slot] = sum of gqT over the offsets of a clip class (what row_loss_t_kernel emits in the
// fused step); classes with nothing out of area stay zero.  One thread per (slot, class).
__global__ void __launch_bounds__(256) plane_class_sums_kernel(const float* gqT, const int32_t* counts, int cap, int KS,
                                                               int P, int K, float* gcls) {
    const int n_slots = min(counts[0], cap);
    const int NC = 2 * K + 1, U = P - K;
    const long long total = (long long)n_slots * NC * NC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int slot = (int)(i % n_slots), c = (int)(i / n_slots);
        const int ca = c / NC, cb = c % NC;
        float acc = 0.f;
        if (ca != K || cb != K) {
            const int dy0 = ca < K ? ca - P : (ca > K ? U + (ca - K) : -U), dy1 = ca == K ? U : dy0;
            const int dx0 = cb < K ? cb - P : (cb > K ? U + (cb - K) : -U), dx1 = cb == K ? U : dx0;
            for (int dy = dy0; dy <= dy1; ++dy)
                for (int dx = dx0; dx <= dx1; ++dx) acc += gqT[qt_index((dy + P) * KS + dx + P, slot, KS * KS)];
        }
        gcls[(long long)c * cap + slot] = acc;
    }
}

struct PlaneFinishParams {
    const float* pad;         // [B][3][Hp][pitch] reflect-padded fp32 SR (image 0 of the padded buffer)
    int Hp, pitch;
    float* gpart;             // [NDXG][B][3][HT][WT]; part 0 is overwritten with the folded padded gradient
    const float* wtab;        // [cap][KW*KW]
    const int32_t* slot_map;
    float* grad;              // [B,3,H,W], overwritten
    int B, H, W, HT, WT, n_parts, cap;
};

// Fold of the padded-domain gradient, one block = 16x16 padded pixels of one image:
//   wsum(Y,X) = sum over the edge pixels p = (Y,X) - (a,b), |a|,|b| <= K, of wtab[p][(a,b)]   (the slots of the
//               (16+2K)^2 pixels around the block are staged once)
//   G(Y,X,c)  = sum of the dx-group partials + 2 * wsum * I(Y,X,c)        -> written over part 0
template <typename Cfg, int NPARTS>
__global__ void __launch_bounds__(256) plane_fold_kernel(PlaneFinishParams p) {
    constexpr int P = Cfg::P, K = Cfg::K, KW = Cfg::KW, T = 16, R = T + 2 * K;
    __shared__ int32_t ss[R][R + 1];
    const int b = blockIdx.z, Y0 = blockIdx.y * T, X0 = blockIdx.x * T;
    const int ty = threadIdx.x / T, tx = threadIdx.x % T;
    const int Y = Y0 + ty, X = X0 + tx;
    const long long plane = (long long)p.HT * p.WT;
    const bool inside = Y < p.Hp && X < p.W + 2 * P;
    // the partials and the image are independent of the slots: all 3 * NPARTS + 3 loads are in flight while the
    // slot tile is staged
    float part[3][NPARTS], img[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long o = ((long long)b * 3 + c) * plane + (long long)Y * p.WT + X;
#pragma unroll
        for (int k = 0; k < NPARTS; ++k) part[c][k] = p.gpart[(long long)k * p.B * 3 * plane + o];
        img[c] = inside ? __ldg(p.pad + (((long long)b * 3 + c) * p.Hp + Y) * p.pitch + X) : 0.f;
    }
    int any = 0;
    for (int i = threadIdx.x; i < R * R; i += 256) {
        const int ry = i / R, rx = i % R;
        const int y = Y0 + ry - K - P, x = X0 + rx - K - P;   // image coordinates of padded (Y0+ry-K, X0+rx-K)
        int slot = -1;
        if (y >= 0 && y < p.H && x >= 0 && x < p.W) slot = p.slot_map[(b * p.H + y) * p.W + x];
        if (slot >= p.cap) slot = -1;
        ss[ry][rx] = slot;
        any |= slot >= 0;
    }
    any = __syncthreads_or(any);
    float w = 0.f;
    if (any) {
        // edge pixel at region (ty + K - a, tx + K - b) reaches this pixel with window offset (a, b); one partial sum
        // per window column keeps the gathers of a window row independent of each other
        float wcol[KW];
#pragma unroll
        for (int bb = 0; bb < KW; ++bb) wcol[bb] = 0.f;
        for (int a = -K; a <= K; ++a)
#pragma unroll
            for (int bb = -K; bb <= K; ++bb) {
                const int slot = ss[ty + K - a][tx + K - bb];
                if (slot >= 0) wcol[bb + K] += p.wtab[(long long)slot * (KW * KW) + (a + K) * KW + bb + K];
            }
#pragma unroll
        for (int bb = 0; bb < KW; ++bb) w += wcol[bb];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long o = ((long long)b * 3 + c) * plane + (long long)Y * p.WT + X;
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < NPARTS; ++k) s += part[c][k];
        if (inside) s = fmaf(2.f * w, img[c], s);
        p.gpart[o] = s;
    }
}

// Adjoint of F.pad(reflect) (similaritywrapper.py:64) as a gather: an image pixel sums the padded pixels that
// mirror onto it (at most 3 per axis).
template <typename Cfg>
__global__ void __launch_bounds__(256) plane_finish_kernel(PlaneFinishParams p) {
    constexpr int P = Cfg::P;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long hw = (long long)p.H * p.W;
    if (idx >= p.B * hw) return;
    const int b = (int)(idx / hw), rem = (int)(idx - b * hw), y = rem / p.W, x = rem - y * p.W;
    int Ys[3], Xs[3], ny = 1, nx = 1;
    Ys[0] = y + P; Xs[0] = x + P;
    if (y >= 1 && y <= P) Ys[ny++] = P - y;
    if (y <= p.H - 2 && y >= p.H - 1 - P) Ys[ny++] = P + 2 * (p.H - 1) - y;
    if (x >= 1 && x <= P) Xs[nx++] = P - x;
    if (x <= p.W - 2 && x >= p.W - 1 - P) Xs[nx++] = P + 2 * (p.W - 1) - x;
    const long long plane = (long long)p.HT * p.WT;
    const float* G = p.gpart + (long long)b * 3 * plane;
    float tot[3] = {0.f, 0.f, 0.f};
    for (int iy = 0; iy < ny; ++iy)
        for (int ix = 0; ix < nx; ++ix) {
            const long long o = (long long)Ys[iy] * p.WT + Xs[ix];
#pragma unroll
            for (int c = 0; c < 3; ++c) tot[c] += G[c * plane + o];
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) p.grad[((long long)b * 3 + c) * hw + rem] = tot[c];
}

}  // namespace sslb
