// Forward "plane" kernels: raw patch distances q(p,d) of every edge pixel of a batch, offset-major.
//
// One CTA owns one forward tile (TYF x TXF edge-pixel positions of one image) and one group of G
// consecutive dx.  Its NWP warp pairs take NWP consecutive dy each step; a sweep thread is one
// image row (lane) of one (dy, dx-group) and walks along X in chunks of 8 columns:
//     D_j[i]   = sum_c ( I(Y, X+i) - I(Y+dy, X+i+dx0+j) )^2          8 x G values in registers
//     S_j[i]   = sum of the last l(dx0+j) values of D_j               (pairwise tree, no subtraction)
// and stores S into a 16-column shared ring, shifted so that every plane is read at column px+K.
// After each chunk the CTA gathers the chunk's edge pixels: lane <-> plane, 4 slots per lane,
//     q(p,d) = sum_{a in A(dy)} S_d[py+a][px+K] + Eout(p, class(dy), class(dx))
// and writes one float4 of qT[d][slot..slot+3].
//
// Reference: GAN-Based-SR/basicsr/losses/similarity/similarity.cu:5-54 (same sums, regrouped);
// the algebra is restated and checked against the oracle in tests/dense_model.py.
#pragma once

#include <type_traits>

#include "plane_geom.cuh"
#include "tma.cuh"

namespace sslb {

struct PlaneFwdParams {
    const float* pad;        // [2][B][3][Hp][pitch] reflect-padded fp32 images (pad.cuh); image 0 = SR, 1 = GT
    int Hp, pitch;
    int img_first;           // padded image set of blockIdx.z == 0 (launches that cover one image of the two)
    float* qT[2];            // KS*KS x cap, panel layout (qt_index)
    const float* eout[2];    // [cap][NCLS*NCLS]
    PlaneLists lists;
    PlaneGeom g;
    int cap;
};

// ---- Eout tables -------------------------------------------------------------------------
// eout[slot][ca*NCLS+cb] = sum over window offsets (a,b) outside A(ca) x A(cb) of sum_c I(p+(a,b))^2.
// One block per unit (8 columns x TYF rows of edge-pixel positions): the squared norms E2 = sum_c I^2 of the
// unit's (TYF + 2K) x (8 + 2K) neighbourhood are formed once in shared memory and shared by all of its slots
// (every E2 value serves up to (2K+1)^2 of them); then ONE THREAD per slot builds the whole table in registers.
// The complement of a clip range is a prefix or a suffix of the window in each direction, so with running sums
// along the window rows (pre / suf / full) and then along the window columns every entry costs one or two
// additions; all terms are non-negative and nothing is subtracted.  Tables leave through shared memory so that the
// global stores are contiguous.
constexpr int kEoutThreads = 64;

template <typename Cfg>
__global__ void __launch_bounds__(kEoutThreads) plane_eout_kernel(PlaneFwdParams p) {
    constexpr int K = Cfg::K, KW = Cfg::KW, NC = Cfg::NCLS, P = Cfg::P;
    constexpr int ER = Cfg::ROWS, EC = 8 + 2 * K, EP = EC + 1;   // region rows, columns, pitch
    constexpr int OP = NC * NC + 1;                              // odd pitch of the staged tables
    __shared__ float sE2[ER * EP];
    __shared__ float sOut[kEoutThreads * OP];
    const int u = blockIdx.x, which = blockIdx.y;
    const int slot0 = p.lists.unit_start[u], slot1 = min(p.lists.unit_start[u + 1], p.cap);
    if (slot0 >= slot1) return;
    int b, ty, tx, cx;
    decode_unit(p.g, u, b, ty, tx, cx);
    // padded coordinates of region (0,0): image (ty*TYF - K, tx*TXF + cx*8 - xs - K)
    const int Y0 = ty * Cfg::TYF - K + P, X0 = tx * Cfg::TXF + cx * 8 - p.g.xs - K + P;
    const int Wp = p.g.W + 2 * P;
    const long long plane = (long long)p.Hp * p.pitch;
    const float* img = p.pad + ((long long)(p.img_first + which) * p.g.B + b) * 3 * plane;
    // four region elements per thread and step: 12 independent loads in flight (the block has only 64 threads)
    constexpr int FU = 4;
    for (int i0 = threadIdx.x; i0 < ER * EC; i0 += FU * kEoutThreads) {
        float v[FU][3];
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            const int i = i0 + u * kEoutThreads;
            const int ry = i / EC, rx = i - ry * EC;
            const int Y = Y0 + ry, X = X0 + rx;
            const bool in = i < ER * EC && Y >= 0 && Y < p.Hp && X >= 0 && X < Wp;
            const float* q = img + (long long)Y * p.pitch + X;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[u][c] = in ? __ldg(q + c * plane) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < FU; ++u) {
            const int i = i0 + u * kEoutThreads;
            if (i < ER * EC) {
                const int ry = i / EC, rx = i - ry * EC;
                sE2[ry * EP + rx] = fmaf(v[u][2], v[u][2], fmaf(v[u][1], v[u][1], v[u][0] * v[u][0]));
            }
        }
    }
    __syncthreads();
    float* eout = const_cast<float*>(which ? p.eout[1] : p.eout[0]);
    for (int base = slot0; base < slot1; base += kEoutThreads) {
        const int slot = base + threadIdx.x;
        const int rc = slot < slot1 ? p.lists.slot_rc[slot] : -1;
        float* mine = sOut + threadIdx.x * OP;
        if (rc >= 0) {
            const int re = rc >> 8, lx = (rc & 255) & 7;   // row lane (tile row + K), column inside the unit
            const float* win = sE2 + (re - K) * EP + lx;   // window (a, b) at win[a * EP + b], a, b in [0, KW)
            // per window row: sum of its first m / last m columns (m = 1..K) and of all of them
            float pre[KW][K + 1], suf[KW][K + 1], full[KW];
#pragma unroll
            for (int a = 0; a < KW; ++a) {
                float e[KW];
#pragma unroll
                for (int bb = 0; bb < KW; ++bb) e[bb] = win[a * EP + bb];
                pre[a][0] = 0.f;
                suf[a][0] = 0.f;
#pragma unroll
                for (int m = 1; m <= K; ++m) {
                    pre[a][m] = pre[a][m - 1] + e[m - 1];
                    suf[a][m] = suf[a][m - 1] + e[KW - m];
                }
                float mid = e[K];
#pragma unroll
                for (int m = 1; m < K; ++m) mid += e[K - m] + e[K + m];          // e[1..KW-2]
                full[a] = (pre[a][1] + mid) + suf[a][1];                        // + e[0] + e[KW-1]
            }
            // rows out of area count in full: running sums of `full` from the top and from the bottom
            float ftop[K + 1], fbot[K + 1];
            ftop[0] = fbot[0] = 0.f;
#pragma unroll
            for (int m = 1; m <= K; ++m) {
                ftop[m] = ftop[m - 1] + full[m - 1];
                fbot[m] = fbot[m - 1] + full[KW - m];
            }
#pragma unroll
            for (int cb = 0; cb < NC; ++cb) {
                // r[a] = the columns of row a that are out of area for column class cb
                float r[KW];
#pragma unroll
                for (int a = 0; a < KW; ++a) r[a] = cb < K ? pre[a][K - cb] : (cb > K ? suf[a][cb - K] : 0.f);
                // rows in area contribute r: sums of r over the rows below the top n / above the bottom n
                float rlow[K + 1], rhigh[K + 1];   // rlow[n] = sum_{a >= n} r[a], rhigh[n] = sum_{a < KW - n} r[a]
                float mid = r[K];
#pragma unroll
                for (int m = 1; m < K; ++m) mid += r[K - m] + r[K + m];          // rows 1..KW-2
                // rows K..KW-1-... build from the middle outwards so that every partial is a plain sum
                float all = (r[0] + mid) + r[KW - 1];
                rlow[0] = all;
                rhigh[0] = all;
                {
                    float t = 0.f;   // rlow[n]: drop the first n rows = sum of rows n..KW-1, accumulated from the bottom
                    float sfx[KW + 1];
                    sfx[KW] = 0.f;
#pragma unroll
                    for (int a = KW - 1; a >= 0; --a) sfx[a] = sfx[a + 1] + r[a];
                    float pfx[KW + 1];
                    pfx[0] = 0.f;
#pragma unroll
                    for (int a = 0; a < KW; ++a) pfx[a + 1] = pfx[a] + r[a];
#pragma unroll
                    for (int n = 1; n <= K; ++n) { rlow[n] = sfx[n]; rhigh[n] = pfx[KW - n]; }
                    (void)t;
                }
#pragma unroll
                for (int ca = 0; ca < NC; ++ca) {
                    const int n_top = ca < K ? K - ca : 0, n_bot = ca > K ? ca - K : 0;
                    // ca < K: the first n_top rows in full, the rest by r; ca > K: the last n_bot rows in full
                    const float v = ca < K ? ftop[n_top] + rlow[n_top] : (ca > K ? fbot[n_bot] + rhigh[n_bot] : all);
                    mine[ca * NC + cb] = v;
                }
            }
        } else {
#pragma unroll 1
            for (int i = 0; i < NC * NC; ++i) mine[i] = 0.f;
        }
        __syncthreads();
        const int n_here = min(kEoutThreads, slot1 - base);
        float* dst = eout + (long long)base * (NC * NC);
        for (int i = threadIdx.x; i < n_here * NC * NC; i += kEoutThreads) dst[i] = sOut[(i / (NC * NC)) * OP + i % (NC * NC)];
        __syncthreads();
    }
}

// ---- shared image tile -------------------------------------------------------------------
// tile[c][row][col]: row 0 <-> padded Y = Ytile0 - K - P, col ICOL0 <-> padded X = Xtile0 - K, filled by ONE
// 3-D tensor copy from the padded images (tma.cuh); the copy engine supplies zeros outside the padded image.
template <typename Cfg>
__device__ __forceinline__ void issue_tile_load(float* tile, const CUtensorMap* map, uint64_t* bar, int Xcol0, int Yrow0,
                                                int plane0) {
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
        mbar_arrive_expect_tx(bar, (uint32_t)(Cfg::TILE_FLOATS * sizeof(float)));
        tma_load_3d(tile, map, bar, Xcol0, Yrow0, plane0);
    }
}

// Sum of the last LEN entries of w ending at index i (static indices; pairwise partial sums are
// shared between the 8 outputs of a chunk by the caller through s2/s4/s8).
template <int LEN>
__device__ __forceinline__ float sum_last(const float* w, const float* s2, const float* s4, const float* s8, int i) {
    if (LEN == 1) return w[i];
    if (LEN == 2) return s2[i];
    if (LEN == 3) return s2[i] + w[i - 2];
    if (LEN == 4) return s4[i];
    if (LEN == 5) return s4[i] + w[i - 4];
    if (LEN == 6) return s4[i] + s2[i - 4];
    if (LEN == 7) return s4[i] + (s2[i - 4] + w[i - 6]);
    if (LEN == 8) return s8[i];
    return s8[i - 1] + w[i];  // LEN == 9
}

// Box sums of one plane over one 8-column chunk: out[i] = sum of the last LEN values ending at column i.
// The pairwise partial sums (pairs, quads, octets) that reach back into the previous chunk are carried
// in registers, so a chunk costs 4 additions per output for LEN = 9; carries a clipped LEN never reads
// are dead code.
struct BoxCarry {
    float w[8], s2[4], s4[4], s8;
};

__device__ __forceinline__ void box_carry_reset(BoxCarry& c) {
#pragma unroll
    for (int i = 0; i < 8; ++i) c.w[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) c.s2[i] = c.s4[i] = 0.f;
    c.s8 = 0.f;
}

template <int LEN, bool PREFIX_SUFFIX = true>
__device__ __forceinline__ void box_last(const float (&cur)[8], BoxCarry& c, float (&out)[8]) {
    if constexpr (PREFIX_SUFFIX && (LEN == 9 || LEN == 8)) {
        // Windows of 8 or 9 over chunks of 8 (van Herk / Gil-Werman): a window that ends at column m of this chunk is
        // a SUFFIX of the previous chunk plus the PREFIX 0..m of this one, so 7 + 7 additions build the prefix and
        // suffix sums of a chunk and 8 more combine them -- 2.75 additions per output instead of 4, still nothing
        // but sums of at most 9 non-negative terms.  c.w carries the previous chunk's suffix sums (zero at the start).
        constexpr int SH = 9 - LEN;                  // the suffix of a LEN-window starts SH columns later
        float pre = 0.f;                             // running prefix: a carried suffix sum is dead once it is used
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            pre = m == 0 ? cur[0] : pre + cur[m];
            out[m] = m + SH < 8 ? c.w[m + SH] + pre : pre;
        }
        c.w[7] = cur[7];
#pragma unroll
        for (int m = 6; m >= 0; --m) c.w[m] = c.w[m + 1] + cur[m];
        return;
    }
    float w[16], s2[16], s4[16], s8[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = c.w[i]; w[8 + i] = cur[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { s2[4 + i] = c.s2[i]; s4[4 + i] = c.s4[i]; }
    s8[7] = c.s8;
#pragma unroll
    for (int i = 8; i < 16; ++i) s2[i] = w[i] + w[i - 1];
#pragma unroll
    for (int i = 8; i < 16; ++i) s4[i] = s2[i] + s2[i - 2];
#pragma unroll
    for (int i = 8; i < 16; ++i) s8[i] = s4[i] + s4[i - 4];
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = sum_last<LEN>(w, s2, s4, s8, 8 + i);
#pragma unroll
    for (int i = 0; i < 8; ++i) c.w[i] = cur[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { c.s2[i] = s2[12 + i]; c.s4[i] = s4[12 + i]; }
    c.s8 = s8[15];
}

template <int J, int JEND>
struct BoxDispatch {
    // compile-time loop over the planes of a group so that LEN is a template argument
    template <typename Cfg, int DX0, typename F>
    static __device__ __forceinline__ void run(F&& f) {
        if constexpr (J < JEND) {
            constexpr int dx = DX0 + J;
            constexpr int len = rng_hi(dx, Cfg::P, Cfg::K) - rng_lo(dx, Cfg::P, Cfg::K) + 1;
            f(std::integral_constant<int, J>{}, std::integral_constant<int, len>{});
            BoxDispatch<J + 1, JEND>::template run<Cfg, DX0>(f);
        }
    }
};

template <typename Cfg, int GI>
struct GroupConsts {
    static constexpr int DX0 = -Cfg::P + GI * Cfg::G;
    static constexpr int GJ = GI == Cfg::NDXG - 1 ? Cfg::KS - GI * Cfg::G : Cfg::G;   // the last group takes the rest
    static constexpr int OFF = ((DX0 % 4) + 4) % 4;
    static constexpr int NV4 = (OFF + 8 + GJ - 1 + 3) / 4;
    static_assert(GJ >= 1 && GJ <= Cfg::GMAX, "group width");
};

// One chunk of one sweep thread: D, box sums, ring store.
template <typename Cfg, int GI>
__device__ __forceinline__ void sweep_chunk_fwd(const float* tile, float* splanes, int r, int dy, int wp, int k,
                                                BoxCarry (&carry)[GroupConsts<Cfg, GI>::GJ]) {
    using GC = GroupConsts<Cfg, GI>;
    constexpr int P = Cfg::P, GJ = GC::GJ, OFF = GC::OFF, NV4 = GC::NV4;
    float d[GJ][8];
#pragma unroll
    for (int j = 0; j < GJ; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) d[j][i] = 0.f;
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
            for (int i = 0; i < 8; ++i) {
                const float t = base[i] - nb[OFF + i + j];
                d[j][i] = fmaf(t, t, d[j][i]);
            }
    }
    // Ring column of plane j = sweep column + (K - hi(dx_j)): every plane is then read at column px + 2K, so
    // the G planes of one edge pixel sit in G consecutive banks for clipped and unclipped dx alike.
    float* spa = splanes + (wp * Cfg::GMAX) * Cfg::SPS + r * Cfg::SRP + (k & 1) * 8;   // ring half of this chunk
    float* spb = splanes + (wp * Cfg::GMAX) * Cfg::SPS + r * Cfg::SRP + ((k & 1) ^ 1) * 8;  // the other half
    BoxDispatch<0, GJ>::template run<Cfg, GC::DX0>([&](auto jc, auto lenc) {
        constexpr int j = decltype(jc)::value, len = decltype(lenc)::value;
        constexpr int shift = Cfg::K - rng_hi(GC::DX0 + j, Cfg::P, Cfg::K);
        float s[8];
        box_last<len>(d[j], carry[j], s);
#ifdef SSLB_EXPERIMENT_NORING     // timing experiment only: no ring stores (one conditional store keeps the math alive)
        if (s[0] + s[7] == 123.456f) spa[j * Cfg::SPS] = s[3];
#else
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i + shift < 8) spa[j * Cfg::SPS + i + shift] = s[i];
            else spb[j * Cfg::SPS + i + shift - 8] = s[i];
        }
#endif
    });
}

// Barrier among the threads of one worker: a named barrier for a warp pair, __syncwarp for one warp.
template <int ROWS>
__device__ __forceinline__ void worker_sync(int wp) {
    if (ROWS == 32) __syncwarp();
    else asm volatile("bar.sync %0, 64;" ::"r"(1 + wp) : "memory");
}

// One worker = one warp pair = 64 image rows of one (dy, dx-group); workers of a CTA only share the
// image tile.  Per chunk: sweep -> barrier(64) -> gather of the previous chunk's edge pixels ->
// barrier(64).  In the gather a thread owns (group of 4 slots, plane j).
template <typename Cfg, int GI>
__device__ __forceinline__ void run_group_fwd(const PlaneFwdParams& p, const float* tile, float* splanes,
                                              const int32_t* ustart, const int32_t* slot_rc, int which) {
    using GC = GroupConsts<Cfg, GI>;
    constexpr int P = Cfg::P, K = Cfg::K, G = Cfg::GMAX, GJ = GC::GJ, NC = Cfg::NCLS;
    constexpr int NGRP = Cfg::ROWS / GJ;  // slot groups handled in parallel by one worker
    const int tid = threadIdx.x;
    const int wp = tid / Cfg::ROWS, r = tid % Cfg::ROWS;
    float* qT = which ? p.qT[1] : p.qT[0];
    const float* eout = which ? p.eout[1] : p.eout[0];
    const int cap = p.cap;
    const int slot0 = ustart[0];  // slot_rc is indexed relative to the tile's first slot
    // gather-side role of this thread
    const int gj = r % GJ, ggrp = r / GJ;
    const bool g_ok = ggrp < NGRP;
    const int g_dx = GC::DX0 + gj;
    const int cb = clip_class(g_dx, P, K);
    const int coff = 2 * K;  // ring columns are stored shifted per plane (sweep_chunk_fwd)
    const float* myplane = splanes + (wp * G + gj) * Cfg::SPS;
    for (int dy = wp - P; dy <= P; dy += Cfg::NWP) {
        const int alo = rng_lo(dy, P, K), ahi = rng_hi(dy, P, K);
        const int ca = clip_class(dy, P, K);
        const bool clipped = ca != K || cb != K;
        const int cls = ca * NC + cb;
        const int qd = (dy + P) * Cfg::KS + g_dx + P;   // offset index of this thread's plane

        BoxCarry carry[GJ];
#pragma unroll
        for (int j = 0; j < GJ; ++j) box_carry_reset(carry[j]);

        for (int k = 0; k < Cfg::NCH; ++k) {
            // Out-of-area terms of the first slot group this thread will gather after the sweep:
            // issued now so that the global-memory latency hides behind the sweep arithmetic.
            const int s0 = k >= 1 ? ustart[k - 1] : 0;
            const int s1 = k >= 1 ? min(ustart[k], cap) : 0;
            float eo[4] = {0.f, 0.f, 0.f, 0.f};
            const int gs_first = s0 + 4 * ggrp;
            if (clipped && g_ok && gs_first < s1) {
#pragma unroll
                for (int e = 0; e < 4; ++e) eo[e] = __ldg(eout + (long long)(gs_first + e) * (NC * NC) + cls);
            }
            sweep_chunk_fwd<Cfg, GI>(tile, splanes, r, dy, wp, k, carry);
            worker_sync<Cfg::ROWS>(wp);
#ifdef SSLB_EXPERIMENT_NOGATHER   // timing experiment only: the sweep without the gather
            if (g_ok && carry[0].s8 == 123.456f) {
#else
            if (g_ok) {
#endif
                for (int gs = gs_first; gs < s1; gs += 4 * NGRP) {
                    const int4 rc4 = *reinterpret_cast<const int4*>(slot_rc + (gs - slot0));
                    const int rcs[4] = {rc4.x, rc4.y, rc4.z, rc4.w};
                    if (gs != gs_first && clipped) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) eo[e] = __ldg(eout + (long long)(gs + e) * (NC * NC) + cls);
                    }
                    // all 4 x (2K+1) ring reads are issued before the first addition (no branch between the
                    // slots, so the loads of one slot do not wait for the sums of the previous one); rows
                    // re-K..re+K always lie inside the worker's rows, the a-range only selects what is added
                    float v[4][2 * K + 1];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int rc = rcs[e] >= 0 ? rcs[e] : (K << 8);  // padding slot: any valid address
                        const int re = rc >> 8, ex = rc & 255;
                        const float* sp = myplane + re * Cfg::SRP + ((ex + coff) & (Cfg::RING - 1));
#pragma unroll
                        for (int a = -K; a <= K; ++a) v[e][a + K] = sp[a * Cfg::SRP];
                    }
                    float out[4];
                    if (ca == K) {
                        // unclipped dy (17 of 25): all 2K+1 rows, summed pairwise (short dependency chains)
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float t[2 * K + 1];
#pragma unroll
                            for (int a = 0; a < 2 * K + 1; ++a) t[a] = v[e][a];
#pragma unroll
                            for (int w = 1; w < 2 * K + 1; w *= 2)
#pragma unroll
                                for (int a = 0; a + w < 2 * K + 1; a += 2 * w) t[a] += t[a + w];
                            out[e] = rcs[e] >= 0 ? t[0] + eo[e] : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float t[2 * K + 1];
#pragma unroll
                            for (int a = -K; a <= K; ++a) t[a + K] = (a >= alo && a <= ahi) ? v[e][a + K] : 0.f;
#pragma unroll
                            for (int w = 1; w < 2 * K + 1; w *= 2)
#pragma unroll
                                for (int a = 0; a + w < 2 * K + 1; a += 2 * w) t[a] += t[a + w];
                            out[e] = rcs[e] >= 0 ? t[0] + eo[e] : 0.f;
                        }
                    }
                    *reinterpret_cast<float4*>(qT + qt_index(qd, gs, Cfg::L)) = make_float4(out[0], out[1], out[2], out[3]);
                }
            }
            worker_sync<Cfg::ROWS>(wp);
        }
    }
}

template <typename Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1) ssg_plane_fwd_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                        PlaneFwdParams p) {
    extern __shared__ __align__(1024) unsigned char plane_smem_raw[];
    float* tile = reinterpret_cast<float*>(plane_smem_raw);
    float* splanes = tile + Cfg::TILE_FLOATS;
    int32_t* rc_s = reinterpret_cast<int32_t*>(splanes + Cfg::NPL * Cfg::SPS + 3);  // keep 16-byte alignment below
    rc_s = reinterpret_cast<int32_t*>((reinterpret_cast<uintptr_t>(rc_s) + 15) & ~(uintptr_t)15);
    __shared__ int32_t ustart_s[Cfg::UNITS_X + 1];
    __shared__ __align__(8) uint64_t tile_bar;
    // the dx-groups of one tile are neighbours in launch order: they share the tile's image data in L2
    const int t = blockIdx.y;
    const int tx = t % p.g.ntx, ty = (t / p.g.ntx) % p.g.nty, b = t / (p.g.ntx * p.g.nty);
    const int unit0 = t * Cfg::UNITS_X;
    const int slot0 = p.lists.unit_start[unit0];
    const int slot1 = min(p.lists.unit_start[unit0 + Cfg::UNITS_X], p.cap);
    if (slot0 >= slot1) return;  // no edge pixel in this tile
    const int which = blockIdx.z;
    // padded coordinates of the tile's first edge-pixel position: (P + ty*TYF, P + tx*TXF - xs); the window's
    // first column is a multiple of 4 (make_geom), as the copy engine requires
    issue_tile_load<Cfg>(tile, &tmap, &tile_bar, Cfg::P + tx * Cfg::TXF - p.g.xs - Cfg::K - Cfg::ICOL0,
                         Cfg::P + ty * Cfg::TYF - Cfg::K - Cfg::P, ((p.img_first + which) * p.g.B + b) * 3);
    if (threadIdx.x <= Cfg::UNITS_X) ustart_s[threadIdx.x] = p.lists.unit_start[unit0 + threadIdx.x];
    const bool staged = slot1 - slot0 <= Cfg::RC_SMEM;
    if (staged)
        for (int i = threadIdx.x; i < slot1 - slot0; i += blockDim.x) rc_s[i] = p.lists.slot_rc[slot0 + i];
    const int32_t* slot_rc = staged ? rc_s : p.lists.slot_rc + slot0;
    __syncthreads();             // barrier initialised (thread 0) and lists staged
    mbar_wait(&tile_bar, 0);     // the tile has landed
    switch (blockIdx.x) {
        case 0: run_group_fwd<Cfg, 0>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 1: if constexpr (Cfg::NDXG > 1) run_group_fwd<Cfg, 1>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 2: if constexpr (Cfg::NDXG > 2) run_group_fwd<Cfg, 2>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 3: if constexpr (Cfg::NDXG > 3) run_group_fwd<Cfg, 3>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 4: if constexpr (Cfg::NDXG > 4) run_group_fwd<Cfg, 4>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 5: if constexpr (Cfg::NDXG > 5) run_group_fwd<Cfg, 5>(p, tile, splanes, ustart_s, slot_rc, which); break;
        case 6: if constexpr (Cfg::NDXG > 6) run_group_fwd<Cfg, 6>(p, tile, splanes, ustart_s, slot_rc, which); break;
        default: break;
    }
}

template <typename Cfg>
constexpr size_t plane_fwd_smem_bytes() {
    return (size_t)(Cfg::TILE_FLOATS + Cfg::NPL * Cfg::SPS + 8) * sizeof(float) +
           (size_t)Cfg::RC_SMEM * sizeof(int32_t);
}

// rows buffer (panel layout) -> rows[i][d] in the reference's row order (i = position in the flat edge list).
__global__ void __launch_bounds__(256) plane_rows_to_reference_kernel(const float* qT, int cap, const int32_t* edges,
                                                                      const int32_t* n_edges_dev, int max_edges,
                                                                      const int32_t* slot_map, int L, float* rows) {
    const int mc = edge_count(n_edges_dev, max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        const int slot = edges[n] >= 0 ? slot_map[edges[n]] : -1;
        for (int d = threadIdx.x; d < L; d += blockDim.x)
            rows[(long long)n * L + d] = slot >= 0 && slot < cap ? qT[qt_index(d, slot, L)] : 0.f;
    }
}

}  // namespace sslb
