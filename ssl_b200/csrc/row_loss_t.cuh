// Row loss on the rows buffers of the plane path (panel layout, plane_geom.cuh: qt_index).
//
// One block owns 16 consecutive slots at a time and keeps both of their rows (SR and GT, KS*KS entries each)
// in shared memory, so the exp / normalise tail of loss_util.py:234-243, the L1 (basic_loss.py:14-16,59-66)
// and KL (basic_loss.py:269-282) numerators and the whole adjoint chain down to dL/dq cost one read of
// each rows buffer and one write (dL/dq overwrites q_sr in place).  It also emits, per slot, the weights of
// the out-of-area terms (similarity.cu:123-124) used by the plane backward: the sums of dL/dq over the clip
// classes, folded into one weight per window offset.
//
// The kernel is a pure stream over 0.9 GB.  Every thread loads its ~20 entries of both rows straight into
// registers (40 independent loads in flight, two full 64-byte segments per warp instruction) and keeps them
// there through the three passes a normalised row needs (row sum, sum of g*s, dL/dq); two blocks per SM overlap
// one block's loads with the other's arithmetic.  (Measured alternatives, profiles/r02_optimization_log.md:
// rows round-tripping through shared memory between the passes, and a TMA producer/consumer pipeline with
// {16 x 128} boxes -- the copy engine handles such 64-byte rows at ~1 per 12 cycles per SM, slower than the
// load/store units.)
#pragma once

#include "plane_geom.cuh"
#include "row_ops.cuh"

namespace sslb {

struct RowLossTParams {
    float* qs;             // L x cap (panel layout) in: q of SR, out: dL/dq (when want_grad)
    const float* qg;       // L x cap q of GT
    const int32_t* slot_pix;
    const int32_t* counts; // counts[0] = slots in use
    int cap;
    float denom, sigma, eps, chain;
    int mode, want_grad;
    float w_l1, w_kl;
    float* wtab;           // [cap][KW*KW] out-of-area weight of every window offset, or NULL
    double* scratch;       // [2 * gridDim.x] block partials
    double* terms;         // [2] += sum|d|, sum KL (added by the last block to finish, in block order)
    unsigned int* done;    // block counter, zero before the launch; the last block resets it
};

constexpr int kRowTThreads = 512;
constexpr int kRowTSlots = 16;                       // slots per group: 64-byte segments of every offset row
constexpr int kRowTPhases = kRowTThreads / kRowTSlots;
static_assert(kRowTSlots == kPanel, "a group of the row loss is one panel of the rows buffers");

// exp(x) for x <= 0 as ex2.approx(x * log2 e): the argument is rounded once (as in the folded division below)
// and ex2.approx is good to 2^-22, i.e. ~2e-7 relative -- far inside the 1e-5 budget, and 4x fewer
// instructions than expf in a kernel that is bound by instruction issue.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// sum of N partials (stride apart) as a balanced tree: log2 N dependent additions instead of N
template <int N>
__device__ __forceinline__ float tree_sum(const float* v, int stride) {
    static_assert((N & (N - 1)) == 0, "power of two");
    float t[N];
#pragma unroll
    for (int k = 0; k < N; ++k) t[k] = v[k * stride];
#pragma unroll
    for (int w = 1; w < N; w *= 2)
#pragma unroll
        for (int k = 0; k + w < N; k += 2 * w) t[k] += t[k + w];
    return t[0];
}

template <int KS, int KW, bool HAS_KL>
__global__ void __launch_bounds__(kRowTThreads, 2) row_loss_t_kernel(RowLossTParams p) {
    extern __shared__ __align__(16) float gqbuf[];           // [L][NS] dL/dq of the current group (pass 4)
    constexpr int NS = kRowTSlots, NPH = kRowTPhases;
    constexpr int L = KS * KS, P = KS / 2, K = KW / 2, NC = 2 * K + 1;
    const int cap = p.cap, mode = p.mode;
    constexpr int NWARP = kRowTThreads / 32;
    __shared__ float red[2][NWARP][NS];             // a warp = two phases x 16 slots: one partial per (warp, slot)
    __shared__ float sG[NC * NC][NS], sT[NC * KW][NS], sR[NC][NS], sW[NS * KW * KW];
    __shared__ double dred[32];
    __shared__ int last_flag;
    const int n_slots = min(p.counts[0], cap);
    const int n_groups = (n_slots + NS - 1) / NS;
    const int s = threadIdx.x % NS, ph = threadIdx.x / NS;
    const int lane_ = threadIdx.x & 31, warp_ = threadIdx.x >> 5;
    // e = exp(-1 * (q / (C kw^2)) / sigma) of loss_util.py:224-225 as 2^(q * nscale2): the two divisions and the
    // change of base are folded into one constant (computed in double), so the argument is rounded once
    const float nscale2 = (float)(-1.4426950408889634 / ((double)p.denom * (double)p.sigma));
    const float w_l1 = p.w_l1, w_kl = p.w_kl, chain = p.chain, eps = p.eps;
    constexpr int gstride = NPH * kPanel;             // global stride between this thread's offsets (panel layout)
    constexpr int sstride = NPH * NS;                 // shared stride
    constexpr int NFULL = L / NPH;                    // every thread handles offsets d = ph + NPH * i, i < NFULL,
    const bool has_tail = ph < L - NFULL * NPH;       // and the first L mod NPH phases one more
    constexpr int NV = NFULL + 1;
    double l1_tot = 0.0, kl_tot = 0.0;
    // The thread's ~20 (slot, offset) pairs are loaded straight into registers -- 2 * NV independent loads in
    // flight per thread, every warp instruction two full 64-byte segments -- and stay there through all
    // passes; shared memory only carries the row reductions and dL/dq for the class sums of pass 4.  The loads of
    // the block's NEXT group are issued as soon as the registers are free (after pass 3), so they fly during pass 4.
    float vs[NV], vt[NV];
    bool valid = false;
    auto fetch = [&](int g) {
        const int slot = g * NS + s;
        valid = g < n_groups && slot < n_slots && p.slot_pix[slot] >= 0;
        const float* gs = p.qs + qt_index(ph, slot, L);   // the group is one panel: L * 64 contiguous bytes
        const float* gg = p.qg + qt_index(ph, slot, L);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const bool on = valid && (i < NFULL || has_tail);
            vs[i] = on ? __ldcs(gs + i * gstride) : 0.f;
            vt[i] = on ? __ldcs(gg + i * gstride) : 0.f;
        }
    };
    fetch(blockIdx.x);
    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const int slot = g * NS + s;
        // pass 1: e, partial row sums
        float zs = 0.f, zt = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const bool on = valid && (i < NFULL || has_tail);
            vs[i] = on ? ex2_approx(vs[i] * nscale2) : 0.f;
            vt[i] = on ? ex2_approx(vt[i] * nscale2) : 0.f;
            zs += vs[i];
            zt += vt[i];
        }
        zs += __shfl_xor_sync(0xffffffffu, zs, 16);   // the warp's two phases
        zt += __shfl_xor_sync(0xffffffffu, zt, 16);
        if (lane_ < NS) { red[0][warp_][s] = zs; red[1][warp_][s] = zt; }
        __syncthreads();
        float rs = 1.f, rt = 1.f;
        if (mode == SSL_B200_ROWS_NORM) {
            rs = 1.0f / (tree_sum<NWARP>(&red[0][0][s], NS) + eps);
            rt = 1.0f / (tree_sum<NWARP>(&red[1][0][s], NS) + eps);
        }
        // pass 2: rows, loss terms, dL/drow; vs <- s, vt <- g
        float l1 = 0.f, kl = 0.f, dot = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float sv = rs * vs[i], tv = rt * vt[i];
            const float df = sv - tv;
            l1 += fabsf(df);
            float gg = df > 0.f ? w_l1 : (df < 0.f ? -w_l1 : 0.f);
            if (HAS_KL) {
                const bool on = valid && (i < NFULL || has_tail);
                const float sc = fmaxf(sv, 1e-10f), tc = fmaxf(tv, 1e-10f);
                if (on) {
                    kl += kl_term(sc, tc) + (mode == SSL_B200_ROWS_NORM ? ((tc - tv) - (sc - sv)) : (tc - sc));
                    if (sv > 1e-10f) gg -= w_kl * tc / sc;
                }
            }
            dot = fmaf(gg, sv, dot);
            vs[i] = sv;
            vt[i] = gg;
        }
        if (valid) { l1_tot += (double)l1; kl_tot += (double)kl; }
        __syncthreads();            // every thread has read red (row sums) before it is reused
        if (p.want_grad) {
            dot += __shfl_xor_sync(0xffffffffu, dot, 16);
            if (lane_ < NS) red[0][warp_][s] = dot;
            __syncthreads();
            float dsum = 0.f;
            if (mode == SSL_B200_ROWS_NORM) {
                dsum = tree_sum<NWARP>(&red[0][0][s], NS);
            }
            // pass 3: dL/dq = chain * s * (g - sum_m g_m s_m)   (EXP rows: chain * e * g)
            {
                float* pw = gqbuf + ph * NS + s;
                float* gs = p.qs + qt_index(ph, slot, L);
                const bool store = slot < n_slots;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (i < NFULL || has_tail) {
                        const float gq = chain * vs[i] * (vt[i] - dsum);
                        pw[i * sstride] = gq;
                        if (store) gs[i * gstride] = gq;
                    }
                }
            }
            fetch(g + gridDim.x);
            __syncthreads();
            // pass 4: weights of the out-of-area terms (similarity.cu:123-124: where the neighbour counts as
            // zero only the centre pixel gets 2*I*g).
            //  (a) sG[class][slot] = sum of dL/dq over the offsets of a clip class (classes with nothing out of
            //      area stay zero);
            //  (b) wtab[slot][(a,b)] = sum of sG over the classes for which window offset (a,b) is out of area:
            //      sums over the column classes first (sR: whole row class out, sT: column b out), then over
            //      the row classes.  Rows of 16 slots are contiguous in wtab: written as one coalesced block.
#ifdef SSLB_EXPERIMENT_NOPASS4   // timing experiment only: the row loss without the out-of-area weights
            if (p.wtab && vs[0] == 123.456f) {
#else
            if (p.wtab) {
#endif
                constexpr int U = P - K;
                // (a) the 4K classes of one clipped dy (or dx) against all unclipped dx (dy) are sums of 2U+1 offsets in
                //     a row (column) of the offset grid: one per phase; the 4K^2 corner classes are single offsets
                if (ph < 4 * K) {
                    const int q = ph % (2 * K);                    // which clipped value
                    const int cc = q < K ? q : q + 1;              // its class (the unclipped class K is skipped)
                    const int t = cc < K ? cc - P : U + (cc - K);  // the clipped dy (rows) or dx (columns)
                    const bool row_class = ph < 2 * K;
                    const int d0 = row_class ? (t + P) * KS + (P - U) : (P - U) * KS + t + P;
                    const int stride = row_class ? NS : KS * NS;
                    const float* src = gqbuf + d0 * NS + s;
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < 2 * U + 1; ++i) acc += src[i * stride];
                    sG[row_class ? cc * NC + K : K * NC + cc][s] = acc;
                } else {
                    for (int j = ph - 4 * K; j < 4 * K * K; j += NPH - 4 * K) {
                        const int qa = j / (2 * K), qb = j % (2 * K);
                        const int ca = qa < K ? qa : qa + 1, cb = qb < K ? qb : qb + 1;
                        const int dy = ca < K ? ca - P : U + (ca - K), dx = cb < K ? cb - P : U + (cb - K);
                        sG[ca * NC + cb][s] = gqbuf[((dy + P) * KS + dx + P) * NS + s];
                    }
                }
                __syncthreads();
                // (b) window column b is out of area for the classes cb < -b (b < 0) or cb > 2K - b (b > 0): running
                //     sums over cb from either end; sR = the whole row class
                if (ph < NC) {
                    const int ca = ph;
                    float g9[NC];
#pragma unroll
                    for (int cb = 0; cb < NC; ++cb) g9[cb] = (ca == K && cb == K) ? 0.f : sG[ca * NC + cb][s];
                    float lo = 0.f, hi = 0.f, all = 0.f;
                    sT[ca * KW + K][s] = 0.f;
#pragma unroll
                    for (int m = 1; m <= K; ++m) {
                        lo += g9[m - 1];
                        hi += g9[NC - m];
                        sT[ca * KW + K - m][s] = lo;      // b = -m
                        sT[ca * KW + K + m][s] = hi;      // b = +m
                    }
#pragma unroll
                    for (int cb = 0; cb < NC; ++cb) all += g9[cb];
                    sR[ca][s] = all;
                }
                __syncthreads();
                // (c) window row a is out of area for the classes ca < -a or ca > 2K - a: those count in full (sR),
                //     the others with their out-of-area columns (sT)
                for (int i = ph; i < KW * KW; i += NPH) {
                    const int a = i / KW - K, b = i % KW;
                    float acc = 0.f;
#pragma unroll
                    for (int ca = 0; ca < NC; ++ca) {
                        const bool out = (a < 0 && ca < -a) || (a > 0 && ca > 2 * K - a);
                        acc += out ? sR[ca][s] : sT[ca * KW + b][s];
                    }
                    sW[s * (KW * KW) + i] = acc;
                }
                __syncthreads();
                {
                    const int n_here = min(NS, n_slots - g * NS);
                    float* dst = p.wtab + (long long)g * NS * (KW * KW);
                    for (int i = threadIdx.x; i < n_here * KW * KW; i += kRowTThreads) dst[i] = sW[i];
                }
            }
        } else {
            fetch(g + gridDim.x);
        }
    }
    // block partials (fixed reduction tree) -> scratch
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        l1_tot = warp_sum(l1_tot);
        kl_tot = warp_sum(kl_tot);
        __syncthreads();
        if (lane == 0) { dred[warp] = l1_tot; dred[16 + warp] = kl_tot; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, b = 0.0;
            for (int w = 0; w < kRowTThreads / 32; ++w) { a += dred[w]; b += dred[16 + w]; }
            p.scratch[2 * blockIdx.x] = a;
            p.scratch[2 * blockIdx.x + 1] = b;
            __threadfence();
            last_flag = atomicAdd(p.done, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        // the last block to finish adds all partials in block order => bitwise reproducible loss
        if (last_flag && threadIdx.x < 32) {
            __threadfence();
            double a = 0.0, b = 0.0;
            for (int i = threadIdx.x; i < (int)gridDim.x; i += 32) {
                a += __ldcg(p.scratch + 2 * i);
                b += __ldcg(p.scratch + 2 * i + 1);
            }
            a = warp_sum(a);
            b = warp_sum(b);
            if (threadIdx.x == 0) {
                p.terms[0] += a;
                p.terms[1] += b;
                *p.done = 0u;
            }
        }
    }
}

}  // namespace sslb
