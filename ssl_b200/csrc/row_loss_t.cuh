// Row loss on the offset-major rows of the plane path: qT[d][slot].
//
// One block owns 16 consecutive slots and keeps both of their rows (SR and GT, KS*KS entries each)
// in shared memory (80 KB: two blocks per SM overlap one block's loads with the other's arithmetic), so the exp / normalise tail of loss_util.py:234-243, the L1 (basic_loss.py:
// 14-16,59-66) and KL (basic_loss.py:269-282) numerators and the whole adjoint chain down to
// dL/dq cost one read of each rows buffer and one write (dL/dq overwrites q_sr in place).
// It also emits, per slot, the sum of dL/dq over every clip class: the weights of the
// out-of-area terms (similarity.cu:123-124) used by the plane backward.
#pragma once

#include "plane_geom.cuh"
#include "row_ops.cuh"

namespace sslb {

struct RowLossTParams {
    float* qs;             // [L][cap] in: q of SR, out: dL/dq (when want_grad)
    const float* qg;       // [L][cap] q of GT
    const int32_t* slot_pix;
    const int32_t* counts; // counts[0] = slots in use
    int cap, L, KS, P, K;
    float denom, sigma, eps, chain;
    int mode, want_grad;
    float w_l1, w_kl;
    float* gcls;           // [(2K+1)^2][cap] or NULL
    double* scratch;       // [2 * gridDim.x]
};

constexpr int kRowTThreads = 512;
constexpr int kRowTSlots = 16;                       // slots per block: 2 rows x 625 x 16 floats = 80 KB => 2 blocks per SM
constexpr int kRowTPhases = kRowTThreads / kRowTSlots;

__global__ void __launch_bounds__(kRowTThreads) row_loss_t_kernel(RowLossTParams p) {
    extern __shared__ float rl_smem[];
    constexpr int NS = kRowTSlots, NPH = kRowTPhases;
    const int L = p.L, cap = p.cap, mode = p.mode;
    float* es = rl_smem;                // [L][NS]
    float* et = es + L * NS;            // [L][NS]
    __shared__ float red[2][NPH][NS];
    __shared__ double dred[32];
    const int s = threadIdx.x % NS, ph = threadIdx.x / NS;
    const int n_slots = min(p.counts[0], cap);
    const float nscale = -1.0f / (p.denom * p.sigma);
    const float w_l1 = p.w_l1, w_kl = p.w_kl, chain = p.chain, eps = p.eps;
    const long long gstride = (long long)NPH * cap;   // global stride between this thread's offsets
    constexpr int sstride = NPH * NS;                 // shared stride
    const int n_mine = (L - ph + NPH - 1) / NPH;      // offsets d = ph, ph + NPH, ... handled by this thread
    double l1_tot = 0.0, kl_tot = 0.0;
    for (int slot0 = blockIdx.x * NS; slot0 < n_slots; slot0 += gridDim.x * NS) {
        const int slot = slot0 + s;
        const bool valid = slot < n_slots && p.slot_pix[slot] >= 0;
        // pass 1: e = exp(-1 * (q / (C kw^2)) / sigma), partial row sums.  Loads are issued in batches
        // of 2*UN so that enough bytes are in flight to cover the HBM latency.
        float zs = 0.f, zt = 0.f;
        constexpr int UN = 8;
        {
            const float* gs = p.qs + (long long)ph * cap + slot;
            const float* gg = p.qg + (long long)ph * cap + slot;
            float* ps = es + ph * NS + s;
            float* pt = et + ph * NS + s;
            for (int i0 = 0; i0 < n_mine; i0 += UN) {
                float qa[UN], qb[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const bool ok = valid && i0 + u < n_mine;
                    qa[u] = ok ? __ldcs(gs + u * gstride) : 0.f;
                    qb[u] = ok ? __ldcs(gg + u * gstride) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    if (i0 + u < n_mine) {
                        // exp(-1 * (q / (C kw^2)) / sigma) of loss_util.py:224-225; the two divisions are folded
                        // into one multiplication (<= 1.5 ulp of the argument, the size of q's own rounding)
                        const float a = valid ? expf(qa[u] * nscale) : 0.f;
                        const float b = valid ? expf(qb[u] * nscale) : 0.f;
                        ps[u * sstride] = a;
                        pt[u * sstride] = b;
                        zs += a;
                        zt += b;
                    }
                }
                gs += UN * gstride; gg += UN * gstride;
                ps += UN * sstride; pt += UN * sstride;
            }
        }
        red[0][ph][s] = zs;
        red[1][ph][s] = zt;
        __syncthreads();
        float rs = 1.f, rt = 1.f;
        if (mode == SSL_B200_ROWS_NORM) {
            float a = 0.f, b = 0.f;
#pragma unroll
            for (int k = 0; k < NPH; ++k) { a += red[0][k][s]; b += red[1][k][s]; }
            rs = 1.0f / (a + eps);
            rt = 1.0f / (b + eps);
        }
        __syncthreads();
        // pass 2: rows, loss terms, dL/drow; es <- s, et <- g
        float l1 = 0.f, kl = 0.f, dot = 0.f;
        {
            float* ps = es + ph * NS + s;
            float* pt = et + ph * NS + s;
            for (int i = 0; i < n_mine; ++i, ps += sstride, pt += sstride) {
                const float sv = rs * *ps, tv = rt * *pt;
                const float df = sv - tv;
                l1 += fabsf(df);
                float g = df > 0.f ? w_l1 : (df < 0.f ? -w_l1 : 0.f);
                if (w_kl != 0.f) {
                    const float sc = fmaxf(sv, 1e-10f), tc = fmaxf(tv, 1e-10f);
                    kl += kl_term(sc, tc) + (mode == SSL_B200_ROWS_NORM ? ((tc - tv) - (sc - sv)) : (tc - sc));
                    if (sv > 1e-10f) g -= w_kl * tc / sc;
                }
                if (!valid) g = 0.f;
                dot = fmaf(g, sv, dot);
                *ps = sv;
                *pt = g;
            }
        }
        if (valid) { l1_tot += (double)l1; kl_tot += (double)kl; }
        if (p.want_grad) {
            red[0][ph][s] = dot;
            __syncthreads();
            float dsum = 0.f;
            if (mode == SSL_B200_ROWS_NORM) {
#pragma unroll
                for (int k = 0; k < NPH; ++k) dsum += red[0][k][s];
            }
            // pass 3: dL/dq = chain * s * (g - sum_m g_m s_m)   (EXP rows: chain * e * g)
            {
                float* ps = es + ph * NS + s;
                const float* pt = et + ph * NS + s;
                float* gs = p.qs + (long long)ph * cap + slot;
                const bool store = slot < n_slots;
                for (int i = 0; i < n_mine; ++i, ps += sstride, pt += sstride, gs += gstride) {
                    const float gq = chain * *ps * (*pt - dsum);
                    *ps = gq;
                    if (store) *gs = gq;
                }
            }
            __syncthreads();
            // pass 4: per clip class sums of dL/dq (classes with nothing out of area are skipped)
            if (p.gcls && slot < n_slots) {
                const int K = p.K, P = p.P, KS = p.KS;
                const int NC = 2 * K + 1, U = P - K;
                for (int c = ph; c < NC * NC; c += NPH) {
                    const int ca = c / NC, cb = c % NC;
                    float acc = 0.f;
                    if (ca != K || cb != K) {
                        const int dy0 = ca < K ? ca - P : (ca > K ? U + (ca - K) : -U);
                        const int dy1 = ca == K ? U : dy0;
                        const int dx0 = cb < K ? cb - P : (cb > K ? U + (cb - K) : -U);
                        const int dx1 = cb == K ? U : dx0;
                        for (int dy = dy0; dy <= dy1; ++dy) {
                            const float* row = es + ((dy + P) * KS + dx0 + P) * NS + s;
                            for (int dx = dx0; dx <= dx1; ++dx, row += NS) acc += *row;
                        }
                    }
                    p.gcls[(long long)c * cap + slot] = acc;
                }
            }
        }
        __syncthreads();
    }
    l1_tot = block_sum(l1_tot, dred);
    kl_tot = block_sum(kl_tot, dred);
    if (threadIdx.x == 0) {
        p.scratch[2 * blockIdx.x] = l1_tot;
        p.scratch[2 * blockIdx.x + 1] = kl_tot;
    }
}

}  // namespace sslb
