// Row-wise pieces of the loss: the L1 / KL between SR and GT similarity rows, the adjoint of the
// exp / normalise tail, and the on-GPU edge mask.
#pragma once

#include "common.cuh"

namespace sslb {

constexpr int kRowThreads = 128;

struct RowLossParams {
    const float* s;   // SR rows
    const float* t;   // GT rows
    float* gq;        // out: dL/dq_sr rows (may be NULL)
    const int32_t* n_edges_dev;
    int max_edges, L;
    float chain;      // -1 / (sigma * C * kw^2)
    int mode;
    float w_l1, w_kl;
    double* scratch;  // [2 * gridDim.x]
};

// One element of KLDistanceLoss (basic_loss.py:269-282), t (log t - log s), in a form that does not
// cancel: with r = (t - s)/s,
//     t log(t/s) = [ s r^2 + t (log1p(r) - r) ] + (t - s).
// The bracket is >= 0 and second-order small when s ~ t, so it is summed as is; the (t - s) parts
// sum to zero over a normalised row up to the 1e-10 clamps, which the caller adds back exactly.
// (Plain logf(t) - logf(s) loses ~1e-7*|log t| / |r| relative accuracy per term.)
__device__ __forceinline__ float kl_term(float sc, float tc) {
    const float r = (tc - sc) / sc;
    if (fabsf(r) < 0.125f) {
        const float h = r * r * (-0.5f + r * (0.33333334f + r * (-0.25f + r * (0.2f + r * (-0.16666667f +
                        r * (0.14285715f + r * (-0.125f + r * 0.11111111f)))))));
        return fmaf(sc * r, r, tc * h);
    }
    return tc * log1pf(r) - (tc - sc);
}

// L1Loss (basic_loss.py:14-16,59-66) and KLDistanceLoss (basic_loss.py:269-282) numerators of a
// block's rows, plus dL/dq of the SR rows:
//   g_s   = w_l1 * sign(s - t) + w_kl * d/ds[ t' (log t' - log s') ],  x' = max(x, 1e-10)
//         = w_l1 * sign(s - t) - w_kl * t'/s' * [s > 1e-10]
//   NORM: g_q = chain * s * (g_s - sum_m g_m s_m)      EXP: g_q = chain * e * g_s
__global__ void __launch_bounds__(kRowThreads) row_loss_kernel(RowLossParams p) {
    __shared__ float red[32];
    __shared__ double dred[32];
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    double l1_tot = 0.0, kl_tot = 0.0;
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        const float* s = p.s + (long long)n * p.L;
        const float* t = p.t + (long long)n * p.L;
        float l1 = 0.f, kl = 0.f, dot = 0.f;
        for (int d = threadIdx.x; d < p.L; d += kRowThreads) {
            const float sv = s[d], tv = t[d];
            const float df = sv - tv;
            l1 += fabsf(df);
            float g = p.w_l1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
            if (p.w_kl != 0.f) {
                const float sc = fmaxf(sv, 1e-10f), tc = fmaxf(tv, 1e-10f);
                // NORM rows sum to one, so sum(t - s) is zero up to the clamps; EXP rows carry it
                kl += kl_term(sc, tc) + (p.mode == SSL_B200_ROWS_NORM ? ((tc - tv) - (sc - sv)) : (tc - sc));
                if (sv > 1e-10f) g -= p.w_kl * tc / sc;
            }
            dot = fmaf(g, sv, dot);
        }
        l1_tot += (double)l1;  // per-thread partials are combined once, after the row loop
        kl_tot += (double)kl;
        if (p.gq) {
            if (p.mode == SSL_B200_ROWS_NORM) dot = block_sum(dot, red); else dot = 0.f;
            float* gq = p.gq + (long long)n * p.L;
            for (int d = threadIdx.x; d < p.L; d += kRowThreads) {
                const float sv = s[d], tv = t[d];
                const float df = sv - tv;
                float g = p.w_l1 * (df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f));
                if (p.w_kl != 0.f && sv > 1e-10f) g -= p.w_kl * fmaxf(tv, 1e-10f) / sv;
                gq[d] = p.chain * sv * (g - dot);
            }
        }
    }
    l1_tot = block_sum(l1_tot, dred);
    kl_tot = block_sum(kl_tot, dred);
    if (threadIdx.x == 0) {
        p.scratch[2 * blockIdx.x] = l1_tot;
        p.scratch[2 * blockIdx.x + 1] = kl_tot;
    }
}

// Fixed-order sum of the per-block partials => bitwise reproducible loss.
__global__ void __launch_bounds__(32) row_loss_finalize_kernel(const double* scratch, int nblocks, double* sums) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) {
        a += scratch[2 * i];
        b += scratch[2 * i + 1];
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) {
        sums[0] += a;
        sums[1] += b;
    }
}

// rows (mode) + dL/drows -> dL/dq, in place over grad_rows.
__global__ void __launch_bounds__(kRowThreads) rows_chain_kernel(const float* rows, float* grows,
                                                                 const int32_t* n_edges_dev, int max_edges, int L,
                                                                 float chain, int mode) {
    __shared__ float red[32];
    const int mc = edge_count(n_edges_dev, max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        const float* s = rows + (long long)n * L;
        float* g = grows + (long long)n * L;
        float dot = 0.f;
        if (mode == SSL_B200_ROWS_NORM) {
            for (int d = threadIdx.x; d < L; d += kRowThreads) dot = fmaf(g[d], s[d], dot);
            dot = block_sum(dot, red);
        }
        for (int d = threadIdx.x; d < L; d += kRowThreads) g[d] = chain * s[d] * (g[d] - dot);
    }
}

// Raw distances -> exp / normalised rows, in place (loss_util.py:234-243); one block per row.
__global__ void __launch_bounds__(kRowThreads) rows_finish_kernel(float* rows, const int32_t* n_edges_dev, int max_edges,
                                                                 int L, float denom, float sigma, float eps, int mode) {
    __shared__ float red[32];
    const int mc = edge_count(n_edges_dev, max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        float* q = rows + (long long)n * L;
        float z = 0.f;
        for (int d = threadIdx.x; d < L; d += kRowThreads) {
            const float e = expf(-1.0f * (q[d] / denom) / sigma);
            q[d] = e;
            z += e;
        }
        if (mode == SSL_B200_ROWS_NORM) {
            z = block_sum(z, red);
            const float r = 1.0f / (z + eps);
            for (int d = threadIdx.x; d < L; d += kRowThreads) q[d] = r * q[d];
        }
    }
}

// generate_mask.py:22-31 on the GT crop.  One thread per pixel.
template <typename T>
__global__ void __launch_bounds__(256) laplacian_mask_kernel(const T* gt, int B, int H, int W, float threshold,
                                                             float* mask) {
    const long long hw = (long long)H * W;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * hw) return;
    const int b = (int)(idx / hw);
    const int r = (int)(idx - b * hw);
    const int y = r / W, x = r - y * W;
    const T* base = gt + (long long)b * 3 * hw;
    auto luma = [&](int yy, int xx) -> int {
        const long long o = (long long)yy * W + xx;
        // dataset tensors are uint8/255; invert that exactly, then PIL's ITU-R 601 integer luma
        const unsigned R = (unsigned)__float2int_rn(fminf(fmaxf(load_as_float(base + o), 0.f), 1.f) * 255.f);
        const unsigned G = (unsigned)__float2int_rn(fminf(fmaxf(load_as_float(base + hw + o), 0.f), 1.f) * 255.f);
        const unsigned Bc = (unsigned)__float2int_rn(fminf(fmaxf(load_as_float(base + 2 * hw + o), 0.f), 1.f) * 255.f);
        return (int)((19595u * R + 38470u * G + 7471u * Bc + 32768u) >> 16);
    };
    // BORDER_REFLECT_101 is the same no-repeat mirror as F.pad(reflect)
    const int yu = reflect_idx(y - 1, H), yd = reflect_idx(y + 1, H);
    const int xl = reflect_idx(x - 1, W), xr = reflect_idx(x + 1, W);
    int v = luma(yu, x) + luma(yd, x) + luma(y, xl) + luma(y, xr) - 4 * luma(y, x);
    v = v < 0 ? 0 : (v > 255 ? 255 : v);  // CV_8U saturation
    mask[idx] = ((float)v > threshold) ? 1.0f : 0.0f;
}

}  // namespace sslb
