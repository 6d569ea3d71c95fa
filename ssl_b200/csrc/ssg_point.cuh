// "Point" kernels: one CTA per edge pixel.
//
// The k_s x k_s x C search tile around the edge pixel is staged once in shared memory (reflect
// pad of loss_util.py:189-191 applied by index mapping, surrounded by K zero rows/columns so the
// zero-padded window unfold of loss_util.py:208-209 -- the "out of area => neighbour = 0" branch
// of similarity.cu:43-47 -- falls out of the data layout instead of a branch).  All k_s^2 patch
// distances of that pixel are then computed from shared memory.
//
//   q[n,i,j] = sum_c sum_{a,b} ( T[c][P+a][P+b] - Tz[c][i+a][j+b] )^2        (similarity.cu:5-54)
//
// with T the tile, Tz the same tile with zeros outside [0,k_s)^2, P = k_s/2, a,b in [-K,K].
//
// These kernels serve any odd k_s / k_w (generic versions) plus a register-tiled forward for the
// shipped configuration k_s=25, k_w=9.  They are the right tool for sparse masks and for the
// reference-convention entry points; dense masks go through the tile-sharing kernels (ssg_tile.cuh).
#pragma once

#include "common.cuh"

namespace sslb {

struct PointParams {
    const void* img[2];   // [B,C,H,W]; second image optional (blockIdx.y selects)
    float* rows[2];       // [max_edges, ks*ks]
    EdgeRef edges;
    const int32_t* n_edges_dev;
    int max_edges;
    int B, C, H, W;
    int ks, kw;
    float sigma, eps;
    int mode;             // SSL_B200_ROWS_*
    // backward only
    const float* gq;      // [max_edges, ks*ks] dL/dq
    float* grad;          // [B,C,H,W] fp32, accumulated with atomics
};

// Stage the search tile of edge pixel (b,py,px) into `tile` [C][TP][pitch] interior.
template <typename T>
__device__ __forceinline__ void load_tile(const PointParams& p, const T* img, float* tile, int b, int py, int px,
                                          int ks, int K, int TP, int pitch) {
    const int P = ks / 2, L = ks * ks;
    for (int idx = threadIdx.x; idx < p.C * L; idx += blockDim.x) {
        const int c = idx / L, r = idx - c * L;
        const int ty = r / ks, tx = r - ty * ks;
        const int sy = reflect_idx(py - P + ty, p.H), sx = reflect_idx(px - P + tx, p.W);
        tile[(c * TP + K + ty) * pitch + K + tx] =
            load_as_float(img + (((long long)b * p.C + c) * p.H + sy) * p.W + sx);
    }
}

// Turn the k_s^2 raw distances held in rowbuf into the requested row and store it.
__device__ __forceinline__ void finish_row(const PointParams& p, float* rowbuf, float* red, float* out, int L) {
    if (p.mode == SSL_B200_ROWS_RAW) {
        for (int d = threadIdx.x; d < L; d += blockDim.x) out[d] = rowbuf[d];
        return;
    }
    // loss_util.py:234-243: q/(C*kw^2), exp(-1*q/sigma), (1/(sum+eps))*q
    const float denom = (float)p.C * (float)(p.kw * p.kw);
    float z = 0.f;
    for (int d = threadIdx.x; d < L; d += blockDim.x) {
        const float e = expf(-1.0f * (rowbuf[d] / denom) / p.sigma);
        rowbuf[d] = e;
        z += e;
    }
    if (p.mode == SSL_B200_ROWS_NORM) {
        z = block_sum(z, red);
        const float r = 1.0f / (z + p.eps);
        for (int d = threadIdx.x; d < L; d += blockDim.x) out[d] = r * rowbuf[d];
    } else {
        for (int d = threadIdx.x; d < L; d += blockDim.x) out[d] = rowbuf[d];
    }
}

// ---- forward, any odd ks/kw ---------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) ssg_point_fwd_generic(PointParams p) {
    extern __shared__ float smem[];
    __shared__ float red[32];
    const int ks = p.ks, kw = p.kw, P = ks / 2, K = kw / 2, TP = ks + 2 * K, pitch = TP | 1, L = ks * ks;
    float* tile = smem;
    float* rowbuf = tile + p.C * TP * pitch;
    const T* img = static_cast<const T*>(blockIdx.y ? p.img[1] : p.img[0]);
    float* rows = blockIdx.y ? p.rows[1] : p.rows[0];
    for (int i = threadIdx.x; i < p.C * TP * pitch; i += blockDim.x) tile[i] = 0.f;
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        int b, py, px;
        decode_edge(p.edges, n, p.H, p.W, b, py, px);
        __syncthreads();  // previous row finished with tile/rowbuf (and the zero fill is visible)
        load_tile(p, img, tile, b, py, px, ks, K, TP, pitch);
        __syncthreads();
        for (int d = threadIdx.x; d < L; d += blockDim.x) {
            const int i = d / ks, j = d - i * ks;
            float acc = 0.f;
            for (int c = 0; c < p.C; ++c) {
                const float* tc = tile + c * TP * pitch;
                for (int aa = 0; aa < kw; ++aa) {
                    const float* cen = tc + (P + aa) * pitch + P;
                    const float* nb = tc + (i + aa) * pitch + j;
                    for (int bb = 0; bb < kw; ++bb) {
                        const float t = cen[bb] - nb[bb];
                        acc = fmaf(t, t, acc);
                    }
                }
            }
            rowbuf[d] = acc;
        }
        __syncthreads();
        finish_row(p, rowbuf, red, rows + (long long)n * L, L);
    }
}

// ---- forward, register-tiled for a fixed (KS, KW) -----------------------------------------
// Thread (i, s) owns search offsets (i, s*RD .. s*RD+RD-1).  Per window row it loads KW centre
// values (warp-uniform => broadcast) and RD+KW-1 neighbour values and feeds RD*KW
// (subtract, fma) pairs from registers: ~0.5 shared loads per term pair instead of 2.
// PITCH is chosen so that PITCH*i + RD*s is a bijection mod 32 over a warp (no bank conflicts).
template <typename T, int KS, int KW, int RD, int PITCH>
__global__ void __launch_bounds__(((KS * (KS / RD) + 31) / 32) * 32) ssg_point_fwd_tiled(PointParams p) {
    static_assert(KS % RD == 0, "strip width must divide k_s");
    constexpr int P = KS / 2, K = KW / 2, TP = KS + 2 * K, NS = KS / RD, L = KS * KS;
    static_assert(PITCH >= TP, "pitch too small");
    extern __shared__ float smem[];
    __shared__ float red[32];
    float* tile = smem;
    float* rowbuf = tile + p.C * TP * PITCH;
    const T* img = static_cast<const T*>(blockIdx.y ? p.img[1] : p.img[0]);
    float* rows = blockIdx.y ? p.rows[1] : p.rows[0];
    for (int i = threadIdx.x; i < p.C * TP * PITCH; i += blockDim.x) tile[i] = 0.f;
    const int ti = threadIdx.x / NS, ts = threadIdx.x - ti * NS;
    const bool active = ti < KS;
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        int b, py, px;
        decode_edge(p.edges, n, p.H, p.W, b, py, px);
        __syncthreads();
        load_tile(p, img, tile, b, py, px, KS, K, TP, PITCH);
        __syncthreads();
        if (active) {
            float acc[RD];
#pragma unroll
            for (int k = 0; k < RD; ++k) acc[k] = 0.f;
            for (int c = 0; c < p.C; ++c) {
                const float* tc = tile + c * TP * PITCH;
                float accc[RD];
#pragma unroll
                for (int k = 0; k < RD; ++k) accc[k] = 0.f;
#pragma unroll
                for (int aa = 0; aa < KW; ++aa) {
                    float cen[KW], nb[RD + KW - 1];
                    const float* cp = tc + (P + aa) * PITCH + P;
                    const float* np = tc + (ti + aa) * PITCH + ts * RD;
#pragma unroll
                    for (int bb = 0; bb < KW; ++bb) cen[bb] = cp[bb];
#pragma unroll
                    for (int m = 0; m < RD + KW - 1; ++m) nb[m] = np[m];
#pragma unroll
                    for (int bb = 0; bb < KW; ++bb)
#pragma unroll
                        for (int k = 0; k < RD; ++k) {
                            const float t = cen[bb] - nb[k + bb];
                            accc[k] = fmaf(t, t, accc[k]);
                        }
                }
#pragma unroll
                for (int k = 0; k < RD; ++k) acc[k] += accc[k];
            }
#pragma unroll
            for (int k = 0; k < RD; ++k) rowbuf[ti * KS + ts * RD + k] = acc[k];
        }
        __syncthreads();
        finish_row(p, rowbuf, red, rows + (long long)n * L, L);
    }
}

// ---- backward, any odd ks/kw --------------------------------------------------------------
// Gather form of similarity.cu:73-131 (no shared-memory atomics, 1 global atomic per tile element):
//   every term t = 2 g[i,j] (T[c][P+a][P+b] - Tz[c][i+a][j+b]) adds +t to the centre-window pixel
//   and -t to the neighbour pixel.  Per tile pixel u (as neighbour, only in-area terms exist):
//       GA[c][u]     = 2 sum_{a,b} g[u - (a,b)] * (T[c][u] - T[c][P+a][P+b])
//   and per centre-window pixel (a,b):
//       GB[c][a][b]  = 2 sum_{i,j} g[i,j] * (T[c][P+a][P+b] - Tz[c][i+a][j+b])
//   (out-of-area neighbours are the zeros of Tz, giving the 2*centre*g branch of similarity.cu:123-124).
// The result is added to the fp32 image gradient at the pixel each tile element mirrors to,
// which is the adjoint of the reflect pad (similaritywrapper.py:64).
template <typename T>
__global__ void __launch_bounds__(256) ssg_point_bwd_generic(PointParams p) {
    extern __shared__ float smem[];
    const int ks = p.ks, kw = p.kw, P = ks / 2, K = kw / 2, TP = ks + 2 * K, pitch = TP | 1, L = ks * ks;
    float* tile = smem;                       // [C][TP][pitch], zero border
    float* gpad = tile + p.C * TP * pitch;    // [TP][pitch], g at (K+i, K+j), zero border
    const T* img = static_cast<const T*>(p.img[0]);
    for (int i = threadIdx.x; i < (p.C + 1) * TP * pitch; i += blockDim.x) smem[i] = 0.f;
    const int mc = edge_count(p.n_edges_dev, p.max_edges);
    for (int n = blockIdx.x; n < mc; n += gridDim.x) {
        int b, py, px;
        decode_edge(p.edges, n, p.H, p.W, b, py, px);
        __syncthreads();
        load_tile(p, img, tile, b, py, px, ks, K, TP, pitch);
        for (int d = threadIdx.x; d < L; d += blockDim.x) {
            const int i = d / ks, j = d - i * ks;
            gpad[(K + i) * pitch + K + j] = __ldg(p.gq + (long long)n * L + d);
        }
        __syncthreads();
        float* gimg = p.grad + (long long)b * p.C * p.H * p.W;
        // GA: neighbour side
        for (int u = threadIdx.x; u < L; u += blockDim.x) {
            const int uy = u / ks, ux = u - uy * ks;
            const int sy = reflect_idx(py - P + uy, p.H), sx = reflect_idx(px - P + ux, p.W);
            for (int c = 0; c < p.C; ++c) {
                const float* tc = tile + c * TP * pitch;
                const float tu = tc[(K + uy) * pitch + K + ux];
                float acc = 0.f;
                for (int aa = 0; aa < kw; ++aa) {
                    const float* cen = tc + (P + aa) * pitch + P;
                    const float* g = gpad + (uy - aa + 2 * K) * pitch + ux + 2 * K;  // g[uy-a][ux-b], b = bb-K
                    for (int bb = 0; bb < kw; ++bb) acc = fmaf(g[-bb], tu - cen[bb], acc);
                }
                atomicAdd(gimg + ((long long)c * p.H + sy) * p.W + sx, 2.f * acc);
            }
        }
        // GB: centre-window side
        for (int o = threadIdx.x; o < p.C * kw * kw; o += blockDim.x) {
            const int c = o / (kw * kw), r = o - c * kw * kw;
            const int aa = r / kw, bb = r - aa * kw;
            const float* tc = tile + c * TP * pitch;
            const float cen = tc[(P + aa) * pitch + P + bb];
            float acc = 0.f;
            for (int i = 0; i < ks; ++i) {
                const float* g = gpad + (K + i) * pitch + K;
                const float* nb = tc + (i + aa) * pitch + bb;
                for (int j = 0; j < ks; ++j) acc = fmaf(g[j], cen - nb[j], acc);
            }
            const int sy = reflect_idx(py + aa - K, p.H), sx = reflect_idx(px + bb - K, p.W);
            atomicAdd(gimg + ((long long)c * p.H + sy) * p.W + sx, 2.f * acc);
        }
    }
}

}  // namespace sslb
