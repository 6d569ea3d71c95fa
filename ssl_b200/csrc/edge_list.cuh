// Edge list of a batch of masks, built on the device with no host round trip.
//
// Replaces, for the batched path, the reference's per-image `nonzero(mask_pad == 1)`
// (similaritywrapper.py:64-68), `torch.where(mask == 1)` (loss_util.py:195-196), the
// `mask_stride` product (realesrganssl_model.py:64-72,385-386) and the `mask.sum() == 0` test
// (:387), each of which costs a device->host sync per image in the reference.
//
// Three launches: per-chunk count -> single-block scan of the chunk counts -> ordered emit.
// Order of the output = ascending flat pixel index = the reference's row order.
#pragma once

#include "common.cuh"

namespace sslb {

constexpr int kElThreads = 256;
constexpr int kElPasses = 8;
constexpr int kElChunk = kElThreads * kElPasses;  // mask pixels per block

struct EdgeListParams {
    const float* mask;
    int mask_channels, H, W, stride;
    long long n_pixels;
    int32_t* edges;
    int capacity;
    int32_t* counts;         // [2 + B]
    int32_t* chunk_counts;   // [n_chunks]
    int32_t* chunk_offsets;  // [n_chunks]
    int n_chunks;
};

__device__ __forceinline__ bool is_edge_pixel(const EdgeListParams& p, long long flat, int& b) {
    const long long hw = (long long)p.H * p.W;
    b = (int)(flat / hw);
    const int r = (int)(flat - (long long)b * hw);
    const float v = __ldg(p.mask + ((long long)b * p.mask_channels) * hw + r);
    if (v != 1.0f) return false;  // exact compare, like the reference
    if (p.stride > 1) {
        const int y = r / p.W, x = r - y * p.W;
        return (y % p.stride) == (x % p.stride);
    }
    return true;
}

__global__ void __launch_bounds__(kElThreads) edge_count_kernel(EdgeListParams p) {
    __shared__ int red[32];
    const long long base = (long long)blockIdx.x * kElChunk;
    int local = 0;
#pragma unroll
    for (int k = 0; k < kElPasses; ++k) {
        const long long flat = base + k * kElThreads + threadIdx.x;
        int b = 0;
        const bool e = flat < p.n_pixels && is_edge_pixel(p, flat, b);
        local += e;
        // per-image totals: one atomic per warp when the warp sits inside one image
        const unsigned ball = __ballot_sync(0xffffffffu, e);
        const int b0 = __shfl_sync(0xffffffffu, b, 0), b31 = __shfl_sync(0xffffffffu, b, 31);
        if (b0 == b31) {
            if ((threadIdx.x & 31) == 0 && ball) atomicAdd(p.counts + 2 + b0, __popc(ball));
        } else if (e) {
            atomicAdd(p.counts + 2 + b, 1);
        }
    }
    // block total
    int v = local;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < kElThreads / 32 ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0) p.chunk_counts[blockIdx.x] = t;
    }
}

// Number of edge pixels only (the host entry sizes its workspace from it): total[0] += count.
__global__ void __launch_bounds__(kElThreads) mask_count_kernel(EdgeListParams p, int32_t* total) {
    __shared__ int red[32];
    int local = 0;
    for (long long flat = (long long)blockIdx.x * kElThreads + threadIdx.x; flat < p.n_pixels;
         flat += (long long)gridDim.x * kElThreads) {
        int b = 0;
        local += is_edge_pixel(p, flat, b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        int t = threadIdx.x < kElThreads / 32 ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (threadIdx.x == 0 && t) atomicAdd(total, t);
    }
}

// One block: exclusive scan of chunk_counts -> chunk_offsets, totals -> counts[0..1].
__global__ void __launch_bounds__(1024) edge_scan_kernel(EdgeListParams p) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int start = 0; start < p.n_chunks; start += 1024) {
        const int i = start + threadIdx.x;
        const int v = i < p.n_chunks ? p.chunk_counts[i] : 0;
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
            warp_tot[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (warp ? warp_tot[warp - 1] : 0) + inc - v;
        if (i < p.n_chunks) p.chunk_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int total = carry_s;
        p.counts[0] = total < p.capacity ? total : p.capacity;
        p.counts[1] = total;
    }
}

__global__ void __launch_bounds__(kElThreads) edge_emit_kernel(EdgeListParams p) {
    __shared__ int warp_tot[kElThreads / 32];
    const long long base = (long long)blockIdx.x * kElChunk;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int running = p.chunk_offsets[blockIdx.x];
    if (p.chunk_counts[blockIdx.x] == 0) return;
#pragma unroll 1
    for (int k = 0; k < kElPasses; ++k) {
        const long long flat = base + k * kElThreads + threadIdx.x;
        int b = 0;
        const bool e = flat < p.n_pixels && is_edge_pixel(p, flat, b);
        const unsigned ball = __ballot_sync(0xffffffffu, e);
        if (lane == 0) warp_tot[warp] = __popc(ball);
        __syncthreads();
        int before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kElThreads / 32; ++w) {
            const int t = warp_tot[w];
            before += w < warp ? t : 0;
            total += t;
        }
        if (e) {
            const int slot = running + before + __popc(ball & ((1u << lane) - 1u));
            if (slot < p.capacity) p.edges[slot] = (int32_t)flat;
        }
        running += total;
        __syncthreads();
    }
}

}  // namespace sslb
