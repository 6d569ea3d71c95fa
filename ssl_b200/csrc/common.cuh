// Shared helpers for the ssl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "ssl_b200.h"

namespace sslb {

// ---- host-side error plumbing -------------------------------------------------------------
inline char* err_buf() {
    static thread_local char buf[512] = "";
    return buf;
}

inline int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

// Kernels launched by this library (ssl_b200_launch_count); bumped at every launch site.
inline std::atomic<uint64_t>& launch_counter() {
    static std::atomic<uint64_t> n{0};
    return n;
}

inline int check_launch(const char* what, int n_kernels = 1) {
    launch_counter().fetch_add((uint64_t)n_kernels, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail((int)e, "%s: %s", what, cudaGetErrorString(e));
    return 0;
}

// ---- optional per-stage timing (ssl_b200_profile_*) ------------------------------------------
// When enabled, every stage of the whole-step entry points is bracketed by CUDA events recorded on
// the launching stream; bench.py reads the per-stage milliseconds of the very launches it timed.
enum Stage {
    kStageEdgeList = 0, kStagePlaneLists, kStageEout, kStagePlaneFwd, kStageRowLoss, kStagePlaneBwdLists,
    kStagePlaneBwd, kStageFinish, kStagePointFwd, kStagePointBwd, kStagePad, kNumStages
};

inline const char* stage_name(int i) {
    static const char* names[kNumStages] = {"edge_list", "plane_lists", "plane_eout", "ssg_plane_fwd", "row_loss",
                                            "plane_bwd_lists", "ssg_plane_bwd", "plane_finish", "ssg_point_fwd",
                                            "ssg_point_bwd", "pad_images"};
    return i >= 0 && i < kNumStages ? names[i] : "?";
}

struct Profiler {
    static constexpr int kMaxRecords = 4096;
    bool enabled = false;
    int n = 0;
    cudaEvent_t start[kMaxRecords], stop[kMaxRecords];
    int stage[kMaxRecords];
    bool created[kMaxRecords] = {};
};

inline Profiler& profiler() {
    static Profiler p;
    return p;
}

struct StageTimer {
    int idx = -1;
    cudaStream_t st;
    StageTimer(int stage, cudaStream_t s) : st(s) {
        Profiler& p = profiler();
        if (!p.enabled || p.n >= Profiler::kMaxRecords) return;
        idx = p.n++;
        if (!p.created[idx]) {
            cudaEventCreate(&p.start[idx]);
            cudaEventCreate(&p.stop[idx]);
            p.created[idx] = true;
        }
        p.stage[idx] = stage;
        cudaEventRecord(p.start[idx], st);
    }
    ~StageTimer() {
        if (idx >= 0) cudaEventRecord(profiler().stop[idx], st);
    }
};

#define SSLB_REQUIRE(cond, ...)                                   \
    do {                                                          \
        if (!(cond)) return sslb::fail(SSL_B200_EINVAL, __VA_ARGS__); \
    } while (0)

#define SSLB_CUDA(call)                                                                  \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) return sslb::fail((int)e_, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

// A second stream of the calling thread on the current device, for small kernels that do not lie on the step's
// critical path (list building next to the padding, the backward's tile lists next to the forward).  Work is
// forked from / joined into the caller's stream with events only, so a step stays capturable into a CUDA graph.
// One lane per (thread, device): the fork / join events are never shared between two callers.
struct SideLane {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};

inline int side_lane(SideLane** out) {
    constexpr int kMaxDev = 64;
    static thread_local SideLane lanes[kMaxDev];
    int dev = 0;
    SSLB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return sslb::fail(SSL_B200_EINVAL, "device ordinal %d out of range", dev);
    SideLane& l = lanes[dev];
    if (!l.stream) {
        SSLB_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        SSLB_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
        SSLB_CUDA(cudaEventCreateWithFlags(&l.join[0], cudaEventDisableTiming));
        SSLB_CUDA(cudaEventCreateWithFlags(&l.join[1], cudaEventDisableTiming));
    }
    *out = &l;
    return 0;
}

struct DeviceInfo {
    int sm_count = 0;
    int max_smem_optin = 0;
};

inline int device_info(DeviceInfo* out) {
    static thread_local int cached_dev = -1;
    static thread_local DeviceInfo cached;
    int dev = 0;
    SSLB_CUDA(cudaGetDevice(&dev));
    if (dev != cached_dev) {
        SSLB_CUDA(cudaDeviceGetAttribute(&cached.sm_count, cudaDevAttrMultiProcessorCount, dev));
        SSLB_CUDA(cudaDeviceGetAttribute(&cached.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        cached_dev = dev;
    }
    *out = cached;
    return 0;
}

// ---- device helpers -----------------------------------------------------------------------

// F.pad(mode="reflect") index map: padded coordinate -> source coordinate (no edge repeat).
__device__ __forceinline__ int reflect_idx(int v, int n) {
    v = v < 0 ? -v : v;
    return v > n - 1 ? 2 * (n - 1) - v : v;
}

template <typename T>
__device__ __forceinline__ float load_as_float(const T* p);
template <>
__device__ __forceinline__ float load_as_float<float>(const float* p) {
    return __ldg(p);
}
template <>
__device__ __forceinline__ float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(__ldg(p));
}
template <>
__device__ __forceinline__ float load_as_float<__half>(const __half* p) {
    return __half2float(__ldg(p));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum with a fixed reduction tree (deterministic).  `red` holds >= 32 elements.
// Every thread gets the total.  Contains two __syncthreads().
template <typename V>
__device__ __forceinline__ V block_sum(V v, V* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();  // protect `red` from a previous use
    if (lane == 0) red[warp] = v;
    __syncthreads();
    V t = lane < nwarp ? red[lane] : V(0);
    return warp_sum(t);
}

// Where an edge pixel is.  Either the batched flat list (b*H*W + y*W + x) or the reference's
// int32 [mc,2] (row, col) list of similaritywrapper.py:67 (single image).
struct EdgeRef {
    const int32_t* flat;
    const int32_t* yx;
};

__device__ __forceinline__ void decode_edge(const EdgeRef& e, int n, int H, int W, int& b, int& y, int& x) {
    if (e.yx) {
        b = 0;
        y = e.yx[2 * n];
        x = e.yx[2 * n + 1];
    } else {
        const int f = e.flat[n];
        const int hw = H * W;
        b = f / hw;
        const int r = f - b * hw;
        y = r / W;
        x = r - y * W;
    }
}

__device__ __forceinline__ int edge_count(const int32_t* n_dev, int max_edges) {
    if (!n_dev) return max_edges;
    const int n = *n_dev;
    return n < max_edges ? n : max_edges;
}

}  // namespace sslb

#define SSLB_DISPATCH_DTYPE(dtype, T, ...)                                          \
    switch (dtype) {                                                                \
        case SSL_B200_F32: { using T = float; __VA_ARGS__; break; }                  \
        case SSL_B200_BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }         \
        case SSL_B200_F16: { using T = __half; __VA_ARGS__; break; }                 \
        default: return sslb::fail(SSL_B200_EINVAL, "unknown dtype %d", (int)(dtype)); \
    }
