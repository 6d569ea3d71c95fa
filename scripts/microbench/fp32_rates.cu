// Measures per-SM issue rates that bound the SSG kernels on sm_100a: scalar vs packed (f32x2)
// FP32 add / sub / fma, the (sub, fma) pair of the patch distance, and shared-memory loads.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rates fp32_rates.cu && ./fp32_rates
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 16

typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 d; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int MODE>
__global__ void __launch_bounds__(512) k(float* out, float x, float y) {
    __shared__ float sm[4096 + 64];
    for (int i = threadIdx.x; i < 4096 + 64; i += blockDim.x) sm[i] = x * i;
    __syncthreads();
    float acc[NACC];
    u64 acc2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = x + i + threadIdx.x; acc2[i] = pack(acc[i], acc[i] + 1.f); }
    const u64 x2 = pack(x, x), y2 = pack(y, y);
    const int t = threadIdx.x;
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 0) {  // scalar FFMA
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], x, y);
        } else if (MODE == 1) {  // packed FFMA2
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc2[i] = fma2(acc2[i], x2, y2);
        } else if (MODE == 2) {  // scalar FADD
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = acc[i] + x;
        } else if (MODE == 3) {  // packed FADD2
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc2[i] = add2(acc2[i], x2);
        } else if (MODE == 4) {  // scalar (sub, fma) pair: acc += (c - v)^2, v varies
#pragma unroll
            for (int i = 0; i < NACC; ++i) { const float d = x - acc[(i + 1) % NACC]; acc[i] = fmaf(d, d, acc[i]); }
        } else if (MODE == 5) {  // packed (sub2, fma2) pair
#pragma unroll
            for (int i = 0; i < NACC; ++i) { const u64 d = sub2(x2, acc2[(i + 1) % NACC]); acc2[i] = fma2(d, d, acc2[i]); }
        } else if (MODE == 6) {  // LDS.32, conflict-free, + 1 FADD each
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += sm[(t + 32 * i + it) & 4095];
        } else if (MODE == 7) {  // LDS.128 + 4 FADD
#pragma unroll
            for (int i = 0; i < NACC / 4; ++i) {
                const float4 v = *reinterpret_cast<const float4*>(&sm[((t * 4 + 128 * i + 4 * it) & 4095)]);
                acc[4 * i] += v.x; acc[4 * i + 1] += v.y; acc[4 * i + 2] += v.z; acc[4 * i + 3] += v.w;
            }
        } else if (MODE == 8) {  // LDS.32 broadcast (all lanes same address) + FADD
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += sm[(32 * i + it) & 4095];
        } else if (MODE == 9) {  // scalar FMUL
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = acc[i] * x;
        } else if (MODE == 10) {  // FADD and FFMA interleaved 1:1 (do they share a pipe?)
#pragma unroll
            for (int i = 0; i < NACC; i += 2) { acc[i] = acc[i] + x; acc[i + 1] = fmaf(acc[i + 1], x, y); }
        } else if (MODE == 11) {  // packed (sub2,fma2) with scalar unpacked neighbour: sub scalar x2 then fma2
#pragma unroll
            for (int i = 0; i < NACC; ++i) {
                const float d0 = x - lo(acc2[(i + 1) % NACC]), d1 = y - hi(acc2[(i + 1) % NACC]);
                const u64 d = pack(d0, d1);
                acc2[i] = fma2(d, d, acc2[i]);
            }
        } else if (MODE == 12) {  // shuffle + FADD
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], 1);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i] + lo(acc2[i]) + hi(acc2[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_thread_iter, float* out) {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const int blocks = sms * 2, threads = 512;
    k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k<MODE><<<blocks, threads>>>(out, 1.0001f, 0.5f);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double ops = (double)blocks * threads * ITERS * ops_per_thread_iter;
    const double per_s = ops / (best * 1e-3);
    printf("%-44s %8.3f ms  %8.2f Gop/s/SM  = %6.1f lane-ops/clk/SM @%d MHz nominal\n", name, best, per_s / sms / 1e9,
           per_s / sms / (khz * 1e3), khz / 1000);
}

int main() {
    float* out;
    cudaMalloc(&out, 148 * 4 * 1024 * sizeof(float));
    run<0>("FFMA scalar (ops = fma)", NACC, out);
    run<1>("FFMA2 packed (ops = fma, 2 per instr)", 2 * NACC, out);
    run<2>("FADD scalar", NACC, out);
    run<3>("FADD2 packed (2 per instr)", 2 * NACC, out);
    run<9>("FMUL scalar", NACC, out);
    run<10>("FADD+FFMA 1:1 scalar (ops = instr)", NACC, out);
    run<4>("(FSUB,FFMA) scalar (ops = pairs)", NACC, out);
    run<5>("(SUB2,FMA2) packed (ops = pairs, 2/instr)", 2 * NACC, out);
    run<11>("(2xFSUB, pack, FMA2) (ops = pairs)", 2 * NACC, out);
    run<6>("LDS.32 + FADD (ops = loads)", NACC, out);
    run<7>("LDS.128 + 4 FADD (ops = floats loaded)", NACC, out);
    run<8>("LDS.32 broadcast + FADD (ops = loads)", NACC, out);
    run<12>("SHFL + FADD (ops = shuffles)", NACC, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
