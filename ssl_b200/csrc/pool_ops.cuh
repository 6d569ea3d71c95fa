// Data movement of the step just upstream of the loss: the paired random crop that carries
// the edge mask (GAN-Based-SR/basicsr/data/transforms.py:93-149) and the training-pair pool
// (GAN-Based-SR/basicsr/models/realesrganssl_model.py:326-367).  Pure copies: results are bit-exact.
#pragma once

#include "common.cuh"

namespace sslb {

// dst[b][c][y][x] = src[b][c][top + y][left + x]   (one launch per tensor; 4-byte elements)
__global__ void __launch_bounds__(256) crop_kernel(const uint32_t* src, uint32_t* dst, int planes, int H, int W, int top,
                                                   int left, int h, int w) {
    const long long n = (long long)planes * h * w;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w), y = (int)((i / w) % h);
        const long long pl = i / ((long long)w * h);
        dst[i] = __ldg(src + (pl * H + top + y) * W + left + x);
    }
}

// Pool exchange of one tensor: for every sample i < b
//     out[i]            = queue[slots[i]]      (skipped when out == NULL: enqueue only)
//     queue[slots[i]]   = in[i]
// `in` may hold ONE channel per sample that is broadcast over the queue's channels (bcast = number of queue
// channels, plane = elements of one channel): the reference allocates the mask queue with the GT's three
// channels and assigns the 1-channel mask into it (realesrganssl_model.py:339-341,357).
__global__ void __launch_bounds__(256) pool_exchange_kernel(uint32_t* queue, const uint32_t* in, uint32_t* out,
                                                            const int32_t* slots, int b, long long sample, int bcast,
                                                            long long plane) {
    const long long n = (long long)b * sample;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = (int)(i / sample);
        const long long r = i - (long long)s * sample;
        uint32_t* q = queue + (long long)slots[s] * sample + r;
        const uint32_t v = bcast > 1 ? __ldg(in + (long long)s * plane + r % plane) : __ldg(in + i);
        if (out) out[i] = *q;
        *q = v;
    }
}

}  // namespace sslb
