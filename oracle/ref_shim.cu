// C-ABI shim around the UNMODIFIED reference CUDA operator -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
//
// oracle/build_ref.py compiles this file together with the reference's own
//   GAN-Based-SR/basicsr/losses/similarity/similarity.cu      (read where it lies under /root/reference)
// into oracle/_ref/libsimilarity_ref.so.  The reference declares its two host launchers in
// similarity.h:2-23 with C++ linkage and no error reporting; the wrappers below give them C names
// ctypes can bind and return cudaGetLastError() so a failed launch is visible.  Nothing of the
// reference is copied: similarity.h is found through the include path at build time.
//
// The reference launches on the legacy default stream with 16-thread blocks (similarity.cu:66-69,
// 144-147); callers that time it must record their events on that stream.
#include <cuda_runtime.h>

#include "similarity.h"

extern "C" int ref_compute_similarity(const float* image, const int* pos, float* out, int mc, int psize, int ksize,
                                      int height, int width, int channel) {
    _compute_similarity(image, pos, out, mc, psize, ksize, height, width, channel);
    return (int)cudaGetLastError();
}

extern "C" int ref_compute_similarity_backward(const float* image, const float* grads, const int* pos,
                                               float* image_grads, int mc, int psize, int ksize, int height,
                                               int width, int channel) {
    _compute_similarity_backward(image, grads, pos, image_grads, mc, psize, ksize, height, width, channel);
    return (int)cudaGetLastError();
}
