// TMA (cp.async.bulk.tensor) + mbarrier plumbing for the plane kernels (sm_100a).
//
// The hot kernels stage their image tile -- [3 channels][rows][cols] of the reflect-padded fp32 copy of
// the batch (pad_images_kernel) -- with ONE 3-D tensor copy per CTA: the copy engine does the address
// arithmetic, fills everything outside the padded image with zeros and signals an mbarrier with the byte
// count, so no thread spends registers or issue slots on the load.  SASS: UTMALDG.3D + SYNCS.
//
// The tensor map is encoded on the host with cuTensorMapEncodeTiled, fetched through
// cudaGetDriverEntryPoint so that the library has no link-time dependency on libcuda.so (it is built, and
// its symbols are checked, on machines without a driver).
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace sslb {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 planes [n_planes][rows][pitch] (pitch in elements, a multiple of 4; `cols` of them valid) seen as a
// 3-D tensor; box = {box_cols, box_rows, box_planes}; out-of-bounds elements read as zero.
inline int make_plane_map(CUtensorMap* map, const float* base, int n_planes, int rows, int cols, int pitch,
                          int box_cols, int box_rows, int box_planes) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(SSL_B200_ENOTSUP, "cuTensorMapEncodeTiled is not available from this driver");
    SSLB_REQUIRE(pitch % 4 == 0 && box_cols % 4 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
                 "tensor map needs 16-byte aligned rows");
    SSLB_REQUIRE(box_cols <= 256 && box_rows <= 256 && box_planes <= 256, "tensor map box too large");
    const cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)n_planes};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * sizeof(float), (cuuint64_t)pitch * rows * sizeof(float)};
    const cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SSL_B200_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// fp32 matrix [rows][pitch] (pitch in elements, a multiple of 4; `cols` of them valid) as a 2-D tensor;
// box = {box_cols, box_rows}; out-of-bounds elements read as zero.
inline int make_matrix_map(CUtensorMap* map, const float* base, int rows, int cols, int pitch, int box_cols,
                           int box_rows) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail(SSL_B200_ENOTSUP, "cuTensorMapEncodeTiled is not available from this driver");
    SSLB_REQUIRE(pitch % 4 == 0 && box_cols % 4 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
                 "tensor map needs 16-byte aligned rows");
    SSLB_REQUIRE(box_cols <= 256 && box_rows <= 256, "tensor map box too large");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SSL_B200_EINVAL, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return 0;
}

// ---- device side ---------------------------------------------------------------------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// makes the initialised barrier visible to the async proxy (the copy engine)
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// orders this thread's generic-proxy accesses to shared memory before later async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// One 3-D tile: box of the map starting at element coordinates (c0 = column, c1 = row, c2 = plane) -> dst.
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 2-D variant (rows buffers of the row-loss kernel)
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}

}  // namespace sslb
