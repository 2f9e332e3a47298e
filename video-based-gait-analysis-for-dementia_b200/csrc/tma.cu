// Host helper: build 2D TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point,
// so the library does not link libcuda).
#include <cuda.h>

#include "common.cuh"

namespace gait {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
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

// Row-major 2D tensor: `dim0` contiguous elements of `elem_bytes` (4 = float32, 8 = 64-bit words) per row, `dim1` rows
// `row_stride_bytes` apart; box = box0 x box1 elements; out-of-bounds elements read as zero.
int make_tensor_map_2d(void* map, int elem_bytes, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t row_stride_bytes,
                       uint32_t box0, uint32_t box1, bool swizzle128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return GAIT_ERR_CUDA;
    }
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elem_bytes == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(map), dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu) stride=%llu box=(%u,%u)", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)row_stride_bytes, box0, box1);
        return GAIT_ERR_CUDA;
    }
    return GAIT_OK;
}

// 3D tensor: dim0 contiguous floats, dim1 / dim2 with byte strides stride1 / stride2; SWIZZLE_128B when asked
// (box0 * 4 bytes must then be <= 128); out-of-bounds elements read as zero.
int make_tensor_map_3d_f32(void* map, const void* ptr, uint64_t dim0, uint64_t dim1, uint64_t dim2, uint64_t stride1_bytes,
                           uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, bool swizzle128) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point unavailable");
        return GAIT_ERR_CUDA;
    }
    cuuint64_t dims[3] = {dim0, dim1, dim2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {box0, box1, box2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(reinterpret_cast<CUtensorMap*>(map), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(3d) failed (%d) dims=(%llu,%llu,%llu) strides=(%llu,%llu) box=(%u,%u,%u)", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)dim2,
                  (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1, box2);
        return GAIT_ERR_CUDA;
    }
    return GAIT_OK;
}

}  // namespace gait
