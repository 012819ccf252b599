// tmap.cu -- host-side TMA tensor-map encoding without linking libcuda.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {
PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}
}  // namespace

int vqb_encode_tmap_2d(CUtensorMap *map, CUtensorMapDataType dtype, const void *base, uint64_t inner,
                       uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                       CUtensorMapSwizzle swizzle) {
    auto enc = get_encode();
    if (!enc) return VQB_ERR_NO_DEVICE;
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(map, dtype, 2, const_cast<void *>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : VQB_ERR_BAD_ARG;
}

int vqb_encode_tmap_4d(CUtensorMap *map, CUtensorMapDataType dtype, const void *base, const uint64_t dims[4],
                       const uint64_t strides_bytes[3], const uint32_t box[4], const uint32_t elem_strides[4],
                       CUtensorMapSwizzle swizzle) {
    auto enc = get_encode();
    if (!enc) return VQB_ERR_NO_DEVICE;
    cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
    cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
    cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
    cuuint32_t e[4] = {elem_strides[0], elem_strides[1], elem_strides[2], elem_strides[3]};
    CUresult r = enc(map, dtype, 4, const_cast<void *>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : VQB_ERR_BAD_ARG;
}

// rank-`rank` (<= 5) tiled map, element strides 1 (the bf16 kernels' 5-D activation views)
int vqb_encode_tmap_nd(CUtensorMap *map, CUtensorMapDataType dtype, const void *base, int rank,
                       const unsigned long long *dims, const unsigned long long *strides_bytes, const uint32_t *box,
                       CUtensorMapSwizzle swizzle) {
    auto enc = get_encode();
    if (!enc) return VQB_ERR_NO_DEVICE;
    if (rank < 1 || rank > 5) return VQB_ERR_BAD_ARG;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = 1; }
    for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
    CUresult r = enc(map, dtype, (cuuint32_t)rank, const_cast<void *>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : VQB_ERR_BAD_ARG;
}
