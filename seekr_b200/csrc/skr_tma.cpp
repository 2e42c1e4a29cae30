#include "skr_tma.h"

#include <mutex>

namespace skr {

namespace {
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                              const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;
std::once_flag g_once;
}  // namespace

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                 CUtensorMapSwizzle swizzle) {
    std::call_once(g_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = (EncodeFn)fn;
    });
    if (!g_encode) return fail(SKR_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    if (((uintptr_t)base & 15) || (row_pitch_bytes & 15))
        return fail(SKR_ERR_ARG, "TMA needs a 16-byte aligned base and row pitch");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    (void)elem_bytes;
    CUresult r = g_encode(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(SKR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return SKR_OK;
}

}  // namespace skr
