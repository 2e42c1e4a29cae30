// K3 + K4 placeholder (filled in by the Pearson milestone).
#include <cuda_runtime.h>

#include "skr_common.h"

extern "C" int64_t skr_pearson_rows_padded(int64_t rows) { return (rows + 127) / 128 * 128; }
extern "C" int64_t skr_pearson_k_padded(int64_t K) { return (K + 63) / 64 * 64; }
extern "C" int skr_pearson_prepare(const void*, int, int64_t, int64_t, int64_t, int, uint16_t*, uint16_t*, float*, void*) {
    return skr::fail(SKR_ERR_ARG, "skr_pearson_prepare: not built yet");
}
extern "C" int skr_pearson_gemm(const uint16_t*, const uint16_t*, const float*, int64_t, const uint16_t*, const uint16_t*,
                                const float*, int64_t, int64_t, double, void*, int, int64_t, void*) {
    return skr::fail(SKR_ERR_ARG, "skr_pearson_gemm: not built yet");
}
