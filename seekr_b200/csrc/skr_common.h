// Shared host-side helpers of libseekr_b200: error text, launch counter, CUDA checks.
#pragma once

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "seekr_b200.h"

namespace skr {

std::string& last_error();
int64_t& launch_counter();
int fail(int code, const char* fmt, ...);

}  // namespace skr

#define SKR_CUDA_CHECK(expr)                                                                         \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return skr::fail(SKR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                             __FILE__, __LINE__);                                                    \
    } while (0)

#define SKR_LAUNCH_CHECK()                       \
    do {                                         \
        skr::launch_counter() += 1;              \
        SKR_CUDA_CHECK(cudaGetLastError());      \
    } while (0)
