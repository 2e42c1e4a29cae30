// Host-side construction of TMA tensor maps without linking libcuda: cuTensorMapEncodeTiled is
// looked up through the runtime's driver entry-point query.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include "skr_common.h"

namespace skr {

// 2-D row-major tensor [outer][inner] with a row pitch in bytes; box = [box_outer][box_inner].
int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer,
                 CUtensorMapSwizzle swizzle);

}  // namespace skr
