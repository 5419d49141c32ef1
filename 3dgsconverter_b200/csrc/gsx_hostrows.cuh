// gsx_hostrows.cuh -- multi-threaded host-side record movement (gsx_hostrows.cu)
#pragma once
#include "gsx_common.cuh"

namespace gsx {
int host_gather_rows(const void* src, int64_t n_rows, int64_t row_bytes, const int64_t* idx, int64_t m, void* dst);
int host_extract_xyz_opacity(const void* src, int64_t n_rows, int64_t row_bytes, int64_t off_x, int64_t off_y,
                             int64_t off_z, int64_t off_op, float* xyz_out, float* op_out);
}  // namespace gsx
