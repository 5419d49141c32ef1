#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t compact_workspace_bytes(int64_t n);
int compact_points(const uint8_t* mask, int64_t n, const float* xyz, const float* opacity, const int32_t* idx,
                   float* xyz_out, float* opacity_out, int32_t* idx_out, int64_t* count_host, void* ws,
                   int64_t ws_bytes, cudaStream_t st);
}
