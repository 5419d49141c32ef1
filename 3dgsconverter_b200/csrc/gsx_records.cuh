#pragma once
#include "gsx_common.cuh"
namespace gsx {
int records_extract_xyz_opacity(const float* rows, int64_t n, int F, int cx, int cy, int cz, int cop, float* xyz,
                                float* opacity, cudaStream_t st);
int records_gather_rows(const float* rows, const int32_t* idx, int64_t m, int F, float* out, cudaStream_t st);
int records_color_rgba8(const float* rows, int64_t n, int F, int c0, int c1, int c2, int cop, float scale, uint8_t* rgba,
                        cudaStream_t st);
int records_scale_exp(const float* rows, int64_t n, int F, int s0, int s1, int s2, float* out, cudaStream_t st);
}
