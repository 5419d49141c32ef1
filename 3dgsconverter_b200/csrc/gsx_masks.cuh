#pragma once
#include "gsx_common.cuh"
namespace gsx {
int bbox_mask(const float* xyz, int64_t n, const float* lohi, uint8_t* mask, cudaStream_t st);
int alpha_mask(const float* opacity, int64_t n, double logit_thresh, uint8_t* mask, cudaStream_t st);
}
