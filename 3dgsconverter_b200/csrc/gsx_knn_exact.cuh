#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t knn_exact_workspace_bytes(int64_t n);
int knn_exact_mean_dists(const float* xyz, int64_t n, int k, float* means, void* ws, int64_t ws_bytes,
                         cudaStream_t st);
}
