#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t kmeans_workspace_bytes(int64_t n_total, int nprob, int K, int D);
void kmeans_set_prefilter(int on);
int kmeans_get_prefilter();
int kmeans_lloyd(const float* X, const int64_t* row_off, int nprob, int K, int D, int max_iter, float* C, int* labels,
                 int* counts, void* ws, int64_t ws_bytes, cudaStream_t st);
}
