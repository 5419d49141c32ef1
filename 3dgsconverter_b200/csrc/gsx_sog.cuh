#pragma once
#include "gsx_common.cuh"
namespace gsx {
int64_t lexsort_workspace_bytes(int64_t n);
int lexsort_zyx(const float* xyz, int64_t n, int32_t* order_out, void* ws, int64_t ws_bytes, cudaStream_t st);
int quantize_to_codebook(const float* vals, int64_t n, const float* codebook_host, int m, uint8_t* labels, void* ws,
                         int64_t ws_bytes, cudaStream_t st);
}
