// gsx_hostcopy.cuh -- multi-threaded pinned staging for copies from / to pageable host buffers (gsx_hostcopy.cu)
#pragma once
#include "gsx_common.cuh"

namespace gsx {
// Enqueue the upload; on return src_host has been read completely and `st` is ordered after the last chunk.
int copy_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
// Download what `st` has produced; BLOCKS until dst_host is complete.
int copy_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);
// Touch the pages of a pageable destination (several threads) while the GPU is still computing; content preserved.
int prefault_host(void* dst_host, size_t bytes);
}  // namespace gsx
