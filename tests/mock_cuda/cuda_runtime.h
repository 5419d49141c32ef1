// mock CUDA runtime for a host-only thread-sanitizer run of gsx_hostcopy.cu: streams execute immediately
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x, y, z; };
static dim3 threadIdx;
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef struct MockStream* cudaStream_t;
typedef struct MockEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1, cudaDevAttrMultiProcessorCount = 16 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };
struct MockStream { int id; };
struct MockEvent { int id; };
static inline const char* cudaGetErrorString(cudaError_t) { return "mock"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorUnknown; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new MockStream{0}; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new MockEvent{0}; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
extern int g_mock_pageable;
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) { a->type = g_mock_pageable ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost; return cudaSuccess; }
