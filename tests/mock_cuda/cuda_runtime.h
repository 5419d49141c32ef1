// Mock CUDA runtime for a HOST-ONLY run of csrc/gsx_hostcopy.cu (tests/test_hostcopy_mock_cpu.py) -- test infrastructure.
// Streams are real FIFO queues drained by one background thread each, with a small artificial delay per operation, so
// that an "async" copy really happens LATER than its enqueue: refilling a pinned chunk before the event of its previous
// DMA, or reading a chunk before its DMA has landed, corrupts the data and the harness sees it.  Events carry the
// (stream, sequence number) of their last record; cudaEventSynchronize / cudaStreamWaitEvent wait for that position.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <mutex>
#include <thread>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline __attribute__((always_inline))
#define __restrict__
struct float4 { float x, y, z, w; };
struct dim3 { unsigned x, y, z; };
static dim3 threadIdx;
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1, cudaDevAttrMultiProcessorCount = 16 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes { cudaMemoryType type; };

struct MockStream {
    std::mutex m;
    std::condition_variable cv_work, cv_done;
    std::deque<std::function<void()>> q;
    unsigned long long enqueued = 0, done = 0;
    MockStream() {
        std::thread([this] {
            for (;;) {
                std::function<void()> op;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv_work.wait(lk, [&] { return !q.empty(); });
                    op = std::move(q.front());
                    q.pop_front();
                }
                std::this_thread::sleep_for(std::chrono::microseconds(150));   // the copy engine is slower than the CPU
                op();
                {
                    std::lock_guard<std::mutex> lk(m);
                    ++done;
                }
                cv_done.notify_all();
            }
        }).detach();
    }
    unsigned long long push(std::function<void()> op) {
        unsigned long long seq;
        {
            std::lock_guard<std::mutex> lk(m);
            q.push_back(std::move(op));
            seq = ++enqueued;
        }
        cv_work.notify_one();
        return seq;
    }
    void wait_for(unsigned long long seq) {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return done >= seq; });
    }
};
typedef MockStream* cudaStream_t;

struct MockEvent {
    std::mutex m;
    MockStream* s = nullptr;        // never recorded: complete
    unsigned long long seq = 0;
};
typedef MockEvent* cudaEvent_t;

static inline MockStream* mock_default_stream() {
    static MockStream* s = new MockStream();
    return s;
}
static inline MockStream* mock_stream(cudaStream_t s) { return s ? s : mock_default_stream(); }

static inline const char* cudaGetErrorString(cudaError_t) { return "mock"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { *p = malloc(n); return *p ? cudaSuccess : cudaErrorUnknown; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new MockStream(); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new MockEvent(); return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
    MockStream* ms = mock_stream(s);
    const unsigned long long seq = ms->push([] {});          // a marker op: complete when everything before it is
    std::lock_guard<std::mutex> lk(e->m);
    e->s = ms;
    e->seq = seq;
    return cudaSuccess;
}
static inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
    MockStream* ms;
    unsigned long long seq;
    {
        std::lock_guard<std::mutex> lk(e->m);
        ms = e->s;
        seq = e->seq;
    }
    if (ms) ms->wait_for(seq);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned) {
    MockStream* es;
    unsigned long long seq;
    {
        std::lock_guard<std::mutex> lk(e->m);
        es = e->s;
        seq = e->seq;
    }
    if (es) mock_stream(s)->push([es, seq] { es->wait_for(seq); });   // the stream stalls until the event's position
    return cudaSuccess;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t s) {
    MockStream* ms = mock_stream(s);
    ms->wait_for(ms->push([] {}));
    return cudaSuccess;
}
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t st) {
    mock_stream(st)->push([d, s, n] { memcpy(d, s, n); });   // happens LATER, when the stream reaches it
    return cudaSuccess;
}
extern int g_mock_pageable;
static inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void*) {
    a->type = g_mock_pageable ? cudaMemoryTypeUnregistered : cudaMemoryTypeHost;
    return cudaSuccess;
}
