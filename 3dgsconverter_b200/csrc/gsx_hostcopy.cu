// gsx_hostcopy.cu -- host <-> device copies for the *_host entry points when the caller's buffer is PAGEABLE.
//
// The reference's call sites hand libgsx ordinary NumPy arrays (np.column_stack at data_processor.py:139, the
// shN block at sog.py:536-549).  cudaMemcpyAsync from pageable memory is staged by the driver through one
// internal pinned buffer on ONE CPU thread: ~11 GB/s measured on the B200 box (profiles/r02_bench_n1.json:
// 19.6 ms end to end from a pageable cloud against 11.1 ms from a pinned one), i.e. the PCIe Gen5 link
// (~55 GB/s) idles 80 % of the time and the upload costs more than the whole filter.
//
// Here T host threads each own two pinned chunks and one copy stream: a thread memcpy()s its next chunk of
// the user's buffer into a free pinned chunk and enqueues the DMA itself, so T memcpys and the DMAs of earlier
// chunks overlap.  D2H is the mirror image (DMA into a pinned chunk, memcpy out while the next DMA runs).
// A buffer that is already pinned / registered / managed goes straight to cudaMemcpyAsync.
#include "gsx_hostcopy.cuh"

#include <algorithm>
#include <atomic>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <thread>
#include <vector>

namespace gsx {

namespace {

constexpr size_t kChunk = 4u << 20;          // 4 MiB per pinned chunk: ~75 us of PCIe time, ~0.4 ms of one memcpy thread
constexpr size_t kStagedMin = 8u << 20;      // below this the plain path is as fast
constexpr int kMaxThreads = 16;

struct Pool {
    std::mutex mu;                 // one staged copy at a time (the pinned chunks are shared state)
    int dev = -1;
    int T = 0;
    char* pinned = nullptr;        // 2 * T chunks
    cudaStream_t streams[kMaxThreads] = {};
    cudaEvent_t ev[2 * kMaxThreads] = {};
    cudaEvent_t ev_start = nullptr;
    bool failed = false;           // creation failed once: stay on the plain path
};
Pool g_pool;

int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s && *s ? atoi(s) : dflt;
}

// (re)create the pool for the current device; false -> use the plain path
bool pool_ready(Pool& p, int dev) {
    if (p.failed) return false;
    if (p.pinned && p.dev == dev) return true;
    if (p.pinned && p.dev != dev) {   // another device became current: streams / events belong to a device
        int cur = dev;
        cudaSetDevice(p.dev);
        for (int t = 0; t < p.T; ++t) cudaStreamDestroy(p.streams[t]);
        for (int i = 0; i < 2 * p.T; ++i) cudaEventDestroy(p.ev[i]);
        cudaEventDestroy(p.ev_start);
        cudaSetDevice(cur);
        p.dev = -1;
    }
    if (!p.pinned) {
        unsigned hw = std::thread::hardware_concurrency();
        int T = env_int("GSX_COPY_THREADS", hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1));
        p.T = std::max(1, std::min(T, kMaxThreads));
        if (cudaHostAlloc((void**)&p.pinned, 2 * (size_t)p.T * kChunk, cudaHostAllocPortable) != cudaSuccess) {
            cudaGetLastError();
            p.pinned = nullptr;
            p.failed = true;
            return false;
        }
    }
    bool ok = true;
    for (int t = 0; t < p.T && ok; ++t) ok = cudaStreamCreateWithFlags(&p.streams[t], cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 * p.T && ok; ++i) ok = cudaEventCreateWithFlags(&p.ev[i], cudaEventDisableTiming) == cudaSuccess;
    ok = ok && cudaEventCreateWithFlags(&p.ev_start, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        p.failed = true;
        return false;
    }
    p.dev = dev;
    return true;
}

// pageable (unregistered) host memory?  Anything the driver already knows (pinned, registered, managed) and any
// failure to tell take the plain path.
bool is_pageable(const void* host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

bool staged_enabled() {
    static int v = -1;
    if (v < 0) v = env_int("GSX_STAGED_COPY", 1) != 0;
    return v != 0;
}

template <typename F>
cudaError_t run_workers(Pool& p, F&& body) {
    std::atomic<int> err{(int)cudaSuccess};
    std::vector<std::thread> th;
    th.reserve(p.T - 1);
    auto wrapped = [&](int t) {
        cudaError_t e = cudaSetDevice(p.dev);
        if (e == cudaSuccess) e = body(t);
        if (e != cudaSuccess) {
            int expect = (int)cudaSuccess;
            err.compare_exchange_strong(expect, (int)e);
        }
    };
    for (int t = 1; t < p.T; ++t) th.emplace_back(wrapped, t);
    wrapped(0);
    for (auto& x : th) x.join();
    return (cudaError_t)err.load();
}

}  // namespace

int copy_h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return GSX_OK;
    int dev = 0;
    GSX_CUDA_CHECK(cudaGetDevice(&dev));
    if (bytes < kStagedMin || !staged_enabled() || !is_pageable(src_host)) {
        GSX_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return GSX_OK;
    }
    Pool& p = g_pool;
    std::lock_guard<std::mutex> lock(p.mu);
    if (!pool_ready(p, dev)) {
        GSX_CUDA_CHECK(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st));
        return GSX_OK;
    }
    GSX_NVTX("gsx_copy_h2d_staged");
    // dst may be a stream-ordered allocation of `st`: the copy streams start after everything queued on st so far
    GSX_CUDA_CHECK(cudaEventRecord(p.ev_start, st));
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    const char* src = (const char*)src_host;
    char* dst = (char*)dst_dev;
    cudaError_t e = run_workers(p, [&](int t) -> cudaError_t {
        cudaError_t r = cudaStreamWaitEvent(p.streams[t], p.ev_start, 0);
        if (r != cudaSuccess) return r;
        int flip = 0;
        for (size_t c = (size_t)t; c < nchunks; c += (size_t)p.T, flip ^= 1) {
            const int slot = 2 * t + flip;
            char* stage = p.pinned + (size_t)slot * kChunk;
            if ((r = cudaEventSynchronize(p.ev[slot])) != cudaSuccess) return r;   // the chunk's previous DMA has drained
            const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
            memcpy(stage, src + off, len);
            if ((r = cudaMemcpyAsync(dst + off, stage, len, cudaMemcpyHostToDevice, p.streams[t])) != cudaSuccess) return r;
            if ((r = cudaEventRecord(p.ev[slot], p.streams[t])) != cudaSuccess) return r;
        }
        return cudaSuccess;
    });
    if (e != cudaSuccess) {
        set_error("staged H2D copy of %zu bytes -> %s", bytes, cudaGetErrorString(e));
        return GSX_ERR_CUDA;
    }
    // st continues once every chunk has landed (the last event of a stream covers its earlier copies)
    for (int i = 0; i < 2 * p.T; ++i) GSX_CUDA_CHECK(cudaStreamWaitEvent(st, p.ev[i], 0));
    return GSX_OK;
}

int copy_d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return GSX_OK;
    int dev = 0;
    GSX_CUDA_CHECK(cudaGetDevice(&dev));
    if (bytes < kStagedMin || !staged_enabled() || !is_pageable(dst_host)) {
        GSX_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaStreamSynchronize(st));
        return GSX_OK;
    }
    Pool& p = g_pool;
    std::lock_guard<std::mutex> lock(p.mu);
    if (!pool_ready(p, dev)) {
        GSX_CUDA_CHECK(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaStreamSynchronize(st));
        return GSX_OK;
    }
    GSX_NVTX("gsx_copy_d2h_staged");
    GSX_CUDA_CHECK(cudaEventRecord(p.ev_start, st));   // the producer of src_dev
    const size_t nchunks = (bytes + kChunk - 1) / kChunk;
    const char* src = (const char*)src_dev;
    char* dst = (char*)dst_host;
    cudaError_t e = run_workers(p, [&](int t) -> cudaError_t {
        cudaError_t r = cudaStreamWaitEvent(p.streams[t], p.ev_start, 0);
        if (r != cudaSuccess) return r;
        int flip = 0;
        long long prev = -1;      // chunk whose DMA is in flight in slot 2t + (flip ^ 1)
        for (size_t c = (size_t)t; c < nchunks; c += (size_t)p.T, flip ^= 1) {
            const int slot = 2 * t + flip;
            const size_t off = c * kChunk, len = std::min(kChunk, bytes - off);
            if ((r = cudaEventSynchronize(p.ev[slot])) != cudaSuccess) return r;   // left over from an earlier call
            if ((r = cudaMemcpyAsync(p.pinned + (size_t)slot * kChunk, src + off, len, cudaMemcpyDeviceToHost,
                                     p.streams[t])) != cudaSuccess) return r;
            if ((r = cudaEventRecord(p.ev[slot], p.streams[t])) != cudaSuccess) return r;
            if (prev >= 0) {       // while that DMA runs, hand the previous chunk to the caller
                const int pslot = 2 * t + (flip ^ 1);
                const size_t poff = (size_t)prev * kChunk, plen = std::min(kChunk, bytes - poff);
                if ((r = cudaEventSynchronize(p.ev[pslot])) != cudaSuccess) return r;
                memcpy(dst + poff, p.pinned + (size_t)pslot * kChunk, plen);
            }
            prev = (long long)c;
        }
        if (prev >= 0) {
            const int pslot = 2 * t + (flip ^ 1);
            const size_t poff = (size_t)prev * kChunk, plen = std::min(kChunk, bytes - poff);
            if ((r = cudaEventSynchronize(p.ev[pslot])) != cudaSuccess) return r;
            memcpy(dst + poff, p.pinned + (size_t)pslot * kChunk, plen);
        }
        return cudaSuccess;
    });
    if (e != cudaSuccess) {
        set_error("staged D2H copy of %zu bytes -> %s", bytes, cudaGetErrorString(e));
        return GSX_ERR_CUDA;
    }
    return GSX_OK;
}

}  // namespace gsx
