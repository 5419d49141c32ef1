// host-only harness: gsx_hostcopy.cu against the mock runtime (asynchronous streams, see cuda_runtime.h), many sizes,
// several caller threads, under -fsanitize=thread
#include "cuda_runtime.h"
int g_mock_pageable = 1;
#include "gsx_hostcopy.cu"
#include <cstdio>
#include <random>
namespace gsx { void set_error(const char*, ...) {} void count_launch() {} int sm_count() { return 148; } }
int main() {
    std::mt19937_64 rng(1);
    const size_t sizes[] = {0, 1, (8u << 20) - 1, 8u << 20, (8u << 20) + 1, (12u << 20) + 13, (64u << 20), (69u << 20) + 7, (133u << 20) + 5};
    auto one = [&](size_t n, unsigned seed) {
        std::vector<unsigned char> a(n), dev(n + 64), b(n);
        std::mt19937 r(seed);
        for (size_t i = 0; i < n; i += 4097) a[i] = (unsigned char)r();
        if (n) a[n - 1] = 0x5a;
        if (gsx::copy_h2d(dev.data(), a.data(), n, nullptr)) return 1;
        // contract of the staged path: on return the source has been read completely -- scribbling on it now must not
        // reach the device (the plain path of a small copy reads the source when the stream gets there: left alone)
        const bool staged = n >= (8u << 20);
        if (staged) a[0] ^= 0xff;
        cudaStreamSynchronize(nullptr);         // what a kernel queued on the same stream would see
        if (staged) a[0] ^= 0xff;
        if (memcmp(dev.data(), a.data(), n)) return 2;
        gsx::prefault_host(b.data(), n);
        if (gsx::copy_d2h(b.data(), dev.data(), n, nullptr)) return 3;
        if (memcmp(b.data(), a.data(), n)) return 4;
        return 0;
    };
    for (int rep = 0; rep < 2; ++rep)
        for (size_t n : sizes) {
            int rc = one(n, (unsigned)n + rep);
            if (rc) { printf("FAIL n=%zu rc=%d\n", n, rc); return 1; }
        }
    // several caller threads at once (the pool serialises them)
    std::vector<std::thread> th;
    std::atomic<int> bad{0};
    for (int t = 0; t < 4; ++t) th.emplace_back([&, t] { for (int k = 0; k < 3; ++k) if (one((20u << 20) + 1000 * t + k, 77 + t)) bad++; });
    for (auto& x : th) x.join();
    printf(bad ? "FAIL concurrent\n" : "hostcopy mock harness OK\n");
    return bad ? 1 : 0;
}
