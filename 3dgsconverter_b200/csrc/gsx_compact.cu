// gsx_compact.cu -- stream compaction of the per-point working set between filters.
//
// Replaces the host-side `vertices[mask]` round trips of the filter chain (data_processor.py:114,149,
// 209,217-224 as driven by converter.py:194-236) for the columns the filters need: xyz (12 B/pt),
// opacity (4 B/pt) and the original row index (4 B/pt).  The 248-byte records stay on the host and
// are gathered ONCE, with the surviving indices, when the caller reads `DataProcessor.data`.
// Stable (order-preserving), like NumPy boolean indexing; one pass (see k_cmp_onepass).
#include "gsx_common.cuh"
#include "gsx_compact.cuh"
#include "gsx_radix.cuh"

namespace gsx {

#define GSX_FULL 0xffffffffu
// One pass, decoupled look-back (the count -> scan -> scatter form read the mask twice, wrote one element per thread
// and took five launches): a tile is 2048 consecutive rows, 8 per thread.  The tile id comes from an atomic counter
// (every lower tile is already running), the tile publishes its survivor count and obtains the number of survivors
// before it by looking back over the published words (count | flag in one 32-bit store: bit 30 = this tile only,
// bit 31 = inclusive prefix), one warp reading 32 predecessors at a time.  Survivors are first placed in shared
// memory in order, then written out as contiguous, fully coalesced runs.  Stable, like NumPy boolean indexing.
constexpr int kCmpThreads = 256, kCmpPer = 8, kCmpTile = kCmpThreads * kCmpPer;   // 2048
constexpr uint32_t kLbAgg = 1u << 30, kLbInc = 1u << 31, kLbVal = (1u << 30) - 1u;
constexpr uint32_t kLbSpinLimit = 1u << 22;

__global__ void __launch_bounds__(kCmpThreads)
    k_cmp_onepass(const uint8_t* __restrict__ mask, int64_t n, int64_t ntiles, const float* __restrict__ xyz,
                  const float* __restrict__ opacity, const int32_t* __restrict__ idx, float* __restrict__ xyz_out,
                  float* __restrict__ opacity_out, int32_t* __restrict__ idx_out, uint32_t* lookback,
                  unsigned int* tile_counter, uint32_t* __restrict__ total_out) {
    __shared__ float s_xyz[kCmpTile * 3];
    __shared__ float s_op[kCmpTile];
    __shared__ int32_t s_idx[kCmpTile];
    __shared__ uint32_t s_wsum[8];
    __shared__ unsigned int s_tile;
    __shared__ uint32_t s_excl;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t i0 = tile * kCmpTile + (int64_t)threadIdx.x * kCmpPer;
    // the thread's 8 mask bytes (one 8-byte load when whole and aligned)
    unsigned keepbits = 0;
    if (i0 + kCmpPer <= n && ((reinterpret_cast<uintptr_t>(mask) + (uintptr_t)i0) & 7) == 0) {
        const uint2 m8 = *reinterpret_cast<const uint2*>(mask + i0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((m8.x >> (8 * e)) & 0xffu) keepbits |= 1u << e;
            if ((m8.y >> (8 * e)) & 0xffu) keepbits |= 1u << (4 + e);
        }
    } else {
#pragma unroll
        for (int e = 0; e < kCmpPer; ++e)
            if (i0 + e < n && mask[i0 + e] != 0) keepbits |= 1u << e;
    }
    const uint32_t mine = __popc(keepbits);
    // exclusive scan of the per-thread counts inside the tile
    uint32_t x = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(GSX_FULL, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) s_wsum[w] = x;
    __syncthreads();
    uint32_t woff = 0, tile_cnt = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t c = s_wsum[k];
        if (k < w) woff += c;
        tile_cnt += c;
    }
    uint32_t lrank = woff + x - mine;   // rank of this thread's first survivor inside the tile
    // publish, look back (warp 0)
    if (w == 0) {
        volatile uint32_t* lb = lookback;
        if (lane == 0) lb[tile] = tile_cnt | (tile == 0 ? kLbInc : kLbAgg);
        uint32_t excl = 0;
        int64_t t = tile - 1;
        uint32_t spins = 0;
        while (t >= 0) {
            const int64_t id = t - lane;
            uint32_t v = (uint32_t)(1u << 31); if (id >= 0) v = lb[id];   // before the first tile: an inclusive prefix of 0
            const unsigned inc = __ballot_sync(GSX_FULL, (v & kLbInc) != 0u);
            const unsigned notready = __ballot_sync(GSX_FULL, (v & (kLbInc | kLbAgg)) == 0u);
            const int li = inc ? __ffs(inc) - 1 : 31;            // nearest predecessor with an inclusive prefix
            const unsigned need = li == 31 ? 0xffffffffu : ((1u << (li + 1)) - 1u);
            if (notready & need) {                               // someone in the window has not published yet
                if (++spins > kLbSpinLimit) __trap();
                continue;
            }
            uint32_t part = lane <= li ? (v & kLbVal) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(GSX_FULL, part, o);
            excl += part;
            if (inc) break;
            t -= 32;
        }
        if (lane == 0) {
            if (tile != 0) lb[tile] = (excl + tile_cnt) | kLbInc;
            s_excl = excl;
            if (tile == ntiles - 1) *total_out = excl + tile_cnt;
        }
    }
    // survivors into shared memory, in order
#pragma unroll
    for (int e = 0; e < kCmpPer; ++e) {
        if ((keepbits >> e) & 1u) {
            const int64_t i = i0 + e;
            s_xyz[3 * lrank] = __ldg(xyz + 3 * i);
            s_xyz[3 * lrank + 1] = __ldg(xyz + 3 * i + 1);
            s_xyz[3 * lrank + 2] = __ldg(xyz + 3 * i + 2);
            if (opacity) s_op[lrank] = __ldg(opacity + i);
            s_idx[lrank] = idx ? __ldg(idx + i) : (int32_t)i;
            ++lrank;
        }
    }
    __syncthreads();
    const size_t obase = s_excl;
    for (uint32_t t = threadIdx.x; t < 3 * tile_cnt; t += kCmpThreads) xyz_out[3 * obase + t] = s_xyz[t];
    for (uint32_t t = threadIdx.x; t < tile_cnt; t += kCmpThreads) {
        if (opacity) opacity_out[obase + t] = s_op[t];
        idx_out[obase + t] = s_idx[t];
    }
}

int64_t compact_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    const int64_t tiles = (n + kCmpTile - 1) / kCmpTile;
    return (int64_t)((size_t)(tiles + 64) * 4 + 1024);
}

int compact_points(const uint8_t* mask, int64_t n, const float* xyz, const float* opacity, const int32_t* idx,
                   float* xyz_out, float* opacity_out, int32_t* idx_out, int64_t* count_host, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
    GSX_NVTX("gsx::compact_points");
    if (n == 0) {
        *count_host = 0;
        return GSX_OK;
    }
    GSX_REQUIRE(n < (1ll << 30), GSX_ERR_ARG, "compact: n out of range");
    GSX_REQUIRE(ws_bytes >= compact_workspace_bytes(n), GSX_ERR_WORKSPACE, "compact: workspace too small");
    GSX_REQUIRE((opacity == nullptr) == (opacity_out == nullptr), GSX_ERR_ARG, "compact: opacity in/out mismatch");
    const int64_t tiles = (n + kCmpTile - 1) / kCmpTile;
    uint32_t* words = (uint32_t*)ws;            // [0] tile counter, [1] total, [16 ..] look-back words
    GSX_CUDA_CHECK(cudaMemsetAsync(words, 0, (size_t)(tiles + 16) * 4, st));
    k_cmp_onepass<<<(unsigned)tiles, kCmpThreads, 0, st>>>(mask, n, tiles, xyz, opacity, idx, xyz_out, opacity_out, idx_out,
                                                           words + 16, words, words + 1);
    GSX_KERNEL_CHECK();
    uint32_t total = 0;
    GSX_CUDA_CHECK(cudaMemcpyAsync(&total, words + 1, 4, cudaMemcpyDeviceToHost, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    *count_host = (int64_t)total;
    return GSX_OK;
}

}  // namespace gsx
