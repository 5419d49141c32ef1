// gsx_compact.cu -- stream compaction of the per-point working set between filters.
//
// Replaces the host-side `vertices[mask]` round trips of the filter chain (data_processor.py:114,149,
// 209,217-224 as driven by converter.py:194-236) for the columns the filters need: xyz (12 B/pt),
// opacity (4 B/pt) and the original row index (4 B/pt).  The 248-byte records stay on the host and
// are gathered ONCE, with the surviving indices, when the caller reads `DataProcessor.data`.
// Stable (order-preserving), like NumPy boolean indexing: block counts -> exclusive scan -> scatter.
// (Round 2 also tried a single-pass form -- 2048-row tiles, decoupled look-back over per-tile survivor counts,
// survivors staged in shared memory: 0.85 ms against this version's 0.74 ms at 64 M rows, 50 % survivors
// (profiles/r02_stream_kernels_probe.json history in DESIGN.md section 10); the two-pass form reads only one extra
// byte per row and its scatter is already sector-exact, so it stays.)
#include "gsx_common.cuh"
#include "gsx_compact.cuh"
#include "gsx_radix.cuh"

namespace gsx {

#define GSX_FULL 0xffffffffu
constexpr int kCmpBlock = 1024;

__global__ void __launch_bounds__(kCmpBlock) k_cmp_count(const uint8_t* __restrict__ mask, int64_t n,
                                                         uint32_t* __restrict__ counts) {
    int64_t i = (int64_t)blockIdx.x * kCmpBlock + threadIdx.x;
    bool keep = i < n && mask[i] != 0;
    unsigned b = __ballot_sync(GSX_FULL, keep);
    __shared__ uint32_t wc[32];
    if ((threadIdx.x & 31) == 0) wc[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t v = wc[threadIdx.x];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(GSX_FULL, v, o);
        if (threadIdx.x == 0) counts[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(kCmpBlock)
    k_cmp_scatter(const uint8_t* __restrict__ mask, int64_t n, const uint32_t* __restrict__ block_off,
                  const float* __restrict__ xyz, const float* __restrict__ opacity, const int32_t* __restrict__ idx,
                  float* __restrict__ xyz_out, float* __restrict__ opacity_out, int32_t* __restrict__ idx_out) {
    int64_t i = (int64_t)blockIdx.x * kCmpBlock + threadIdx.x;
    bool keep = i < n && mask[i] != 0;
    unsigned b = __ballot_sync(GSX_FULL, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __shared__ uint32_t wc[32];
    if (lane == 0) wc[w] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {  // exclusive scan of the 32 warp counts
        uint32_t v = wc[threadIdx.x], x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(GSX_FULL, x, o);
            if (lane >= o) x += y;
        }
        wc[threadIdx.x] = x - v;
    }
    __syncthreads();
    if (!keep) return;
    const uint32_t pos = block_off[blockIdx.x] + wc[w] + __popc(b & ((1u << lane) - 1u));
    xyz_out[3 * (size_t)pos] = xyz[3 * i];
    xyz_out[3 * (size_t)pos + 1] = xyz[3 * i + 1];
    xyz_out[3 * (size_t)pos + 2] = xyz[3 * i + 2];
    if (opacity) opacity_out[pos] = opacity[i];
    idx_out[pos] = idx ? idx[i] : (int32_t)i;
}

int64_t compact_workspace_bytes(int64_t n) {
    if (n < 1) n = 1;
    int64_t blocks = (n + kCmpBlock - 1) / kCmpBlock;
    return (int64_t)((size_t)(blocks + 64) * 4 + scan_workspace_bytes(blocks) + 1024);
}

int compact_points(const uint8_t* mask, int64_t n, const float* xyz, const float* opacity, const int32_t* idx,
                   float* xyz_out, float* opacity_out, int32_t* idx_out, int64_t* count_host, void* ws,
                   int64_t ws_bytes, cudaStream_t st) {
    GSX_NVTX("gsx::compact_points");
    if (n == 0) {
        *count_host = 0;
        return GSX_OK;
    }
    GSX_REQUIRE(ws_bytes >= compact_workspace_bytes(n), GSX_ERR_WORKSPACE, "compact: workspace too small");
    GSX_REQUIRE((opacity == nullptr) == (opacity_out == nullptr), GSX_ERR_ARG, "compact: opacity in/out mismatch");
    const int64_t blocks = (n + kCmpBlock - 1) / kCmpBlock;
    uint32_t* counts = (uint32_t*)ws;          // blocks + 1 (total in the extra slot after the scan)
    uint32_t* sws = counts + blocks + 64;
    GSX_CUDA_CHECK(cudaMemsetAsync(counts + blocks, 0, 4, st));
    k_cmp_count<<<(unsigned)blocks, kCmpBlock, 0, st>>>(mask, n, counts);
    GSX_KERNEL_CHECK();
    int rc = exclusive_scan_u32_ws(counts, blocks + 1, sws, st);  // counts[blocks] becomes the total
    if (rc) return rc;
    k_cmp_scatter<<<(unsigned)blocks, kCmpBlock, 0, st>>>(mask, n, counts, xyz, opacity, idx, xyz_out, opacity_out,
                                                          idx_out);
    GSX_KERNEL_CHECK();
    uint32_t total = 0;
    GSX_CUDA_CHECK(cudaMemcpyAsync(&total, counts + blocks, 4, cudaMemcpyDeviceToHost, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    *count_host = (int64_t)total;
    return GSX_OK;
}

}  // namespace gsx
