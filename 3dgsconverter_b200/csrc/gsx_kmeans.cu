// gsx_kmeans.cu -- batched Lloyd K-Means (SOG codebooks) for sm_100a.
//
// Replaces gpu_ops.py:57-73 (k_means_assign), :75-96 (k_means_update) and the loop at :186-188.
// Arithmetic contract: SURVEY A.5 -- strict float32, no fma: dist = (((0 + d0^2) + d1^2) + ...),
// strict '<' with the lowest centroid index winning ties, min_dist = 1e20, update = float32 sums
// accumulated serially in point-index order, inv = 1.0f/cnt, empty clusters collapse to 0.
//
// Design:
//   * all independent problems (the <= 64 spatial chunks of sog.py:527-549) run in ONE launch per
//     phase; a CTA works on a tile of points of one problem.
//   * assign: x rows live in registers (D is a template parameter), centroids are staged through
//     shared memory in tiles and read as broadcast float4; every thread runs P points x 2 centroids
//     = 2P independent accumulation chains, so the dependent FADD chain of the contract does not
//     stall the FP32 pipes.  CUDA cores, not tensor cores: the contract's rounding sequence is not
//     a GEMM (DESIGN.md discusses the GEMM-prefilter idea for a later round).
//   * update: one warp per (problem, cluster) streams the problem's labels in index order, ballots
//     the members and accumulates their rows lane-per-dimension.  This reproduces the oracle's
//     serial index-order sum bit-for-bit and is run-to-run deterministic (the reference's float
//     atomics are neither); labels are L2-resident, each X row is read once.
#include "gsx_common.cuh"
#include "gsx_kmeans.cuh"

#include <vector>

namespace gsx {

#define GSX_FULL 0xffffffffu
constexpr int kAssignThreads = 128;
constexpr int kCentTile = 64;  // centroids per shared-memory tile

struct KmProb {
    long long row0;  // first row of the problem in X
    long long rows;  // number of rows
    int tile0;       // first assign tile of the problem
};

template <int D>
struct PointsPerThread {
    static constexpr int value = D <= 4 ? 4 : (D <= 24 ? 2 : 2);
};

template <int D, int P>
__global__ void __launch_bounds__(kAssignThreads)
    k_kmeans_assign(const float* __restrict__ X, const float* __restrict__ C, int* __restrict__ labels,
                    const KmProb* __restrict__ probs, int nprob, int K) {
    constexpr int DP = (D + 3) / 4 * 4;  // padded row stride in shared memory (float4 aligned)
    constexpr int G = DP / 4;
    __shared__ __align__(16) float sc[kCentTile * DP];

    // which problem does this tile belong to?
    int lo = 0, hi = nprob - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (probs[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const KmProb pr = probs[lo];
    const long long tile = (long long)blockIdx.x - pr.tile0;
    const float* Cp = C + (size_t)lo * K * D;

    float x[P][DP];
    long long row[P];
    bool live[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        long long r = tile * (kAssignThreads * P) + p * kAssignThreads + threadIdx.x;
        live[p] = r < pr.rows;
        row[p] = pr.row0 + (live[p] ? r : 0);
        const float* xr = X + (size_t)row[p] * D;
#pragma unroll
        for (int d = 0; d < DP; ++d) x[p][d] = d < D ? xr[d] : 0.f;
    }
    float best_d[P];
    int best_k[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        best_d[p] = 1e20f;
        best_k[p] = -1;
    }

    for (int k0 = 0; k0 < K; k0 += kCentTile) {
        const int kt = K - k0 < kCentTile ? K - k0 : kCentTile;
        __syncthreads();
        for (int t = threadIdx.x; t < kCentTile * DP; t += kAssignThreads) {
            int c = t / DP, d = t - c * DP;
            sc[t] = (c < kt && d < D) ? Cp[(size_t)(k0 + c) * D + d] : 0.f;
        }
        __syncthreads();
        // two centroids per step; an odd tail centroid is paired with a zero row and ignored
        for (int c = 0; c < kt; c += 2) {
            float acc0[P], acc1[P];
#pragma unroll
            for (int p = 0; p < P; ++p) acc0[p] = acc1[p] = 0.f;
            const float4* r0 = reinterpret_cast<const float4*>(sc + c * DP);
            const float4* r1 = reinterpret_cast<const float4*>(sc + (c + 1 < kCentTile ? c + 1 : c) * DP);
#pragma unroll
            for (int g = 0; g < G; ++g) {
                const float4 a = r0[g], b = r1[g];
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    if (g * 4 + e < D) {
#pragma unroll
                        for (int p = 0; p < P; ++p) {
                            float d0 = __fsub_rn(x[p][g * 4 + e], av[e]);
                            acc0[p] = __fadd_rn(acc0[p], __fmul_rn(d0, d0));
                            float d1 = __fsub_rn(x[p][g * 4 + e], bv[e]);
                            acc1[p] = __fadd_rn(acc1[p], __fmul_rn(d1, d1));
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < P; ++p) {
                if (acc0[p] < best_d[p]) {
                    best_d[p] = acc0[p];
                    best_k[p] = k0 + c;
                }
                if (c + 1 < kt && acc1[p] < best_d[p]) {
                    best_d[p] = acc1[p];
                    best_k[p] = k0 + c + 1;
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
        if (live[p]) labels[row[p]] = best_k[p];
}

// any D: one point per thread, x re-read through L1 (slow path for unusual dimensions)
__global__ void __launch_bounds__(kAssignThreads)
    k_kmeans_assign_generic(const float* __restrict__ X, const float* __restrict__ C, int* __restrict__ labels,
                            const KmProb* __restrict__ probs, int nprob, int K, int D) {
    int lo = 0, hi = nprob - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (probs[mid].tile0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const KmProb pr = probs[lo];
    long long r = ((long long)blockIdx.x - pr.tile0) * kAssignThreads + threadIdx.x;
    if (r >= pr.rows) return;
    const float* xr = X + (size_t)(pr.row0 + r) * D;
    const float* Cp = C + (size_t)lo * K * D;
    float best = 1e20f;
    int bk = -1;
    for (int c = 0; c < K; ++c) {
        float acc = 0.f;
        for (int d = 0; d < D; ++d) {
            float df = __fsub_rn(xr[d], __ldg(Cp + (size_t)c * D + d));
            acc = __fadd_rn(acc, __fmul_rn(df, df));
        }
        if (acc < best) {
            best = acc;
            bk = c;
        }
    }
    labels[pr.row0 + r] = bk;
}

// one warp per (problem, cluster): serial index-order float32 sums, then the 1/cnt scaling
__global__ void __launch_bounds__(256)
    k_kmeans_update(const float* __restrict__ X, float* __restrict__ C, const int* __restrict__ labels,
                    int* __restrict__ counts, const KmProb* __restrict__ probs, int nprob, int K, int D) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (long long)nprob * K) return;
    const int p = (int)(wid / K), c = (int)(wid - (long long)p * K);
    const KmProb pr = probs[p];
    const int* lab = labels + pr.row0;
    const float* Xp = X + (size_t)pr.row0 * D;
    float* out = C + ((size_t)p * K + c) * D;
    int cnt = 0;
    for (int d0 = 0; d0 < D; d0 += 64) {  // lanes cover dims d0+lane and d0+32+lane
        const int da = d0 + lane, db = d0 + 32 + lane;
        float sa = 0.f, sb = 0.f;
        int n_here = 0;
        for (long long base = 0; base < pr.rows; base += 32) {
            long long i = base + lane;
            bool hit = i < pr.rows && __ldg(lab + i) == c;
            unsigned m = __ballot_sync(GSX_FULL, hit);
            n_here += __popc(m);
            while (m) {
                int src = __ffs(m) - 1;
                m &= m - 1;
                const float* xr = Xp + (size_t)(base + src) * D;
                if (da < D) sa = __fadd_rn(sa, __ldg(xr + da));
                if (db < D) sb = __fadd_rn(sb, __ldg(xr + db));
            }
        }
        cnt = n_here;
        if (cnt > 0) {
            float inv = __fdiv_rn(1.0f, (float)cnt);
            sa = __fmul_rn(sa, inv);
            sb = __fmul_rn(sb, inv);
        }
        if (da < D) out[da] = sa;
        if (db < D) out[db] = sb;
    }
    if (lane == 0) counts[(size_t)p * K + c] = cnt;
}

int64_t kmeans_workspace_bytes(int64_t n_total, int nprob, int K, int D) {
    (void)n_total;
    (void)K;
    (void)D;
    return (int64_t)align_up((size_t)(nprob > 0 ? nprob : 1) * sizeof(KmProb), 256) + 1024;
}

template <int D>
static void launch_assign(const float* X, const float* C, int* labels, const KmProb* probs, int nprob, int K,
                          int tiles, cudaStream_t st) {
    constexpr int P = PointsPerThread<D>::value;
    k_kmeans_assign<D, P><<<tiles, kAssignThreads, 0, st>>>(X, C, labels, probs, nprob, K);
}

static int points_per_thread(int D) {
    switch (D) {
        case 1: return PointsPerThread<1>::value;
        case 2: return PointsPerThread<2>::value;
        case 3: return PointsPerThread<3>::value;
        case 4: return PointsPerThread<4>::value;
        case 9: return PointsPerThread<9>::value;
        case 24: return PointsPerThread<24>::value;
        case 45: return PointsPerThread<45>::value;
        default: return 1;
    }
}

int kmeans_lloyd(const float* X, const int64_t* row_off, int nprob, int K, int D, int max_iter, float* C, int* labels,
                 int* counts, void* ws, int64_t ws_bytes, cudaStream_t st) {
    GSX_REQUIRE(nprob >= 1 && K >= 1 && D >= 1 && max_iter >= 0, GSX_ERR_ARG, "kmeans: bad shape");
    GSX_REQUIRE(ws_bytes >= kmeans_workspace_bytes(row_off[nprob] - row_off[0], nprob, K, D), GSX_ERR_WORKSPACE,
                "kmeans: workspace too small");
    const int per_tile = kAssignThreads * points_per_thread(D);
    std::vector<KmProb> hp(nprob);
    long long tiles = 0;
    for (int p = 0; p < nprob; ++p) {
        hp[p].row0 = row_off[p];
        hp[p].rows = row_off[p + 1] - row_off[p];
        GSX_REQUIRE(hp[p].rows >= 1, GSX_ERR_ARG, "kmeans: empty problem %d", p);
        hp[p].tile0 = (int)tiles;
        tiles += (hp[p].rows + per_tile - 1) / per_tile;
    }
    GSX_REQUIRE(tiles < 2147483647ll, GSX_ERR_UNSUPPORTED, "kmeans: too many tiles");
    KmProb* dp = (KmProb*)ws;
    GSX_CUDA_CHECK(cudaMemcpyAsync(dp, hp.data(), hp.size() * sizeof(KmProb), cudaMemcpyHostToDevice, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));  // hp is a local pageable buffer
    const long long uwarps = (long long)nprob * K;
    const int ublocks = (int)((uwarps * 32 + 255) / 256);
    for (int it = 0; it < max_iter; ++it) {
        switch (D) {
            case 1: launch_assign<1>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 2: launch_assign<2>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 3: launch_assign<3>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 4: launch_assign<4>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 9: launch_assign<9>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 24: launch_assign<24>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            case 45: launch_assign<45>(X, C, labels, dp, nprob, K, (int)tiles, st); break;
            default:
                k_kmeans_assign_generic<<<(int)tiles, kAssignThreads, 0, st>>>(X, C, labels, dp, nprob, K, D);
        }
        GSX_KERNEL_CHECK();
        k_kmeans_update<<<ublocks, 256, 0, st>>>(X, C, labels, counts, dp, nprob, K, D);
        GSX_KERNEL_CHECK();
    }
    return GSX_OK;
}

}  // namespace gsx
