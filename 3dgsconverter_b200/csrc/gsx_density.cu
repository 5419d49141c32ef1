// gsx_density.cu -- voxel histogram and voxel-membership mask of the density filter.
//
// Replaces data_processor.py:38-52 (quantise, np.unique(axis=0) with counts, dense = counts >=
// min_points) and :111-114 (per-point membership of the kept voxels).  The connected-component
// step in between (:59-106) stays on the host: it runs over at most N/min_points dense voxels
// and its tie-breaking depends on CPython set iteration order.
//
// Instead of the reference's lexicographic sort of N int64 triples, voxels are counted with
// atomics: on a dense int32 grid over the voxel bounding box when that fits the workspace,
// otherwise in an open-addressing hash table keyed by the packed relative voxel coordinate.
// The thread whose increment makes a voxel reach the density threshold appends it to the
// dense list, so no pass over the grid is needed.  q = floor(x / f32(voxel)) uses the float32
// division of the reference (NumPy-2 weak-scalar semantics).
#include "gsx_density.cuh"
#include "gsx_sor.cuh"

#include <math.h>
#include <stdlib.h>
#include <vector>

namespace gsx {

constexpr int kAxisBits = 21;
constexpr long long kAxisLim = 1ll << kAxisBits;

__host__ __device__ __forceinline__ long long voxel_of(float v, float voxel) {
#ifdef __CUDA_ARCH__
    return (long long)floorf(__fdiv_rn(v, voxel));
#else
    return (long long)floorf(v / voxel);
#endif
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}
static inline uint64_t mix64_host(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

struct VoxGrid {
    long long q0[3];   // voxel-space origin
    long long dim[3];  // extent in voxels
};

int64_t density_workspace_bytes(int64_t n, int64_t cap) {
    if (n < 1) n = 1;
    if (cap < 1) cap = 1;
    // hash path: table of >= 2n slots (power of two), 8-byte key + 4-byte count; plus minmax scratch
    size_t slots = 64;
    while (slots < (size_t)2 * n) slots <<= 1;
    return (int64_t)(slots * 12 + 6 * 1024 * 4 + (size_t)cap * 28 + 8192);
}

// ---------------------------------------------------------------- dense-grid path
__global__ void __launch_bounds__(256) k_vox_count_grid(const float* __restrict__ xyz, int64_t n, float voxel,
                                                        VoxGrid g, int thr, int* __restrict__ grid,
                                                        unsigned long long* __restrict__ counters,
                                                        long long* __restrict__ dense_vox, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long qx = voxel_of(xyz[3 * i], voxel), qy = voxel_of(xyz[3 * i + 1], voxel),
              qz = voxel_of(xyz[3 * i + 2], voxel);
    size_t idx = ((size_t)(qx - g.q0[0]) * g.dim[1] + (size_t)(qy - g.q0[1])) * g.dim[2] + (size_t)(qz - g.q0[2]);
    int old = atomicAdd(grid + idx, 1);
    if (old == 0) atomicAdd(counters + 1, 1ull);  // number of distinct voxels
    if (old + 1 == thr) {
        unsigned long long slot = atomicAdd(counters, 1ull);
        if ((int64_t)slot < cap) {
            dense_vox[3 * slot] = qx;
            dense_vox[3 * slot + 1] = qy;
            dense_vox[3 * slot + 2] = qz;
        }
    }
}

// Same histogram with per-block aggregation: clustered clouds put 10^5..10^6 points into a handful of voxels and
// same-address global atomics serialise in L2.  Each block first counts its 2048 points in a 1024-slot shared-memory
// hash (cell index -> count) and then issues ONE global atomicAdd per distinct cell; the increment may be > 1, so the
// threshold crossing is detected as old < thr <= old + c (still exactly one block sees it).
constexpr int kAggItems = 8, kAggSlots = 1024;
__global__ void __launch_bounds__(256) k_vox_count_grid_agg(const float* __restrict__ xyz, int64_t n, float voxel,
                                                            VoxGrid g, int thr, int* __restrict__ grid,
                                                            unsigned long long* __restrict__ counters,
                                                            long long* __restrict__ dense_vox, int64_t cap) {
    __shared__ unsigned int skey[kAggSlots];
    __shared__ int sval[kAggSlots];
    for (int t = threadIdx.x; t < kAggSlots; t += 256) {
        skey[t] = 0xffffffffu;
        sval[t] = 0;
    }
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * (256 * kAggItems);
    auto commit = [&](unsigned int idx, int c) {
        int old = atomicAdd(grid + idx, c);
        if (old == 0) atomicAdd(counters + 1, 1ull);
        if (old < thr && old + c >= thr) {
            unsigned long long slot = atomicAdd(counters, 1ull);
            if ((int64_t)slot < cap) {
                size_t cz = idx % (size_t)g.dim[2], cy = (idx / (size_t)g.dim[2]) % (size_t)g.dim[1],
                       cx = idx / ((size_t)g.dim[2] * (size_t)g.dim[1]);
                dense_vox[3 * slot] = (long long)cx + g.q0[0];
                dense_vox[3 * slot + 1] = (long long)cy + g.q0[1];
                dense_vox[3 * slot + 2] = (long long)cz + g.q0[2];
            }
        }
    };
#pragma unroll 2
    for (int e = 0; e < kAggItems; ++e) {
        const int64_t i = base + (int64_t)e * 256 + threadIdx.x;
        if (i >= n) break;
        long long qx = voxel_of(xyz[3 * i], voxel), qy = voxel_of(xyz[3 * i + 1], voxel),
                  qz = voxel_of(xyz[3 * i + 2], voxel);
        const unsigned int idx =
            (unsigned int)(((size_t)(qx - g.q0[0]) * g.dim[1] + (size_t)(qy - g.q0[1])) * g.dim[2] + (size_t)(qz - g.q0[2]));
        unsigned int slot = (idx * 2654435761u) >> 22;  // 10 bits
        bool placed = false;
#pragma unroll 1
        for (int probe = 0; probe < 8 && !placed; ++probe) {
            unsigned int cur = atomicCAS(&skey[slot], 0xffffffffu, idx);
            if (cur == 0xffffffffu || cur == idx) {
                atomicAdd(&sval[slot], 1);
                placed = true;
            }
            slot = (slot + 1) & (kAggSlots - 1);
        }
        if (!placed) commit(idx, 1);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < kAggSlots; t += 256)
        if (sval[t] > 0) commit(skey[t], sval[t]);
}

// Small voxel boxes (<= kSmemCells cells: --density_sensitivity 0.5 over a 24-unit cloud is 22^3 = 10.6 k): the whole
// histogram lives in shared memory.  Persistent CTAs count a contiguous slice of the cloud with shared-memory atomics
// (4 splats = three 128-bit streaming loads per thread) and flush their non-zero bins with ONE global atomicAdd each:
// <= cells per CTA instead of one same-address global atomic per splat (clustered clouds put 10^5..10^6 splats into a
// handful of voxels).  thr > 0: also the dense-voxel detection of the one-shot path (the CTA whose add crosses thr).
constexpr int kSmemCells = 24 * 1024;
__global__ void __launch_bounds__(512)
    k_vox_count_smem(const float* __restrict__ xyz, int64_t n, float voxel, VoxGrid g, int ncell, int thr,
                     int* __restrict__ grid, unsigned long long* __restrict__ counters,
                     long long* __restrict__ dense_vox, int64_t cap, unsigned long long* __restrict__ oob) {
    extern __shared__ int sh_hist[];
    for (int t = threadIdx.x; t < ncell; t += blockDim.x) sh_hist[t] = 0;
    __syncthreads();
    const int64_t n4 = n / 4;                       // groups of 4 splats = 3 float4
    const int64_t per = (n4 + gridDim.x - 1) / gridDim.x;
    const int64_t g0 = (int64_t)blockIdx.x * per, g1 = g0 + per < n4 ? g0 + per : n4;
    const bool vec = (reinterpret_cast<uintptr_t>(xyz) & 15) == 0;
    unsigned long long bad = 0;
    auto add = [&](float x, float y, float z) {
        const long long rx = voxel_of(x, voxel) - g.q0[0], ry = voxel_of(y, voxel) - g.q0[1], rz = voxel_of(z, voxel) - g.q0[2];
        if (rx < 0 || ry < 0 || rz < 0 || rx >= g.dim[0] || ry >= g.dim[1] || rz >= g.dim[2]) {
            ++bad;
            return;
        }
        atomicAdd(&sh_hist[(int)((rx * g.dim[1] + ry) * g.dim[2] + rz)], 1);
    };
    for (int64_t t = g0 + threadIdx.x; t < g1; t += blockDim.x) {
        if (vec) {
            const float4* p4 = reinterpret_cast<const float4*>(xyz) + 3 * t;
            const float4 a = ld_stream_f4(p4), b = ld_stream_f4(p4 + 1), c = ld_stream_f4(p4 + 2);
            add(a.x, a.y, a.z);
            add(a.w, b.x, b.y);
            add(b.z, b.w, c.x);
            add(c.y, c.z, c.w);
        } else {
            for (int e = 0; e < 4; ++e) add(xyz[12 * t + 3 * e], xyz[12 * t + 3 * e + 1], xyz[12 * t + 3 * e + 2]);
        }
    }
    if (blockIdx.x == 0)   // ragged tail (< 4 splats)
        for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) add(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    if (bad && oob) atomicAdd(oob, bad);
    __syncthreads();
    for (int t = threadIdx.x; t < ncell; t += blockDim.x) {
        const int c = sh_hist[t];
        if (c == 0) continue;
        const int old = atomicAdd(grid + t, c);
        if (thr > 0) {
            if (old == 0) atomicAdd(counters + 1, 1ull);
            if (old < thr && old + c >= thr) {
                const unsigned long long slot = atomicAdd(counters, 1ull);
                if ((int64_t)slot < cap) {
                    const long long cz = t % g.dim[2], cy = (t / g.dim[2]) % g.dim[1], cx = t / (g.dim[2] * g.dim[1]);
                    dense_vox[3 * slot] = cx + g.q0[0];
                    dense_vox[3 * slot + 1] = cy + g.q0[1];
                    dense_vox[3 * slot + 2] = cz + g.q0[2];
                }
            }
        }
    }
}

static int launch_vox_count_smem(const float* xyz, int64_t n, float voxel, const VoxGrid& g, size_t ncell, int thr, int* grid,
                                 unsigned long long* counters, long long* dvox, int64_t cap, unsigned long long* oob,
                                 cudaStream_t st) {
    const size_t smem = ncell * sizeof(int);
    static bool attr_done = false;
    if (!attr_done) {
        GSX_CUDA_CHECK(cudaFuncSetAttribute(k_vox_count_smem, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(kSmemCells * sizeof(int))));
        attr_done = true;
    }
    int blocks = sm_count() * 2;
    const int64_t n4 = n / 4;
    if ((int64_t)blocks > (n4 + 511) / 512) blocks = (int)((n4 + 511) / 512);
    if (blocks < 1) blocks = 1;
    k_vox_count_smem<<<blocks, 512, smem, st>>>(xyz, n, voxel, g, (int)ncell, thr, grid, counters, dvox, cap, oob);
    return GSX_OK;
}

__global__ void k_vox_dense_counts_grid(const long long* __restrict__ dense_vox, int64_t nd, VoxGrid g,
                                        const int* __restrict__ grid, int* __restrict__ dense_cnt) {
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nd) return;
    size_t idx = ((size_t)(dense_vox[3 * s] - g.q0[0]) * g.dim[1] + (size_t)(dense_vox[3 * s + 1] - g.q0[1])) * g.dim[2] +
                 (size_t)(dense_vox[3 * s + 2] - g.q0[2]);
    dense_cnt[s] = grid[idx];
}

// ---------------------------------------------------------------- hash path
__device__ __forceinline__ uint64_t pack_rel(long long rx, long long ry, long long rz) {
    return (((uint64_t)rx << (2 * kAxisBits)) | ((uint64_t)ry << kAxisBits) | (uint64_t)rz) + 1ull;  // 0 = empty
}

__global__ void __launch_bounds__(256) k_vox_count_hash(const float* __restrict__ xyz, int64_t n, float voxel,
                                                        VoxGrid g, int thr, unsigned long long* hkeys,
                                                        int* hcnt, uint64_t slot_mask,
                                                        unsigned long long* __restrict__ counters,
                                                        long long* __restrict__ dense_vox, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long qx = voxel_of(xyz[3 * i], voxel), qy = voxel_of(xyz[3 * i + 1], voxel),
              qz = voxel_of(xyz[3 * i + 2], voxel);
    uint64_t key = pack_rel(qx - g.q0[0], qy - g.q0[1], qz - g.q0[2]);
    uint64_t s = mix64(key) & slot_mask;
    for (;;) {
        unsigned long long cur = hkeys[s];
        if (cur == 0ull) {
            unsigned long long prev = atomicCAS(hkeys + s, 0ull, (unsigned long long)key);
            cur = prev == 0ull ? (unsigned long long)key : prev;
        }
        if (cur == key) break;
        s = (s + 1) & slot_mask;
    }
    int old = atomicAdd(hcnt + s, 1);
    if (old == 0) atomicAdd(counters + 1, 1ull);
    if (old + 1 == thr) {
        unsigned long long slot = atomicAdd(counters, 1ull);
        if ((int64_t)slot < cap) {
            dense_vox[3 * slot] = qx;
            dense_vox[3 * slot + 1] = qy;
            dense_vox[3 * slot + 2] = qz;
        }
    }
}

__global__ void k_vox_dense_counts_hash(const long long* __restrict__ dense_vox, int64_t nd, VoxGrid g,
                                        const unsigned long long* __restrict__ hkeys, const int* __restrict__ hcnt,
                                        uint64_t slot_mask, int* __restrict__ dense_cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nd) return;
    uint64_t key = pack_rel(dense_vox[3 * t] - g.q0[0], dense_vox[3 * t + 1] - g.q0[1], dense_vox[3 * t + 2] - g.q0[2]);
    uint64_t s = mix64(key) & slot_mask;
    while (hkeys[s] != key) s = (s + 1) & slot_mask;
    dense_cnt[t] = hcnt[s];
}

// Wide-key variant for voxel boxes that span 2^21 or more voxels on some axis (one far-away flyer with a small voxel):
// the slot key is the pair a = (rx << 32 | ry) + 1, b = rz + 1 (extents < 2^31).  A slot is claimed by a CAS on `a`;
// the owner then publishes `b`; a thread that finds its own `a` waits for `b` (independent thread scheduling makes
// the intra-warp wait safe) and moves on if it differs (same x,y column, other z).  Exact for any extent < 2^31.
__device__ __forceinline__ void wide_key(long long rx, long long ry, long long rz, unsigned long long& a,
                                         unsigned long long& b) {
    a = (((unsigned long long)rx << 32) | (unsigned long long)ry) + 1ull;
    b = (unsigned long long)rz + 1ull;
}

__device__ __forceinline__ uint64_t wide_find_or_insert(unsigned long long a, unsigned long long b,
                                                        unsigned long long* ha, unsigned long long* hb,
                                                        uint64_t slot_mask, bool insert) {
    uint64_t s = mix64(a ^ mix64(b)) & slot_mask;
    for (;;) {
        unsigned long long cur = *((volatile unsigned long long*)(ha + s));
        if (cur == 0ull) {
            if (!insert) return ~0ull;
            unsigned long long prev = atomicCAS(ha + s, 0ull, a);
            if (prev == 0ull) {
                atomicExch(hb + s, b);
                return s;
            }
            cur = prev;
        }
        if (cur == a) {
            unsigned long long bv;
            do {
                bv = *((volatile unsigned long long*)(hb + s));
            } while (bv == 0ull);
            if (bv == b) return s;
        }
        s = (s + 1) & slot_mask;
    }
}

__global__ void __launch_bounds__(256) k_vox_count_hash_wide(const float* __restrict__ xyz, int64_t n, float voxel,
                                                             VoxGrid g, int thr, unsigned long long* ha,
                                                             unsigned long long* hb, int* hcnt, uint64_t slot_mask,
                                                             unsigned long long* __restrict__ counters,
                                                             long long* __restrict__ dense_vox, int64_t cap) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long qx = voxel_of(xyz[3 * i], voxel), qy = voxel_of(xyz[3 * i + 1], voxel),
              qz = voxel_of(xyz[3 * i + 2], voxel);
    unsigned long long a, b;
    wide_key(qx - g.q0[0], qy - g.q0[1], qz - g.q0[2], a, b);
    uint64_t s = wide_find_or_insert(a, b, ha, hb, slot_mask, true);
    int old = atomicAdd(hcnt + s, 1);
    if (old == 0) atomicAdd(counters + 1, 1ull);
    if (old + 1 == thr) {
        unsigned long long slot = atomicAdd(counters, 1ull);
        if ((int64_t)slot < cap) {
            dense_vox[3 * slot] = qx;
            dense_vox[3 * slot + 1] = qy;
            dense_vox[3 * slot + 2] = qz;
        }
    }
}

__global__ void k_vox_dense_counts_hash_wide(const long long* __restrict__ dense_vox, int64_t nd, VoxGrid g,
                                             unsigned long long* ha, unsigned long long* hb,
                                             const int* __restrict__ hcnt, uint64_t slot_mask,
                                             int* __restrict__ dense_cnt) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nd) return;
    unsigned long long a, b;
    wide_key(dense_vox[3 * t] - g.q0[0], dense_vox[3 * t + 1] - g.q0[1], dense_vox[3 * t + 2] - g.q0[2], a, b);
    uint64_t s = wide_find_or_insert(a, b, ha, hb, slot_mask, false);
    dense_cnt[t] = s == ~0ull ? 0 : hcnt[s];
}

int density_voxel_count(const float* xyz, int64_t n, float voxel, int64_t min_points, int64_t* dense_vox_host,
                        int32_t* dense_cnt_host, int64_t cap, int64_t* n_dense_host, int64_t* n_voxels_host, void* ws,
                        int64_t ws_bytes, cudaStream_t st) {
    GSX_NVTX("gsx::density_voxel_count");
    GSX_REQUIRE(n >= 1, GSX_ERR_ARG, "density: n must be >= 1");
    GSX_REQUIRE(voxel > 0.f, GSX_ERR_ARG, "density: voxel size must be > 0");
    GSX_REQUIRE(ws_bytes >= density_workspace_bytes(n, cap), GSX_ERR_WORKSPACE, "density: workspace too small");
    GSX_REQUIRE(cap >= 1, GSX_ERR_ARG, "density: cap must be >= 1");
    Carver c(ws, (size_t)ws_bytes);
    float* partial = c.take<float>(6 * 1024);
    float* minmax = c.take<float>(8);
    unsigned long long* counters = c.take<unsigned long long>(4);
    long long* dvox = c.take<long long>(3 * (size_t)cap);
    int* dcnt = c.take<int>((size_t)cap);
    size_t used = align_up(c.off, 256);
    GSX_REQUIRE(c.ok() && used < (size_t)ws_bytes, GSX_ERR_WORKSPACE, "density: workspace too small for cap=%lld",
                (long long)cap);
    char* blob = (char*)ws + used;
    size_t blob_bytes = (size_t)ws_bytes - used;

    int rc = sor_minmax(xyz, n, minmax, partial, st);
    if (rc) return rc;
    float mm[6];
    GSX_CUDA_CHECK(cudaMemcpyAsync(mm, minmax, sizeof(mm), cudaMemcpyDeviceToHost, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    VoxGrid g;
    double cells = 1.0;
    bool wide = false;
    for (int a = 0; a < 3; ++a) {
        g.q0[a] = voxel_of(mm[a], voxel);  // floor(x/voxel) is monotone in x: min/max commute with it
        long long q1 = voxel_of(mm[3 + a], voxel);
        g.dim[a] = q1 - g.q0[a] + 1;
        GSX_REQUIRE(g.dim[a] >= 1 && g.dim[a] < (1ll << 31), GSX_ERR_UNSUPPORTED,
                    "density: voxel grid extent %lld on axis %d exceeds 2^31 voxels", g.dim[a], a);
        if (g.dim[a] >= kAxisLim) wide = true;  // the packed 3 x 21-bit key does not fit: two-word keys
        cells *= (double)g.dim[a];
    }
    long long thr_ll = min_points < 1 ? 1 : min_points;
    GSX_REQUIRE(thr_ll < 2147483647ll, GSX_ERR_ARG, "density: min_points too large");
    int thr = (int)thr_ll;
    GSX_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), st));
    int blocks = (int)((n + 255) / 256);
    bool use_grid = cells * 4.0 <= (double)blob_bytes;
    uint64_t slot_mask = 0;
    unsigned long long* hkeys = nullptr;
    unsigned long long* hkeys_b = nullptr;
    int* hcnt = nullptr;
    if (use_grid) {
        size_t ncell = (size_t)g.dim[0] * g.dim[1] * g.dim[2];
        GSX_CUDA_CHECK(cudaMemsetAsync(blob, 0, ncell * 4, st));
        if (ncell <= (size_t)kSmemCells) {
            int rc2 = launch_vox_count_smem(xyz, n, voxel, g, ncell, thr, (int*)blob, counters, dvox, cap, nullptr, st);
            if (rc2) return rc2;
        } else if (ncell < 0xfffffff0ull) {
            int ablocks = (int)((n + 256 * kAggItems - 1) / (256 * kAggItems));
            k_vox_count_grid_agg<<<ablocks, 256, 0, st>>>(xyz, n, voxel, g, thr, (int*)blob, counters, dvox, cap);
        } else {
            k_vox_count_grid<<<blocks, 256, 0, st>>>(xyz, n, voxel, g, thr, (int*)blob, counters, dvox, cap);
        }
    } else {
        size_t slots = 64;
        while (slots < (size_t)2 * n) slots <<= 1;
        GSX_REQUIRE(slots * 12 <= blob_bytes, GSX_ERR_WORKSPACE, "density: workspace too small for the hash table");
        hkeys = (unsigned long long*)blob;
        hcnt = (int*)(blob + slots * 8);
        slot_mask = slots - 1;
        if (!wide) {
            GSX_CUDA_CHECK(cudaMemsetAsync(blob, 0, slots * 12, st));
            k_vox_count_hash<<<blocks, 256, 0, st>>>(xyz, n, voxel, g, thr, hkeys, hcnt, slot_mask, counters, dvox, cap);
        } else {  // two-word keys in the same budget: half the slots (still >= n), 20 bytes each
            slots >>= 1;
            hkeys_b = hkeys + slots;
            hcnt = (int*)(blob + slots * 16);
            slot_mask = slots - 1;
            GSX_CUDA_CHECK(cudaMemsetAsync(blob, 0, slots * 20, st));
            k_vox_count_hash_wide<<<blocks, 256, 0, st>>>(xyz, n, voxel, g, thr, hkeys, hkeys_b, hcnt, slot_mask,
                                                          counters, dvox, cap);
        }
    }
    GSX_KERNEL_CHECK();
    unsigned long long hc[2];
    GSX_CUDA_CHECK(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    *n_dense_host = (int64_t)hc[0];
    if (n_voxels_host) *n_voxels_host = (int64_t)hc[1];
    GSX_REQUIRE((int64_t)hc[0] <= cap, GSX_ERR_WORKSPACE, "density: %llu dense voxels exceed cap %lld", hc[0],
                (long long)cap);
    int64_t nd = (int64_t)hc[0];
    if (nd > 0) {
        int b2 = (int)((nd + 127) / 128);
        if (use_grid) k_vox_dense_counts_grid<<<b2, 128, 0, st>>>(dvox, nd, g, (const int*)blob, dcnt);
        else if (!wide) k_vox_dense_counts_hash<<<b2, 128, 0, st>>>(dvox, nd, g, hkeys, hcnt, slot_mask, dcnt);
        else k_vox_dense_counts_hash_wide<<<b2, 128, 0, st>>>(dvox, nd, g, hkeys, hkeys_b, hcnt, slot_mask, dcnt);
        GSX_KERNEL_CHECK();
        GSX_CUDA_CHECK(cudaMemcpyAsync(dense_vox_host, dvox, (size_t)nd * 24, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaMemcpyAsync(dense_cnt_host, dcnt, (size_t)nd * 4, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return GSX_OK;
}

// ---------------------------------------------------------------- staged dense-grid API (multi-GPU)
// Rank-local histogram into a caller-owned int32 grid over a caller-chosen voxel box (the global one);
// the caller all-reduces the grid and then extracts the dense voxels.
__global__ void __launch_bounds__(256) k_vox_count_grid_only(const float* __restrict__ xyz, int64_t n, float voxel,
                                                             VoxGrid g, int* __restrict__ grid,
                                                             unsigned long long* __restrict__ oob) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    long long rx = voxel_of(xyz[3 * i], voxel) - g.q0[0], ry = voxel_of(xyz[3 * i + 1], voxel) - g.q0[1],
              rz = voxel_of(xyz[3 * i + 2], voxel) - g.q0[2];
    if (rx < 0 || ry < 0 || rz < 0 || rx >= g.dim[0] || ry >= g.dim[1] || rz >= g.dim[2]) {
        atomicAdd(oob, 1ull);
        return;
    }
    atomicAdd(grid + ((size_t)rx * g.dim[1] + (size_t)ry) * g.dim[2] + (size_t)rz, 1);
}

__global__ void __launch_bounds__(256) k_vox_grid_dense(const int* __restrict__ grid, size_t ncell, VoxGrid g, int thr,
                                                        unsigned long long* __restrict__ counters,
                                                        long long* __restrict__ dense_vox,
                                                        int* __restrict__ dense_cnt, int64_t cap) {
    size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int v = grid[c];
    if (v > 0) atomicAdd(counters + 1, 1ull);
    if (v >= thr) {
        unsigned long long slot = atomicAdd(counters, 1ull);
        if ((int64_t)slot < cap) {
            long long z = (long long)(c % (size_t)g.dim[2]);
            long long y = (long long)((c / (size_t)g.dim[2]) % (size_t)g.dim[1]);
            long long x = (long long)(c / ((size_t)g.dim[2] * (size_t)g.dim[1]));
            dense_vox[3 * slot] = x + g.q0[0];
            dense_vox[3 * slot + 1] = y + g.q0[1];
            dense_vox[3 * slot + 2] = z + g.q0[2];
            dense_cnt[slot] = v;
        }
    }
}

static int make_grid(const int64_t* q0, const int64_t* dim, VoxGrid& g, size_t& ncell) {
    double cells = 1.0;
    for (int a = 0; a < 3; ++a) {
        g.q0[a] = q0[a];
        g.dim[a] = dim[a];
        GSX_REQUIRE(dim[a] >= 1, GSX_ERR_ARG, "density: bad grid extent on axis %d", a);
        cells *= (double)dim[a];
    }
    GSX_REQUIRE(cells < 4.0e9, GSX_ERR_UNSUPPORTED, "density: grid too large for the dense path");
    ncell = (size_t)dim[0] * dim[1] * dim[2];
    return GSX_OK;
}

void density_voxel_range(const float* minmax_host, float voxel, int64_t* q0, int64_t* dim) {
    for (int a = 0; a < 3; ++a) {
        q0[a] = voxel_of(minmax_host[a], voxel);
        dim[a] = voxel_of(minmax_host[3 + a], voxel) - q0[a] + 1;
    }
}

int density_grid_count(const float* xyz, int64_t n, float voxel, const int64_t* q0, const int64_t* dim, int* grid_dev,
                       unsigned long long* oob_dev, cudaStream_t st) {
    if (n == 0) return GSX_OK;
    GSX_REQUIRE(voxel > 0.f, GSX_ERR_ARG, "density: voxel size must be > 0");
    VoxGrid g;
    size_t ncell;
    int rc = make_grid(q0, dim, g, ncell);
    if (rc) return rc;
    if (ncell <= (size_t)kSmemCells) {
        rc = launch_vox_count_smem(xyz, n, voxel, g, ncell, 0, grid_dev, nullptr, nullptr, 0, oob_dev, st);
        if (rc) return rc;
    } else {
        k_vox_count_grid_only<<<(int)((n + 255) / 256), 256, 0, st>>>(xyz, n, voxel, g, grid_dev, oob_dev);
    }
    GSX_KERNEL_CHECK();
    return GSX_OK;
}

int density_grid_dense(const int* grid_dev, const int64_t* q0, const int64_t* dim, int64_t min_points,
                       int64_t* dense_vox_host, int32_t* dense_cnt_host, int64_t cap, int64_t* n_dense_host,
                       int64_t* n_voxels_host, void* ws, int64_t ws_bytes, cudaStream_t st) {
    VoxGrid g;
    size_t ncell;
    int rc = make_grid(q0, dim, g, ncell);
    if (rc) return rc;
    GSX_REQUIRE(cap >= 1, GSX_ERR_ARG, "density: cap must be >= 1");
    Carver c(ws, (size_t)ws_bytes);
    unsigned long long* counters = c.take<unsigned long long>(4);
    long long* dvox = c.take<long long>(3 * (size_t)cap);
    int* dcnt = c.take<int>((size_t)cap);
    GSX_REQUIRE(c.ok(), GSX_ERR_WORKSPACE, "density: workspace too small for cap=%lld", (long long)cap);
    long long thr_ll = min_points < 1 ? 1 : min_points;
    GSX_REQUIRE(thr_ll < 2147483647ll, GSX_ERR_ARG, "density: min_points too large");
    GSX_CUDA_CHECK(cudaMemsetAsync(counters, 0, 4 * sizeof(unsigned long long), st));
    k_vox_grid_dense<<<(unsigned)((ncell + 255) / 256), 256, 0, st>>>(grid_dev, ncell, g, (int)thr_ll, counters, dvox,
                                                                      dcnt, cap);
    GSX_KERNEL_CHECK();
    unsigned long long hc[2];
    GSX_CUDA_CHECK(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    *n_dense_host = (int64_t)hc[0];
    if (n_voxels_host) *n_voxels_host = (int64_t)hc[1];
    GSX_REQUIRE((int64_t)hc[0] <= cap, GSX_ERR_WORKSPACE, "density: %llu dense voxels exceed cap %lld", hc[0],
                (long long)cap);
    if (hc[0] > 0) {
        GSX_CUDA_CHECK(cudaMemcpyAsync(dense_vox_host, dvox, (size_t)hc[0] * 24, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaMemcpyAsync(dense_cnt_host, dcnt, (size_t)hc[0] * 4, cudaMemcpyDeviceToHost, st));
        GSX_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    return GSX_OK;
}

// ---------------------------------------------------------------- membership mask
template <bool WIDE>
__device__ __forceinline__ uint8_t vox_member_one(float x, float y, float z, float voxel, long long ox,
                                                  long long oy, long long oz,
                                                  const unsigned long long* __restrict__ set, uint64_t slot_mask) {
    long long rx = voxel_of(x, voxel) - ox, ry = voxel_of(y, voxel) - oy, rz = voxel_of(z, voxel) - oz;
    if (WIDE) {  // two-word keys {a, b} at set[2s], set[2s+1]: kept voxels that span >= 2^21 on some axis
        if (rx < 0 || ry < 0 || rz < 0 || rx >= (1ll << 31) || ry >= (1ll << 31) || rz >= (1ll << 31)) return 0;
        unsigned long long a, b;
        wide_key(rx, ry, rz, a, b);
        uint64_t s = mix64(a ^ mix64(b)) & slot_mask;
        for (;;) {
            unsigned long long ca = __ldg(set + 2 * s);
            if (ca == 0ull) return 0;
            if (ca == a && __ldg(set + 2 * s + 1) == b) return 1;
            s = (s + 1) & slot_mask;
        }
    }
    if (rx < 0 || ry < 0 || rz < 0 || rx >= kAxisLim || ry >= kAxisLim || rz >= kAxisLim) return 0;
    uint64_t key = pack_rel(rx, ry, rz);
    uint64_t s = mix64(key) & slot_mask;
    for (;;) {
        unsigned long long cur = __ldg(set + s);
        if (cur == key) return 1;
        if (cur == 0ull) return 0;
        s = (s + 1) & slot_mask;
    }
}

template <bool WIDE>
__global__ void __launch_bounds__(256) k_vox_member(const float* __restrict__ xyz, int64_t begin, int64_t n,
                                                    float voxel, long long ox, long long oy, long long oz,
                                                    const unsigned long long* __restrict__ set, uint64_t slot_mask,
                                                    uint8_t* __restrict__ mask) {
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mask[i] = vox_member_one<WIDE>(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], voxel, ox, oy, oz, set, slot_mask);
}

// 4 points (3 x 128-bit loads) per thread, uchar4 store
template <bool WIDE>
__global__ void __launch_bounds__(256) k_vox_member4(const float4* __restrict__ xyz4, int64_t n4, float voxel,
                                                     long long ox, long long oy, long long oz,
                                                     const unsigned long long* __restrict__ set, uint64_t slot_mask,
                                                     uchar4* __restrict__ mask4) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    float4 a = ld_stream_f4(xyz4 + 3 * t), b = ld_stream_f4(xyz4 + 3 * t + 1), c = ld_stream_f4(xyz4 + 3 * t + 2);
    uchar4 o;
    o.x = vox_member_one<WIDE>(a.x, a.y, a.z, voxel, ox, oy, oz, set, slot_mask);
    o.y = vox_member_one<WIDE>(a.w, b.x, b.y, voxel, ox, oy, oz, set, slot_mask);
    o.z = vox_member_one<WIDE>(b.z, b.w, c.x, voxel, ox, oy, oz, set, slot_mask);
    o.w = vox_member_one<WIDE>(c.y, c.z, c.w, voxel, ox, oy, oz, set, slot_mask);
    mask4[t] = o;
}

// Bitmap form of the keep set: when the bounding box of the kept voxels is small (the usual case -- a few clusters of
// voxels of one scene unit) one bit per voxel of that box replaces the hash probe: the whole set is a few KiB that stay
// in L1, and the kernel streams at the speed of the bbox mask instead of waiting on dependent table reads.
struct VoxBits {
    long long ox, oy, oz;
    unsigned long long dx, dy, dz;
};

__device__ __forceinline__ uint8_t vox_member_bit(float x, float y, float z, float voxel, const VoxBits& g,
                                                  const uint32_t* __restrict__ bits) {
    const unsigned long long rx = (unsigned long long)(voxel_of(x, voxel) - g.ox),
                             ry = (unsigned long long)(voxel_of(y, voxel) - g.oy),
                             rz = (unsigned long long)(voxel_of(z, voxel) - g.oz);
    if (rx >= g.dx || ry >= g.dy || rz >= g.dz) return 0;    // (negative differences wrap to huge values)
    const uint32_t idx = (uint32_t)((rx * g.dy + ry) * g.dz + rz);
    return (uint8_t)((__ldg(bits + (idx >> 5)) >> (idx & 31u)) & 1u);
}

__global__ void __launch_bounds__(256) k_vox_member_bits(const float* __restrict__ xyz, int64_t begin, int64_t n,
                                                         float voxel, VoxBits g, const uint32_t* __restrict__ bits,
                                                         uint8_t* __restrict__ mask) {
    int64_t i = begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    mask[i] = vox_member_bit(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], voxel, g, bits);
}

__global__ void __launch_bounds__(256) k_vox_member_bits4(const float4* __restrict__ xyz4, int64_t n4, float voxel,
                                                          VoxBits g, const uint32_t* __restrict__ bits,
                                                          uchar4* __restrict__ mask4) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n4) return;
    float4 a = ld_stream_f4(xyz4 + 3 * t), b = ld_stream_f4(xyz4 + 3 * t + 1), c = ld_stream_f4(xyz4 + 3 * t + 2);
    uchar4 o;
    o.x = vox_member_bit(a.x, a.y, a.z, voxel, g, bits);
    o.y = vox_member_bit(a.w, b.x, b.y, voxel, g, bits);
    o.z = vox_member_bit(b.z, b.w, c.x, voxel, g, bits);
    o.w = vox_member_bit(c.y, c.z, c.w, voxel, g, bits);
    mask4[t] = o;
}

constexpr unsigned long long kMaxKeepBits = 1ull << 27;   // 16 MiB of bitmap at most; larger boxes use the hash set

static bool keep_bitmap_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("GSX_DENSITY_BITMAP");
        v = !(e && e[0] == '0');
    }
    return v != 0;
}

int density_member_mask(const float* xyz, int64_t n, float voxel, const int64_t* keep, int64_t n_keep, uint8_t* mask,
                        void* ws, int64_t ws_bytes, cudaStream_t st) {
    GSX_NVTX("gsx::density_member_mask");
    if (n == 0) return GSX_OK;
    GSX_REQUIRE(voxel > 0.f, GSX_ERR_ARG, "density: voxel size must be > 0");
    if (n_keep == 0) {
        GSX_CUDA_CHECK(cudaMemsetAsync(mask, 0, (size_t)n, st));
        return GSX_OK;
    }
    long long o[3] = {keep[0], keep[1], keep[2]}, hi[3] = {keep[0], keep[1], keep[2]};
    for (int64_t t = 1; t < n_keep; ++t)
        for (int a = 0; a < 3; ++a) {
            if (keep[3 * t + a] < o[a]) o[a] = keep[3 * t + a];
            if (keep[3 * t + a] > hi[a]) hi[a] = keep[3 * t + a];
        }
    bool wide = false;
    for (int a = 0; a < 3; ++a) {
        GSX_REQUIRE(hi[a] - o[a] < (1ll << 31), GSX_ERR_UNSUPPORTED, "density: kept voxels span more than 2^31 on axis %d", a);
        if (hi[a] - o[a] >= kAxisLim) wide = true;
    }
    {   // bitmap over the bounding box of the kept voxels, when that box is small enough
        const unsigned long long dx = (unsigned long long)(hi[0] - o[0] + 1), dy = (unsigned long long)(hi[1] - o[1] + 1),
                                 dz = (unsigned long long)(hi[2] - o[2] + 1);
        // (each factor < 2^31: the products below cannot overflow before the comparisons reject them)
        const bool small = dx <= kMaxKeepBits && dy <= kMaxKeepBits && dx * dy <= kMaxKeepBits &&
                           dx * dy * dz <= kMaxKeepBits;
        const size_t nwords = small ? (size_t)((dx * dy * dz + 31) / 32) : 0;
        if (small && keep_bitmap_enabled() && nwords * 4 <= (size_t)ws_bytes) {
            std::vector<uint32_t> bm(nwords, 0u);
            for (int64_t t = 0; t < n_keep; ++t) {
                const unsigned long long idx = ((unsigned long long)(keep[3 * t] - o[0]) * dy +
                                                (unsigned long long)(keep[3 * t + 1] - o[1])) * dz +
                                               (unsigned long long)(keep[3 * t + 2] - o[2]);
                bm[idx >> 5] |= 1u << (idx & 31);
            }
            GSX_CUDA_CHECK(cudaMemcpyAsync(ws, bm.data(), nwords * 4, cudaMemcpyHostToDevice, st));
            GSX_CUDA_CHECK(cudaStreamSynchronize(st));  // bm is a stack-owned pageable buffer
            const VoxBits g{o[0], o[1], o[2], dx, dy, dz};
            const uint32_t* bits = (const uint32_t*)ws;
            int64_t n4 = 0;
            if (((uintptr_t)xyz % 16 == 0) && ((uintptr_t)mask % 4 == 0)) {
                n4 = n / 4;
                if (n4 > 0) {
                    k_vox_member_bits4<<<(int)((n4 + 255) / 256), 256, 0, st>>>((const float4*)xyz, n4, voxel, g, bits,
                                                                               (uchar4*)mask);
                    GSX_KERNEL_CHECK();
                }
            }
            if (n - 4 * n4 > 0) {
                k_vox_member_bits<<<(int)((n - 4 * n4 + 255) / 256), 256, 0, st>>>(xyz, 4 * n4, n, voxel, g, bits, mask);
                GSX_KERNEL_CHECK();
            }
            return GSX_OK;
        }
    }
    size_t slots = 64;
    while (slots < (size_t)2 * n_keep) slots <<= 1;
    const size_t words = wide ? 2 : 1;
    GSX_REQUIRE(slots * 8 * words <= (size_t)ws_bytes, GSX_ERR_WORKSPACE, "density: workspace too small for the keep set");
    std::vector<unsigned long long> tab(slots * words, 0ull);
    for (int64_t t = 0; t < n_keep; ++t) {
        const uint64_t rx = (uint64_t)(keep[3 * t] - o[0]), ry = (uint64_t)(keep[3 * t + 1] - o[1]),
                       rz = (uint64_t)(keep[3 * t + 2] - o[2]);
        if (!wide) {
            uint64_t key = ((rx << (2 * kAxisBits)) | (ry << kAxisBits) | rz) + 1ull;
            uint64_t s = mix64_host(key) & (slots - 1);
            while (tab[s] != 0ull && tab[s] != key) s = (s + 1) & (slots - 1);
            tab[s] = key;
        } else {
            uint64_t a = ((rx << 32) | ry) + 1ull, b = rz + 1ull;
            uint64_t s = mix64_host(a ^ mix64_host(b)) & (slots - 1);
            while (tab[2 * s] != 0ull && !(tab[2 * s] == a && tab[2 * s + 1] == b)) s = (s + 1) & (slots - 1);
            tab[2 * s] = a;
            tab[2 * s + 1] = b;
        }
    }
    GSX_CUDA_CHECK(cudaMemcpyAsync(ws, tab.data(), slots * 8 * words, cudaMemcpyHostToDevice, st));
    GSX_CUDA_CHECK(cudaStreamSynchronize(st));  // tab is a stack-owned pageable buffer
    const unsigned long long* set = (const unsigned long long*)ws;
    int64_t n4 = 0;
    if (((uintptr_t)xyz % 16 == 0) && ((uintptr_t)mask % 4 == 0)) {
        n4 = n / 4;
        if (n4 > 0) {
            const int b4 = (int)((n4 + 255) / 256);
            if (wide) k_vox_member4<true><<<b4, 256, 0, st>>>((const float4*)xyz, n4, voxel, o[0], o[1], o[2], set, slots - 1, (uchar4*)mask);
            else k_vox_member4<false><<<b4, 256, 0, st>>>((const float4*)xyz, n4, voxel, o[0], o[1], o[2], set, slots - 1, (uchar4*)mask);
            GSX_KERNEL_CHECK();
        }
    }
    if (n - 4 * n4 > 0) {
        const int b1 = (int)((n - 4 * n4 + 255) / 256);
        if (wide) k_vox_member<true><<<b1, 256, 0, st>>>(xyz, 4 * n4, n, voxel, o[0], o[1], o[2], set, slots - 1, mask);
        else k_vox_member<false><<<b1, 256, 0, st>>>(xyz, 4 * n4, n, voxel, o[0], o[1], o[2], set, slots - 1, mask);
        GSX_KERNEL_CHECK();
    }
    return GSX_OK;
}

}  // namespace gsx
