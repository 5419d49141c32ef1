/*
 * gsx_oracle.c -- CPU restatement of the reference's device kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.
 *
 * PARITY STATUS: "parity unpinned" by the reference itself -- the reference ships
 * no tests, golden vectors or fixtures and its Taichi kernels cannot
 * be run in this container (taichi is not installable).  The contract restated
 * here is the strict-IEEE reading of SURVEY.md Appendix A; the NumPy-visible
 * parts (pairwise mean/std) ARE pinned bit-for-bit against NumPy 2.3.5 in
 * tests/test_oracle_numpy_pins.py.
 *
 * Reference lines followed (relative to /root/reference/gsconverter/processing/):
 *   orc_sor_mean_dists     gpu_ops.py:98-176   (Taichi kernel sor_compute_mean_dists)
 *   orc_kmeans_assign      gpu_ops.py:57-73    (k_means_assign)
 *   orc_kmeans_update      gpu_ops.py:75-96    (k_means_update; serial f32 order, A.5)
 *   orc_pairwise_sum_f32   NumPy's pairwise float32 add.reduce,
 *   orc_mean_std_f32       as used by gpu_ops.py:259-260 / data_processor.py:176-177
 *   orc_pairwise_mean_f64_rows   np.mean(dists[:,1:],axis=1) data_processor.py:172
 *
 * Build: gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -shared -fPIC (see Makefile).
 * All float arithmetic is binary32 round-to-nearest, no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_KMAX 50

/* Python-style (floor) modulo, result in [0, n). */
static inline int64_t pymod64(int64_t a, int64_t n) {
    int64_t r = a % n;
    return r < 0 ? r + n : r;
}

/* probe hash, gpu_ops.py:130-132.  mode 0 = Taichi default-int (i32, wrapping
 * products) -- the faithful reading; mode 1 = int64 products -- what the
 * host-side table build uses (gpu_ops.py:222-223). */
static inline int32_t probe_hash(int32_t nx, int32_t ny, int32_t nz, int32_t hash_size, int mode) {
    if (mode == 0) {
        int32_t a = (int32_t)((uint32_t)nx * 73856093u);
        int32_t b = (int32_t)((uint32_t)ny * 19349663u);
        int32_t c = (int32_t)((uint32_t)nz * 83492791u);
        int32_t h = a ^ b ^ c;
        return (int32_t)pymod64((int64_t)h, (int64_t)hash_size);
    } else {
        int64_t h = ((int64_t)nx * 73856093LL) ^ ((int64_t)ny * 19349663LL) ^ ((int64_t)nz * 83492791LL);
        return (int32_t)pymod64(h, (int64_t)hash_size);
    }
}

/* gpu_ops.py:98-176.  pos is the hash-sorted float32[N,3] array. visits (may be
 * NULL) receives the per-query candidate-visit count V_i of SURVEY §8(d). */
void orc_sor_mean_dists(const float *pos, const int32_t *cell_start, const int32_t *cell_count,
                        float *mean_dists, float bminx, float bminy, float bminz, float cell_size,
                        int32_t hash_size, int64_t N, int32_t K, int hash_mode, int64_t *visits) {
    const float eps_cell = 1e-8f, eps_self = 1.0e-12f, big = 1.0e10f, valid_lim = 0.9e10f;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t i = 0; i < N; ++i) {
        const float px = pos[3 * i], py = pos[3 * i + 1], pz = pos[3 * i + 2];
        int64_t nv = 0;
        if (cell_size > eps_cell) {
            int32_t gx = (int32_t)floorf((px - bminx) / cell_size);
            int32_t gy = (int32_t)floorf((py - bminy) / cell_size);
            int32_t gz = (int32_t)floorf((pz - bminz) / cell_size);
            float dists[ORC_KMAX];
            for (int t = 0; t < ORC_KMAX; ++t) dists[t] = big;
            for (int dx = -1; dx < 2; ++dx)
                for (int dy = -1; dy < 2; ++dy)
                    for (int dz = -1; dz < 2; ++dz) {
                        int32_t h = probe_hash(gx + dx, gy + dy, gz + dz, hash_size, hash_mode);
                        int32_t start = cell_start[h], cnt = cell_count[h];
                        if (start == -1) continue;
                        nv += cnt;
                        for (int32_t j = start; j < start + cnt; ++j) {
                            float ax = px - pos[3 * (int64_t)j];
                            float ay = py - pos[3 * (int64_t)j + 1];
                            float az = pz - pos[3 * (int64_t)j + 2];
                            float d2 = (ax * ax + ay * ay) + az * az;
                            if (d2 > eps_self) {
                                float d = sqrtf(d2);
                                if (d < dists[K - 1]) {
                                    int ins = K - 1;
                                    while (ins > 0 && dists[ins - 1] > d) {
                                        dists[ins] = dists[ins - 1];
                                        --ins;
                                    }
                                    dists[ins] = d;
                                }
                            }
                        }
                    }
            float sum = 0.0f;
            int valid = 0;
            for (int ki = 0; ki < K; ++ki) {
                float v = dists[ki];
                if (v < valid_lim) {
                    sum += v;
                    ++valid;
                }
            }
            mean_dists[i] = valid > 0 ? sum / (float)valid : 0.0f;
        } else {
            mean_dists[i] = 0.0f;
        }
        if (visits) visits[i] = nv;
    }
}

/* ---------------------------------------------------------------- NumPy pairwise */

/* NumPy's float32 pairwise summation (loops_utils.h pairwise_sum, PW_BLOCKSIZE 128). */
float orc_pairwise_sum_f32(const float *a, int64_t n) {
    if (n < 8) {
        float res = 0.0f;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return orc_pairwise_sum_f32(a, n2) + orc_pairwise_sum_f32(a + n2, n - n2);
    }
}

static double pairwise_sum_f64(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.0;
        for (int64_t i = 0; i < n; ++i) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum_f64(a, n2) + pairwise_sum_f64(a + n2, n - n2);
    }
}

/* np.mean / np.std of a float32 vector (float32 accumulators), SURVEY A.1 step 9.
 * out[0] = mean, out[1] = std. */
void orc_mean_std_f32(const float *a, int64_t n, float *out) {
    /* the float32 sum divided by the element count: NumPy promotes float32 / intp to float64 and rounds the quotient
     * back to float32 (numpy/_core/_methods.py _mean / _var); identical to a float32 division only while n <= 2^24 */
    float mean = (float)((double)orc_pairwise_sum_f32(a, n) / (double)n);
    float *x = (float *)malloc((size_t)(n > 0 ? n : 1) * sizeof(float));
    for (int64_t i = 0; i < n; ++i) {
        float t = a[i] - mean;
        x[i] = t * t;
    }
    float var = (float)((double)orc_pairwise_sum_f32(x, n) / (double)n);
    free(x);
    out[0] = mean;
    out[1] = sqrtf(var);
}

/* np.mean(d[:,1:], axis=1) in float64 for a row-major [rows, k+1] matrix
 * (data_processor.py:172); the result is stored to float32 as :173 does. */
void orc_pairwise_mean_f64_rows(const double *d, int64_t rows, int64_t kp1, float *out) {
    for (int64_t r = 0; r < rows; ++r)
        out[r] = (float)(pairwise_sum_f64(d + r * kp1 + 1, kp1 - 1) / (double)(kp1 - 1));
}

/* ---------------------------------------------------------------- K-Means (A.5) */

/* gpu_ops.py:57-73 */
void orc_kmeans_assign(const float *X, const float *C, int32_t *labels, int64_t N, int32_t K, int32_t D) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        float min_dist = 1e20f;
        int32_t best = -1;
        const float *x = X + i * D;
        for (int32_t c = 0; c < K; ++c) {
            const float *cc = C + (int64_t)c * D;
            float dist = 0.0f;
            for (int32_t dim = 0; dim < D; ++dim) {
                float diff = x[dim] - cc[dim];
                dist += diff * diff;
            }
            if (dist < min_dist) {
                min_dist = dist;
                best = c;
            }
        }
        labels[i] = best;
    }
}

/* gpu_ops.py:75-96 with the serial index-order float32 accumulation fixed by
 * SURVEY A.5 (the reference uses float atomics in arbitrary order). */
void orc_kmeans_update(const float *X, float *C, const int32_t *labels, int32_t *counts, int64_t N, int32_t K,
                       int32_t D) {
    memset(C, 0, (size_t)K * D * sizeof(float));
    memset(counts, 0, (size_t)K * sizeof(int32_t));
    for (int64_t i = 0; i < N; ++i) {
        int32_t l = labels[i];
        if (l < 0) continue; /* never happens for finite data; guards NaN rows */
        float *c = C + (int64_t)l * D;
        const float *x = X + i * D;
        for (int32_t dim = 0; dim < D; ++dim) c[dim] += x[dim];
        counts[l] += 1;
    }
    for (int32_t c = 0; c < K; ++c) {
        int32_t cnt = counts[c];
        if (cnt > 0) {
            float inv = 1.0f / (float)cnt;
            for (int32_t dim = 0; dim < D; ++dim) C[(int64_t)c * D + dim] *= inv;
        }
    }
}

/* gpu_ops.py:186-188 loop: max_iter x (assign; update).  C holds the injected
 * initial centroids on entry and the last-updated centroids on exit; labels are the
 * ones of the last assign (one update behind C). */
void orc_kmeans_lloyd(const float *X, float *C, int32_t *labels, int32_t *counts, int64_t N, int32_t K, int32_t D,
                      int32_t max_iter) {
    for (int32_t it = 0; it < max_iter; ++it) {
        orc_kmeans_assign(X, C, labels, N, K, D);
        orc_kmeans_update(X, C, labels, counts, N, K, D);
    }
}

/* float32 cell size of gpu_ops.py:205-213 computed the way a C host would
 * (powf).  Kept here so the C-ABI's host helper can be checked against NumPy. */
float orc_powf(float a, float b) { return powf(a, b); }
