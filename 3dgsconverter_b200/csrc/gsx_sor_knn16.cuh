// gsx_sor_knn16.cuh -- the K <= 16 form of the neighbour search (included by gsx_sor.cu after k_sor_knn).
//
// k_sor_knn gives a whole warp to one query: its sorted K-list lives one rank per lane, so with K = 16 (the benchmark
// configuration, and every --sor_intensity up to 2.3) half of every list instruction is wasted, and ~60 % of the ~700
// warp instructions per query are list maintenance (serial inserts at ~11, 32-lane bitonic merges at ~95).  Here a warp
// carries TWO queries, one per 16-lane half: rank r of a half's list lives in lane r of that half, a scan step looks at
// 16 candidates per half, a serial insert serves both halves with one instruction stream, and the merge network is the
// 16-lane one (10 + 4 compare-exchange stages instead of 15 + 5).  Each half owns its own batch of 16 consecutive
// hash-sorted positions and the two halves walk their batches in lock step (query t of both batches together); every
// shuffle / vote / redux names only the half's lanes, so the halves may diverge freely where their searches differ
// (different probe order, different pruning) and run together where they do not.
//
// Semantics are those of k_sor_knn (gpu_ops.py:98-176): same candidate set per query (the buckets of the 27 probes),
// same float32 op sequence for d^2, exact box pruning, and only the multiset of the K smallest d^2 matters -- the
// visiting order is free.  With 16 lanes a lane carries two of the 27 probes (p and p + 16) and, inside a long bucket,
// two of the 32 chunk boxes of a super.

struct TopK16 {
    float v;     // lane hl holds rank hl of the ascending d^2 list
    float tau;   // rank K-1, uniform inside the half
    int K;
    __device__ __forceinline__ void init(int k) {
        K = k;
        v = tau = __uint_as_float(GSX_D2LIM_BITS);
    }
    __device__ __forceinline__ void refresh_tau(unsigned hmask, int base) { tau = __shfl_sync(hmask, v, base + K - 1); }
    // insert x (uniform inside the half, < tau)
    __device__ __forceinline__ void insert(float x, unsigned hmask, int base, int hl) {
        float up = __shfl_up_sync(hmask, v, 1, 16);
        if (hl == 0) up = 0.f;
        if (v > x) v = fmaxf(up, x);
        refresh_tau(hmask, base);
    }
    // merge one candidate per lane (sentinel where there is none): bitonic sort of the 16 new values, reversed,
    // lane-wise min with the ascending list = the 16 smallest of the union as a bitonic sequence, 4 stages sort it
    __device__ __forceinline__ void merge16(float nv, unsigned hmask, int base, int hl) {
#pragma unroll
        for (int k = 2; k <= 16; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const float o = __shfl_xor_sync(hmask, nv, j);
                nv = (((hl & k) == 0) == ((hl & j) == 0)) ? fminf(nv, o) : fmaxf(nv, o);
            }
        }
        const float r = __shfl_sync(hmask, nv, base + 15 - hl);
        float m = fminf(v, r);
#pragma unroll
        for (int j = 8; j > 0; j >>= 1) {
            const float o = __shfl_xor_sync(hmask, m, j);
            m = ((hl & j) == 0) ? fminf(m, o) : fmaxf(m, o);
        }
        v = m;
        refresh_tau(hmask, base);
    }
};

#ifndef GSX_MERGE16_THRESHOLD
#define GSX_MERGE16_THRESHOLD 5
#endif

// 16 candidates of one half: positions j (valid where inside the bucket)
template <bool STATS>
__device__ __forceinline__ void scan16(const float4* __restrict__ spos, pos_t j, bool valid, float qx, float qy, float qz,
                                       TopK16& tk, unsigned hmask, int base, int hl, unsigned long long& n_scanned) {
    float d2 = INFINITY;
    if (valid) {
        const float4 c = __ldg(spos + j);
        const float ax = __fsub_rn(qx, c.x), ay = __fsub_rn(qy, c.y), az = __fsub_rn(qz, c.z);
        d2 = __fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az));
    }
    if (STATS) n_scanned += __popc(__ballot_sync(hmask, valid));
    const bool pass = valid && d2 > 1.0e-12f && d2 < tk.tau;
    unsigned m = __ballot_sync(hmask, pass);
    if (__popc(m) >= GSX_MERGE16_THRESHOLD) {
        tk.merge16(pass ? d2 : __uint_as_float(GSX_D2LIM_BITS), hmask, base, hl);
        return;
    }
    while (m) {
        const int src = __ffs(m) - 1;   // absolute lane (the vote only has this half's bits)
        m &= m - 1;
        const float x = __shfl_sync(hmask, d2, src);
        if (x < tk.tau) tk.insert(x, hmask, base, hl);
    }
}

#ifndef GSX_KNN16_MINBLOCKS
#define GSX_KNN16_MINBLOCKS 6
#endif
template <bool STATS>
__global__ void __launch_bounds__(256, GSX_KNN16_MINBLOCKS)
    k_sor_knn16(const float4* __restrict__ spos, const int2* __restrict__ tab_se, const float4* __restrict__ tab_box,
                const uint32_t* __restrict__ cellbits, const float4* __restrict__ caabb, const float4* __restrict__ saabb,
                float* __restrict__ final_means, unsigned int* __restrict__ work, int64_t q_begin, int64_t q_end,
                int q_stride, int q_phase, int K, int hash_mode, float bx, float by, float bz, float cell, uint32_t n,
                uint64_t M, unsigned long long* __restrict__ stats) {
    static_assert(kQueryBatch == 16, "a half-warp batch is one 16-position run of a cellbits word");
    const int lane = lane_id();
    const int base = lane & 16, hl = lane & 15;
    const unsigned hmask = base ? 0xffff0000u : 0x0000ffffu;
    unsigned long long st_visits = 0, st_scanned = 0, st_boxes = 0, st_queries = 0;
    // probes of lane hl: p0 = hl (always < 27) and p1 = hl + 16 (< 27 for hl < 11), reference loop order (dx outer)
    const int p1 = hl + 16;
    const bool has1 = p1 < 27;
    const int dx0 = hl / 9 - 1, dy0 = (hl / 3) % 3 - 1, dz0 = hl % 3 - 1;
    const int dx1 = p1 / 9 - 1, dy1 = (p1 / 3) % 3 - 1, dz1 = p1 % 3 - 1;

    for (;;) {
        unsigned int b0 = 0;
        if (lane == 0) b0 = atomicAdd(work, 2u * (unsigned)kQueryBatch);
        b0 = __shfl_sync(GSX_FULL, b0, 0);
        // the warp took two consecutive batches of this launch; half h works on batch b0/16 + h (global batch number
        // (.)*q_stride + q_phase, see k_sor_knn)
        const int64_t qb_lo = q_begin + ((int64_t)(b0 / kQueryBatch) * q_stride + q_phase) * kQueryBatch;
        if (qb_lo >= q_end) break;   // the lower batch is the earlier one: nothing left for either half
        const int64_t qb = q_begin + ((int64_t)(b0 / kQueryBatch + (base >> 4)) * q_stride + q_phase) * kQueryBatch;
        const int64_t qe = qb + kQueryBatch < q_end ? qb + kQueryBatch : q_end;
        const uint32_t cellword = qb < q_end ? __ldg(cellbits + (qb >> 5)) : 0u;
        int ps0 = 0, pc0 = 0, ps1 = 0, pc1 = 0;
        float l0x = 0.f, l0y = 0.f, l0z = 0.f, h0x = 0.f, h0y = 0.f, h0z = 0.f;
        float l1x = 0.f, l1y = 0.f, l1z = 0.f, h1x = 0.f, h1y = 0.f, h1z = 0.f;
#pragma unroll 1
        for (int64_t i = qb; i < qe; ++i) {
            const float4 q = __ldg(spos + i);
            const uint32_t w_i = (i >> 5) == (qb >> 5) ? cellword : __ldg(cellbits + (i >> 5));
            if (i == qb || ((w_i >> (i & 31)) & 1u)) {   // uniform inside the half
                const int gx = (int)floorf(__fdiv_rn(__fsub_rn(q.x, bx), cell));
                const int gy = (int)floorf(__fdiv_rn(__fsub_rn(q.y, by), cell));
                const int gz = (int)floorf(__fdiv_rn(__fsub_rn(q.z, bz), cell));
                {
                    const uint32_t h = probe_hash(gx + dx0, gy + dy0, gz + dz0, n, M, hash_mode);
                    const int2 se = __ldg(tab_se + h);
                    ps0 = se.x, pc0 = se.y - se.x;
                    if (pc0 > 0) {
                        const float4 a = __ldg(tab_box + 2 * (size_t)h), b = __ldg(tab_box + 2 * (size_t)h + 1);
                        l0x = a.x, l0y = a.y, l0z = a.z, h0x = b.x, h0y = b.y, h0z = b.z;
                    }
                }
                ps1 = 0, pc1 = 0;
                if (has1) {
                    const uint32_t h = probe_hash(gx + dx1, gy + dy1, gz + dz1, n, M, hash_mode);
                    const int2 se = __ldg(tab_se + h);
                    ps1 = se.x, pc1 = se.y - se.x;
                    if (pc1 > 0) {
                        const float4 a = __ldg(tab_box + 2 * (size_t)h), b = __ldg(tab_box + 2 * (size_t)h + 1);
                        l1x = a.x, l1y = a.y, l1z = a.z, h1x = b.x, h1y = b.y, h1z = b.z;
                    }
                }
            }
            if (STATS) {
                int tot = pc0 + pc1;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) tot += __shfl_xor_sync(hmask, tot, o);
                if (hl == 0) {
                    st_visits += (unsigned long long)tot;
                    st_queries += 1;
                }
            }
            TopK16 tk;
            tk.init(K);
            // lower bounds of d^2 to the probes' boxes (same monotone op sequence as d^2, see box_lb)
            unsigned pk0 = 0xffffffffu, pk1 = 0xffffffffu;
            if (pc0 > 0) {
                const float dx = fmaxf(fmaxf(__fsub_rn(l0x, q.x), __fsub_rn(q.x, h0x)), 0.f);
                const float dy = fmaxf(fmaxf(__fsub_rn(l0y, q.y), __fsub_rn(q.y, h0y)), 0.f);
                const float dz = fmaxf(fmaxf(__fsub_rn(l0z, q.z), __fsub_rn(q.z, h0z)), 0.f);
                pk0 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            }
            if (pc1 > 0) {
                const float dx = fmaxf(fmaxf(__fsub_rn(l1x, q.x), __fsub_rn(q.x, h1x)), 0.f);
                const float dy = fmaxf(fmaxf(__fsub_rn(l1y, q.y), __fsub_rn(q.y, h1y)), 0.f);
                const float dz = fmaxf(fmaxf(__fsub_rn(l1z, q.z), __fsub_rn(q.z, h1z)), 0.f);
                pk1 = __float_as_uint(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            }

            // seed from the 32-position chunk that holds the query when its own bucket is a long one (legal only if
            // the centre probe -- lane 13, slot 0 -- really reaches the query's bucket range)
            int skip_chunk = -1;
            {
                const int s13 = __shfl_sync(hmask, ps0, base + 13), c13 = __shfl_sync(hmask, pc0, base + 13);
                if (c13 > kSmallBucket && i >= s13 && i < (int64_t)s13 + c13) {
                    skip_chunk = (int)(i >> 5);
                    const pos_t j0 = ((pos_t)skip_chunk << 5) + hl;
                    scan16<STATS>(spos, j0, j0 >= s13 && j0 < (pos_t)s13 + c13, q.x, q.y, q.z, tk, hmask, base, hl, st_scanned);
                    scan16<STATS>(spos, j0 + 16, j0 + 16 >= s13 && j0 + 16 < (pos_t)s13 + c13, q.x, q.y, q.z, tk, hmask, base,
                                  hl, st_scanned);
                }
            }

#pragma unroll 1
            for (;;) {
                const unsigned mine = pk0 < pk1 ? pk0 : pk1;
                const unsigned mp = __reduce_min_sync(hmask, mine);
                if (mp == 0xffffffffu || !(__uint_as_float(mp) < tk.tau)) break;
                const int pl = __ffs(__ballot_sync(hmask, mine == mp)) - 1;   // absolute lane holding the nearest probe
                const bool slot0 = pk0 == mp;                                 // (meaningful in lane pl)
                const int s = __shfl_sync(hmask, slot0 ? ps0 : ps1, pl), c = __shfl_sync(hmask, slot0 ? pc0 : pc1, pl);
                const bool centre = __shfl_sync(hmask, (int)(slot0 && hl == 13), pl) != 0;
                if (lane == pl) {
                    if (slot0) pk0 = 0xffffffffu;
                    else pk1 = 0xffffffffu;
                }
                const pos_t e = (pos_t)s + c;
                if (c <= kSmallBucket) {
#pragma unroll 1
                    for (pos_t b = s; b < e; b += 16)
                        scan16<STATS>(spos, b + hl, b + hl < e, q.x, q.y, q.z, tk, hmask, base, hl, st_scanned);
                    continue;
                }
                // long bucket: supers (1024 positions) nearest box first, then their chunks (32 positions)
                const int skip = centre ? skip_chunk : -1;
                const int fc = s >> 5, lc = (int)((e - 1) >> 5);
                const int fs = fc >> 5, ls = lc >> 5;
                for (int sb = fs; sb <= ls; sb += 16) {
                    const int sid = sb + hl;
                    unsigned skey = 0xffffffffu;
                    if (sid <= ls) {
                        const float lb = box_lb(saabb, sid, q.x, q.y, q.z);
                        if (lb < tk.tau) skey = __float_as_uint(lb);
                    }
                    if (STATS) st_boxes += __popc(__ballot_sync(hmask, sid <= ls));
                    for (;;) {
                        const unsigned ms = __reduce_min_sync(hmask, skey);
                        if (ms == 0xffffffffu || !(__uint_as_float(ms) < tk.tau)) break;
                        const int sl = __ffs(__ballot_sync(hmask, skey == ms)) - 1;
                        if (lane == sl) skey = 0xffffffffu;
                        const int sup = sb + (sl & 15);
                        // the 32 chunks of the super: lane hl holds chunks hl and hl + 16
                        const int cid0 = sup * 32 + hl, cid1 = cid0 + 16;
                        const bool cv0 = cid0 >= fc && cid0 <= lc && cid0 != skip;
                        const bool cv1 = cid1 >= fc && cid1 <= lc && cid1 != skip;
                        unsigned ck0 = 0xffffffffu, ck1 = 0xffffffffu;
                        if (cv0) {
                            const float lb = box_lb(caabb, cid0, q.x, q.y, q.z);
                            if (lb < tk.tau) ck0 = __float_as_uint(lb);
                        }
                        if (cv1) {
                            const float lb = box_lb(caabb, cid1, q.x, q.y, q.z);
                            if (lb < tk.tau) ck1 = __float_as_uint(lb);
                        }
                        if (STATS) st_boxes += __popc(__ballot_sync(hmask, cv0)) + __popc(__ballot_sync(hmask, cv1));
                        for (;;) {
                            const unsigned cmine = ck0 < ck1 ? ck0 : ck1;
                            const unsigned mc = __reduce_min_sync(hmask, cmine);
                            if (mc == 0xffffffffu || !(__uint_as_float(mc) < tk.tau)) break;
                            const int cl = __ffs(__ballot_sync(hmask, cmine == mc)) - 1;
                            const bool cs0 = ck0 == mc;
                            const int chunk = __shfl_sync(hmask, cs0 ? cid0 : cid1, cl);
                            if (lane == cl) {
                                if (cs0) ck0 = 0xffffffffu;
                                else ck1 = 0xffffffffu;
                            }
                            const pos_t j0 = ((pos_t)chunk << 5) + hl;
                            scan16<STATS>(spos, j0, j0 >= s && j0 < e, q.x, q.y, q.z, tk, hmask, base, hl, st_scanned);
                            scan16<STATS>(spos, j0 + 16, j0 + 16 >= s && j0 + 16 < e, q.x, q.y, q.z, tk, hmask, base, hl,
                                          st_scanned);
                        }
                    }
                }
            }

            // gpu_ops.py:163-174: ascending serial float32 sum of the valid (< 0.9e10) distances (a prefix of the list)
            const float d = __fsqrt_rn(tk.v);
            const int valid = __popc(__ballot_sync(hmask, hl < K && d < 0.9e10f));
            float sum = 0.f;
            for (int r = 0; r < valid; ++r) sum = __fadd_rn(sum, __shfl_sync(hmask, d, base + r));
            if (hl == 0) final_means[__float_as_int(q.w)] = valid > 0 ? __fdiv_rn(sum, (float)valid) : 0.f;
        }
    }
    if (STATS && hl == 0) {
        atomicAdd(stats + 0, st_visits);
        atomicAdd(stats + 1, st_scanned);
        atomicAdd(stats + 2, st_boxes);
        atomicAdd(stats + 3, st_queries);
    }
}
