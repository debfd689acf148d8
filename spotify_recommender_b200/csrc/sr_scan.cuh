// sr_scan.cuh -- the hot kernel: fused cosine score + threshold filter + streaming
// exact top-K over the device-resident song store (replaces the reference's
// cublasSgemv + computeNormsKernel + normalizeSimilaritiesKernel + D2H + host
// heap, Recommender.cu:184-254 and :293-315, for a whole batch of queries).
//
// Work = (query tile x song tile) units dealt in contiguous runs to a persistent
// grid.  Per song tile every thread keeps S songs of the PRE-NORMALISED store in
// registers (as S/2 packed pairs) and walks the query tile held in shared memory:
//
//   filter (hot, 12 FMA per pair, issued as FFMA2 = two songs per instruction; the
//   query value is a UNIFORM-register scalar operand fed from constant memory, so an
//   FFMA2 reads only the song pair and the accumulator pair from the register file
//   -- measured on B200 this is what lifts the loop from ~70% to >85% of the FP32 pipe):
//       acc = -T' + sum_j fhat_j * qhat_j        T' = (exact running K-th best) - kEps
//     sign(acc) == 0  <=>  the pair MAY belong to the exact top-K  (DESIGN.md
//     "filter slack": |acc - (oracle score - T')| < kEps for regular rows; NaN
//     rows/queries always pass).  One LOP3 tree over the sign bits per S songs.
//   hit (rare): the song id is appended, unscored, to the query's candidate buffer.
//   settle (rare, warp-cooperative, at tile borders): unscored candidates are
//     scored in the reference's own arithmetic (Recommender.cu:263-271, unfused
//     mul/add, sqrt*qn, IEEE divide, clamp) from the RAW store; a radix select
//     keeps the best K exact keys (score desc, id asc), which raises T' and is
//     published to the other CTAs working on the same queries (g_best).
//   rescan (pathological tiles only: mass ties, NaN queries): when a tile yields
//     more hits than the buffer holds, the owning warp scores the tile exactly.
//
// The filter only ever discards pairs that provably are not in the exact top-K,
// so results are bit-identical to the oracle whatever the thresholds were.
#pragma once
#include "sr_device.cuh"

namespace sr {

// Normalised query rows of the current query group: the FFMA2 scalar operand comes
// from here through LDCU -> uniform register.  61440 B of the 64 KB constant bank.
constexpr int kConstQueries = 1280;
__constant__ float4 c_qhat[kConstQueries * 3];

struct ScanArgs {
    const float *hat;        // normalised rows, n_pad x 12 (pad rows 0, irregular rows NaN)
    const float *raw;        // raw rows, n_pad x 12
    const float *nf;         // exact row norms (reference order), n_pad
    int64_t n;               // valid local rows
    int32_t id_base;         // global id of local row 0
    int n_tiles;             // ceil(n / TS)
    const float *qraw;       // [nq][12] raw query rows            (all per-query arrays are
    const float *qn;         // [nq] exact query norms               already offset to the first
    const int32_t *exclude;  // [nq] global id to skip or -1         query of this group)
    int nq;                  // queries in this group (<= kConstQueries), rows of c_qhat
    int qt;                  // queries per tile (<= kQTMax)
    int K, prune_at, bufcap;
    uint64_t *cta_buf;       // [grid][qt][bufcap] candidate buffers (L2 resident)
    uint32_t *g_best;        // [nq] orderable exact K-th best score seen so far by anyone
    uint64_t *pool;          // [nq][segs*K] exact keys handed to finalize
    int32_t *pool_cnt;       // [nq]
    int segs;
    unsigned long long *stats;  // [0] hits [1] settles [2] rescans [3] rescored
};

__host__ __device__ inline size_t scan_smem_bytes(int qt, int warps)
{
    return (size_t)qt * (kF + 6) * 4 + (size_t)warps * 256 * 4;
}

__host__ __device__ inline int scan_segs(int grid, int nqt) { return (grid + nqt - 1) / nqt + 2; }

// -T' for the filter record; never +-0 (a -0 accumulator would read as "below").
__device__ __forceinline__ float neg_threshold(uint32_t best_ord)
{
    float t = ord2f(best_ord) - kEps;  // -inf stays -inf => +inf record => everything passes
    float nt = -t;
    return (nt == 0.0f) ? 1.0e-30f : nt;
}

struct QueryCtx {  // shared-memory views of one query tile
    float *nthr, *qraw, *qn;
    uint32_t *best;
    int *cnt, *excl, *qid;
};

// exact key of one (query, row) pair, 0 when the row is the excluded song
__device__ __forceinline__ uint64_t exact_key(const ScanArgs &a, int64_t row, const float *q, float qn, int32_t ex)
{
    const int32_t gid = a.id_base + (int32_t)row;
    if (gid == ex) return 0ull;
    float f[kF];
    load_row12(a.raw, row, f);
    return make_key(exact_score(f, __ldg(a.nf + row), q, qn), (uint32_t)gid);
}

// Keep the best K valid keys of buf[0,cnt) at the front (unordered).  Returns the
// new count; *kth = K-th best key when K valid keys exist, else 0.
__device__ __forceinline__ int warp_keep_topk(uint64_t *buf, int cnt, int K, uint32_t *hist, uint64_t *kth)
{
    const int lane = threadIdx.x & 31;
    uint64_t cut = 1ull;  // keep every valid key
    if (cnt > K) {
        int rank;
        const uint32_t vK = warp_radix_select(
            cnt, K, hist, [&](int i, uint32_t *w) { *w = (uint32_t)(__ldcg(buf + i) >> 32); return true; }, &rank);
        if (vK != 0u) {  // at least K valid keys: break the tie group by id
            int r2;
            const uint32_t lo = warp_radix_select(
                cnt, rank, hist,
                [&](int i, uint32_t *w) {
                    const uint64_t k = __ldcg(buf + i);
                    *w = (uint32_t)k;
                    return (uint32_t)(k >> 32) == vK;
                },
                &r2);
            cut = ((uint64_t)vK << 32) | lo;
        }
    }
    int newcnt = 0;
    uint64_t mn = ~0ull;
    for (int base = 0; base < cnt; base += 32) {
        const int i = base + lane;
        const uint64_t key = (i < cnt) ? __ldcg(buf + i) : 0ull;
        const bool p = key >= cut;
        const uint32_t b = __ballot_sync(0xffffffffu, p);
        const int pos = newcnt + __popc(b & ((1u << lane) - 1u));
        if (p) { __stcg(buf + pos, key); mn = key < mn ? key : mn; }
        newcnt += __popc(b);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, mn, off);
        mn = o < mn ? o : mn;
    }
    __syncwarp();
    *kth = (newcnt >= K) ? mn : 0ull;
    return newcnt;
}

// Settle one query's candidate buffer (whole warp): score unscored entries exactly,
// keep the best K, on overflow re-scan tile rows [tile_lo, tile_hi) exactly.
__device__ __forceinline__ void warp_settle(const ScanArgs &a, const QueryCtx &c, uint64_t *buf, int ql, int64_t tile_lo,
                                            int64_t tile_hi, uint32_t *hist)
{
    const int lane = threadIdx.x & 31;
    const int raw_cnt = c.cnt[ql];
    const bool overflow = raw_cnt > a.bufcap;
    int cnt = overflow ? a.bufcap : raw_cnt;
    float q[kF];
#pragma unroll
    for (int j = 0; j < kF; ++j) q[j] = c.qraw[ql * kF + j];
    const float qn = c.qn[ql];
    const int32_t ex = c.excl[ql];
    unsigned rescored = 0;
    for (int i = lane; i < cnt; i += 32) {
        const uint64_t k = __ldcg(buf + i);
        if ((uint32_t)(k >> 32) == kUnscored) {
            const int64_t row = (int64_t)key_id(k) - a.id_base;
            // entries of an overflowed tile are dropped here and found again by the re-scan
            const bool drop = overflow && row >= tile_lo && row < tile_hi;
            __stcg(buf + i, drop ? 0ull : exact_key(a, row, q, qn, ex));
            ++rescored;
        }
    }
    __syncwarp();
    uint64_t kth;
    cnt = warp_keep_topk(buf, cnt, a.K, hist, &kth);
    if (overflow) {
        const int chunk = ((a.bufcap - a.K) / 32) * 32;
        for (int64_t base = tile_lo; base < tile_hi; base += chunk) {
            const int m = (int)min((int64_t)chunk, tile_hi - base);
            for (int i = lane; i < m; i += 32) __stcg(buf + cnt + i, exact_key(a, base + i, q, qn, ex));
            rescored += (m + 31 - lane) / 32;
            __syncwarp();
            cnt = warp_keep_topk(buf, cnt + m, a.K, hist, &kth);
        }
    }
    if (lane == 0) {
        c.cnt[ql] = cnt;
        if (kth != 0ull) {
            const uint32_t b = (uint32_t)(kth >> 32);
            if (b > c.best[ql]) {
                c.best[ql] = b;
                atomicMax(a.g_best + c.qid[ql], b);
                c.nthr[ql] = neg_threshold(b);
            }
        }
        if (a.stats) {
            atomicAdd(a.stats + 1, 1ull);
            if (overflow) atomicAdd(a.stats + 2, 1ull);
        }
    }
    if (a.stats && rescored) atomicAdd(a.stats + 3, (unsigned long long)rescored);
    __syncwarp();
}

// One filter pass of a thread's S songs against query record `ql` of the tile.
template <int S>
__device__ __forceinline__ uint32_t filter_query(const float2 (&fp)[S / 2][kF], int cq, float nt, float2 (&acc)[S / 2])
{
    const float4 *r = c_qhat + cq * 3;
    const float4 q0 = r[0], q1 = r[1], q2 = r[2];
    const float q[kF] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
    for (int p = 0; p < S / 2; ++p) acc[p] = make_float2(nt, nt);
#pragma unroll
    for (int j = 0; j < kF; ++j)
#pragma unroll
        for (int p = 0; p < S / 2; ++p) acc[p] = __ffma2_rn(fp[p][j], make_float2(q[j], q[j]), acc[p]);
    uint32_t m = 0xffffffffu;
#pragma unroll
    for (int p = 0; p < S / 2; ++p) m &= __float_as_uint(acc[p].x) & __float_as_uint(acc[p].y);
    return m;  // sign bit clear <=> at least one song passes the filter
}

template <int S, int THREADS, int MINB, bool DEFER>
__global__ void __launch_bounds__(THREADS, MINB) scan_kernel(const ScanArgs a)
{
    constexpr int TS = S * THREADS;
    constexpr int WARPS = THREADS / 32;
    static_assert(S % 2 == 0, "songs per thread must be even");
    static_assert(kQTMax <= 32 * WARPS, "one lane per owned query in the tile epilogue");
    static_assert(kRowPad % TS == 0, "store padding must cover whole tiles");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    QueryCtx c;
    c.qraw = reinterpret_cast<float *>(smem_raw);
    c.nthr = c.qraw + a.qt * kF;
    c.qn = c.nthr + a.qt;
    c.best = reinterpret_cast<uint32_t *>(c.qn + a.qt);
    c.cnt = reinterpret_cast<int *>(c.best + a.qt);
    c.excl = c.cnt + a.qt;
    c.qid = c.excl + a.qt;
    uint32_t *s_hist = reinterpret_cast<uint32_t *>(c.qid + a.qt);

    const int G = gridDim.x;
    const int nqt = (a.nq + a.qt - 1) / a.qt;
    const int64_t U = (int64_t)nqt * a.n_tiles;
    int64_t u = (U * blockIdx.x) / G;
    const int64_t u_end = (U * (blockIdx.x + 1)) / G;
    uint64_t *mybuf = a.cta_buf + (size_t)blockIdx.x * a.qt * a.bufcap;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t *my_hist = s_hist + warp * 256;
    unsigned long long hits = 0;

    while (u < u_end) {
        const int qtile = (int)(u / a.n_tiles);
        const int t0 = (int)(u % a.n_tiles);
        const int t1 = (int)min((int64_t)a.n_tiles, (int64_t)t0 + (u_end - u));
        const int q0 = qtile * a.qt;
        const int nql = min(a.qt, a.nq - q0);

        // ---- segment prologue: bring the query tile's state into shared memory
        for (int ql = tid; ql < nql; ql += THREADS) {
            const int qid = q0 + ql;
            c.qid[ql] = qid;
            c.excl[ql] = a.exclude[qid];
            c.cnt[ql] = 0;
            c.qn[ql] = a.qn[qid];
            const uint32_t b = __ldcg(a.g_best + qid);
            c.best[ql] = b;
            c.nthr[ql] = neg_threshold(b);
        }
        for (int i = tid; i < nql * kF; i += THREADS) c.qraw[i] = a.qraw[(size_t)q0 * kF + i];
        __syncthreads();

        for (int tile = t0; tile < t1; ++tile) {
            const int64_t row0 = (int64_t)tile * TS + tid;

            // ---- S songs of the normalised store into registers, as packed pairs
            float2 fp[S / 2][kF];
#pragma unroll
            for (int p = 0; p < S / 2; ++p) {
                float r0[kF], r1[kF];
                load_row12(a.hat, row0 + (int64_t)(2 * p) * THREADS, r0);
                load_row12(a.hat, row0 + (int64_t)(2 * p + 1) * THREADS, r1);
#pragma unroll
                for (int j = 0; j < kF; ++j) fp[p][j] = make_float2(r0[j], r1[j]);
            }

            auto append = [&](int ql, const float2 (&acc)[S / 2]) {
#pragma unroll
                for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float v = h ? acc[p].y : acc[p].x;
                        if ((int)__float_as_uint(v) >= 0) {
                            const int64_t row = row0 + (int64_t)(2 * p + h) * THREADS;
                            if (row < a.n) {
                                const int slot = atomicAdd(&c.cnt[ql], 1);
                                if (slot < a.bufcap)
                                    __stcg(mybuf + (size_t)ql * a.bufcap + slot,
                                           ((uint64_t)kUnscored << 32) |
                                               (uint64_t)(0xFFFFFFFFu - (uint32_t)(a.id_base + (int32_t)row)));
                                ++hits;
                            }
                        }
                    }
                }
            };

            // ---- the hot loop: one query per iteration
            if (DEFER) {
                // the sign test of query ql is consumed one iteration later, so its LOP3 tree
                // overlaps the next query's FFMA2 stream; a hit re-runs the filter (rare path)
                uint32_t m_prev = 0xffffffffu;
#pragma unroll 2
                for (int ql = 0; ql < nql; ++ql) {
                    if ((int)m_prev >= 0) {
                        float2 acc2[S / 2];
                        filter_query<S>(fp, q0 + ql - 1, c.nthr[ql - 1], acc2);
                        append(ql - 1, acc2);
                    }
                    float2 acc[S / 2];
                    m_prev = filter_query<S>(fp, q0 + ql, c.nthr[ql], acc);
                }
                if ((int)m_prev >= 0) {
                    float2 acc2[S / 2];
                    filter_query<S>(fp, q0 + nql - 1, c.nthr[nql - 1], acc2);
                    append(nql - 1, acc2);
                }
            } else {
#pragma unroll 2
                for (int ql = 0; ql < nql; ++ql) {
                    float2 acc[S / 2];
                    const uint32_t m = filter_query<S>(fp, q0 + ql, c.nthr[ql], acc);
                    if ((int)m >= 0) append(ql, acc);
                }
            }
            __syncthreads();

            // ---- tile epilogue: warp w looks after queries ql == w (mod WARPS), one lane each:
            // adopt thresholds published by other CTAs, settle buffers that filled up.
            {
                const int ql_mine = warp + WARPS * lane;
                bool need = false;
                if (ql_mine < nql) {
                    const uint32_t g = __ldcg(a.g_best + c.qid[ql_mine]);
                    if (g > c.best[ql_mine]) {
                        c.best[ql_mine] = g;
                        c.nthr[ql_mine] = neg_threshold(g);
                    }
                    need = c.cnt[ql_mine] > a.prune_at;
                }
                uint32_t todo = __ballot_sync(0xffffffffu, need);
                const int64_t tile_lo = (int64_t)tile * TS;
                const int64_t tile_hi = min(a.n, tile_lo + TS);
                while (todo) {
                    const int l = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ql = warp + WARPS * l;
                    warp_settle(a, c, mybuf + (size_t)ql * a.bufcap, ql, tile_lo, tile_hi, my_hist);
                }
            }
            __syncthreads();
        }

        // ---- segment epilogue: hand the exact survivors to the per-query pool
        for (int ql = warp; ql < nql; ql += WARPS) {
            if (c.cnt[ql] == 0) continue;
            uint64_t *buf = mybuf + (size_t)ql * a.bufcap;
            warp_settle(a, c, buf, ql, 0, 0, my_hist);
            const int cnt = c.cnt[ql];
            if (cnt == 0) continue;
            int base = 0;
            if (lane == 0) base = atomicAdd(a.pool_cnt + q0 + ql, cnt);
            base = __shfl_sync(0xffffffffu, base, 0);
            uint64_t *slab = a.pool + (size_t)(q0 + ql) * a.segs * a.K;
            for (int i = lane; i < cnt; i += 32) slab[base + i] = __ldcg(buf + i);
        }
        __syncthreads();
        u += (t1 - t0);
    }
    if (a.stats && hits) atomicAdd(a.stats + 0, hits);
}

}  // namespace sr
