// sr_scan.cuh -- the hot kernel: fused cosine score + threshold filter + streaming
// exact top-K over the device-resident song store (replaces the reference's
// cublasSgemv + computeNormsKernel + normalizeSimilaritiesKernel + D2H + host
// heap, Recommender.cu:184-254 and :293-315, for a whole batch of queries).
//
// Work = (query tile x song tile) units on a persistent grid: a CTA serves one query tile
// (<= 256 queries, state in shared memory) and claims song tiles from a per-query-tile counter
// (DYN shapes) or walks a contiguous run of units (static shapes).  Per song tile every thread
// keeps S songs of the PRE-NORMALISED store in registers (as S/2 packed pairs), loaded straight
// from global memory or, in the small-batch shape, staged through shared memory by TMA:
//
//   filter (hot, 12 FMA per pair, issued as FFMA2 = two songs per instruction; the
//   query value is a UNIFORM-register scalar operand fed from constant memory, so an
//   FFMA2 reads only the song pair and the accumulator pair from the register file
//   -- measured on B200 this is what lifts the loop from ~70% to >85% of the FP32 pipe):
//       acc = -T' + sum_j fhat_j * qhat_j        T' = (best known bound of the K-th best) - kEps
//     sign(acc) == 0  <=>  the pair MAY belong to the exact top-K  (DESIGN.md
//     "filter slack": |acc - (oracle score - T')| < kEps for regular rows; NaN
//     rows/queries always pass).  One LOP3 tree over the sign bits per S songs, one bit
//     per query in a mask: the loop is branch-free.
//   hit (rare): picked up after every 32 queries; the filter is re-run for the flagged query
//     and the song id is appended, unscored, to the query's hit buffer (shared memory).
//   settle (rare, warp-cooperative, at tile borders): pending ids are scored in the
//     reference's own arithmetic (Recommender.cu:263-271, unfused mul/add, sqrt*qn, IEEE
//     divide, clamp) from the RAW store and merged into this CTA's exact top-K list of the
//     query (shared memory; score desc, id asc).  Threshold feedback is lock-free: a full
//     list's K-th key, and the K-th largest of the query's global residue slots (gslot:
//     atomicMax of every exact score into slot id mod nslot, i.e. the bests of nslot disjoint
//     sets of songs), are lower bounds of the final K-th best; the better one raises T' and is
//     published (g_best, atomicMax).
//   overflow: when a tile yields more hits than the buffer holds (clustered stores), the
//     buffered hits raise the threshold and the CTA re-filters the tile against it; if the
//     threshold did not move (mass ties, NaN queries) the owning warp scores the tile exactly.
//   flush: after its last tile of a query tile the CTA appends the list entries that can
//     still matter to the per-query pool -- and, in the DYN shapes, the exact keys of the hits
//     still unsettled at that point, scored 32 at a time across all of a warp's queries;
//     finalize_kernel merges the pool into the ordered top-K.
//   bound pass (bound_kernel, before the scan): starting thresholds at filter speed -- this
//     store's own, or (row shards of one store) from block maxima max-reduced across the shards,
//     which bound the K-th best of the WHOLE store: a shard's list may then end up shorter than K.
//   ceilings (a.ceil): pass p > 0 of a K > 1024 query admits only keys below the last key of pass p - 1.
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
    const float *hat;        // normalised rows in the kernel shape's pair-interleaved tile layout
                             // (hat_offset(); pad rows 0, irregular rows NaN)
    const float *raw;        // raw rows, n_pad x 12
    const float *nf;         // exact row norms (reference order), n_pad
    int64_t n;               // valid local rows
    int32_t id_base;         // global id of local row 0
    int n_tiles;             // song tiles this launch visits ...
    int tile_stride;         // ... tile t of the launch is store tile t * tile_stride (1: every tile)
    int upc, extra;          // work units per CTA: CTA b owns upc + (b < extra) units, in order
    int cpq;                 // DYN shapes: CTAs per query tile (the first grid - cpq*nqt tiles get one more)
    int *tile_ctr;           // DYN shapes: [nqt] next unclaimed song tile of each query tile
    int slab;                // keys per query in the pool (segs x K, plus room for a segment's unsettled hits in DYN shapes)
    uint64_t *list_ws;       // non-null: the CTAs' top-K lists live here ([grid][qt][K], L2-resident) instead of in shared
                             // memory -- long lists would otherwise halve the query tile
    int *visit_ctr;          // DYN shapes: [nqt] CTAs that joined a query tile after finishing their own
    int steal_max;           // DYN shapes: how many such late joiners a query tile admits (sizes the pool slab)
    const float *qraw;       // [nq][12] raw query rows            (all per-query arrays are
    const float *qn;         // [nq] exact query norms               already offset to the first
    const int32_t *exclude;  // [nq] global id to skip or -1         query of this group)
    int nq;                  // queries in this group (<= kConstQueries), rows of c_qhat
    int qt;                  // queries per tile (<= kQTMax)
    int K;
    int cap;                 // hit-buffer entries per query (shared memory)
    int settle_at;           // a settle phase scores every hit buffer holding at least this many ids ...
    int trigger_at;          // ... and is triggered when some buffer reaches this many (>= settle_at)
    int refresh_every;       // tiles between two looks at the thresholds other CTAs published (an L2 round
                             // trip: every tile for large query tiles, rarer when a tile is only a few queries)
    // per-query exact survivors of every CTA segment, merged by finalize_kernel
    uint64_t *pool;          // [nq][slab] exact keys
    int32_t *pool_cnt;       // [nq] keys in the pool slab
    int segs;                // slab capacity in segments (scan_segs())
    uint32_t *g_best;        // [nq] orderable score: best known lower bound of the final K-th best
    uint32_t *gslot;         // [nq][nslot] lock-free global feedback: slot (id mod nslot) holds the best exact
                             // score (orderable) any CTA has found among the songs with that residue
    int nslot;               // >= K, multiple of 32
    int prefetch;            // DYN shapes: bulk-prefetch the next song tile into L2 while the current one is multiplied
    int bound_finish;        // bound_kernel: 1 = fold the block maxima into g_best (0: they are exchanged between shards first)
    const uint64_t *ceil;    // null, or [nq] per-query ceiling keys: only keys BELOW the ceiling are admitted (pass p > 0 of a
                             // K > kKMax query continues below the last key of pass p - 1); offset like the other per-query arrays
    unsigned long long *stats;  // [0] hits [1] settles [2] rescans [3] rescored [4] refilters; [5..12] cycle counters (-DSR_SCAN_TIMING)
};

// `stage_bytes`: size of the TMA staging buffers (0 for unstaged shapes)
__host__ __device__ inline size_t scan_smem_bytes(int qt, int cap, int K, size_t stage_bytes = 0, bool lists_in_smem = true)
{
    return (stage_bytes ? stage_bytes + 16 : 0) + (lists_in_smem ? (size_t)qt * K * 8 : 0) + (size_t)qt * (kF + 10 + cap) * 4 + 64;
}

// ---- TMA (bulk async copy) staging of song tiles: global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_load_tile(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    // the buffer was last read through the generic proxy (LDS): order those reads before the async-proxy write
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// CTA runs (segments) that can touch one query tile = slab capacity of the pool, in lists:
// a query tile is n_tiles consecutive units and every CTA owns at least `upc` of them
__host__ __device__ inline int scan_segs(int n_tiles, int upc) { return (n_tiles + upc - 1) / upc + 2; }

// -T' for the filter; never +-0 (a -0 accumulator would read as "below").
__device__ __forceinline__ float neg_threshold(uint32_t best_ord)
{
    float t = ord2f(best_ord) - kEps;  // -inf stays -inf => +inf => everything passes
    float nt = -t;
    return (nt == 0.0f) ? 1.0e-30f : nt;
}

struct QueryCtx {  // shared-memory views of one query tile
    uint64_t *list;  // [qt][K] this CTA's exact top-K so far (unordered), per query
    float *nthr, *qraw, *qn;
    uint32_t *best;
    int *cnt, *lcnt, *excl, *qid;
    int *dirty;      // [qt] a tile was re-filtered / re-scanned for this query in this segment: pending hits may repeat list entries
    uint32_t *hit;   // [qt][cap] global ids that passed the filter, not yet scored
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

// ---- the CTA-local per-query list (shared memory, one warp at a time per query) ---------
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = __shfl_xor_sync(0xffffffffu, v, off);
        v = o < v ? o : v;
    }
    return v;
}

// Up to 128 exact keys (four per lane, 0 = none) merged into the list of query ql: the best
// K keys survive; duplicates (a song met twice after a tile re-scan) are ignored.  For
// K <= 128 the list is pulled into registers (4 keys per lane), merged there and written
// back once; longer lists are updated in place, one candidate at a time.  Returns the
// list's minimum when it is full (the exact K-th best of the songs this CTA has seen), else 0.
__device__ __forceinline__ uint64_t list_merge4(const ScanArgs &a, const QueryCtx &c, int ql, const uint64_t (&key)[4])
{
    const int lane = threadIdx.x & 31;
    uint64_t *list = c.list + (size_t)ql * a.K;
    int n = c.lcnt[ql];
    uint32_t cand[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) cand[r] = __ballot_sync(0xffffffffu, key[r] != 0ull);
    uint64_t g = 0ull;  // minimum of the full list
    if (a.K <= 128) {
        uint64_t s[4];  // slot j of lane l <-> list[j * 32 + l]; empty slots hold ~0 (never the minimum)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[j] = (j * 32 + lane < n) ? list[j * 32 + lane] : ~0ull;
        auto list_min = [&]() {
            uint64_t m = s[0] < s[1] ? s[0] : s[1];
            const uint64_t m23 = s[2] < s[3] ? s[2] : s[3];
            return warp_min_u64(m < m23 ? m : m23);
        };
        if (n == a.K) g = list_min();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t cm = cand[r];
            while (cm) {
                const int l = __ffs(cm) - 1;
                cm &= cm - 1;
                const uint64_t k = __shfl_sync(0xffffffffu, key[r], l);
                if (n == a.K && k <= g) continue;
                const bool dup = (s[0] == k) | (s[1] == k) | (s[2] == k) | (s[3] == k);
                if (__any_sync(0xffffffffu, dup)) continue;
                if (n < a.K) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (n == j * 32 + lane) s[j] = k;
                    ++n;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (s[j] == g) s[j] = k;  // keys are unique: one slot of one lane
                }
                if (n == a.K) g = list_min();
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (s[j] != ~0ull) list[j * 32 + lane] = s[j];
    } else {
        auto scan_min = [&](uint64_t probe, bool *dup, int *pos) {  // list minimum, its position, and whether probe is present
            uint64_t m = ~0ull;
            int p = -1;
            bool d = false;
            for (int i = lane; i < n; i += 32) {
                const uint64_t v = list[i];
                d |= (v == probe);
                if (v < m) { m = v; p = i; }
            }
            const uint64_t gm = warp_min_u64(m);
            *dup = __any_sync(0xffffffffu, d);
            const uint32_t owner = __ballot_sync(0xffffffffu, m == gm && p >= 0);
            *pos = __shfl_sync(0xffffffffu, p, owner ? __ffs(owner) - 1 : 0);
            return gm;
        };
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t cm = cand[r];
            while (cm) {
                const int l = __ffs(cm) - 1;
                cm &= cm - 1;
                const uint64_t k = __shfl_sync(0xffffffffu, key[r], l);
                if (n == a.K && g != 0ull && k <= g) continue;
                bool dup;
                int pos;
                const uint64_t m = scan_min(k, &dup, &pos);
                if (dup) continue;
                if (n < a.K) {
                    if (lane == 0) list[n] = k;
                    ++n;
                    g = 0ull;  // recomputed below / on the next full-list insertion
                } else if (k > m) {
                    if (lane == 0) list[pos] = k;
                    g = 0ull;
                } else {
                    g = m;
                }
                __syncwarp();
            }
        }
        if (n == a.K && g == 0ull) {
            bool dup;
            int pos;
            g = scan_min(0ull, &dup, &pos);
        }
    }
    __syncwarp();
    if (lane == 0) c.lcnt[ql] = n;
    return (n == a.K) ? g : 0ull;
}

// Settle one query's hit buffer (whole warp): score the pending hits in the reference's
// arithmetic (four independent load chains per lane) and merge them into the CTA's list;
// when the buffer overflowed during this tile, score tile rows [tile_lo, tile_hi)
// exhaustively instead (nothing is ever lost).  A full list's minimum is the exact K-th
// best of the songs this CTA has seen -- a lower bound of the final K-th best -- and is
// published to every CTA working on the same query (g_best).
// An overflowed buffer (more than `cap` hits in this tile) is handled in two ways: normally
// the buffered hits alone raise the threshold (they contain >= cap candidates), the query is
// put on `redo` and the whole CTA re-filters the tile -- still in registers -- against the
// raised threshold; if the threshold did not move (mass ties) the tile is scored exhaustively.
__device__ __forceinline__ void warp_settle(const ScanArgs &a, const QueryCtx &c, int ql, int64_t tile_lo,
                                            int64_t tile_hi, int *redo, int *redo_cnt)
{
    const int lane = threadIdx.x & 31;
    const int raw_cnt = c.cnt[ql];
    const bool overflow = raw_cnt > a.cap;
    const int cnt = overflow ? a.cap : raw_cnt;
    float q[kF];
#pragma unroll
    for (int j = 0; j < kF; ++j) q[j] = c.qraw[ql * kF + j];
    const float qn = c.qn[ql];
    const int32_t ex = c.excl[ql];
    const uint32_t *hit = c.hit + (size_t)ql * a.cap;
    // only keys above the current threshold can enter (they are compared again inside the merge)
    const uint32_t best_before = c.best[ql];
    const uint64_t floor_key = (uint64_t)best_before << 32;
    const uint64_t ceil_key = a.ceil ? __ldg(a.ceil + c.qid[ql]) : ~0ull;
    uint64_t kth = 0ull;
    for (int base = 0; base < cnt; base += 128) {
        uint64_t key[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int i = base + r * 32 + lane;
            key[r] = (i < cnt) ? exact_key(a, (int64_t)hit[i] - a.id_base, q, qn, ex) : 0ull;
            if (key[r] < floor_key || key[r] >= ceil_key) key[r] = 0ull;
            if (key[r]) atomicMax(a.gslot + (size_t)c.qid[ql] * a.nslot + key_id(key[r]) % (uint32_t)a.nslot, (uint32_t)(key[r] >> 32));
        }
        kth = list_merge4(a, c, ql, key);
    }
    // A full list's minimum is this CTA's exact K-th best.  Global feedback without locks: the
    // residue slots hold the best scores of nslot DISJOINT sets of songs, found by any CTA, so
    // the K-th largest slot value is a lower bound of the K-th best over everything scored so
    // far by anyone -- with nslot >= 2.5 K it is close to that K-th best itself, and far
    // tighter than this CTA's own view when 37 CTAs share a query.  The better of the two bounds
    // is published.  (K-th largest by binary search over the 32 bits of the orderable scores.)
    auto publish = [&](uint64_t kth_key) {
        const uint32_t *slots = a.gslot + (size_t)c.qid[ql] * a.nslot;
        uint32_t sv[8];
        const int per_lane = a.nslot / 32;  // <= 8 on the register path
        uint32_t slot_kth = 0;
        if (per_lane <= 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) sv[j] = (j < per_lane) ? __ldcg(slots + j * 32 + lane) : 0u;
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t cand = slot_kth | (1u << bit);
                int cgt = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) cgt += (sv[j] >= cand);
                if ((int)__reduce_add_sync(0xffffffffu, (unsigned)cgt) >= a.K) slot_kth = cand;
            }
        } else {
            for (int bit = 31; bit >= 0; --bit) {
                const uint32_t cand = slot_kth | (1u << bit);
                int cgt = 0;
                for (int i = lane; i < a.nslot; i += 32) cgt += (__ldcg(slots + i) >= cand);
                if ((int)__reduce_add_sync(0xffffffffu, (unsigned)cgt) >= a.K) slot_kth = cand;
            }
        }
        if (lane == 0) {
            uint32_t b = __ldcg(a.g_best + c.qid[ql]);
            const uint32_t mine = max((uint32_t)(kth_key >> 32), slot_kth);
            if (mine > b) {
                atomicMax(a.g_best + c.qid[ql], mine);
                b = mine;
            }
            if (b > c.best[ql]) {
                c.best[ql] = b;
                c.nthr[ql] = neg_threshold(b);
            }
        }
        __syncwarp();
    };
    publish(kth);
    bool rescanned = false;
    if (overflow) {
        if (lane == 0) c.dirty[ql] = 1;
        if (c.best[ql] > best_before) {
            if (lane == 0) redo[atomicAdd(redo_cnt, 1)] = ql;  // re-filter the tile against the raised threshold
        } else {
            rescanned = true;
            for (int64_t base = tile_lo; base < tile_hi; base += 128) {
                uint64_t key[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int64_t row = base + r * 32 + lane;
                    key[r] = (row < tile_hi) ? exact_key(a, row, q, qn, ex) : 0ull;
                    if (key[r] < floor_key || key[r] >= ceil_key) key[r] = 0ull;
                    if (key[r]) atomicMax(a.gslot + (size_t)c.qid[ql] * a.nslot + key_id(key[r]) % (uint32_t)a.nslot, (uint32_t)(key[r] >> 32));
                }
                kth = list_merge4(a, c, ql, key);
            }
            publish(kth);
        }
    }
    if (lane == 0) {
        c.cnt[ql] = 0;
        if (a.stats) {
            atomicAdd(a.stats + 1, 1ull);
            if (rescanned) atomicAdd(a.stats + 2, 1ull);
            if (overflow && !rescanned) atomicAdd(a.stats + 4, 1ull);
            atomicAdd(a.stats + 3, (unsigned long long)(cnt + (rescanned ? (tile_hi - tile_lo) : 0)));
        }
    }
    __syncwarp();
}

// Layout of the normalised store (S songs per thread, kLT = 256 "layout threads"): a layout
// tile holds S*kLT consecutive songs; layout thread t owns songs t + s*kLT and multiplies them
// in pairs (2p, 2p+1).  A CTA of THREADS = m*kLT threads works on m consecutive layout tiles
// (its tile), so every kernel shape with the same S reads the same store.  The two songs of a pair are stored interleaved,
// [a0 b0 a1 b1 ... a11 b11], so one 128-bit load yields two ready FFMA2 operands and a
// warp's loads cover a contiguous 3 KB.  Float offset of feature j of local row `row`:
constexpr int kLT = 256;
__host__ __device__ __forceinline__ int64_t hat_offset(int64_t row, int j, int S, int THREADS = kLT)
{
    const int64_t TS = (int64_t)S * THREADS;
    const int64_t tile = row / TS;
    const int r = (int)(row - tile * TS);
    const int s = r / THREADS, t = r - s * THREADS;
    return ((tile * (S / 2) + (s >> 1)) * THREADS + t) * 24 + 2 * j + (s & 1);
}

// One filter pass of a thread's S songs against query record `ql` of the tile.
// Development instrumentation (-DSR_SCAN_TIMING): thread 0 of every CTA accumulates the clock
// cycles it spends in the hot loop / in settle phases / in the kernel into stats[5..7].
#ifdef SR_SCAN_TIMING
#define SR_TIME_BEGIN(slot) do { if (threadIdx.x == 0) s_time[3] = (int)clock(); } while (0)
#define SR_TIME_END(slot) do { if (threadIdx.x == 0) s_time[slot] += (int)clock() - s_time[3]; } while (0)
#else
#define SR_TIME_BEGIN(slot) do { } while (0)
#define SR_TIME_END(slot) do { } while (0)
#endif

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

// STAGE: song tiles are brought into shared memory by the bulk-copy (TMA) engine one tile
// ahead -- issued as soon as the previous tile has been copied into registers -- so the HBM
// stream never waits for the arithmetic (the shape for small, HBM-bound batches; the large-batch
// shape spends its shared memory on 256 queries' lists and hit buffers instead and hides the
// 3 % its tile loads cost behind nothing).
// DYN: instead of a fixed run of units, a CTA serves ONE query tile and claims its song tiles
// one at a time from a per-query-tile counter (the claim for the tile after next is issued
// before the current tile's arithmetic, so its latency is hidden).  CTAs that meet clusters,
// ties or many settles simply claim fewer tiles: measured SM idle time at the end of a launch
// drops from 9 % to ~1 % (large batches) and from 23 % to ~3 % (16 queries, STAGE + DYN).
// STAGE + DYN: the tile the bulk copy fetches next is the one claimed a tile ago (s_next), and a
// CTA that joins another query tile starts that tile's first copy itself.
// The shapes the engine selects (sr_engine.cu, kVariants; template <S, THREADS, CTAs/SM, DEFER, staging buffers, DYN>).
template <int S, int THREADS, int MINB, bool DEFER, int NB, bool DYN>
__global__ void __launch_bounds__(THREADS, MINB) scan_kernel(const ScanArgs a)
{
    constexpr bool STAGE = NB > 0;  // NB: TMA staging buffers per CTA (0: tiles go straight into registers)
    static_assert(NB <= 1 || DYN, "two staging buffers need the dynamic claim ring");
    constexpr int TS = S * THREADS;
    constexpr int WARPS = THREADS / 32;
    constexpr int SUB = THREADS / kLT;  // layout tiles per CTA tile
    static_assert(THREADS % kLT == 0, "CTA size must be a multiple of the layout tile's thread count");
    static_assert(S % 2 == 0, "songs per thread must be even");
    static_assert(kQTMax <= 32 * WARPS, "one lane per owned query in the tile epilogue");
    static_assert(kRowPad % TS == 0, "store padding must cover whole tiles");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t kTileBytes = (uint32_t)TS * kF * 4;
    // TWO staging buffers (one CTA per SM then: 2 x 96 KB): the copy of tile t + 2 starts the moment tile t has left its
    // buffer, so a copy is in flight at every moment of the CTA's life and the arithmetic never waits for one that was
    // issued too late (two single-buffered CTAs per SM fall into step with each other: both wait, then both compute).
    // (Compiled with MINB = 2 although only one such CTA fits an SM: without the 128-register cap ptxas keeps the query
    // operands of the hot loop in ordinary registers -- 0 of 768 FFMA2 with a uniform-register operand at 245 registers.)
    constexpr int NBUF = NB > 1 ? 2 : 1;
    constexpr int NS = NBUF + 2;  // DYN: ring of claimed tiles -- the current one, the next NBUF (already known), the one being claimed
    float4 *s_tile = reinterpret_cast<float4 *>(smem_raw);                                 // STAGE only: [NBUF] tiles
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(smem_raw + NBUF * kTileBytes);           // STAGE only: [NBUF]
    QueryCtx c;
    c.list = reinterpret_cast<uint64_t *>(smem_raw + (STAGE ? NBUF * kTileBytes + 16 : 0));
    c.qraw = reinterpret_cast<float *>(c.list + (a.list_ws ? 0 : (size_t)a.qt * a.K));
    if (a.list_ws) c.list = a.list_ws + (size_t)blockIdx.x * a.qt * a.K;
    c.nthr = c.qraw + a.qt * kF;
    c.qn = c.nthr + a.qt;
    c.best = reinterpret_cast<uint32_t *>(c.qn + a.qt);
    c.cnt = reinterpret_cast<int *>(c.best + a.qt);
    c.lcnt = c.cnt + a.qt;
    c.excl = c.lcnt + a.qt;
    c.qid = c.excl + a.qt;
    c.dirty = c.qid + a.qt;
    c.hit = reinterpret_cast<uint32_t *>(c.dirty + a.qt);
    // s_flag[tile % 3] != 0: some hit buffer filled up during that tile (set by the appending
    // thread).  Three slots make the protocol race-free with ONE barrier per tile: slot t%3 is
    // read right after tile t's barrier, cleared after tile t+1's barrier (every read is done),
    // and next written during tile t+3, which starts after tile t+2's barrier.
    int *s_flag = reinterpret_cast<int *>(c.hit + (size_t)a.qt * a.cap);
    int *s_redo_cnt = s_flag + 4;                 // [2]
    int *s_redo = s_redo_cnt + 4;                 // [2][qt] queries whose tile must be re-filtered
    int *s_next = s_redo + 2 * a.qt;              // DYN: [0..NS-1] claimed song tiles (a ring: this one and the next NS - 2), [NS] next query tile
    // DYN: the per-tile barrier is split (arrive ... wait) so the next tile's loads overlap the wait
    __shared__ uint64_t s_tbar;
    uint32_t tbar_phase = 0;

    if (threadIdx.x < 4) s_flag[threadIdx.x] = 0;
    if (threadIdx.x < 2) s_redo_cnt[threadIdx.x] = 0;
#ifdef SR_SCAN_TIMING
    __shared__ int s_time[8];  // hot loop, settle phases, barrier wait, scratch, final settle, flush, join, prologue
    __shared__ long long s_t0;
    if (threadIdx.x < 8) s_time[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_t0 = clock64();
#endif
    int tphase = 0;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // this CTA's contiguous run of (query tile, song tile) units; kept in uniform
    // arithmetic (no division) so the constant-bank query index stays warp-uniform and
    // the FFMA2 query operand can live in a uniform register
    int u = (int)blockIdx.x * a.upc + min((int)blockIdx.x, a.extra);
    int u_end = u + a.upc + ((int)blockIdx.x < a.extra ? 1 : 0);
    int qtile = 0, t0 = u;
    const int nqt_d = (a.nq + a.qt - 1) / a.qt;
    if (DYN) {
        // first segment: this CTA's home query tile (uniform arithmetic again)
        int bb = (int)blockIdx.x;
        if (bb >= a.cpq * nqt_d) qtile = bb - a.cpq * nqt_d;
        else while (bb >= a.cpq) { bb -= a.cpq; ++qtile; }
        u = 0; u_end = 1; t0 = 0;
        if (tid == 0) {
            for (int i = 0; i < NS - 1; ++i) s_next[i] = atomicAdd(a.tile_ctr + qtile, 1);
            mbar_init(&s_tbar, WARPS);
        }
    } else {
        while (t0 >= a.n_tiles) { t0 -= a.n_tiles; ++qtile; }
    }
    uint32_t sphase = 0;
    if (STAGE) {
        if (tid == 0) {
            // (DYN: the first claimed tile(s); a claim at or beyond n_tiles means the query tile is already exhausted)
            for (int b = 0; b < NBUF; ++b) {
                mbar_init(s_bar + b, 1);
                const int first = DYN ? s_next[b] : t0;
                if (u < u_end && first < a.n_tiles)
                    tma_load_tile(s_tile + (size_t)b * (kTileBytes / 16), a.hat + (int64_t)first * a.tile_stride * (TS * kF), kTileBytes, s_bar + b);
            }
        }
        __syncthreads();
    }

    while (u < u_end) {
        const int t1 = DYN ? a.n_tiles : min(a.n_tiles, t0 + (u_end - u));
        const int q0 = qtile * a.qt;
        const int nql = min(a.qt, a.nq - q0);

        // ---- segment prologue: bring the query tile's state into shared memory
        SR_TIME_BEGIN(7);
        for (int ql = tid; ql < nql; ql += THREADS) {
            const int qid = q0 + ql;
            c.qid[ql] = qid;
            c.excl[ql] = a.exclude[qid];
            c.cnt[ql] = 0;
            c.lcnt[ql] = 0;
            c.dirty[ql] = 0;
            c.qn[ql] = a.qn[qid];
            const uint32_t b = __ldcg(a.g_best + qid);
            c.best[ql] = b;
            c.nthr[ql] = neg_threshold(b);
        }
        for (int i = tid; i < nql * kF; i += THREADS) c.qraw[i] = a.qraw[(size_t)q0 * kF + i];
        __syncthreads();

        SR_TIME_END(7);
        int it = 0, slot = 0;
        // this thread's S songs of store tile `t` -> registers: S/2 interleaved pairs, six 128-bit
        // loads each, every load two ready FFMA2 operands
        float2 fp[S / 2][kF];
        auto load_songs = [&](int t) {
            const int64_t lt = (int64_t)t * a.tile_stride * SUB + tid / kLT;
            const float4 *src = reinterpret_cast<const float4 *>(a.hat) + (lt * (S / 2) * kLT + tid % kLT) * 6;
#pragma unroll
            for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                for (int c4 = 0; c4 < 6; ++c4) {
                    const float4 v = __ldg(src + (int64_t)p * kLT * 6 + c4);
                    fp[p][2 * c4] = make_float2(v.x, v.y);
                    fp[p][2 * c4 + 1] = make_float2(v.z, v.w);
                }
            }
        };
        // (redux.sync hands the claimed index back as a warp-uniform value, which keeps the tile loop --
        // and with it the hot loop's uniform-register operands -- in uniform control flow)
        // DYN claims two tiles ahead through three slots: the slot written during tile t was last
        // read during tile t-2, and the NEXT tile's index is already visible while tile t runs, so
        // its songs are loaded between arriving at tile t's barrier and waiting on it.
        int tile = DYN ? __reduce_max_sync(0xffffffffu, s_next[0]) : t0;
        if (DYN && !STAGE) load_songs(min(tile, a.n_tiles - 1));
        for (; tile < t1; ++it, slot = (slot == NS - 1 ? 0 : slot + 1)) {
            // (the claim for the tile after next: issued now, stored after the hot loop, so thread 0
            // does not start every tile a global round trip late)
            int claimed = 0;
            if (DYN && tid == 0) claimed = atomicAdd(a.tile_ctr + qtile, 1);
            if (DYN && !STAGE && a.prefetch && lane == 0) {
                // the next tile of this CTA (claimed one tile ago) starts its way from HBM into L2 now, so the loads
                // issued between this tile's arrive and wait find it there: with few queries per tile (mid-size
                // batches) the tile load is otherwise 15-25 % of the tile's time.  One slice per warp.
                const int nt = s_next[slot == NS - 1 ? 0 : slot + 1];
                if (nt < a.n_tiles) {
                    const char *src = reinterpret_cast<const char *>(a.hat) + ((int64_t)nt * a.tile_stride * TS * kF * 4) + (size_t)warp * (kTileBytes / WARPS);
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(kTileBytes / WARPS) : "memory");
                }
            }
            const int64_t stile = (int64_t)tile * a.tile_stride;  // store tile
            const int64_t ltile = stile * SUB + tid / kLT;      // this thread's layout tile
            const int row0 = (int)(ltile * (S * kLT)) + tid % kLT;  // its songs: row0 + s * kLT (ids are 32-bit)

            if (STAGE) {
                const int buf = NBUF == 2 ? (it & 1) : 0;
                mbar_wait(s_bar + buf, (sphase >> buf) & 1u);  // this tile has landed in shared memory
                sphase ^= 1u << buf;
                const float4 *src = s_tile + (size_t)buf * (kTileBytes / 16) + ((tid / kLT) * (S / 2) * kLT + tid % kLT) * 6;
#pragma unroll
                for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                    for (int c4 = 0; c4 < 6; ++c4) {
                        const float4 v = src[p * kLT * 6 + c4];
                        fp[p][2 * c4] = make_float2(v.x, v.y);
                        fp[p][2 * c4 + 1] = make_float2(v.z, v.w);
                    }
                }
                // (a full barrier: an arrive-only barrier with thread 0 alone waiting was tried -- the spin loop inside the
                // divergent branch makes ptxas drop the hot loop's uniform-register operands)
                __syncthreads();  // every thread has copied its songs out: the buffer is free
                if (tid == 0) {   // next tile of this run: same query tile, or tile 0 of the next one
                    int64_t nxt = -1;
                    if (DYN) {    // (the tile NBUF ahead, claimed a tile ago; at or beyond n_tiles: this query tile is exhausted)
                        const int nt = s_next[(slot + NBUF) % NS];
                        if (nt < a.n_tiles) nxt = (int64_t)nt * a.tile_stride;
                    } else if (tile + 1 < t1) nxt = stile + a.tile_stride;
                    else if (u + (t1 - t0) < u_end) nxt = 0;
                    if (nxt >= 0) tma_load_tile(s_tile + (size_t)buf * (kTileBytes / 16), a.hat + nxt * (TS * kF), kTileBytes, s_bar + buf);
                }
            } else if (!DYN) {
                load_songs(tile);
            }

            // a settle phase follows this tile if some hit buffer fills up (flagged by the thread
            // whose append crosses the mark), and always after the last tile of the segment
            const bool forced = !DYN && (tile == t1 - 1);

            auto append = [&](int ql, const float2 (&acc)[S / 2]) {
#pragma unroll
                for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float v = h ? acc[p].y : acc[p].x;
                        if ((int)__float_as_uint(v) >= 0) {
                            const int row = row0 + (2 * p + h) * kLT;
                            if ((int64_t)row < a.n) {
                                const int slot = atomicAdd(&c.cnt[ql], 1);
                                if (slot < a.cap) c.hit[(size_t)ql * a.cap + slot] = (uint32_t)(a.id_base + row);
                                if (slot + 1 >= a.trigger_at) s_flag[tphase] = 1;
                                atomicAdd(&s_flag[3], 1);  // statistics only
                            }
                        }
                    }
                }
            };

            // ---- the hot loop: one query per iteration, branch-free.  Iteration i only
            // records whether any of the thread's S songs passed (one bit); the rare hits
            // are picked up after every 32 queries by re-running the filter for the flagged
            // queries, so the FFMA2 stream of consecutive queries is never split by a branch
            // and every operand that depends on the query stays in uniform registers.
            SR_TIME_BEGIN(0);
            for (int qb = 0; qb < nql; qb += 32) {
                const int qe = min(32, nql - qb);
                uint32_t mask = 0;
                if (DEFER) {
                    // one funnel shift per query collects the sign of the AND: after qe queries bit
                    // (qe - 1 - i) of `signs` is SET when none of the thread's songs passed query qb + i
                    uint32_t signs = 0;
#pragma unroll 16
                    for (int i = 0; i < qe; ++i) {
                        float2 acc[S / 2];
                        const uint32_t m = filter_query<S>(fp, q0 + qb + i, c.nthr[qb + i], acc);
                        signs = __funnelshift_l(m, signs, 1);
                    }
                    mask = __brev(~signs) >> (32 - qe);  // bit i set <=> a song passed query qb + i
                } else {
#pragma unroll 2
                    for (int i = 0; i < qe; ++i) {
                        float2 acc[S / 2];
                        const uint32_t m = filter_query<S>(fp, q0 + qb + i, c.nthr[qb + i], acc);
                        mask |= ((~m) >> 31) << i;
                    }
                }
                while (mask) {  // rare: this thread has a passing song for query qb + b
                    const int b = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const int ql = qb + b;
                    float2 acc[S / 2];
                    filter_query<S>(fp, q0 + ql, c.nthr[ql], acc);
                    append(ql, acc);
                }
            }
            SR_TIME_END(0);
            // ---- tile epilogue: warp w looks after queries ql == w (mod WARPS), one lane each:
            // adopt thresholds published by other CTAs and settle hit buffers that filled up -- every non-empty
            // one after a segment's first and last tile.  One barrier when there is nothing to
            // settle (threshold updates racing with other warps' reads are benign: any published
            // threshold is a valid lower bound), two when there is.
            auto refresh = [&]() {
                const int ql_mine = warp + WARPS * lane;
                if (ql_mine < nql && it % a.refresh_every == 0) {
                    const uint32_t g = __ldcg(a.g_best + c.qid[ql_mine]);
                    if (g > c.best[ql_mine]) {
                        c.best[ql_mine] = g;
                        c.nthr[ql_mine] = neg_threshold(g);
                    }
                }
            };
            int next_tile = tile + 1;
            if (DYN) {
                if (tid == 0) s_next[slot == 0 ? NS - 1 : slot - 1] = claimed;  // (claimed at the top of the tile)
                next_tile = __reduce_max_sync(0xffffffffu, s_next[slot == NS - 1 ? 0 : slot + 1]);
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_tbar);
                // in flight while the slower warps finish: the next tile's songs (unless the bulk-copy engine is
                // already bringing them into shared memory), the published thresholds
                if (!STAGE) load_songs(min(next_tile, a.n_tiles - 1));
                refresh();
                SR_TIME_BEGIN(2);
                mbar_wait(&s_tbar, tbar_phase);
                SR_TIME_END(2);
                tbar_phase ^= 1u;
            } else {
                refresh();
                __syncthreads();  // all hits of this tile are in the buffers
            }
            const bool settle_phase = forced || (s_flag[tphase] != 0);
            if (tid == 0) s_flag[tphase == 0 ? 2 : tphase - 1] = 0;  // the previous tile's slot
            tphase = (tphase == 2) ? 0 : tphase + 1;
            if (settle_phase) {
                SR_TIME_BEGIN(1);
                bool need = false;
                {
                    const int ql_mine = warp + WARPS * lane;
                    if (ql_mine < nql) {
                        const int cn = c.cnt[ql_mine];
                        need = cn >= a.settle_at || (cn > 0 && forced);
                    }
                }
                uint32_t todo = __ballot_sync(0xffffffffu, need);
                const int64_t tile_lo = stile * TS;
                const int64_t tile_hi = min(a.n, tile_lo + TS);
                while (todo) {
                    const int l = __ffs(todo) - 1;
                    todo &= todo - 1;
                    warp_settle(a, c, warp + WARPS * l, tile_lo, tile_hi, s_redo, s_redo_cnt);
                }
                __syncthreads();
                // re-filter rounds for queries whose buffer overflowed: the tile is still in
                // registers, one filter iteration per flagged query against the raised threshold
                for (int rp = 0;; rp ^= 1) {
                    const int nr = s_redo_cnt[rp];
                    if (nr == 0) break;
                    const int *list = s_redo + rp * a.qt;
                    {
                        // the tile is re-read (L2-hot) rather than kept alive in registers across
                        // the settle code, which would spill
                        float2 fr[S / 2][kF];
                        const float4 *src = reinterpret_cast<const float4 *>(a.hat) + (ltile * (S / 2) * kLT + tid % kLT) * 6;
#pragma unroll
                        for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                            for (int c4 = 0; c4 < 6; ++c4) {
                                const float4 v = __ldcg(src + (int64_t)p * kLT * 6 + c4);
                                fr[p][2 * c4] = make_float2(v.x, v.y);
                                fr[p][2 * c4 + 1] = make_float2(v.z, v.w);
                            }
                        }
                        for (int r = 0; r < nr; ++r) {
                            const int ql = list[r];
                            float2 acc[S / 2];
                            const uint32_t m = filter_query<S>(fr, q0 + ql, c.nthr[ql], acc);
                            if ((int)m >= 0) append(ql, acc);
                        }
                    }
                    __syncthreads();
                    if (tid == 0) s_redo_cnt[rp] = 0;
                    for (int r = 0; r < nr; ++r) {
                        const int ql = list[r];
                        if (ql % WARPS == warp) warp_settle(a, c, ql, tile_lo, tile_hi, s_redo + (rp ^ 1) * a.qt, s_redo_cnt + (rp ^ 1));
                    }
                    __syncthreads();
                }
                // (the prefetched songs were not kept alive across the settle code: fetch them again, L2-hot)
                if (DYN && !STAGE) load_songs(min(next_tile, a.n_tiles - 1));
                SR_TIME_END(1);
            }
            tile = next_tile;
        }
        SR_TIME_BEGIN(4);
        if (DYN) {
            // The last tile is not known in advance, so hits are still pending now (fewer than trigger_at
            // per query).  Settling them one query at a time is a chain of dependent global round trips
            // per query (rows, slots, thresholds) that nothing overlaps at the end of a segment.  Instead
            // every warp walks the pending hits of ALL its queries 32 at a time -- one round of
            // independent row loads -- and appends the exact keys that can still matter straight to the
            // pool; the list is bypassed (finalize sorts the pool anyway).  Queries whose tile was
            // re-filtered in this segment may hold hits that repeat list entries: they take the
            // duplicate-checking settle.
            const int ql_mine = warp + WARPS * lane;
            const bool mine = ql_mine < nql;
            const bool slow = mine && c.dirty[ql_mine] && c.cnt[ql_mine] > 0;
            uint32_t todo = __ballot_sync(0xffffffffu, slow);
            while (todo) {
                const int l = __ffs(todo) - 1;
                todo &= todo - 1;
                warp_settle(a, c, warp + WARPS * l, 0, 0, s_redo, s_redo_cnt);  // cnt <= cap here: no overflow
            }
            const int my_cnt = (mine && !c.dirty[ql_mine]) ? min(c.cnt[ql_mine], a.cap) : 0;
            int offs = my_cnt;  // inclusive, then exclusive, prefix sum over the warp's queries
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int o = __shfl_up_sync(0xffffffffu, offs, d);
                if (lane >= d) offs += o;
            }
            const int total = __shfl_sync(0xffffffffu, offs, 31);
            offs -= my_cnt;
            for (int base = 0; base < total; base += 32) {
                const int idx = base + lane;
                int L = 0;  // owner = the last lane whose range starts at or before idx
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int o = __shfl_sync(0xffffffffu, offs, L + step);
                    if (o <= idx) L += step;
                }
                const int o_own = __shfl_sync(0xffffffffu, offs, L);
                if (idx < total) {
                    const int ql = warp + WARPS * L;
                    float q[kF];
#pragma unroll
                    for (int j = 0; j < kF; ++j) q[j] = c.qraw[ql * kF + j];
                    const uint64_t key = exact_key(a, (int64_t)c.hit[(size_t)ql * a.cap + (idx - o_own)] - a.id_base, q, c.qn[ql], c.excl[ql]);
                    const uint64_t floor_key = (uint64_t)max(c.best[ql], __ldcg(a.g_best + c.qid[ql])) << 32;
                    const uint64_t ceil_key = a.ceil ? __ldg(a.ceil + c.qid[ql]) : ~0ull;
                    if (key != 0ull && key >= floor_key && key < ceil_key) {
                        const int at = atomicAdd(a.pool_cnt + q0 + ql, 1);  // (finalize reports a count beyond the slab)
                        if (at < a.slab) a.pool[(size_t)(q0 + ql) * a.slab + at] = key;
                    }
                }
            }
            if (my_cnt) c.cnt[ql_mine] = 0;
            if (a.stats && lane == 0 && total) atomicAdd(a.stats + 3, (unsigned long long)total);
            __syncthreads();
        }
        SR_TIME_END(4);
        SR_TIME_BEGIN(5);
        // ---- segment epilogue: hand this CTA's exact survivors to the per-query pool
        // (the last tile's settle phase and its closing barrier have just run)
        // (only keys that can still make the final top-K: score >= the best known bound)
        for (int ql = warp; ql < nql; ql += WARPS) {
            const int n = c.lcnt[ql];
            if (n == 0) continue;
            const uint64_t floor_key = (uint64_t)max(c.best[ql], __ldcg(a.g_best + c.qid[ql])) << 32;
            uint64_t *slab = a.pool + (size_t)(q0 + ql) * a.slab;
            const uint64_t *list = c.list + (size_t)ql * a.K;
            for (int b0 = 0; b0 < n; b0 += 32) {
                const int i = b0 + lane;
                const uint64_t k = (i < n) ? list[i] : 0ull;
                const bool keep = k != 0ull && k >= floor_key;
                const uint32_t m = __ballot_sync(0xffffffffu, keep);
                if (!m) continue;
                int base = 0;
                if (lane == 0) base = atomicAdd(a.pool_cnt + q0 + ql, __popc(m));
                base = __shfl_sync(0xffffffffu, base, 0);
                const int at = base + __popc(m & ((1u << lane) - 1u));
                if (keep && at < a.slab) slab[at] = k;
            }
        }
        __syncthreads();
        SR_TIME_END(5);
        SR_TIME_BEGIN(6);
        if (DYN) {
            // this query tile has no unclaimed song tiles left: join another one that has (CTAs
            // per query tile rarely divide the grid evenly, and settle work differs between tiles)
            if (tid == 0) {
                int cand = -1, c2 = qtile;
                for (int k = 1; k < nqt_d && cand < 0; ++k) {
                    c2 = (c2 + 1 == nqt_d) ? 0 : c2 + 1;
                    if (*(volatile int *)(a.tile_ctr + c2) < a.n_tiles && atomicAdd(a.visit_ctr + c2, 1) < a.steal_max) cand = c2;
                }
                s_next[NS] = cand;
                if (cand >= 0) {
                    for (int i = 0; i < NS - 1; ++i) s_next[i] = atomicAdd(a.tile_ctr + cand, 1);
                    if (STAGE) {
                        for (int b = 0; b < NBUF; ++b)
                            if (s_next[b] < a.n_tiles)
                                tma_load_tile(s_tile + (size_t)b * (kTileBytes / 16), a.hat + (int64_t)s_next[b] * a.tile_stride * (TS * kF), kTileBytes, s_bar + b);
                    }
                }
            }
            __syncthreads();
            const int cand = __reduce_max_sync(0xffffffffu, s_next[NS]);
            SR_TIME_END(6);
            if (cand < 0) break;
            qtile = cand;
        } else {
            u += (t1 - t0);
            t0 = 0;
            ++qtile;
        }
    }
    __syncthreads();
    if (a.stats && tid == 0 && s_flag[3]) atomicAdd(a.stats + 0, (unsigned long long)s_flag[3]);
#ifdef SR_SCAN_TIMING
    if (a.stats && tid == 0) {
        atomicAdd(a.stats + 5, (unsigned long long)(unsigned)s_time[0]);
        atomicAdd(a.stats + 6, (unsigned long long)(unsigned)s_time[1]);
        atomicAdd(a.stats + 7, (unsigned long long)(clock64() - s_t0));
        atomicAdd(a.stats + 8, (unsigned long long)(unsigned)s_time[2]);
        for (int i = 4; i < 8; ++i) atomicAdd(a.stats + 5 + i, (unsigned long long)(unsigned)s_time[i]);
    }
#endif
}

// ---- bound pass ---------------------------------------------------------------------------
// Threshold bootstrap at filter speed.  `n_sample` evenly spaced FULL tiles are scored for
// every query of the group (12 FFMA per pair, same loop as the scan, a max instead of a sign
// test).  Their songs are split into `nblk` >= K+1 disjoint blocks by the residue of the owning
// layout thread, so every block spans all sample tiles (and with them every cluster of the store
// that is at least store/n_sample long -- e.g. every genre of a genre-sorted store).  Each block's
// best filter score belongs to a distinct song, at most one of them the query song itself, so the
// (K+1)-th LARGEST of the block maxima is a lower bound of the K-th best filter score among real
// candidates: no sorting of songs, ~1-2.5 % of a full pass.  With nblk well above K+1 two of the
// sample's best songs rarely share a block and the bound approaches the sample's exact K-th best
// (with nblk = K+1 it is the minimum of the maxima, about H(K+1) = 3-5 times further down the ranks).
// Unit u = (query tile, sample tile); each CTA takes whole units; the CTA that completes a query
// tile's last unit folds the maxima into the starting thresholds g_best.
template <int S, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) bound_kernel(const ScanArgs a, int nblk, int n_sample, int stride,
                                                              uint32_t *gmax, int *done_ctr)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_max = reinterpret_cast<uint32_t *>(smem_raw);  // [qt][nblk] orderable block maxima of this CTA's run
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int blk = (tid % kLT) % nblk;
    const int nqt = (a.nq + a.qt - 1) / a.qt;
    // a contiguous run of units (query-tile major): the maxima stay in shared memory across the sample tiles of
    // one query tile and reach the global array once per query tile of the run
    const int total = nqt * n_sample, per = total / (int)gridDim.x, extra = total % (int)gridDim.x;
    int u = (int)blockIdx.x * per + min((int)blockIdx.x, extra);
    const int u_end = u + per + ((int)blockIdx.x < extra ? 1 : 0);
    int qtile = 0, j = u;
    while (j >= n_sample) { j -= n_sample; ++qtile; }
    while (u < u_end) {
        const int q0 = qtile * a.qt;
        const int nql = min(a.qt, a.nq - q0);
        const int j_end = min(n_sample, j + (u_end - u));
        for (int i = tid; i < nql * nblk; i += THREADS) s_max[i] = 0u;
        __syncthreads();
        for (int jj = j; jj < j_end; ++jj) {
            const int64_t ltile = (int64_t)jj * stride * (THREADS / kLT) + tid / kLT;
            float2 fp[S / 2][kF];
            {
                const float4 *src = reinterpret_cast<const float4 *>(a.hat) + (ltile * (S / 2) * kLT + tid % kLT) * 6;
#pragma unroll
                for (int p = 0; p < S / 2; ++p) {
#pragma unroll
                    for (int c4 = 0; c4 < 6; ++c4) {
                        const float4 v = __ldg(src + (int64_t)p * kLT * 6 + c4);
                        fp[p][2 * c4] = make_float2(v.x, v.y);
                        fp[p][2 * c4 + 1] = make_float2(v.z, v.w);
                    }
                }
            }
#pragma unroll 2
            for (int ql = 0; ql < nql; ++ql) {
                float2 acc[S / 2];
                filter_query<S>(fp, q0 + ql, 0.0f, acc);
                float m = -__int_as_float(0x7f800000);
#pragma unroll
                for (int p = 0; p < S / 2; ++p) m = fmaxf(m, fmaxf(acc[p].x, acc[p].y));  // NaN (irregular) rows are ignored
                const uint32_t o = f2ord(m);
                if (o > s_max[ql * nblk + blk]) atomicMax(&s_max[ql * nblk + blk], o);
            }
        }
        __syncthreads();
        for (int i = tid; i < nql * nblk; i += THREADS)  // (a read first: most maxima of a later run already stand)
            if (s_max[i] > __ldcg(gmax + (size_t)q0 * nblk + i)) atomicMax(gmax + (size_t)q0 * nblk + i, s_max[i]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = a.bound_finish && (atomicAdd(done_ctr + qtile, j_end - j) + (j_end - j) == n_sample);
        __syncthreads();
        if (s_last) {
            // every sample tile of this query tile is in: (K+1)-th largest block maximum per query, one warp
            // per query (binary search over the 32 bits of the orderable values, nblk / 32 values per lane)
            __threadfence();
            const int lane = tid & 31, warp = tid >> 5;
            for (int ql = warp; ql < nql; ql += THREADS / 32) {
                const uint32_t *row = gmax + (size_t)(q0 + ql) * nblk;
                uint32_t v[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) v[r] = (r * 32 + lane < nblk) ? __ldcg(row + r * 32 + lane) : 0u;
                uint32_t kth = 0;
                for (int bit = 31; bit >= 0; --bit) {
                    const uint32_t cand = kth | (1u << bit);
                    int c = 0;
#pragma unroll
                    for (int r = 0; r < 8; ++r) c += (v[r] >= cand);
                    if ((int)__reduce_add_sync(0xffffffffu, (unsigned)c) >= a.K + 1) kth = cand;
                }
                if (lane == 0 && kth != 0u) {  // (0: fewer than K+1 blocks saw a regular song -- no bound)
                    const uint32_t o = f2ord(ord2f(kth) - kBoundSlack);
                    if (o > a.g_best[q0 + ql]) a.g_best[q0 + ql] = o;
                }
            }
        }
        __syncthreads();
        u += j_end - j;
        j = 0;
        ++qtile;
    }
}

}  // namespace sr
