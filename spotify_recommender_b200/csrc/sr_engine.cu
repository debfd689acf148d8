// sr_engine.cu -- host side of the C ABI declared in include/sr_engine.h: device
// store management, batch workspace, kernel launches.  sm_100a only; there is no
// CPU fallback anywhere in this file (north_star (1)).
#include "../../include/sr_engine.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "sr_aux.cuh"
#include "sr_scan.cuh"

namespace {

using namespace sr;

thread_local std::string g_create_error;

// ---- scan kernel shapes --------------------------------------------------------
typedef cudaError_t (*ScanLaunch)(const ScanArgs &, int grid, size_t smem, cudaStream_t st);
typedef cudaError_t (*ScanOcc)(int *ctas_per_sm, size_t smem);
typedef cudaError_t (*BoundLaunch)(const ScanArgs &, int nblk, int n_sample, int stride, uint32_t *gmax, int *done_ctr, int grid,
                                   size_t smem, cudaStream_t st);

template <int S, int T, int M, bool D, int G, bool Y>
cudaError_t launch_scan(const ScanArgs &a, int grid, size_t smem, cudaStream_t st)
{
    auto k = scan_kernel<S, T, M, D, G, Y>;
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k<<<grid, T, smem, st>>>(a);
    return cudaGetLastError();
}
template <int S, int T, int M, bool D, int G, bool Y>
cudaError_t occ_scan(int *ctas, size_t smem)
{
    auto k = scan_kernel<S, T, M, D, G, Y>;
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas, k, T, smem);
}

template <int S, int T, int M>
cudaError_t launch_bound(const ScanArgs &a, int nblk, int n_sample, int stride, uint32_t *gmax, int *done_ctr, int grid, size_t smem,
                         cudaStream_t st)
{
    auto k = bound_kernel<S, T, M>;
    cudaError_t err = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    k<<<grid, T, smem, st>>>(a, nblk, n_sample, stride, gmax, done_ctr);
    return cudaGetLastError();
}

struct Variant {
    const char *name;
    int S, threads, ctas;  // ctas: the kernel's __launch_bounds__ minimum (its register cap)
    int nbuf;              // TMA staging buffers per CTA (0: tiles are loaded straight into registers)
    bool dynamic;
    bool staged() const { return nbuf > 0; }
    // CTAs per SM the shared-memory budget is shared by: one when the staging buffers alone exceed half an SM's shared memory
    int smem_ctas() const { return (size_t)nbuf * S * threads * kF * 4 > (size_t)110 * 1024 ? 1 : ctas; }
    ScanLaunch launch;
    ScanOcc occ;
    BoundLaunch bound;
};
#define SR_VARIANT(S, T, M, D) \
    {"S" #S "xT" #T "x" #M "-" #D, S, T, M, 0, false, launch_scan<S, T, M, D, 0, false>, occ_scan<S, T, M, D, 0, false>, launch_bound<S, T, M>}
#define SR_VARIANT_TMA(S, T, M) \
    {"S" #S "xT" #T "x" #M "-tma", S, T, M, 1, false, launch_scan<S, T, M, true, 1, false>, occ_scan<S, T, M, true, 1, false>, launch_bound<S, T, M>}
#define SR_VARIANT_DYN(S, T, M) \
    {"S" #S "xT" #T "x" #M "-dyn", S, T, M, 0, true, launch_scan<S, T, M, true, 0, true>, occ_scan<S, T, M, true, 0, true>, launch_bound<S, T, M>}
#define SR_VARIANT_DTMA(S, T, M) \
    {"S" #S "xT" #T "x" #M "-dyn-tma", S, T, M, 1, true, launch_scan<S, T, M, true, 1, true>, occ_scan<S, T, M, true, 1, true>, launch_bound<S, T, M>}
#define SR_VARIANT_DTMA2(S, T, M) \
    {"S" #S "xT" #T "-dyn-tma2", S, T, M, 2, true, launch_scan<S, T, M, true, 2, true>, occ_scan<S, T, M, true, 2, true>, launch_bound<S, T, M>}
const Variant kVariants[] = {
    SR_VARIANT(8, 256, 2, false),   // 0  plain hit branch in the loop
    SR_VARIANT(8, 256, 2, true),    // 1  branch-free loop, deferred hits
    SR_VARIANT(8, 512, 1, false),   // 2
    SR_VARIANT(8, 512, 1, true),    // 3  static runs of units (fallback of 5 when query tiles outnumber CTAs)
    SR_VARIANT_DTMA(8, 256, 2),     // 4  small batches (HBM-bound): TMA-staged tiles, dynamic tile claiming
    SR_VARIANT_DYN(8, 512, 1),      // 5  large batches: dynamic tile claiming, tiles loaded straight into registers
    SR_VARIANT_DTMA(8, 512, 1),     // 6  64-query tiles, the next song tile staged by TMA meanwhile (an experiment kept for reference: +2 % at
                                    //    64 queries, slower above -- copying 196 KB out of shared memory costs what the hidden load saves)
    SR_VARIANT_TMA(8, 256, 2),      // 7  the static form of 4 (contiguous runs of units)
    SR_VARIANT_DYN(8, 256, 2),      // 8  mid-size batches, short lists: two 256-thread CTAs per SM loading straight into registers
    SR_VARIANT_DTMA2(8, 256, 2),    // 9  one 256-thread CTA per SM with TWO TMA staging buffers (a copy in flight at every moment)
    // (tried and dropped: SR_VARIANT_DTMA2(4, 256, 2) -- 4 songs per thread, two CTAs per SM, two 48 KB staging buffers each: with
    // half the registers taken by songs ptxas keeps the query operands in ordinary registers (LDC, not LDCU): 1 query 74.8 us
    // (0.994 of the copy bandwidth) but 16 queries 119 us against 99 us.)
    // (tried and dropped: SR_VARIANT_DTMA(4, 256, 3) -- 4 songs per thread, three CTAs per SM, 49 KB TMA stages, on the SAME
    // store (an 8-song layout tile is two contiguous 4-song layout tiles).  At 85 registers ptxas keeps none of the FFMA2 query
    // operands in uniform registers and the shape loses from 8 queries up: 16 queries 118 us against 99 us.)
};
constexpr int kNumVariants = (int)(sizeof(kVariants) / sizeof(kVariants[0]));
constexpr int kAutoSmall = 4, kAutoLarge = 5, kAutoMid = 8, kStaticLarge = 3, kAutoS = 8;

enum KernelId { kPrep = 0, kSample, kScan, kFinalize, kMerge, kBound, kNumKernels };
const char *const kKernelNames[kNumKernels] = {"prep", "sample", "scan", "finalize", "merge", "bound"};

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct sr_engine {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // store
    float *d_raw = nullptr, *d_hat = nullptr, *d_nf = nullptr;
    int64_t n = 0, n_pad = 0;
    int32_t id_base = 0;
    int64_t irregular = 0;
    int hat_S = -1;  // songs per thread d_hat is laid out for
    int last_variant = kAutoLarge;

    // options
    int variant = -1;  // -1: chosen per pass by batch size (same store layout for every S = 8 shape)
    int qt_opt = kQTMax;
    int batch = 8192;
    int sample = -1;  // -1: automatic
    bool bound = true;  // bound pass (filter-speed threshold bootstrap)
    int settle_at = 0;  // 0: cap / 32
    int trigger_at = 0; // 0: cap / 4
    int hit_cap = 0;   // hit-buffer entries per query in shared memory (0: sized from K)
    int list_ws_opt = 1;    // allow the CTAs' lists in an L2-resident workspace when that keeps the query tile at full size
    int list_ws_kmax = 72;  // ... for k up to this
    int small_max = 32;     // batches of at most this many queries take the TMA-staged small-batch shape
    int mid_max = 192;      // ... and up to this many (k <= 16) two dynamic 256-thread CTAs per SM (measured against the
                            // one-CTA shape: +10 % at 40 queries, +7 % at 64, +4 % at 128, -1 % at 256)
    int refresh_every = 0;  // tiles between two looks at the thresholds other CTAs published (0: automatic)
    int bound_blocks = 0;   // disjoint sample blocks of the bound pass (0: 64 / 128 / 256 by k)
    int bound_cap_div = 16; // the bound pass samples at most 1 / this of the store's tiles
    int prefetch = 0;       // 1: L2 bulk prefetch of the next song tile in the dynamic register-loading shape (no measured gain)
    int bound_tiles = 0;    // layout tiles (of S x 256 songs) the bound pass samples (0: 48 for k <= 16, else 128)
    bool profile = false;

    // batch workspace (grow-only)
    DevBuf qraw, qn, qhat, excl, gbest, gbound, gslot, ctr, pool_cnt, pool, out, out2, qin, minmax, list_ws, ceil;
    unsigned long long *d_stats = nullptr;  // [16]
    unsigned long long *d_irregular = nullptr;
    int32_t *d_flag = nullptr;   // bit 0: a query id this store does not own; bit 1: a pool slab overflowed
    int32_t *h_flag = nullptr;   // pinned mirror
    float *d_cbank = nullptr;    // device address of the constant bank c_qhat
    void *h_pin = nullptr;
    size_t h_pin_cap = 0;
    cudaStream_t copy_stream = nullptr;  // all-pairs: result copies overlap the next batch
    cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
    int scan_grid = 0;
    int last_lists_in_smem = 1;

    // Small single-group passes are replayed as CUDA graphs: prep + bound + scan + finalize captured once per
    // (buffers, nq, k, shape) and launched with one call -- the reference's only real use is one query per call
    // (main.cpp:71,82), where the four launches around an 80 us kernel are what a user waits for.
    struct GraphKey {
        const void *p[10];
        int nq, K, stride, col;
        bool operator<(const GraphKey &o) const { return memcmp(this, &o, sizeof(GraphKey)) < 0; }
    };
    struct GraphEntry {
        cudaGraphExec_t exec;
        int launches;
    };
    std::map<GraphKey, GraphEntry> graphs;
    int64_t epoch = 0, graphs_epoch = 0;   // any reallocation, store load or option change invalidates the cache
    int use_graphs = 1;
    int64_t graph_replays = 0;

    // counters
    int64_t launches = 0, queries = 0;
    struct Timed {
        cudaEvent_t a, b;
        int kernel;
    };
    std::vector<Timed> pending;
    double ms_total[kNumKernels] = {0, 0, 0, 0, 0, 0};
    int64_t ms_count[kNumKernels] = {0, 0, 0, 0, 0, 0};
    int64_t device_bytes = 0;
};

namespace {

int fail(sr_engine *e, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_create_error = buf;
    return code;
}

#define SR_CUDA(call)                                                                                       \
    do {                                                                                                    \
        cudaError_t err_ = (call);                                                                          \
        if (err_ != cudaSuccess)                                                                            \
            return fail(e, err_ == cudaErrorMemoryAllocation ? SR_ENOMEM : SR_ECUDA, "%s failed: %s (%s:%d)", \
                        #call, cudaGetErrorString(err_), __FILE__, __LINE__);                               \
    } while (0)

int ensure(sr_engine *e, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SR_OK;
    if (b.p) {
        SR_CUDA(cudaDeviceSynchronize());  // earlier passes may still run on a caller's stream
        SR_CUDA(cudaFree(b.p));
        e->device_bytes -= (int64_t)b.cap;
        b.p = nullptr;
        b.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    SR_CUDA(cudaMalloc(&b.p, want));
    b.cap = want;
    e->device_bytes += (int64_t)want;
    ++e->epoch;
    return SR_OK;
}

int ensure_pinned(sr_engine *e, size_t bytes)
{
    if (bytes <= e->h_pin_cap) return SR_OK;
    if (e->h_pin) {
        SR_CUDA(cudaStreamSynchronize(e->stream));
        SR_CUDA(cudaFreeHost(e->h_pin));
        e->h_pin = nullptr;
        e->h_pin_cap = 0;
    }
    size_t want = bytes + bytes / 4 + 4096;
    SR_CUDA(cudaMallocHost(&e->h_pin, want));
    e->h_pin_cap = want;
    return SR_OK;
}

struct Scope {  // event bracket around one kernel when profiling
    sr_engine *e;
    cudaStream_t st;
    int kernel;
    cudaEvent_t a = nullptr, b = nullptr;
    Scope(sr_engine *e_, cudaStream_t st_, int k) : e(e_), st(st_), kernel(k)
    {
        ++e->launches;
        if (!e->profile) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; return; }
        cudaEventRecord(a, st);
    }
    ~Scope()
    {
        if (!a) return;
        cudaEventRecord(b, st);
        e->pending.push_back({a, b, kernel});
    }
};

int resolve_timings(sr_engine *e)
{
    for (auto &t : e->pending) {
        SR_CUDA(cudaEventSynchronize(t.b));
        float ms = 0.f;
        SR_CUDA(cudaEventElapsedTime(&ms, t.a, t.b));
        e->ms_total[t.kernel] += ms;
        e->ms_count[t.kernel] += 1;
        cudaEventDestroy(t.a);
        cudaEventDestroy(t.b);
    }
    e->pending.clear();
    return SR_OK;
}

int pow2_floor(int64_t v)
{
    int p = 1;
    while ((int64_t)p * 2 <= v) p *= 2;
    return p;
}

int build_store(sr_engine *e)
{
    // d_raw holds n rows; pad rows are already zero
    SR_CUDA(cudaMemsetAsync(e->d_irregular, 0, 8, e->stream));
    const int threads = 256;
    const int64_t blocks = (e->n_pad + threads - 1) / threads;
    build_store_kernel<<<(unsigned)blocks, threads, 0, e->stream>>>(e->d_raw, e->n, e->n_pad, e->d_nf, e->d_hat, kAutoS, e->d_irregular);
    e->hat_S = kAutoS;
    SR_CUDA(cudaGetLastError());
    ++e->launches;
    unsigned long long irr = 0;
    SR_CUDA(cudaMemcpyAsync(&irr, e->d_irregular, 8, cudaMemcpyDeviceToHost, e->stream));
    SR_CUDA(cudaStreamSynchronize(e->stream));
    e->irregular = (int64_t)irr;
    return SR_OK;
}

int alloc_store(sr_engine *e, int64_t n, int64_t id_base)
{
    if (n <= 0) return fail(e, SR_EINVAL, "load_features: n must be positive (got %lld)", (long long)n);
    if (id_base < 0 || id_base + n > 0x7fffffffLL)
        return fail(e, SR_EINVAL, "load_features: ids [%lld, %lld) do not fit 32 bits", (long long)id_base,
                    (long long)(id_base + n));
    SR_CUDA(cudaStreamSynchronize(e->stream));
    if (e->d_raw) {
        cudaFree(e->d_raw); cudaFree(e->d_hat); cudaFree(e->d_nf);
        e->device_bytes -= e->n_pad * (int64_t)(2 * kF + 1) * 4;
        e->d_raw = e->d_hat = e->d_nf = nullptr;
        e->n = e->n_pad = 0;
    }
    const int64_t n_pad = (n + kRowPad - 1) / kRowPad * kRowPad;
    SR_CUDA(cudaMalloc(&e->d_raw, (size_t)n_pad * kF * 4));
    SR_CUDA(cudaMalloc(&e->d_hat, (size_t)n_pad * kF * 4));
    SR_CUDA(cudaMalloc(&e->d_nf, (size_t)n_pad * 4));
    e->device_bytes += n_pad * (int64_t)(2 * kF + 1) * 4;
    e->n = n;
    e->n_pad = n_pad;
    e->id_base = (int32_t)id_base;
    ++e->epoch;
    SR_CUDA(cudaMemsetAsync(e->d_raw + (size_t)n * kF, 0, (size_t)(n_pad - n) * kF * 4, e->stream));
    return SR_OK;
}

// The constant bank holding the current group's normalised queries is one per device
// (per CUDA context), shared by every engine of the process on that device: a group's
// upload is ordered behind the last scan that read the bank, on whichever stream it ran.
std::mutex g_bank_mutex;
cudaEvent_t g_bank_event[64] = {nullptr};

// where one internal pass delivers its rows
struct PassOut {
    int32_t *idx = nullptr;     // [nq][stride] or null
    float *score = nullptr;     // [nq][stride] or null
    uint64_t *keys = nullptr;   // [nq][stride] packed keys or null
    int stride = 0, col = 0;    // the pass fills columns [col, col + K)
    const uint64_t *ceil_in = nullptr;  // [nq] only keys below these are admitted (or null)
    uint64_t *ceil_out = nullptr;       // [nq] receives the K-th key of the pass (or null)
    // The bound pass shared between the row shards of one store: every shard samples 1 / sample_div of
    // what a single store would, the block maxima are max-reduced across the shards (one small all-reduce), and every
    // shard starts from thresholds that bound the K-th best of the WHOLE store.
    int sample_div = 1;
    float *blocks_out = nullptr;        // non-null: bound pass only -- [nq][bound blocks] block maxima (float, -inf = empty)
    const float *blocks_in = nullptr;   // non-null: the max-reduced maxima; no bound pass of its own
};

// blocks of the bound pass for lists of K (the same on every shard: it sizes the exchange)
int bound_block_count(const sr_engine *e, int K)
{
    const int nblk = e->bound_blocks > 0 ? std::max(e->bound_blocks, K + 1) : (K < 16 ? 64 : (K < 64 ? 128 : 256));
    return (e->bound && K + 1 <= nblk && nblk <= kLT) ? nblk : 0;
}

// One internal pass: nq <= e->batch queries, K <= kKMax, everything on device, stream-ordered.
int run_pass(sr_engine *e, const int32_t *d_qidx, const float *d_qrows, const int32_t *d_excl, int nq, int K,
             const PassOut &out, cudaStream_t st)
{
    // small batches are HBM-bound: two 256-thread CTAs per SM, song tiles staged through shared
    // memory by TMA one tile ahead; large batches are FP32-bound: one 512-thread CTA, bigger
    // tiles, shared memory spent on 256 queries' lists and hit buffers
    // in between (short lists): two 256-thread CTAs per SM loading straight into registers -- with one query tile per
    // launch a tile's load is 15-25 % of its time, and one CTA's load hides behind the other's arithmetic
    int vi = e->variant >= 0 ? e->variant : (nq <= e->small_max ? kAutoSmall : (nq <= e->mid_max && K <= 16 ? kAutoMid : kAutoLarge));
    const Variant *vp = nullptr;
    int TS = 0, n_tiles = 0, groups = 0, gsize = 0, qt_cap = 0, cap = 0;
    bool lists_in_smem = true;
    size_t stage_bytes = 0;
    for (int attempt = 0; attempt < 2 && !qt_cap; ++attempt) {
        if (attempt == 1) {  // the staged shapes have little shared memory left for long lists
            if (e->variant >= 0 || vi == kAutoLarge) break;
            vi = kAutoLarge;
        }
        vp = &kVariants[vi];
        if (vp->staged() && vp->dynamic && (nq + std::min(e->qt_opt, kQTMax) - 1) / std::min(e->qt_opt, kQTMax) > e->sm_count * vp->smem_ctas()) {
            vi = kAutoLarge;  // more query tiles than CTAs (only with a tiny "qt" option): the shape with a static fallback
            vp = &kVariants[vi];
        }
        TS = vp->S * vp->threads;
        n_tiles = (int)((e->n + TS - 1) / TS);
        // query groups (what fits the constant bank), evenly filled; then query tiles inside a group
        groups = (nq + kConstQueries - 1) / kConstQueries;
        gsize = (nq + groups - 1) / groups;
        // Queries per tile (qt) and hit-buffer entries per query (cap), both in shared memory next to
        // the CTA's exact top-K lists (qt x K keys).  Large qt amortises the per-tile costs; cap of a
        // few K lets an overflowing buffer alone lift the threshold far enough for the re-filter
        // round to converge.  First combination that fits wins.
        stage_bytes = (size_t)vp->nbuf * TS * kF * 4;
        const size_t smem_budget = vp->staged() && vp->smem_ctas() == 1 ? (size_t)230000 : (size_t)216 * 1024 / vp->smem_ctas();
        const int qt_max = std::max(1, std::min({e->qt_opt, kQTMax, vp->staged() && vp->smem_ctas() == 1 ? 64 : kQTMax}));
        const int kk = std::max(K, 32);
        for (int qtry = qt_max; qtry >= 1 && !qt_cap; qtry = (qtry > 8 ? qtry / 2 : qtry - 1)) {
            const int caps[4] = {e->hit_cap > 0 ? e->hit_cap : std::max(128, 4 * kk), std::max(128, 2 * kk),
                                 std::max(128, 3 * kk / 2), qtry <= 64 ? std::max(32, kk) : 0};
            // lists in shared memory first; for a full-size query tile of a one-CTA shape also with the
            // lists in an L2-resident workspace (a settle then costs two global round trips, a query
            // tile of half the size costs 6 % of the whole scan; measured: wins up to k ~ 72, beyond
            // that twice as many CTAs per query with long lists of their own cost more)
            for (int in_smem = 1; in_smem >= 0 && !qt_cap; --in_smem) {
                if (!in_smem && !(e->list_ws_opt && K <= e->list_ws_kmax && qtry == qt_max && qtry >= 128 && vp->ctas == 1 && !vp->staged())) break;
                for (int ci = 0; ci < 4 && !qt_cap; ++ci) {
                    const int ctry = std::min(1024, (caps[ci] + 31) / 32 * 32);
                    if (ctry > 0 && scan_smem_bytes(qtry, ctry, K, stage_bytes, in_smem != 0) <= smem_budget) {
                        qt_cap = qtry; cap = ctry; lists_in_smem = in_smem != 0;
                    }
                }
            }
        }
    }
    e->last_variant = vi;
    const Variant &v = *vp;
    if (!qt_cap) return fail(e, SR_EINVAL, "k = %d does not fit the scan kernel's shared memory", K);
    const int nqt0 = (gsize + qt_cap - 1) / qt_cap;
    const int qt = (gsize + nqt0 - 1) / nqt0;
    const int nqt = (gsize + qt - 1) / qt;
    if (!lists_in_smem && scan_smem_bytes(qt, cap, K, stage_bytes, true) <= (size_t)216 * 1024 / v.ctas) lists_in_smem = true;  // few queries: they fit after all
    if (v.dynamic && v.staged() && nqt > e->sm_count * v.smem_ctas()) return fail(e, SR_EINVAL, "kernel shape %s: %d query tiles exceed the grid; raise the \"qt\" option", v.name, nqt);
    e->last_lists_in_smem = lists_in_smem ? 1 : 0;
    const size_t smem = scan_smem_bytes(qt, cap, K, stage_bytes, lists_in_smem);
    int ctas = 0;
    SR_CUDA(v.occ(&ctas, smem));
    if (ctas < 1) return fail(e, SR_ECUDA, "scan kernel %s does not fit one SM (smem %zu)", v.name, smem);
    const int64_t units = (int64_t)nqt * n_tiles;
    if (units > 0x7fffffffLL) return fail(e, SR_EINVAL, "store too large for one pass (%lld work units)", (long long)units);
    const int grid = (int)std::min<int64_t>((int64_t)e->sm_count * ctas, units);
    e->scan_grid = grid;
    // Pool slab of a query = the most CTA segments any of this pass's groups can put on one query tile: static
    // shapes deal contiguous runs of at least upc units; dynamic shapes have their home CTAs plus the late
    // joiners a tile admits (a short last group has fewer query tiles, hence more CTAs on each)
    const int steal_max = std::min(grid, 16);
    int upc_min = 0x7fffffff, cpq_max = 0;
    for (int g0 = 0; g0 < nq; g0 += gsize) {
        const int gnqt = (std::min(gsize, nq - g0) + qt - 1) / qt;
        const int64_t gunits = (int64_t)gnqt * n_tiles;
        const int64_t ggrid = std::min<int64_t>(grid, gunits);
        upc_min = (int)std::min<int64_t>(upc_min, std::max<int64_t>(1, gunits / ggrid));
        cpq_max = std::max(cpq_max, (int)(ggrid / gnqt));
    }
    const int segs = std::max(scan_segs(n_tiles, upc_min), cpq_max + 3 + (v.dynamic ? steal_max : 0));

    // threshold bootstrap: the bound pass (filter speed, per query group) when the store has
    // enough full tiles, else the exact sample.  Neither is valid under a ceiling (they bound the
    // K-th best of ALL songs, a ceiling pass wants the K-th best below the ceiling).
    const int64_t full_tiles = e->n / TS;
    // disjoint blocks of sample songs: at least K + 1, and several times that where it is cheap, so that the
    // (K+1)-th largest block maximum comes close to the sample's exact K-th best
    const bool shared = out.blocks_out || out.blocks_in;
    const int nblk = shared ? bound_block_count(e, K) : (e->bound_blocks > 0 ? std::max(e->bound_blocks, K + 1) : (K < 16 ? 64 : (K < 64 ? 128 : 256)));
    if (shared && (nblk == 0 || out.ceil_in)) return fail(e, SR_EINVAL, "a shared bound pass needs k <= 255, the \"bound\" option on and no ceiling");
    // sample tiles: 48 (short lists) or 96 layout tiles on large stores, never more than ~6 % of the store; a shard of
    // a store that shares its bound pass samples its part of them (at least one tile, none at all if it has none)
    // (small and mid-size batches: the pass is latency-bound there, so one tile per SM costs what 48 do and starts the scan
    // with a third of the filter hits -- 16 queries over 10 M songs: 950 -> 317 hits per query, scan 104 -> 98.5 us; 64
    // queries: 1123 -> 330, call 344 -> 317 us)
    const bool few = nq <= std::max(e->small_max, e->mid_max) && K <= 16;
    // (up to ~640 queries one or two query tiles are shared by all the CTAs, each with a list of its own: a four times larger
    // sample -- 4 % of the scan's work -- halves the hits twice over and pays: 256 / 512 queries 1.095 / 2.104 -> 1.071 / 2.060 ms)
    const int64_t want_tiles = (int64_t)(e->bound_tiles > 0 ? e->bound_tiles : (few ? std::max(48, e->sm_count) : (K <= 16 ? (nq <= 640 ? 192 : 48) : 96))) * (kAutoS / v.S) / (v.threads / kLT);
    const int n_sample = shared ? (int)std::min<int64_t>(std::max<int64_t>(1, (want_tiles + out.sample_div - 1) / std::max(1, out.sample_div)), full_tiles)
                                : (int)std::min<int64_t>({want_tiles, std::max<int64_t>(4, full_tiles / (e->bound_cap_div > 0 ? e->bound_cap_div : 16)), full_tiles / 4});
    const bool use_bound = shared ? (out.blocks_out != nullptr) : (e->bound && !out.ceil_in && K + 1 <= nblk && nblk <= kLT && n_sample >= 4 && (int64_t)n_sample * TS >= 16 * (int64_t)nblk);
    // the bound pass's own (finer) query tiles: the block maxima of a tile live in shared memory.  (Larger tiles -- 64 ... 160
    // queries with 256 blocks -- were measured SLOWER, 525 -> 545 ... 655 us per 4096-query top-100 batch: the shared-memory
    // atomics and the flush, not the sample tile's load, are what the pass spends its time on.)
    const int bqt = std::max(1, std::min({qt, 64, (int)(48 * 1024 / (nblk * 4))}));
    const int bnqt_max = (gsize + bqt - 1) / bqt;

    int rc;
    if ((rc = ensure(e, e->qraw, (size_t)nq * kF * 4))) return rc;
    if ((rc = ensure(e, e->qn, (size_t)nq * 4))) return rc;
    if ((rc = ensure(e, e->qhat, (size_t)nq * kF * 4))) return rc;
    if ((rc = ensure(e, e->excl, (size_t)nq * 4))) return rc;
    if ((rc = ensure(e, e->gbest, (size_t)nq * 4))) return rc;
    if ((rc = ensure(e, e->gbound, (size_t)nq * nblk * 4))) return rc;
    const int nslot = std::max(256, (K + 31) / 32 * 32);  // residue slots per query (global threshold feedback)
    if ((rc = ensure(e, e->gslot, (size_t)nq * nslot * 4))) return rc;
    const int ctr_per_group = 2 * nqt + bnqt_max + 1;     // tile counters, visit counters, bound-pass completion per bound query tile
    if ((rc = ensure(e, e->ctr, (size_t)groups * ctr_per_group * 4))) return rc;
    if ((rc = ensure(e, e->pool_cnt, (size_t)nq * 4))) return rc;
    // settle triggers: a quarter-full buffer starts a settle phase (earlier settles = tighter thresholds = fewer
    // hits to re-run the filter for: 863 -> 474 per query at 10 M songs, 1.7 % of the scan), which takes every buffer
    // holding at least cap/32 ids along
    // pool slab of a query: K keys per segment, plus (dynamic shapes) the hits a segment can leave unsettled
    const int settle_at_eff = e->settle_at > 0 ? std::min(e->settle_at, cap) : std::max(4, cap / 32);
    const int trigger_at_eff = e->trigger_at > 0 ? std::min(std::max(e->trigger_at, settle_at_eff), cap) : std::max(settle_at_eff, cap / 4);
    const int64_t slab64 = (int64_t)segs * (K + (v.dynamic ? trigger_at_eff : 0));
    if (slab64 * 8 * nq > ((int64_t)8 << 30)) return fail(e, SR_EINVAL, "k = %d with %d queries per pass needs a %lld-byte pool: lower the \"batch\" option", K, nq, (long long)(slab64 * 8 * nq));
    const int slab = (int)slab64;
    if ((rc = ensure(e, e->pool, (size_t)nq * slab * 8))) return rc;
    if (!lists_in_smem && (rc = ensure(e, e->list_ws, (size_t)grid * qt * K * 8))) return rc;

    // the constant bank is free once the last scan that read it has finished (on whichever stream)
    std::lock_guard<std::mutex> lock(g_bank_mutex);
    cudaEvent_t &ev = g_bank_event[e->device & 63];
    if (!ev) SR_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    else SR_CUDA(cudaStreamWaitEvent(st, ev, 0));
    // everything the pass enqueues, on `st` -- directly, or into a stream capture (no event operations then)
    auto enqueue = [&](cudaStream_t st, bool capturing) -> int {
    {
        // one kernel resets the pass's workspace, prepares the queries and -- for a single-group pass --
        // writes the normalised rows straight into the constant bank
        PrepArgs p;
        p.raw_store = e->d_raw; p.n = e->n; p.id_base = e->id_base;
        p.qidx = d_qidx; p.qrows_in = d_qrows; p.excl_in = d_excl; p.nq = nq;
        p.qraw = (float *)e->qraw.p; p.qn = (float *)e->qn.p; p.qhat = (float *)e->qhat.p;
        p.excl = (int32_t *)e->excl.p; p.pool_cnt = (int32_t *)e->pool_cnt.p;
        p.g_best = (uint32_t *)e->gbest.p;
        p.flag = e->d_flag;
        p.gslot = (uint32_t *)e->gslot.p; p.gslot_words = (int64_t)nq * nslot;
        p.gbound = use_bound ? (uint32_t *)e->gbound.p : nullptr; p.gbound_words = use_bound ? (int64_t)nq * nblk : 0;
        p.ctr = (int *)e->ctr.p; p.ctr_words = groups * ctr_per_group;
        p.cbank = groups == 1 ? e->d_cbank : nullptr; p.cbank_q = nq;
        const int64_t zero_words = p.gslot_words / 4 + p.gbound_words + p.ctr_words;
        const int blocks = (int)std::max<int64_t>((nq + 127) / 128, std::min<int64_t>((int64_t)e->sm_count * 4, (zero_words + 128 * 8 - 1) / (128 * 8)));
        Scope sc(e, st, kPrep);
        prep_queries_kernel<<<blocks, 128, 0, st>>>(p);
        SR_CUDA(cudaGetLastError());
    }
    // threshold bootstrap: the bound pass (filter speed, per query group, below) when the store
    // has enough full tiles, else / additionally the exact sample
    int m = e->sample;
    if (out.ceil_in || shared) m = 0;
    if (m < 0 && use_bound) m = 0;
    if (m < 0) m = std::min(kSortCap, std::max(1024, 2 * pow2_floor((int64_t)K * 32 - 1)));
    if (m > 0) {
        m = std::min(m, pow2_floor(e->n / 4));  // only worth it on stores much larger than the sample
        if (m >= 2 * K && m >= 64) {
            SampleArgs s;
            s.raw = e->d_raw; s.nf = e->d_nf; s.n = e->n; s.id_base = e->id_base;
            s.qraw = (float *)e->qraw.p; s.qn = (float *)e->qn.p; s.exclude = (int32_t *)e->excl.p;
            s.nq = nq; s.m = m; s.K = K; s.g_best = (uint32_t *)e->gbest.p;
            Scope sc(e, st, kSample);
            sample_threshold_kernel<256><<<nq, 256, 0, st>>>(s);
            SR_CUDA(cudaGetLastError());
        }
    }
    if (out.blocks_in) {
        blocks_finish_kernel<<<(nq + 7) / 8, 256, 0, st>>>(out.blocks_in, nblk, K, (uint32_t *)e->gbest.p, nq);
        SR_CUDA(cudaGetLastError());
        ++e->launches;
    }
    int gi = 0;
    for (int g0 = 0; g0 < nq; g0 += gsize, ++gi) {
        const int gq = std::min(gsize, nq - g0);
        ScanArgs a;
        a.hat = e->d_hat; a.raw = e->d_raw; a.nf = e->d_nf; a.n = e->n; a.id_base = e->id_base;
        a.n_tiles = n_tiles;
        a.qraw = (float *)e->qraw.p + (size_t)g0 * kF; a.qn = (float *)e->qn.p + g0;
        a.exclude = (int32_t *)e->excl.p + g0; a.nq = gq; a.qt = qt;
        a.K = K; a.cap = cap; a.settle_at = settle_at_eff;
        // (dynamic shapes: every tile -- a CTA's tiles come from all over the store, it lives on what the others found)
        a.refresh_every = e->refresh_every > 0 ? e->refresh_every : (v.dynamic ? 1 : std::max(1, std::min(8, 64 / qt)));
        a.trigger_at = trigger_at_eff;
        a.gslot = (uint32_t *)e->gslot.p + (size_t)g0 * nslot;
        a.nslot = nslot;
        a.prefetch = e->prefetch > 0 ? 1 : 0;  // (measured three times, from 64 to 4096 queries per batch and at top-100: no change)
        a.ceil = out.ceil_in ? out.ceil_in + g0 : nullptr;
        a.pool = (uint64_t *)e->pool.p + (size_t)g0 * slab; a.pool_cnt = (int32_t *)e->pool_cnt.p + g0;
        a.segs = segs; a.slab = slab;
        a.g_best = (uint32_t *)e->gbest.p + g0;
        a.list_ws = lists_in_smem ? nullptr : (uint64_t *)e->list_ws.p;
        a.stats = e->d_stats;
        const int gnqt = (gq + qt - 1) / qt;
        int *gctr = (int *)e->ctr.p + (size_t)gi * ctr_per_group;
        if (groups > 1) {  // (stream order keeps the upload behind the previous group's scan)
            SR_CUDA(cudaMemcpyToSymbolAsync(c_qhat, (const float *)e->qhat.p + (size_t)g0 * kF, (size_t)gq * kF * 4, 0,
                                            cudaMemcpyDeviceToDevice, st));
        }
        a.bound_finish = out.blocks_out ? 0 : 1;
        if (use_bound && n_sample > 0) {
            ScanArgs b = a;
            b.qt = bqt;
            const int bnqt = (gq + b.qt - 1) / b.qt;
            const int bgrid = (int)std::min<int64_t>((int64_t)e->sm_count * v.ctas, (int64_t)bnqt * n_sample);
            uint32_t *gmax = (uint32_t *)e->gbound.p + (size_t)g0 * nblk;
            Scope sc(e, st, kBound);
            SR_CUDA(v.bound(b, nblk, n_sample, (int)(full_tiles / n_sample), gmax, gctr + 2 * nqt, bgrid, (size_t)b.qt * nblk * 4, st));
        }
        if (!out.blocks_out) {
            a.n_tiles = n_tiles;
            a.tile_stride = 1;
            const int64_t gunits = (int64_t)gnqt * n_tiles;
            const int ggrid = (int)std::min<int64_t>(grid, gunits);
            a.upc = (int)(gunits / ggrid);
            a.extra = (int)(gunits % ggrid);
            a.cpq = 0;
            a.tile_ctr = gctr;
            a.visit_ctr = gctr + nqt;
            a.steal_max = steal_max;
            const Variant *vl = &v;
            if (v.dynamic && gnqt > ggrid) vl = &kVariants[v.threads == 512 ? kStaticLarge : 1];  // more query tiles than CTAs: static runs (same S, threads, smem)
            if (vl->dynamic) a.cpq = ggrid / gnqt;
            Scope sc(e, st, kScan);
            SR_CUDA(vl->launch(a, ggrid, smem, st));
        }
        if (!capturing) SR_CUDA(cudaEventRecord(ev, st));
    }
    if (out.blocks_out) {  // bound pass only: hand the block maxima over as floats (what the all-reduce maximises)
        const int64_t count = (int64_t)nq * nblk;
        blocks_to_float_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>((const uint32_t *)e->gbound.p, count, out.blocks_out);
        SR_CUDA(cudaGetLastError());
        ++e->launches;
        return SR_OK;
    }
    {
        FinalArgs f;
        f.pool = (uint64_t *)e->pool.p; f.pool_cnt = (int32_t *)e->pool_cnt.p;
        f.nq = nq; f.K = K; f.slab = slab;
        f.stride = out.stride > 0 ? out.stride : K; f.col = out.col;
        f.out_idx = out.idx; f.out_score = out.score; f.out_keys = out.keys;
        f.ceil_out = out.ceil_out; f.flag = e->d_flag;
        Scope sc(e, st, kFinalize);
        finalize_kernel<256><<<nq, 256, 0, st>>>(f);
        SR_CUDA(cudaGetLastError());
    }
    return SR_OK;
    };

    const bool graphable = e->use_graphs && !e->profile && groups == 1 && nq <= 256;
    if (!graphable) {
        if ((rc = enqueue(st, false))) return rc;
    } else {
        if (e->graphs_epoch != e->epoch || e->graphs.size() > 64) {
            for (auto &kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
            e->graphs.clear();
            e->graphs_epoch = e->epoch;
        }
        sr_engine::GraphKey key;
        memset(&key, 0, sizeof key);
        const void *ptrs[10] = {d_qidx, d_qrows, d_excl, out.idx, out.score, out.keys, out.ceil_in, out.ceil_out, out.blocks_out, out.blocks_in};
        memcpy(key.p, ptrs, sizeof ptrs);
        key.nq = nq; key.K = K; key.stride = out.stride; key.col = out.col + (out.sample_div << 16);
        auto it = e->graphs.find(key);
        if (it == e->graphs.end()) {
            // captured on the engine's own stream (the caller's may be the legacy stream, which cannot capture);
            // a capture only records, it does not run or wait for anything
            const int64_t launches_before = e->launches;
            SR_CUDA(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
            rc = enqueue(e->stream, true);
            cudaGraph_t g = nullptr;
            const cudaError_t ce = cudaStreamEndCapture(e->stream, &g);
            const int captured = (int)(e->launches - launches_before);
            e->launches = launches_before;
            if (rc || ce != cudaSuccess) {
                if (g) cudaGraphDestroy(g);
                cudaGetLastError();
                return rc ? rc : fail(e, SR_ECUDA, "stream capture of a small pass failed: %s", cudaGetErrorString(ce));
            }
            cudaGraphExec_t ex = nullptr;
            const cudaError_t ie = cudaGraphInstantiate(&ex, g, 0);
            cudaGraphDestroy(g);
            if (ie != cudaSuccess) return fail(e, SR_ECUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
            it = e->graphs.emplace(key, sr_engine::GraphEntry{ex, captured}).first;
        } else {
            ++e->graph_replays;
        }
        SR_CUDA(cudaGraphLaunch(it->second.exec, st));
        e->launches += it->second.launches;
        SR_CUDA(cudaEventRecord(ev, st));
    }
    if (!out.blocks_out) e->queries += nq;
    return SR_OK;
}

// SR_ENGINE_OWN_STREAM selects the engine's stream; anything else is a cudaStream_t
// (NULL being CUDA's legacy default stream, which is what torch's default stream is).
cudaStream_t pick_stream(sr_engine *e, void *stream)
{
    return stream == SR_ENGINE_OWN_STREAM ? e->stream : (cudaStream_t)stream;
}

int check_query_args(sr_engine *e, const void *q, int nq, int k, const void *out_idx)
{
    if (!e) return SR_EINVAL;
    if (!e->d_raw) return fail(e, SR_ESTATE, "no store loaded: call sr_engine_load_features first");
    if (!q || !out_idx) return fail(e, SR_EINVAL, "null query or output pointer");
    if (nq <= 0) return fail(e, SR_EINVAL, "nq must be positive (got %d)", nq);
    if (k <= 0) return fail(e, SR_EINVAL, "k must be positive (got %d)", k);
    return SR_OK;
}

// A batch of any size and any k: internal passes of <= e->batch queries; a k beyond kKMax is served kKMax
// results at a time, each pass admitting only keys below the last key of the pass before (the ceiling), so the
// caller sees min(k, songs - 1) results per query like the reference (Recommender.cu:300-315).
int run_device(sr_engine *e, const int32_t *d_qidx, const float *d_qrows, const int32_t *d_excl, int nq, int k,
               int32_t *d_out_idx, float *d_out_score, uint64_t *d_out_keys, const uint64_t *d_ceil, cudaStream_t st,
               const float *d_blocks = nullptr)
{
    const int chunks = (k + kKMax - 1) / kKMax;
    const int blk_stride = d_blocks ? bound_block_count(e, k) : 0;
    for (int done = 0; done < nq; done += e->batch) {
        const int cur = std::min(e->batch, nq - done);
        int rc;
        if (chunks > 1 && (rc = ensure(e, e->ceil, (size_t)cur * 8))) return rc;
        for (int c = 0; c < chunks; ++c) {
            PassOut o;
            o.stride = k; o.col = c * kKMax;
            o.idx = d_out_idx ? d_out_idx + (size_t)done * k : nullptr;
            o.score = d_out_score ? d_out_score + (size_t)done * k : nullptr;
            o.keys = d_out_keys ? d_out_keys + (size_t)done * k : nullptr;
            o.ceil_in = c > 0 ? (const uint64_t *)e->ceil.p : (d_ceil ? d_ceil + done : nullptr);
            o.ceil_out = c + 1 < chunks ? (uint64_t *)e->ceil.p : nullptr;
            o.blocks_in = d_blocks ? d_blocks + (size_t)done * blk_stride : nullptr;
            rc = run_pass(e, d_qidx ? d_qidx + done : nullptr, d_qrows ? d_qrows + (size_t)done * kF : nullptr,
                          d_excl ? d_excl + done : nullptr, cur, std::min(kKMax, k - c * kKMax), o, st);
            if (rc) return rc;
        }
    }
    return SR_OK;
}

// the sticky device flag, read at a synchronising call (clears it)
int check_flag(sr_engine *e, cudaStream_t st)
{
    SR_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, 4, cudaMemcpyDeviceToHost, st));
    SR_CUDA(cudaStreamSynchronize(st));
    const int f = *e->h_flag;
    if (!f) return SR_OK;
    SR_CUDA(cudaMemsetAsync(e->d_flag, 0, 4, st));
    if (f & 2) return fail(e, SR_ECUDA, "internal error: a result pool overflowed (results of the last batch are incomplete)");
    return fail(e, SR_EINVAL, "a query id is not a song of this store [%d, %lld): its result row is -1",
                e->id_base, (long long)(e->id_base + e->n));
}

// host-buffer front end shared by query_by_index / query_by_vector
int run_host(sr_engine *e, const int32_t *qidx, const float *qrows, const int32_t *exclude, int nq, int k,
             int32_t *out_idx, float *out_score)
{
    const size_t in_q = qidx ? (size_t)nq * 4 : (size_t)nq * kF * 4;
    const size_t in_x = exclude ? (size_t)nq * 4 : 0;
    const size_t out_i = (size_t)nq * k * 4;
    const size_t out_s = out_score ? (size_t)nq * k * 4 : 0;
    int rc;
    if ((rc = ensure_pinned(e, in_q + in_x + out_i + out_s))) return rc;
    if ((rc = ensure(e, e->qin, in_q + in_x))) return rc;
    if ((rc = ensure(e, e->out, out_i + out_s))) return rc;
    char *pin = (char *)e->h_pin;
    memcpy(pin, qidx ? (const void *)qidx : (const void *)qrows, in_q);
    if (in_x) memcpy(pin + in_q, exclude, in_x);
    SR_CUDA(cudaMemcpyAsync(e->qin.p, pin, in_q + in_x, cudaMemcpyHostToDevice, e->stream));  // one copy in ...
    char *d_in = (char *)e->qin.p, *d_out = (char *)e->out.p;
    rc = run_device(e, qidx ? (int32_t *)d_in : nullptr, qidx ? nullptr : (float *)d_in,
                    in_x ? (int32_t *)(d_in + in_q) : nullptr, nq, k, (int32_t *)d_out,
                    out_s ? (float *)(d_out + out_i) : nullptr, nullptr, nullptr, e->stream);
    if (rc) return rc;
    char *pout = pin + in_q + in_x;
    SR_CUDA(cudaMemcpyAsync(pout, d_out, out_i + out_s, cudaMemcpyDeviceToHost, e->stream));  // ... one copy out
    if ((rc = check_flag(e, e->stream))) return rc;
    memcpy(out_idx, pout, out_i);
    if (out_s) memcpy(out_score, pout + out_i, out_s);
    return SR_OK;
}

__global__ void iota_kernel(int32_t *out, int32_t start, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = start + i;
}

}  // namespace

extern "C" {

int sr_engine_create(sr_engine **out, int device)
{
    sr_engine *e = nullptr;
    if (!out) return fail(nullptr, SR_EINVAL, "create: null output pointer");
    *out = nullptr;
    int count = 0;
    cudaError_t err = cudaGetDeviceCount(&count);
    if (err != cudaSuccess || count == 0)
        return fail(nullptr, SR_ENODEVICE, "no CUDA device: %s (this engine has no CPU fallback)",
                    err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) device = 0;
    }
    if (device >= count) return fail(nullptr, SR_ENODEVICE, "device %d out of range (%d present)", device, count);
    cudaDeviceProp prop;
    if ((err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, SR_ECUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(err));
    if (prop.major != 10)
        return fail(nullptr, SR_ENODEVICE, "device %d (%s) is sm_%d%d; this engine is built for sm_100a only", device,
                    prop.name, prop.major, prop.minor);
    if ((err = cudaSetDevice(device)) != cudaSuccess)
        return fail(nullptr, SR_ECUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(err));
    e = new sr_engine();
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    if ((err = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (err = cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (err = cudaEventCreateWithFlags(&e->ev_done[0], cudaEventDisableTiming)) != cudaSuccess ||
        (err = cudaEventCreateWithFlags(&e->ev_done[1], cudaEventDisableTiming)) != cudaSuccess ||
        (err = cudaEventCreateWithFlags(&e->ev_copied[0], cudaEventDisableTiming)) != cudaSuccess ||
        (err = cudaEventCreateWithFlags(&e->ev_copied[1], cudaEventDisableTiming)) != cudaSuccess ||
        (err = cudaGetSymbolAddress((void **)&e->d_cbank, c_qhat)) != cudaSuccess ||
        (err = cudaMallocHost((void **)&e->h_flag, 4)) != cudaSuccess ||
        (err = cudaMalloc(&e->d_stats, 16 * 8)) != cudaSuccess || (err = cudaMalloc(&e->d_irregular, 8)) != cudaSuccess ||
        (err = cudaMalloc(&e->d_flag, 4)) != cudaSuccess || (err = cudaMemset(e->d_stats, 0, 128)) != cudaSuccess ||
        (err = cudaMemset(e->d_flag, 0, 4)) != cudaSuccess) {
        int rc = fail(nullptr, SR_ECUDA, "engine setup: %s", cudaGetErrorString(err));
        sr_engine_destroy(e);
        return rc;
    }
    *out = e;
    return SR_OK;
}

void sr_engine_destroy(sr_engine *e)
{
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    for (auto &t : e->pending) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    for (auto &kv : e->graphs) cudaGraphExecDestroy(kv.second.exec);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    DevBuf *bufs[] = {&e->qraw, &e->qn, &e->qhat, &e->excl, &e->gbest, &e->gbound, &e->gslot, &e->ctr, &e->pool_cnt, &e->pool, &e->minmax, &e->list_ws,
                      &e->out, &e->out2, &e->qin, &e->ceil};
    for (DevBuf *b : bufs)
        if (b->p) cudaFree(b->p);
    if (e->d_raw) cudaFree(e->d_raw);
    if (e->d_hat) cudaFree(e->d_hat);
    if (e->d_nf) cudaFree(e->d_nf);
    if (e->d_stats) cudaFree(e->d_stats);
    if (e->d_irregular) cudaFree(e->d_irregular);
    if (e->d_flag) cudaFree(e->d_flag);
    if (e->h_pin) cudaFreeHost(e->h_pin);
    if (e->h_flag) cudaFreeHost(e->h_flag);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_done[i]) cudaEventDestroy(e->ev_done[i]);
        if (e->ev_copied[i]) cudaEventDestroy(e->ev_copied[i]);
    }
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char *sr_engine_last_error(const sr_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int sr_engine_load_features(sr_engine *e, const float *rows, int64_t n, int64_t id_base)
{
    if (!e) return SR_EINVAL;
    if (!rows) return fail(e, SR_EINVAL, "load_features: null rows");
    SR_CUDA(cudaSetDevice(e->device));
    int rc = alloc_store(e, n, id_base);
    if (rc) return rc;
    SR_CUDA(cudaMemcpyAsync(e->d_raw, rows, (size_t)n * kF * 4, cudaMemcpyHostToDevice, e->stream));
    return build_store(e);
}

int sr_engine_load_features_device(sr_engine *e, const float *d_rows, int64_t n, int64_t id_base)
{
    if (!e) return SR_EINVAL;
    if (!d_rows) return fail(e, SR_EINVAL, "load_features_device: null rows");
    SR_CUDA(cudaSetDevice(e->device));
    int rc = alloc_store(e, n, id_base);
    if (rc) return rc;
    SR_CUDA(cudaMemcpyAsync(e->d_raw, d_rows, (size_t)n * kF * 4, cudaMemcpyDeviceToDevice, e->stream));
    return build_store(e);
}

int64_t sr_engine_song_count(const sr_engine *e) { return e ? e->n : 0; }

int sr_engine_query_by_index(sr_engine *e, const int32_t *qidx, int nq, int k, int32_t *out_idx, float *out_score)
{
    int rc = check_query_args(e, qidx, nq, k, out_idx);
    if (rc) return rc;
    for (int i = 0; i < nq; ++i) {
        const int64_t local = (int64_t)qidx[i] - e->id_base;
        if (local < 0 || local >= e->n)
            return fail(e, SR_EINVAL, "query %d: song id %d is not in this store [%d, %lld)", i, qidx[i], e->id_base,
                        (long long)(e->id_base + e->n));
    }
    SR_CUDA(cudaSetDevice(e->device));
    return run_host(e, qidx, nullptr, nullptr, nq, k, out_idx, out_score);
}

int sr_engine_query_by_vector(sr_engine *e, const float *qrows, const int32_t *exclude, int nq, int k,
                              int32_t *out_idx, float *out_score)
{
    int rc = check_query_args(e, qrows, nq, k, out_idx);
    if (rc) return rc;
    SR_CUDA(cudaSetDevice(e->device));
    return run_host(e, nullptr, qrows, exclude, nq, k, out_idx, out_score);
}

int sr_engine_query_by_index_dev(sr_engine *e, const int32_t *d_qidx, int nq, int k, int32_t *d_out_idx,
                                 float *d_out_score, void *stream)
{
    int rc = check_query_args(e, d_qidx, nq, k, d_out_idx);
    if (rc) return rc;
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    return run_device(e, d_qidx, nullptr, nullptr, nq, k, d_out_idx, d_out_score, nullptr, nullptr, st);
}

int sr_engine_query_by_vector_dev(sr_engine *e, const float *d_qrows, const int32_t *d_exclude, int nq, int k,
                                  int32_t *d_out_idx, float *d_out_score, void *stream)
{
    int rc = check_query_args(e, d_qrows, nq, k, d_out_idx);
    if (rc) return rc;
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    return run_device(e, nullptr, d_qrows, d_exclude, nq, k, d_out_idx, d_out_score, nullptr, nullptr, st);
}

int sr_engine_query_keys_by_vector_dev(sr_engine *e, const float *d_qrows, const int32_t *d_exclude, int nq, int k,
                                       const uint64_t *d_ceil, const float *d_blocks, uint64_t *d_out_keys, void *stream)
{
    int rc = check_query_args(e, d_qrows, nq, k, d_out_keys);
    if (rc) return rc;
    if (k > kKMax) return fail(e, SR_EINVAL, "query_keys: k must be in [1, %d] (got %d); longer lists go %d at a time under a ceiling", kKMax, k, kKMax);
    if (d_blocks && d_ceil) return fail(e, SR_EINVAL, "query_keys: shared block maxima do not apply under a ceiling");
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    return run_device(e, nullptr, d_qrows, d_exclude, nq, k, nullptr, nullptr, d_out_keys, d_ceil, st, d_blocks);
}

int sr_engine_bound_block_count(sr_engine *e, int k)
{
    return (e && k >= 1 && k <= kKMax) ? bound_block_count(e, k) : 0;
}

int sr_engine_bound_blocks_dev(sr_engine *e, const float *d_qrows, int nq, int k, int shards, float *d_blocks, void *stream)
{
    int rc = check_query_args(e, d_qrows, nq, k, d_blocks);
    if (rc) return rc;
    const int nblk = k <= kKMax ? bound_block_count(e, k) : 0;
    if (!nblk) return fail(e, SR_EINVAL, "bound_blocks: no bound pass for k = %d (sr_engine_bound_block_count() is 0)", k);
    if (shards < 1) return fail(e, SR_EINVAL, "bound_blocks: shards must be positive");
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    for (int done = 0; done < nq; done += e->batch) {
        const int cur = std::min(e->batch, nq - done);
        PassOut o;
        o.sample_div = shards;
        o.blocks_out = d_blocks + (size_t)done * nblk;
        if ((rc = run_pass(e, nullptr, d_qrows + (size_t)done * kF, nullptr, cur, k, o, st))) return rc;
    }
    return SR_OK;
}

namespace {
int merge_common(sr_engine *e, const uint64_t *d_keys, const uint64_t *const *d_part_keys, const int32_t *d_idx, const float *d_score,
                 int parts, int nq, int k, int32_t *d_out_idx, float *d_out_score, int stride, int col, uint64_t *d_ceil_out, void *stream)
{
    if (!e) return SR_EINVAL;
    if ((!d_keys && !d_part_keys && !(d_idx && d_score)) || !d_out_idx) return fail(e, SR_EINVAL, "merge: null pointer");
    if (parts <= 0 || nq <= 0 || k <= 0 || k > kKMax) return fail(e, SR_EINVAL, "merge: bad parts/nq/k (k must be in [1, %d])", kKMax);
    if (stride < col + k) return fail(e, SR_EINVAL, "merge: output window [%d, %d) exceeds the row stride %d", col, col + k, stride);
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    Scope sc(e, st, kMerge);
    merge_parts_kernel<256><<<nq, 256, 0, st>>>(d_keys, d_part_keys, d_idx, d_score, parts, nq, k, d_out_idx, d_out_score, stride, col, d_ceil_out);
    SR_CUDA(cudaGetLastError());
    return SR_OK;
}
}  // namespace

int sr_engine_merge_topk_dev(sr_engine *e, const int32_t *d_idx, const float *d_score, int parts, int nq, int k,
                             int32_t *d_out_idx, float *d_out_score, void *stream)
{
    return merge_common(e, nullptr, nullptr, d_idx, d_score, parts, nq, k, d_out_idx, d_out_score, k, 0, nullptr, stream);
}

int sr_engine_merge_keys_dev(sr_engine *e, const uint64_t *d_keys, int parts, int nq, int k, int32_t *d_out_idx,
                             float *d_out_score, int stride, int col, uint64_t *d_ceil_out, void *stream)
{
    return merge_common(e, d_keys, nullptr, nullptr, nullptr, parts, nq, k, d_out_idx, d_out_score, stride > 0 ? stride : k, col, d_ceil_out, stream);
}

int sr_engine_gather_rows_dev(sr_engine *e, const int32_t *d_ids, int count, float *d_out, void *stream)
{
    if (!e) return SR_EINVAL;
    if (!e->d_raw) return fail(e, SR_ESTATE, "no store loaded: call sr_engine_load_features first");
    if (!d_ids || !d_out || count <= 0) return fail(e, SR_EINVAL, "gather_rows: bad arguments");
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    gather_rows_kernel<<<(count + 127) / 128, 128, 0, st>>>(e->d_raw, e->n, e->id_base, d_ids, count, d_out);
    SR_CUDA(cudaGetLastError());
    ++e->launches;
    return SR_OK;
}

// genre name -> id.  The reference hands out ids first-come inside an OpenMP
// critical section (DataManager.cpp:244-250), so they depend on thread scheduling; mode 0
// restates what it does with ONE thread (order of first appearance), mode 1 is the
// deterministic replacement (rank of the name in sorted order).  Host only.
int sr_genre_ids(const char *const *names, int64_t n, int mode, int32_t *ids, int32_t *n_genres)
{
    if (!names || !ids || n < 0 || (mode != 0 && mode != 1)) return SR_EINVAL;
    std::map<std::string, int32_t> seen;
    for (int64_t i = 0; i < n; ++i) {
        if (!names[i]) return SR_EINVAL;
        auto it = seen.find(names[i]);
        if (it == seen.end()) it = seen.emplace(names[i], (int32_t)seen.size()).first;
        ids[i] = it->second;
    }
    if (mode == 1) {
        std::vector<int32_t> rank(seen.size());
        int32_t r = 0;
        for (const auto &kv : seen) rank[kv.second] = r++;  // std::map iterates in sorted key order
        for (int64_t i = 0; i < n; ++i) ids[i] = rank[ids[i]];
    }
    if (n_genres) *n_genres = (int32_t)seen.size();
    return SR_OK;
}

// the normalisation step of the reference's preprocessing (DataManager.cpp:270-301) on the GPU
int sr_engine_normalize_features_dev(sr_engine *e, const float *d_raw11, const int32_t *d_genre, int64_t n, int32_t n_genres,
                                     float *d_out, float *d_minmax, void *stream)
{
    if (!e) return SR_EINVAL;
    if (!d_raw11 || !d_genre || !d_out) return fail(e, SR_EINVAL, "normalize_features: null pointer");
    if (n <= 0 || n > 0x7fffffffLL) return fail(e, SR_EINVAL, "normalize_features: n must be in [1, 2^31) (got %lld)", (long long)n);
    if (n_genres < 1) return fail(e, SR_EINVAL, "normalize_features: n_genres must be positive (got %d)", n_genres);
    SR_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = pick_stream(e, stream);
    int rc;
    if ((rc = ensure(e, e->minmax, 2 * kRawF * 4))) return rc;
    SR_CUDA(cudaMemsetAsync(e->minmax.p, 0, 2 * kRawF * 4, st));  // "empty" in the kernels' encoding
    const int64_t want = (n * kRawF / 4 + 256 * 4 - 1) / (256 * 4);  // ~4 128-bit loads per thread
    const int blocks = (int)std::max<int64_t>(kRawF, std::min<int64_t>((int64_t)e->sm_count * 8 / kRawF * kRawF, (want + kRawF - 1) / kRawF * kRawF));
    minmax_kernel<<<blocks, 256, 0, st>>>(d_raw11, n, (uint32_t *)e->minmax.p);
    SR_CUDA(cudaGetLastError());
    const float genre_den = (float)std::max(1, n_genres - 1);
    const int nblocks = (int)std::min<int64_t>((int64_t)e->sm_count * 8, (n + kNormRows - 1) / kNormRows);
    normalize_kernel<<<nblocks, kNormRows, 0, st>>>(d_raw11, d_genre, n, genre_den, (const uint32_t *)e->minmax.p, d_out, d_minmax);
    SR_CUDA(cudaGetLastError());
    e->launches += 2;
    return SR_OK;
}

int sr_engine_normalize_features(sr_engine *e, const float *raw11, const int32_t *genre_id, int64_t n, int32_t n_genres,
                                 float *out, float *minmax_out)
{
    if (!e) return SR_EINVAL;
    if (!raw11 || !genre_id || !out) return fail(e, SR_EINVAL, "normalize_features: null pointer");
    if (n <= 0 || n > 0x7fffffffLL) return fail(e, SR_EINVAL, "normalize_features: n must be in [1, 2^31) (got %lld)", (long long)n);
    SR_CUDA(cudaSetDevice(e->device));
    float *d_raw = nullptr, *d_out = nullptr, *d_mm = nullptr;
    int32_t *d_genre = nullptr;
    cudaError_t err;
    int rc = SR_OK;
    if ((err = cudaMalloc(&d_raw, (size_t)n * kRawF * 4)) != cudaSuccess || (err = cudaMalloc(&d_out, (size_t)n * kF * 4)) != cudaSuccess ||
        (err = cudaMalloc(&d_genre, (size_t)n * 4)) != cudaSuccess || (err = cudaMalloc(&d_mm, 2 * kRawF * 4)) != cudaSuccess ||
        (err = cudaMemcpyAsync(d_raw, raw11, (size_t)n * kRawF * 4, cudaMemcpyHostToDevice, e->stream)) != cudaSuccess ||
        (err = cudaMemcpyAsync(d_genre, genre_id, (size_t)n * 4, cudaMemcpyHostToDevice, e->stream)) != cudaSuccess)
        rc = fail(e, SR_ECUDA, "normalize_features: %s", cudaGetErrorString(err));
    if (!rc) rc = sr_engine_normalize_features_dev(e, d_raw, d_genre, n, n_genres, d_out, d_mm, SR_ENGINE_OWN_STREAM);
    if (!rc && ((err = cudaMemcpyAsync(out, d_out, (size_t)n * kF * 4, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess ||
                (minmax_out && (err = cudaMemcpyAsync(minmax_out, d_mm, 2 * kRawF * 4, cudaMemcpyDeviceToHost, e->stream)) != cudaSuccess) ||
                (err = cudaStreamSynchronize(e->stream)) != cudaSuccess))
        rc = fail(e, SR_ECUDA, "normalize_features: %s", cudaGetErrorString(err));
    cudaFree(d_raw); cudaFree(d_out); cudaFree(d_genre); cudaFree(d_mm);
    return rc;
}

int sr_engine_all_pairs_topk(sr_engine *e, int64_t q_lo, int64_t q_hi, int k, int32_t *out_idx, float *out_score)
{
    if (!e) return SR_EINVAL;
    if (!e->d_raw) return fail(e, SR_ESTATE, "no store loaded: call sr_engine_load_features first");
    if (!out_idx) return fail(e, SR_EINVAL, "all_pairs: null output");
    if (k <= 0) return fail(e, SR_EINVAL, "k must be positive (got %d)", k);
    if (q_lo < e->id_base || q_hi > e->id_base + e->n || q_lo >= q_hi)
        return fail(e, SR_EINVAL, "all_pairs: [%lld, %lld) is not inside this store", (long long)q_lo, (long long)q_hi);
    SR_CUDA(cudaSetDevice(e->device));
    // Batches of e->batch queries; the results of batch b travel to the host (copy stream, pinned halves) while
    // batch b + 1 is being scored, and are unpacked into the caller's table while batch b + 2 is enqueued.
    const int B = (int)std::min<int64_t>(e->batch, q_hi - q_lo);
    const size_t half = (size_t)B * k * (out_score ? 8 : 4);
    int rc;
    if ((rc = ensure(e, e->qin, (size_t)B * 4))) return rc;
    if ((rc = ensure(e, e->out, half))) return rc;
    if ((rc = ensure(e, e->out2, half))) return rc;
    if ((rc = ensure_pinned(e, 2 * half))) return rc;
    char *pin = (char *)e->h_pin;
    struct Batch { int64_t lo; int cur; } inflight[2] = {{0, 0}, {0, 0}};
    auto unpack = [&](int sl) {
        const Batch &b = inflight[sl];
        const size_t bytes = (size_t)b.cur * k * 4;
        memcpy(out_idx + (size_t)(b.lo - q_lo) * k, pin + sl * half, bytes);
        if (out_score) memcpy(out_score + (size_t)(b.lo - q_lo) * k, pin + sl * half + bytes, bytes);
    };
    int nb = 0;
    for (int64_t lo = q_lo; lo < q_hi; lo += B, ++nb) {
        const int sl = nb & 1;
        const int cur = (int)std::min<int64_t>(B, q_hi - lo);
        char *d_out = (char *)(sl ? e->out2.p : e->out.p);
        if (nb >= 2) {  // slot reuse: batch nb - 2 must have left the device buffer and the pinned half
            SR_CUDA(cudaEventSynchronize(e->ev_copied[sl]));
            unpack(sl);
        }
        iota_kernel<<<(cur + 255) / 256, 256, 0, e->stream>>>((int32_t *)e->qin.p, (int32_t)lo, cur);
        SR_CUDA(cudaGetLastError());
        ++e->launches;
        const size_t bytes = (size_t)cur * k * 4;
        rc = run_device(e, (int32_t *)e->qin.p, nullptr, nullptr, cur, k, (int32_t *)d_out,
                        out_score ? (float *)(d_out + bytes) : nullptr, nullptr, nullptr, e->stream);
        if (rc) return rc;
        SR_CUDA(cudaEventRecord(e->ev_done[sl], e->stream));
        SR_CUDA(cudaStreamWaitEvent(e->copy_stream, e->ev_done[sl], 0));
        SR_CUDA(cudaMemcpyAsync(pin + sl * half, d_out, bytes * (out_score ? 2 : 1), cudaMemcpyDeviceToHost, e->copy_stream));
        SR_CUDA(cudaEventRecord(e->ev_copied[sl], e->copy_stream));
        inflight[sl] = {lo, cur};
    }
    for (int i = std::max(0, nb - 2); i < nb; ++i) {
        SR_CUDA(cudaEventSynchronize(e->ev_copied[i & 1]));
        unpack(i & 1);
    }
    return check_flag(e, e->stream);
}

int sr_engine_set_option(sr_engine *e, const char *key, int64_t value)
{
    if (!e || !key) return SR_EINVAL;
    ++e->epoch;  // captured passes were planned under the old options
    if (!strcmp(key, "variant")) {
        if (value < -1 || value >= kNumVariants) return fail(e, SR_EINVAL, "variant must be -1 (auto) or in [0, %d)", kNumVariants);
        e->variant = (int)value;
        if (e->d_raw && e->hat_S != kAutoS) {  // (every shape of the table reads the 8-song layout: a 4-song tile is half of one)
            SR_CUDA(cudaSetDevice(e->device));
            SR_CUDA(cudaStreamSynchronize(e->stream));
            int rc = build_store(e);
            if (rc) return rc;
        }
    } else if (!strcmp(key, "qt")) {
        if (value < 1 || value > kQTMax) return fail(e, SR_EINVAL, "qt must be in [1, %d]", kQTMax);
        e->qt_opt = (int)value;
    } else if (!strcmp(key, "batch")) {
        if (value < 1 || value > (1 << 20)) return fail(e, SR_EINVAL, "batch must be in [1, 2^20]");
        e->batch = (int)value;
    } else if (!strcmp(key, "sample")) {
        if (value > kSortCap || (value > 0 && (value & (value - 1))))
            return fail(e, SR_EINVAL, "sample must be 0, negative (auto) or a power of two <= %d", kSortCap);
        e->sample = (int)value;
    } else if (!strcmp(key, "settle_at")) {
        if (value < 0 || value > 1024) return fail(e, SR_EINVAL, "settle_at must be in [0, 1024]");
        e->settle_at = (int)value;
    } else if (!strcmp(key, "trigger_at")) {
        if (value < 0 || value > 1024) return fail(e, SR_EINVAL, "trigger_at must be in [0, 1024]");
        e->trigger_at = (int)value;
    } else if (!strcmp(key, "bound")) {
        e->bound = value != 0;
    } else if (!strcmp(key, "list_ws")) {
        e->list_ws_opt = value != 0;
    } else if (!strcmp(key, "list_ws_kmax")) {
        if (value < 0 || value > kKMax) return fail(e, SR_EINVAL, "list_ws_kmax must be in [0, %d]", kKMax);
        e->list_ws_kmax = (int)value;
    } else if (!strcmp(key, "bound_cap_div")) {
        if (value < 4 || value > 1024) return fail(e, SR_EINVAL, "bound_cap_div must be in [4, 1024]");
        e->bound_cap_div = (int)value;
    } else if (!strcmp(key, "bound_blocks")) {
        if (value != 0 && (value < 2 || value > kLT)) return fail(e, SR_EINVAL, "bound_blocks must be 0 (auto) or in [2, %d]", kLT);
        e->bound_blocks = (int)value;
    } else if (!strcmp(key, "prefetch")) {
        e->prefetch = value != 0;
    } else if (!strcmp(key, "graphs")) {
        e->use_graphs = value != 0;
    } else if (!strcmp(key, "refresh_every")) {
        if (value < 0 || value > 64) return fail(e, SR_EINVAL, "refresh_every must be in [0, 64]");
        e->refresh_every = (int)value;
    } else if (!strcmp(key, "mid_max")) {
        if (value < 0 || value > (1 << 20)) return fail(e, SR_EINVAL, "mid_max must be in [0, 2^20]");
        e->mid_max = (int)value;
    } else if (!strcmp(key, "small_max")) {
        if (value < 0 || value > kConstQueries) return fail(e, SR_EINVAL, "small_max must be in [0, %d]", kConstQueries);
        e->small_max = (int)value;
    } else if (!strcmp(key, "bound_tiles")) {
        if (value != 0 && (value < 8 || value > 1024)) return fail(e, SR_EINVAL, "bound_tiles must be 0 (auto) or in [8, 1024]");
        e->bound_tiles = (int)value;
    } else if (!strcmp(key, "hit_cap")) {
        if (value != 0 && (value < 32 || value > 1024 || value % 32)) return fail(e, SR_EINVAL, "hit_cap must be 0 (auto) or a multiple of 32 in [32, 1024]");
        e->hit_cap = (int)value;
    } else if (!strcmp(key, "profile")) {
        e->profile = value != 0;
    } else if (!strcmp(key, "reset")) {
        SR_CUDA(cudaSetDevice(e->device));
        SR_CUDA(cudaStreamSynchronize(e->stream));
        int rc = resolve_timings(e);
        if (rc) return rc;
        SR_CUDA(cudaMemset(e->d_stats, 0, 128));
        e->launches = e->queries = 0;
        for (int i = 0; i < kNumKernels; ++i) { e->ms_total[i] = 0; e->ms_count[i] = 0; }
    } else {
        return fail(e, SR_EINVAL, "unknown option '%s'", key);
    }
    return SR_OK;
}

int sr_engine_get_stat(sr_engine *e, const char *key, int64_t *value)
{
    if (!e || !key || !value) return SR_EINVAL;
    SR_CUDA(cudaSetDevice(e->device));
    static const char *const dev_keys[] = {"filter_hits", "settles", "rescans", "rescored", "refilters",
                                           "hot_cycles", "settle_cycles", "cta_cycles", "wait_cycles",
                                           "final_settle_cycles", "flush_cycles", "join_cycles", "prologue_cycles"};  // from hot_cycles on: -DSR_SCAN_TIMING builds only
    for (int i = 0; i < 13; ++i) {
        if (!strcmp(key, dev_keys[i])) {
            unsigned long long h[16];
            SR_CUDA(cudaStreamSynchronize(e->stream));
            SR_CUDA(cudaMemcpy(h, e->d_stats, 128, cudaMemcpyDeviceToHost));
            *value = (int64_t)h[i];
            return SR_OK;
        }
    }
    if (!strcmp(key, "kernel_launches")) *value = e->launches;
    else if (!strcmp(key, "queries")) *value = e->queries;
    else if (!strcmp(key, "irregular_songs")) *value = e->irregular;
    else if (!strcmp(key, "sm_count")) *value = e->sm_count;
    else if (!strcmp(key, "scan_grid")) *value = e->scan_grid;
    else if (!strcmp(key, "scan_tile_songs")) *value = kVariants[e->last_variant].S * kVariants[e->last_variant].threads;
    else if (!strcmp(key, "device_bytes")) *value = e->device_bytes;
    else if (!strcmp(key, "variant")) *value = e->last_variant;
    else if (!strcmp(key, "qt")) *value = e->qt_opt;
    else if (!strcmp(key, "lists_in_smem")) *value = e->last_lists_in_smem;
    else if (!strcmp(key, "graph_replays")) *value = e->graph_replays;
    else if (!strcmp(key, "bad_index")) {  // reads (and clears) the sticky flag the device-pointer calls leave behind
        SR_CUDA(cudaMemcpyAsync(e->h_flag, e->d_flag, 4, cudaMemcpyDeviceToHost, e->stream));
        SR_CUDA(cudaStreamSynchronize(e->stream));
        *value = *e->h_flag;
        if (*e->h_flag) SR_CUDA(cudaMemsetAsync(e->d_flag, 0, 4, e->stream));
    }
    else return fail(e, SR_EINVAL, "unknown stat '%s'", key);
    return SR_OK;
}

int sr_engine_get_timing(sr_engine *e, const char *kernel, double *ms_total, int64_t *launches)
{
    if (!e || !kernel || !ms_total) return SR_EINVAL;
    SR_CUDA(cudaSetDevice(e->device));
    int rc = resolve_timings(e);
    if (rc) return rc;
    for (int i = 0; i < kNumKernels; ++i) {
        if (!strcmp(kernel, kKernelNames[i])) {
            *ms_total = e->ms_total[i];
            if (launches) *launches = e->ms_count[i];
            return SR_OK;
        }
    }
    return fail(e, SR_EINVAL, "unknown kernel '%s'", kernel);
}

const char *sr_engine_variant_name(int i) { return (i >= 0 && i < kNumVariants) ? kVariants[i].name : nullptr; }

int sr_engine_measure_fp32(sr_engine *e, int variant, double *tflops)
{
    if (!e || !tflops) return SR_EINVAL;
    if (variant < 0 || variant > 2) return fail(e, SR_EINVAL, "fp32 variant must be 0, 1 or 2");
    SR_CUDA(cudaSetDevice(e->device));
    float *d_out = nullptr;
    SR_CUDA(cudaMalloc(&d_out, 4));
    const int iters = 20000, grid = e->sm_count * 8;
    cudaEvent_t a, b;
    SR_CUDA(cudaEventCreate(&a));
    SR_CUDA(cudaEventCreate(&b));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 warms up
        SR_CUDA(cudaEventRecord(a, e->stream));
        if (variant == 0) fp32_pipe_kernel<0><<<grid, 256, 0, e->stream>>>(d_out, iters, 1e-9f);
        else if (variant == 1) fp32_pipe_kernel<1><<<grid, 256, 0, e->stream>>>(d_out, iters, 1e-9f);
        else fp32_pipe_kernel<2><<<grid, 256, 0, e->stream>>>(d_out, iters, 1e-9f);
        SR_CUDA(cudaGetLastError());
        SR_CUDA(cudaEventRecord(b, e->stream));
        SR_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        SR_CUDA(cudaEventElapsedTime(&ms, a, b));
        if (rep > 0 && ms < best) best = ms;
        ++e->launches;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(d_out);
    // one "step" = 8 accumulators x 12 FMA-equivalents x 2 flop
    *tflops = (double)grid * 256.0 * iters * 12.0 * 8.0 * 2.0 / (best * 1e-3) / 1e12;
    return SR_OK;
}

int sr_engine_selftest_div(sr_engine *e, const float *a, const float *b, int n, float *out)
{
    if (!e || !a || !b || !out || n <= 0) return SR_EINVAL;
    SR_CUDA(cudaSetDevice(e->device));
    float *d = nullptr;
    SR_CUDA(cudaMalloc(&d, (size_t)n * 12));
    cudaError_t err;
    if ((err = cudaMemcpy(d, a, (size_t)n * 4, cudaMemcpyHostToDevice)) == cudaSuccess &&
        (err = cudaMemcpy(d + n, b, (size_t)n * 4, cudaMemcpyHostToDevice)) == cudaSuccess) {
        div_selftest_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(d, d + n, d + 2 * (size_t)n, n);
        if ((err = cudaGetLastError()) == cudaSuccess && (err = cudaStreamSynchronize(e->stream)) == cudaSuccess)
            err = cudaMemcpy(out, d + 2 * (size_t)n, (size_t)n * 4, cudaMemcpyDeviceToHost);
    }
    cudaFree(d);
    if (err != cudaSuccess) return fail(e, SR_ECUDA, "selftest_div: %s", cudaGetErrorString(err));
    return SR_OK;
}

int sr_engine_synchronize(sr_engine *e)
{
    if (!e) return SR_EINVAL;
    SR_CUDA(cudaSetDevice(e->device));
    return check_flag(e, e->stream);
}

}  // extern "C"

#include "sr_sharded.cuh"
