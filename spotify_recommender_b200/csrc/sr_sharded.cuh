// sr_sharded.cuh -- several engines behind one handle, in ONE process (SURVEY 8e behind the C ABI, so that the
// reference's C++ host -- main.cpp:59-82 through class Recommender -- reaches every GPU of the box).
// Included at the end of sr_engine.cu (same translation unit: it works on the engines' own buffers).
//
//   row-sharded store (north_star (4)): shard s owns rows [s * per, (s + 1) * per) and answers with global ids.
//   Per batch and per shard, on the shard's own stream: the query rows are read from their owners' raw stores
//   over NVLink peer access (no exchange step for the queries); the threshold bound pass is SHARED -- every shard
//   samples 1/G of the tiles, the block maxima are max-reduced by a kernel that reads the other shards' arrays over
//   peer access, every shard scans with thresholds that bound the K-th best of the whole store; the fused scan
//   produces the shard's candidates as packed 64-bit keys; the root shard's merge kernel then reads the G key lists straight out of the shards'
//   buffers over peer access -- gather and merge are ONE kernel, nothing is staged -- and the merged rows go to
//   the host.  Cross-device ordering is by CUDA events only.  (The multi-PROCESS host, one rank per GPU under
//   torchrun, exchanges the same keys with one NCCL all-gather: spotify_recommender_b200/sharded.py.)
//
//   replicated store (BASELINE config 5, all-pairs): every shard holds all rows and serves a contiguous slice of
//   the QUERIES; no exchange at all, each shard writes its slice of the caller's table.
//
// `devices` may name a GPU more than once: several shards then share it (peer access to oneself is trivially
// there), which is how the whole path is tested on a single-GPU box.
#pragma once

#include <thread>

struct sr_sharded {
    std::vector<sr_engine *> eng;
    std::vector<int> dev;
    std::string err;
    int64_t n = 0, per = 0;
    bool replicated = false;
    int batch = 8192;
    // per shard
    struct Shard {
        int32_t *d_q = nullptr;        // [batch] query ids
        float *d_qrows = nullptr;      // [batch][12]
        uint64_t *d_keys = nullptr;    // [batch][kKMax-capped k] local top-K keys
        size_t keys_cap = 0;
        float *d_blocks = nullptr;     // [batch][256] this shard's block maxima of the shared bound pass ...
        float *d_blocks_max = nullptr; // ... and their maximum over all shards (read over peer access)
        const float **d_block_ptrs = nullptr;  // [G] every shard's d_blocks, on this shard's device
        cudaEvent_t ev_blocks = nullptr;       // this shard's block maxima are complete
        const float **d_raw_ptrs = nullptr;  // [G] raw stores of every shard (peer pointers), on this shard's device
        cudaEvent_t ev_keys = nullptr;       // this shard's keys of the current pass are complete
    };
    std::vector<Shard> sh;
    // root (shard 0)
    const uint64_t **d_part_ptrs = nullptr;  // [G] the shards' key buffers
    uint64_t *d_ceil = nullptr;              // [batch] ceilings of a k > kKMax query (read by every shard)
    char *d_out = nullptr;                   // merged rows: idx then scores
    size_t out_cap = 0;
    char *h_pin = nullptr;
    size_t pin_cap = 0;
    cudaEvent_t ev_merged = nullptr;         // the root's merge has consumed every shard's keys
};

namespace {

thread_local std::string g_sharded_create_error;

int sfail(sr_sharded *s, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (s) s->err = buf; else g_sharded_create_error = buf;
    return code;
}

#define SRS_CUDA(call)                                                                                        \
    do {                                                                                                      \
        cudaError_t err_ = (call);                                                                            \
        if (err_ != cudaSuccess)                                                                              \
            return sfail(s, err_ == cudaErrorMemoryAllocation ? SR_ENOMEM : SR_ECUDA, "%s failed: %s (%s:%d)", \
                         #call, cudaGetErrorString(err_), __FILE__, __LINE__);                                \
    } while (0)
#define SRS_ENGINE(g, call)                                                                       \
    do {                                                                                          \
        int rc_ = (call);                                                                         \
        if (rc_) return sfail(s, rc_, "shard %d: %s", (g), sr_engine_last_error(s->eng[(g)]));     \
    } while (0)

void sharded_free_buffers(sr_sharded *s)
{
    for (size_t g = 0; g < s->sh.size(); ++g) {
        cudaSetDevice(s->dev[g]);
        sr_sharded::Shard &h = s->sh[g];
        if (h.d_q) cudaFree(h.d_q);
        if (h.d_qrows) cudaFree(h.d_qrows);
        if (h.d_keys) cudaFree(h.d_keys);
        if (h.d_raw_ptrs) cudaFree((void *)h.d_raw_ptrs);
        if (h.d_blocks) cudaFree(h.d_blocks);
        if (h.d_blocks_max) cudaFree(h.d_blocks_max);
        if (h.d_block_ptrs) cudaFree((void *)h.d_block_ptrs);
        h.d_q = nullptr; h.d_qrows = nullptr; h.d_keys = nullptr; h.d_raw_ptrs = nullptr; h.keys_cap = 0;
        h.d_blocks = h.d_blocks_max = nullptr; h.d_block_ptrs = nullptr;
    }
}

int sharded_ensure_keys(sr_sharded *s, int kc)
{
    const size_t want = (size_t)s->batch * kc * 8;
    const int G = (int)s->eng.size();
    bool grew = false;
    for (int g = 0; g < G; ++g) {
        sr_sharded::Shard &h = s->sh[g];
        if (h.keys_cap >= want) continue;
        SRS_CUDA(cudaSetDevice(s->dev[g]));
        SRS_CUDA(cudaDeviceSynchronize());
        if (h.d_keys) SRS_CUDA(cudaFree(h.d_keys));
        h.d_keys = nullptr; h.keys_cap = 0;
        SRS_CUDA(cudaMalloc(&h.d_keys, want));
        h.keys_cap = want;
        grew = true;
    }
    if (grew) {
        std::vector<const uint64_t *> ptrs(G);
        for (int g = 0; g < G; ++g) ptrs[g] = s->sh[g].d_keys;
        SRS_CUDA(cudaSetDevice(s->dev[0]));
        SRS_CUDA(cudaMemcpy((void *)s->d_part_ptrs, ptrs.data(), G * sizeof(void *), cudaMemcpyHostToDevice));
    }
    return SR_OK;
}

}  // namespace

extern "C" {

int sr_sharded_create(sr_sharded **out, const int *devices, int n_devices)
{
    sr_sharded *s = nullptr;
    if (!out) return sfail(nullptr, SR_EINVAL, "sharded create: null output pointer");
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        return sfail(nullptr, SR_ENODEVICE, "no CUDA device (this engine has no CPU fallback)");
    std::vector<int> dev;
    if (devices && n_devices > 0) dev.assign(devices, devices + n_devices);
    else for (int d = 0; d < count; ++d) dev.push_back(d);  // every visible GPU
    if (dev.size() > 64) return sfail(nullptr, SR_EINVAL, "at most 64 shards");
    s = new sr_sharded();
    s->dev = dev;
    const int G = (int)dev.size();
    for (int g = 0; g < G; ++g) {
        sr_engine *e = nullptr;
        int rc = sr_engine_create(&e, dev[g]);
        if (rc) {
            sfail(nullptr, rc, "shard %d (device %d): %s", g, dev[g], sr_engine_last_error(nullptr));
            sr_sharded_destroy(s);
            return rc;
        }
        s->eng.push_back(e);
    }
    s->sh.resize(G);
    // peer access between every pair of distinct devices (NVLink / NVSwitch on the 8 x B200 box)
    for (int a = 0; a < G; ++a) {
        for (int b = 0; b < G; ++b) {
            if (dev[a] == dev[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, dev[a], dev[b]);
            if (!can) {
                sfail(nullptr, SR_ENODEVICE, "device %d cannot access device %d's memory (no NVLink / PCIe peer path): the single-process "
                      "multi-GPU host needs peer access", dev[a], dev[b]);
                sr_sharded_destroy(s);
                return SR_ENODEVICE;
            }
            cudaSetDevice(dev[a]);
            cudaError_t err = cudaDeviceEnablePeerAccess(dev[b], 0);
            if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) {
                sfail(nullptr, SR_ECUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dev[a], dev[b], cudaGetErrorString(err));
                sr_sharded_destroy(s);
                return SR_ECUDA;
            }
            cudaGetLastError();
        }
    }
    cudaError_t err = cudaSuccess;
    for (int g = 0; g < G && err == cudaSuccess; ++g) {
        cudaSetDevice(dev[g]);
        sr_sharded::Shard &h = s->sh[g];
        if ((err = cudaMalloc(&h.d_q, (size_t)s->batch * 4)) != cudaSuccess) break;
        if ((err = cudaMalloc(&h.d_qrows, (size_t)s->batch * kF * 4)) != cudaSuccess) break;
        if ((err = cudaMalloc((void **)&h.d_raw_ptrs, G * sizeof(void *))) != cudaSuccess) break;
        if ((err = cudaMalloc(&h.d_blocks, (size_t)s->batch * kLT * 4)) != cudaSuccess) break;
        if ((err = cudaMalloc(&h.d_blocks_max, (size_t)s->batch * kLT * 4)) != cudaSuccess) break;
        if ((err = cudaMalloc((void **)&h.d_block_ptrs, G * sizeof(void *))) != cudaSuccess) break;
        if ((err = cudaEventCreateWithFlags(&h.ev_blocks, cudaEventDisableTiming)) != cudaSuccess) break;
        err = cudaEventCreateWithFlags(&h.ev_keys, cudaEventDisableTiming);
    }
    if (err == cudaSuccess) {
        std::vector<const float *> bp(G);
        for (int g = 0; g < G; ++g) bp[g] = s->sh[g].d_blocks;
        for (int g = 0; g < G && err == cudaSuccess; ++g) {
            cudaSetDevice(dev[g]);
            err = cudaMemcpy((void *)s->sh[g].d_block_ptrs, bp.data(), G * sizeof(void *), cudaMemcpyHostToDevice);
        }
    }
    if (err == cudaSuccess) {
        cudaSetDevice(dev[0]);
        if ((err = cudaMalloc((void **)&s->d_part_ptrs, G * sizeof(void *))) == cudaSuccess &&
            (err = cudaMalloc(&s->d_ceil, (size_t)s->batch * 8)) == cudaSuccess)
            err = cudaEventCreateWithFlags(&s->ev_merged, cudaEventDisableTiming);
    }
    if (err != cudaSuccess) {
        sfail(nullptr, SR_ECUDA, "sharded setup: %s", cudaGetErrorString(err));
        sr_sharded_destroy(s);
        return SR_ECUDA;
    }
    *out = s;
    return SR_OK;
}

void sr_sharded_destroy(sr_sharded *s)
{
    if (!s) return;
    for (size_t g = 0; g < s->eng.size(); ++g) {
        cudaSetDevice(s->dev[g]);
        cudaDeviceSynchronize();
    }
    sharded_free_buffers(s);
    for (size_t g = 0; g < s->sh.size(); ++g) {
        cudaSetDevice(s->dev[g]);
        if (s->sh[g].ev_keys) cudaEventDestroy(s->sh[g].ev_keys);
        if (s->sh[g].ev_blocks) cudaEventDestroy(s->sh[g].ev_blocks);
    }
    if (!s->dev.empty()) cudaSetDevice(s->dev[0]);
    if (s->d_part_ptrs) cudaFree((void *)s->d_part_ptrs);
    if (s->d_ceil) cudaFree(s->d_ceil);
    if (s->d_out) cudaFree(s->d_out);
    if (s->h_pin) cudaFreeHost(s->h_pin);
    if (s->ev_merged) cudaEventDestroy(s->ev_merged);
    for (sr_engine *e : s->eng) sr_engine_destroy(e);
    delete s;
}

const char *sr_sharded_last_error(const sr_sharded *s) { return s ? s->err.c_str() : g_sharded_create_error.c_str(); }
int64_t sr_sharded_song_count(const sr_sharded *s) { return s ? s->n : 0; }
int sr_sharded_shard_count(const sr_sharded *s) { return s ? (int)s->eng.size() : 0; }
sr_engine *sr_sharded_engine(sr_sharded *s, int i) { return (s && i >= 0 && i < (int)s->eng.size()) ? s->eng[i] : nullptr; }

int sr_sharded_load_features(sr_sharded *s, const float *rows, int64_t n, int replicate)
{
    if (!s) return SR_EINVAL;
    if (!rows || n <= 0) return sfail(s, SR_EINVAL, "load_features: null rows or n <= 0");
    if (n > 0x7fffffffLL) return sfail(s, SR_EINVAL, "load_features: %lld songs do not fit 32-bit ids", (long long)n);
    const int G = (int)s->eng.size();
    s->n = 0;
    s->replicated = replicate != 0;
    s->per = s->replicated ? n : (n + G - 1) / G;
    if (!s->replicated && s->per * (G - 1) >= n && G > 1)
        return sfail(s, SR_EINVAL, "load_features: %lld songs are too few for %d row shards", (long long)n, G);
    for (int g = 0; g < G; ++g) {
        const int64_t lo = s->replicated ? 0 : g * s->per;
        const int64_t cnt = s->replicated ? n : std::min(n, lo + s->per) - lo;
        SRS_ENGINE(g, sr_engine_load_features(s->eng[g], rows + lo * kF, cnt, lo));
    }
    std::vector<const float *> raw(G);
    for (int g = 0; g < G; ++g) raw[g] = s->eng[g]->d_raw;
    for (int g = 0; g < G; ++g) {
        SRS_CUDA(cudaSetDevice(s->dev[g]));
        SRS_CUDA(cudaMemcpy((void *)s->sh[g].d_raw_ptrs, raw.data(), G * sizeof(void *), cudaMemcpyHostToDevice));
    }
    s->n = n;
    return SR_OK;
}

int sr_sharded_query_by_index(sr_sharded *s, const int32_t *qidx, int nq, int k, int32_t *out_idx, float *out_score)
{
    if (!s) return SR_EINVAL;
    if (!s->n) return sfail(s, SR_ESTATE, "no store loaded: call sr_sharded_load_features first");
    if (!qidx || !out_idx) return sfail(s, SR_EINVAL, "null query or output pointer");
    if (nq <= 0 || k <= 0) return sfail(s, SR_EINVAL, "nq and k must be positive (got %d, %d)", nq, k);
    for (int i = 0; i < nq; ++i)
        if (qidx[i] < 0 || qidx[i] >= s->n) return sfail(s, SR_EINVAL, "query %d: song id %d is not in [0, %lld)", i, qidx[i], (long long)s->n);
    const int G = (int)s->eng.size();
    if (s->replicated) {
        // every shard holds the whole store: split the QUERIES, one host thread per shard, no exchange
        const int per_q = (nq + G - 1) / G;
        std::vector<int> rcs(G, SR_OK);
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) {
            const int lo = std::min(nq, g * per_q), cnt = std::min(nq, lo + per_q) - lo;
            if (cnt <= 0) continue;
            th.emplace_back([=, &rcs] {
                rcs[g] = sr_engine_query_by_index(s->eng[g], qidx + lo, cnt, k, out_idx + (size_t)lo * k,
                                                  out_score ? out_score + (size_t)lo * k : nullptr);
            });
        }
        for (auto &t : th) t.join();
        for (int g = 0; g < G; ++g)
            if (rcs[g]) return sfail(s, rcs[g], "shard %d: %s", g, sr_engine_last_error(s->eng[g]));
        return SR_OK;
    }
    const int kc_max = std::min(k, kKMax);
    int rc = sharded_ensure_keys(s, kc_max);
    if (rc) return rc;
    const int B = s->batch;
    const size_t row_bytes = (size_t)k * (out_score ? 8 : 4);
    const size_t need = (size_t)std::min(nq, B) * row_bytes;
    SRS_CUDA(cudaSetDevice(s->dev[0]));
    if (need > s->out_cap) {
        SRS_CUDA(cudaDeviceSynchronize());
        if (s->d_out) SRS_CUDA(cudaFree(s->d_out));
        s->d_out = nullptr; s->out_cap = 0;
        SRS_CUDA(cudaMalloc(&s->d_out, need));
        s->out_cap = need;
    }
    const size_t pin_need = (size_t)B * 4 + need;
    if (pin_need > s->pin_cap) {
        SRS_CUDA(cudaDeviceSynchronize());
        if (s->h_pin) SRS_CUDA(cudaFreeHost(s->h_pin));
        s->h_pin = nullptr; s->pin_cap = 0;
        SRS_CUDA(cudaMallocHost(&s->h_pin, pin_need));
        s->pin_cap = pin_need;
    }
    const int chunks = (k + kKMax - 1) / kKMax;
    sr_engine *root = s->eng[0];
    for (int done = 0; done < nq; done += B) {
        const int cur = std::min(B, nq - done);
        memcpy(s->h_pin, qidx + done, (size_t)cur * 4);
        int32_t *d_oi = (int32_t *)s->d_out;
        float *d_os = out_score ? (float *)(s->d_out + (size_t)cur * k * 4) : nullptr;
        for (int c = 0; c < chunks; ++c) {
            const int kc = std::min(kKMax, k - c * kKMax);
            // the bound pass is shared between the shards (first pass of a list only: ceilings have none): every shard
            // samples 1/G of the tiles, the block maxima are max-reduced over peer access, every shard scans with
            // thresholds that bound the k-th best of the WHOLE store
            const int nblk = (c == 0 && G > 1) ? sr_engine_bound_block_count(root, kc) : 0;
            for (int g = 0; g < G; ++g) {
                sr_engine *e = s->eng[g];
                sr_sharded::Shard &h = s->sh[g];
                SRS_CUDA(cudaSetDevice(s->dev[g]));
                SRS_CUDA(cudaStreamWaitEvent(e->stream, s->ev_merged, 0));  // the previous merge has read this shard's keys (and written the ceilings)
                if (c == 0) {
                    SRS_CUDA(cudaMemcpyAsync(h.d_q, s->h_pin, (size_t)cur * 4, cudaMemcpyHostToDevice, e->stream));
                    gather_rows_p2p_kernel<<<(cur + 127) / 128, 128, 0, e->stream>>>(h.d_raw_ptrs, s->per, s->n, h.d_q, cur, h.d_qrows);
                    SRS_CUDA(cudaGetLastError());
                    ++e->launches;
                }
                if (nblk) {
                    SRS_ENGINE(g, sr_engine_bound_blocks_dev(e, h.d_qrows, cur, kc, G, h.d_blocks, SR_ENGINE_OWN_STREAM));
                    SRS_CUDA(cudaEventRecord(h.ev_blocks, e->stream));
                }
            }
            for (int g = 0; g < G; ++g) {
                sr_engine *e = s->eng[g];
                sr_sharded::Shard &h = s->sh[g];
                SRS_CUDA(cudaSetDevice(s->dev[g]));
                if (nblk) {
                    for (int o = 0; o < G; ++o)
                        if (o != g) SRS_CUDA(cudaStreamWaitEvent(e->stream, s->sh[o].ev_blocks, 0));
                    const int64_t count = (int64_t)cur * nblk;
                    blocks_max_p2p_kernel<<<(unsigned)((count + 255) / 256), 256, 0, e->stream>>>(h.d_block_ptrs, G, count, h.d_blocks_max);
                    SRS_CUDA(cudaGetLastError());
                    ++e->launches;
                }
                SRS_ENGINE(g, sr_engine_query_keys_by_vector_dev(e, h.d_qrows, h.d_q, cur, kc, c > 0 ? s->d_ceil : nullptr,
                                                                 nblk ? h.d_blocks_max : nullptr, h.d_keys, SR_ENGINE_OWN_STREAM));
                SRS_CUDA(cudaEventRecord(h.ev_keys, e->stream));
            }
            SRS_CUDA(cudaSetDevice(s->dev[0]));
            for (int g = 1; g < G; ++g) SRS_CUDA(cudaStreamWaitEvent(root->stream, s->sh[g].ev_keys, 0));
            rc = merge_common(root, nullptr, s->d_part_ptrs, nullptr, nullptr, G, cur, kc, d_oi, d_os, k, c * kKMax,
                              c + 1 < chunks ? s->d_ceil : nullptr, SR_ENGINE_OWN_STREAM);
            if (rc) return sfail(s, rc, "merge: %s", sr_engine_last_error(root));
            SRS_CUDA(cudaEventRecord(s->ev_merged, root->stream));
        }
        char *pout = s->h_pin + (size_t)B * 4;
        SRS_CUDA(cudaMemcpyAsync(pout, s->d_out, (size_t)cur * row_bytes, cudaMemcpyDeviceToHost, root->stream));
        SRS_CUDA(cudaStreamSynchronize(root->stream));
        memcpy(out_idx + (size_t)done * k, pout, (size_t)cur * k * 4);
        if (out_score) memcpy(out_score + (size_t)done * k, pout + (size_t)cur * k * 4, (size_t)cur * k * 4);
    }
    for (int g = 0; g < G; ++g) SRS_ENGINE(g, sr_engine_synchronize(s->eng[g]));
    return SR_OK;
}

int sr_sharded_all_pairs_topk(sr_sharded *s, int k, int32_t *out_idx, float *out_score)
{
    if (!s) return SR_EINVAL;
    if (!s->n) return sfail(s, SR_ESTATE, "no store loaded: call sr_sharded_load_features first");
    if (!out_idx || k <= 0) return sfail(s, SR_EINVAL, "all_pairs: null output or k <= 0");
    const int G = (int)s->eng.size();
    if (!s->replicated) {
        // row shards: every batch of queries visits every shard (8 x the exchange of the replicated form)
        const int64_t B = 1 << 16;
        std::vector<int32_t> q((size_t)std::min<int64_t>(B, s->n));
        for (int64_t lo = 0; lo < s->n; lo += B) {
            const int cur = (int)std::min<int64_t>(B, s->n - lo);
            for (int i = 0; i < cur; ++i) q[i] = (int32_t)(lo + i);
            int rc = sr_sharded_query_by_index(s, q.data(), cur, k, out_idx + (size_t)lo * k, out_score ? out_score + (size_t)lo * k : nullptr);
            if (rc) return rc;
        }
        return SR_OK;
    }
    const int64_t per_q = (s->n + G - 1) / G;
    std::vector<int> rcs(G, SR_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        const int64_t lo = std::min<int64_t>(s->n, g * per_q), hi = std::min<int64_t>(s->n, lo + per_q);
        if (hi <= lo) continue;
        th.emplace_back([=, &rcs] {
            rcs[g] = sr_engine_all_pairs_topk(s->eng[g], lo, hi, k, out_idx + (size_t)lo * k, out_score ? out_score + (size_t)lo * k : nullptr);
        });
    }
    for (auto &t : th) t.join();
    for (int g = 0; g < G; ++g)
        if (rcs[g]) return sfail(s, rcs[g], "shard %d: %s", g, sr_engine_last_error(s->eng[g]));
    return SR_OK;
}

}  // extern "C"
