// ref_harness.cpp -- TEST INFRASTRUCTURE.  A C-ABI driver around the UNMODIFIED
// reference `Recommender` class (/root/reference/Recommender.{h,cu}), compiled by
// oracle/Makefile straight from the read-only reference tree into oracle/_ref/.
// It exists so tests and bench.py can (1) pin the C oracle against the real
// reference implementation and (2) time the reference's own code path.  Nothing
// here is reference source: it only calls the reference's public API, plus the
// private `calculateSimilarities` (Recommender.h:114) to read raw scores, which
// the public API never returns.
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

// std headers are all in before access control is opened for the one class.
#define private public
#include "Recommender.h"
#undef private

namespace {
struct Quiet {  // the reference prints banners from initialize(); keep logs clean
    std::streambuf *old;
    std::ostringstream sink;
    Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

// Build a reference Recommender over a dense row-major n x 12 matrix.
// Song strings are synthetic ("id<i>", "Track <i>").  Returns nullptr on failure.
void *ref_create(const float *features, int64_t n)
{
    std::vector<Song> songs((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        Song &s = songs[(size_t)i];
        s.track_id = "id" + std::to_string(i);
        s.track_name = "Track " + std::to_string(i);
        s.artists = "Artist";
        s.genre_id = 0;
        std::memcpy(s.features, features + i * FEATURE_COUNT, sizeof(s.features));
    }
    Recommender *r = new Recommender();
    Quiet q;
    if (!r->initialize(songs)) {
        delete r;
        return nullptr;
    }
    return r;
}

void ref_destroy(void *h) { delete static_cast<Recommender *>(h); }

int ref_gpu_enabled(void *h) { return static_cast<Recommender *>(h)->isGPUEnabled() ? 1 : 0; }

int64_t ref_song_count(void *h) { return static_cast<Recommender *>(h)->getSongCount(); }

// reference recommendByIndex (Recommender.cu:275-318); returns result count.
int ref_recommend_by_index(void *h, int idx, int k, int32_t *out)
{
    std::vector<int> r = static_cast<Recommender *>(h)->recommendByIndex(idx, k);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}

int ref_recommend_by_name(void *h, const char *name, int k, int32_t *out)
{
    std::vector<int> r = static_cast<Recommender *>(h)->recommendByName(name, k);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}

int ref_recommend_by_id(void *h, const char *id, int k, int32_t *out)
{
    std::vector<int> r = static_cast<Recommender *>(h)->recommend(id, k);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}

// Raw similarity scores of one query (CPU build: Recommender.cu:256-273;
// GPU build: cuBLAS SGEMV + the two kernels, :184-254).
void ref_scores(void *h, int idx, float *out)
{
    static_cast<Recommender *>(h)->calculateSimilarities(idx, out);
}

// A batch the way the reference would serve it: Q sequential recommendByIndex
// calls.  threads > 1 runs calls concurrently (only valid on the CPU build,
// whose per-query state is call-local).  out: nq x k, -1 padded.
void ref_batch_by_index(void *h, const int32_t *qidx, int nq, int k, int32_t *out, int threads)
{
    Recommender *r = static_cast<Recommender *>(h);
    if (threads > 1 && r->isGPUEnabled()) threads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1)
#endif
    for (int qi = 0; qi < nq; ++qi) {
        std::vector<int> res = r->recommendByIndex(qidx[qi], k);
        for (int j = 0; j < k; ++j) out[(size_t)qi * k + j] = j < (int)res.size() ? res[j] : -1;
    }
}

// The survey's D1/D2 generator: std::mt19937(seed) +
// std::uniform_real_distribution<float>(0,1), filled song-major.
void ref_gen_mt19937_uniform(int64_t count, uint32_t seed, float *out)
{
    std::mt19937 gen(seed);
    std::uniform_real_distribution<float> dist(0.0f, 1.0f);
    for (int64_t i = 0; i < count; ++i) out[i] = dist(gen);
}

int ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
