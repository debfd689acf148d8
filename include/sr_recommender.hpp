/*
 * sr_recommender.hpp -- drop-in replacement for the reference's `class Recommender`
 * (Recommender.h:28-82).  The public section is the reference's, member for member;
 * everything behind it is the B200 engine reached through the C ABI of sr_engine.h.
 *
 * To build the reference CLI on this engine, put a one-line `Recommender.h`
 * (`#include "sr_recommender.hpp"`) ahead of the reference directory on the include
 * path and link libsr_recommender.so instead of Recommender.o.
 *
 * Differences that are deliberate (north_star):
 *   - no CPU fallback: initialize() fails without an sm_100 device, and
 *     isGPUEnabled() == isInitialized();
 *   - exact ties are ordered by lower song index (the reference's order on ties is
 *     an artefact of its heap, Recommender.cu:293-315);
 *   - topN <= 0 returns {} (the reference dereferences an empty heap); any other topN yields min(topN, N - 1)
 *     results like the reference (Recommender.cu:300-315);
 *   - more than one GPU: a store of >= 20 M songs, or SR_DEVICES="0,1,..." in the environment, row-shards the songs
 *     over several devices inside this process (the reference is hard-wired to device 0, Recommender.cu:124); the
 *     results do not depend on the number of devices;
 *   - track-id and name lookups use indexes built once in initialize() but keep the
 *     reference's first-match rules (Recommender.cu:320-354).
 */
#ifndef SR_RECOMMENDER_HPP
#define SR_RECOMMENDER_HPP
/* Also claim the reference header's include guard: a translation unit that sees this
 * file first skips the reference's own class declaration (Recommender.h:1-2), which is
 * how the reference's main.cpp is compiled against this class unmodified. */
#ifndef RECOMMENDER_H
#define RECOMMENDER_H
#endif

#include <string>
#include <vector>

#include "sr_song.h"

/* reference Recommender.h:12-22 (result record; operator< orders a min-heap) */
struct Recommendation {
    int songIndex;
    float similarity;
    Recommendation() : songIndex(-1), similarity(0.0f) {}
    Recommendation(int idx, float sim) : songIndex(idx), similarity(sim) {}
    bool operator<(const Recommendation &other) const { return similarity > other.similarity; }
};

class Recommender {
public:
    Recommender();
    ~Recommender();
    Recommender(const Recommender &) = delete;
    Recommender &operator=(const Recommender &) = delete;

    /* reference Recommender.h:40 / Recommender.cu:100-182 */
    bool initialize(const std::vector<Song> &songs);
    /* reference Recommender.h:49 / Recommender.cu:356-363 */
    std::vector<int> recommend(const std::string &trackId, int topN);
    /* reference Recommender.h:58 / Recommender.cu:365-372 */
    std::vector<int> recommendByName(const std::string &trackName, int topN);
    /* reference Recommender.h:67 / Recommender.cu:275-318 */
    std::vector<int> recommendByIndex(int songIndex, int topN);

    bool isInitialized() const { return initialized; }
    bool isGPUEnabled() const { return gpuEnabled; }
    int getSongCount() const { return numSongs; }

    /* ---- additions (not in the reference) ------------------------------------- */
    /* Same scoring for a batch of query songs in one pass over the store; scores are
     * the reference's similarity values (Recommender.cu:256-273), bit for bit. */
    std::vector<std::vector<Recommendation> > recommendBatch(const std::vector<int> &songIndices, int topN);
    /* Dense entry point for embedders that do not hold a vector<Song>. */
    bool initializeDense(const float *features, long long count);
    /* Load path without the vector<Song> detour: reads a songs_data.bin
     * written by the reference's DataManager::preprocessData (DataManager.cpp:321-342,
     * Song.h:35-54 -- native size_t / int / float, no magic) straight into the dense
     * feature matrix and the lookup indexes.  Unlike DataManager::loadData
     * (DataManager.cpp:363-409) every length is bounds-checked; a truncated or
     * implausible file fails with a message instead of reading garbage. */
    bool initializeFromFile(const std::string &binaryPath);
    /* Lookups as the reference resolves them; -1 when absent. */
    int findSongByTrackId(const std::string &trackId) const;
    int findSongByName(const std::string &trackName) const;

private:
    struct Impl;
    bool initialized;
    int numSongs;
    bool gpuEnabled;
    Impl *impl;
};

#endif /* SR_RECOMMENDER_HPP */
