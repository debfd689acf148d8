// recommender_host.cpp -- the C++ host class of include/sr_recommender.hpp on top of
// the C ABI (include/sr_engine.h), plus a small C API around it for FFI users and tests.
// Mirrors the reference's operator interface for the hot path: same names, argument
// meaning and error behaviour (messages on std::cerr, empty vector / false on failure;
// reference Recommender.cu:100-107, :276-284, :356-372).
#include "sr_recommender.hpp"

#include <algorithm>
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <unordered_map>

#include "sr_engine.h"

struct Recommender::Impl {
    sr_engine *engine = nullptr;     // one GPU ...
    sr_sharded *sharded = nullptr;   // ... or a row-sharded store over several
    // query resolution indexes: first occurrence wins, as the reference's
    // linear scans return the first match (Recommender.cu:320-327, :336-354)
    std::unordered_map<std::string, int> by_id;
    std::unordered_map<std::string, int> by_lower_name;
    std::string lower_blob;            // all lower-cased names, '\0'-separated (a '\0' inside a name is stored as '\1')
    std::vector<size_t> lower_off;     // start of name i in lower_blob (size n+1)
    void reset_index(size_t n)
    {
        by_id.clear();
        by_lower_name.clear();
        lower_blob.clear();
        lower_off.assign(1, (size_t)0);
        by_id.reserve(n * 2);
        by_lower_name.reserve(n * 2);
    }
    void index_song(size_t i, const std::string &id, const std::string &name);
};

namespace {
std::string lower(const std::string &s)
{
    std::string r = s;
    std::transform(r.begin(), r.end(), r.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    return r;
}
// Which GPUs serve the store.  SR_DEVICES="0,1,2,3" (ordinals, may repeat) decides; without it a store of at
// least 20 M songs is row-sharded over every visible device and anything smaller stays on the current one.
std::vector<int> pick_devices(long long count)
{
    std::vector<int> dev;
    if (const char *env = std::getenv("SR_DEVICES")) {
        for (const char *p = env; *p;) {
            char *end = nullptr;
            long v = std::strtol(p, &end, 10);
            if (end == p) break;
            dev.push_back((int)v);
            p = (*end == ',') ? end + 1 : end;
        }
        return dev;
    }
    if (count >= 20000000LL) return dev;  // empty = "every visible device", resolved by sr_sharded_create
    dev.push_back(-1);                    // the current device
    return dev;
}
}  // namespace

void Recommender::Impl::index_song(size_t i, const std::string &id, const std::string &name)
{
    by_id.emplace(id, (int)i);  // emplace keeps the first
    std::string l = lower(name);
    by_lower_name.emplace(l, (int)i);
    for (char &ch : l)
        if (ch == '\0') ch = '\1';  // keep "a hit lies inside one name" true for the substring scan
    lower_blob.append(l);
    lower_blob.push_back('\0');
    lower_off.push_back(lower_blob.size());
}

Recommender::Recommender() : initialized(false), numSongs(0), gpuEnabled(false), impl(new Impl()) {}

Recommender::~Recommender()
{
    if (impl) {
        if (impl->engine) sr_engine_destroy(impl->engine);
        if (impl->sharded) sr_sharded_destroy(impl->sharded);
        delete impl;
    }
}

bool Recommender::initializeDense(const float *features, long long count)
{
    if (!features || count <= 0) {
        std::cerr << "Error: Cannot initialize with empty song database" << std::endl;
        return false;
    }
    if (count > 0x7fffffffLL) {
        std::cerr << "Error: too many songs for 32-bit song indices: " << count << std::endl;
        return false;
    }
    const std::vector<int> dev = pick_devices(count);
    const bool want_sharded = dev.size() != 1;
    if (impl->engine && want_sharded) { sr_engine_destroy(impl->engine); impl->engine = nullptr; }
    if (impl->sharded && !want_sharded) { sr_sharded_destroy(impl->sharded); impl->sharded = nullptr; }
    if (want_sharded) {
        if (!impl->sharded && sr_sharded_create(&impl->sharded, dev.empty() ? nullptr : dev.data(), (int)dev.size()) != SR_OK) {
            std::cerr << "Error: GPU engines unavailable: " << sr_sharded_last_error(nullptr)
                      << " (no CPU fallback in this build)" << std::endl;
            impl->sharded = nullptr;
            return false;
        }
        if (sr_sharded_load_features(impl->sharded, features, count, 0) != SR_OK) {
            std::cerr << "Error: failed to upload features: " << sr_sharded_last_error(impl->sharded) << std::endl;
            return false;
        }
    } else {
        if (!impl->engine && sr_engine_create(&impl->engine, dev[0]) != SR_OK) {
            std::cerr << "Error: GPU engine unavailable: " << sr_engine_last_error(nullptr)
                      << " (no CPU fallback in this build)" << std::endl;
            impl->engine = nullptr;
            return false;
        }
        if (sr_engine_load_features(impl->engine, features, count, 0) != SR_OK) {
            std::cerr << "Error: failed to upload features: " << sr_engine_last_error(impl->engine) << std::endl;
            return false;
        }
    }
    numSongs = (int)count;
    initialized = true;
    gpuEnabled = true;
    return true;
}

namespace {
// bounds-checked cursor over the file image
struct Cursor {
    const unsigned char *p, *end;
    bool ok = true;
    bool take(void *dst, size_t n)
    {
        if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
        std::memcpy(dst, p, n);
        p += n;
        return true;
    }
    bool take_string(std::string *s, size_t limit)
    {
        uint64_t len = 0;
        if (!take(&len, sizeof len) || len > limit || (uint64_t)(end - p) < len) { ok = false; return false; }
        if (s) s->assign(reinterpret_cast<const char *>(p), (size_t)len);
        p += len;
        return true;
    }
};
}  // namespace

bool Recommender::initializeFromFile(const std::string &binaryPath)
{
    std::cout << "Initializing recommender (B200 engine) from " << binaryPath << "..." << std::endl;
    initialized = false;
    gpuEnabled = false;
    std::ifstream in(binaryPath, std::ios::binary | std::ios::ate);
    if (!in.is_open()) {
        std::cerr << "Error: Could not open binary file: " << binaryPath << std::endl;
        return false;
    }
    const std::streamoff size = in.tellg();
    std::vector<unsigned char> image((size_t)std::max<std::streamoff>(size, 0));
    in.seekg(0);
    if (size <= 0 || !in.read(reinterpret_cast<char *>(image.data()), size)) {
        std::cerr << "Error: Could not read binary file: " << binaryPath << std::endl;
        return false;
    }
    Cursor cur{image.data(), image.data() + image.size()};
    uint64_t numSongs = 0, numGenres = 0;
    cur.take(&numSongs, sizeof numSongs);   // DataManager.cpp:321-323
    cur.take(&numGenres, sizeof numGenres); // DataManager.cpp:325-327
    // every song takes at least 3 lengths + genre id + 12 floats = 76 bytes
    if (!cur.ok || numSongs == 0 || numSongs > 0x7fffffffULL || numSongs > image.size() / 76 || numGenres > image.size() / 12) {
        std::cerr << "Error: " << binaryPath << " is not a songs_data.bin (implausible header)" << std::endl;
        return false;
    }
    for (uint64_t g = 0; g < numGenres && cur.ok; ++g) {  // genre table (DataManager.cpp:329-337): skipped
        int32_t id;
        cur.take(&id, sizeof id);
        cur.take_string(nullptr, 1 << 20);
    }
    const size_t n = (size_t)numSongs;
    std::vector<float> dense(n * FEATURE_COUNT);
    impl->reset_index(n);
    std::string id, name;
    for (size_t i = 0; i < n && cur.ok; ++i) {  // Song::serialize order (Song.h:35-54)
        int32_t genre;
        cur.take_string(&id, 1 << 20);
        cur.take_string(&name, 1 << 20);
        cur.take_string(nullptr, 1 << 20);  // artists: not needed by the scoring engine
        cur.take(&genre, sizeof genre);
        cur.take(&dense[i * FEATURE_COUNT], sizeof(float) * FEATURE_COUNT);
        if (!cur.ok) break;
        impl->index_song(i, id, name);
    }
    if (!cur.ok) {
        std::cerr << "Error: " << binaryPath << " is truncated or corrupt" << std::endl;
        return false;
    }
    if (!initializeDense(dense.data(), (long long)n)) return false;
    std::cout << "Recommender initialized on GPU: " << numSongs << " songs resident" << std::endl;
    return true;
}

bool Recommender::initialize(const std::vector<Song> &songs)
{
    std::cout << "Initializing recommender (B200 engine)..." << std::endl;
    if (songs.empty()) {
        std::cerr << "Error: Cannot initialize with empty song database" << std::endl;
        return false;
    }
    initialized = false;
    gpuEnabled = false;
    const size_t n = songs.size();
    // AoS -> dense row-major n x 12 (what Recommender.cu:162-166 packs), then one upload
    std::vector<float> dense(n * FEATURE_COUNT);
    for (size_t i = 0; i < n; ++i) std::memcpy(&dense[i * FEATURE_COUNT], songs[i].features, sizeof(float) * FEATURE_COUNT);
    impl->reset_index(n);
    for (size_t i = 0; i < n; ++i) impl->index_song(i, songs[i].track_id, songs[i].track_name);
    if (!initializeDense(dense.data(), (long long)n)) return false;
    std::cout << "Recommender initialized on GPU: " << numSongs << " songs resident" << std::endl;
    return true;
}

std::vector<std::vector<Recommendation> > Recommender::recommendBatch(const std::vector<int> &idx, int topN)
{
    std::vector<std::vector<Recommendation> > out;
    if (!initialized) {
        std::cerr << "Error: Recommender not initialized" << std::endl;
        return out;
    }
    if (topN <= 0 || idx.empty()) return out;
    for (int q : idx) {
        if (q < 0 || q >= numSongs) {
            std::cerr << "Error: Invalid song index: " << q << std::endl;
            return out;
        }
    }
    // like the reference (Recommender.cu:300-315) a query yields min(topN, songs - 1) results, whatever topN
    const int k = std::max(1, std::min(topN, numSongs - 1));
    std::vector<int32_t> qi(idx.begin(), idx.end());
    std::vector<int32_t> oi(qi.size() * (size_t)k);
    std::vector<float> os(qi.size() * (size_t)k);
    const int rc = impl->sharded ? sr_sharded_query_by_index(impl->sharded, qi.data(), (int)qi.size(), k, oi.data(), os.data())
                                 : sr_engine_query_by_index(impl->engine, qi.data(), (int)qi.size(), k, oi.data(), os.data());
    if (rc != SR_OK) {
        std::cerr << "Error: GPU query failed: "
                  << (impl->sharded ? sr_sharded_last_error(impl->sharded) : sr_engine_last_error(impl->engine)) << std::endl;
        return out;
    }
    out.resize(qi.size());
    for (size_t q = 0; q < qi.size(); ++q)
        for (int r = 0; r < k && oi[q * k + r] >= 0; ++r) out[q].emplace_back(oi[q * k + r], os[q * k + r]);
    return out;
}

std::vector<int> Recommender::recommendByIndex(int songIndex, int topN)
{
    if (!initialized) {
        std::cerr << "Error: Recommender not initialized" << std::endl;
        return {};
    }
    if (songIndex < 0 || songIndex >= numSongs) {
        std::cerr << "Error: Invalid song index: " << songIndex << std::endl;
        return {};
    }
    std::vector<int> res;
    std::vector<std::vector<Recommendation> > b = recommendBatch(std::vector<int>(1, songIndex), topN);
    if (b.empty()) return res;
    res.reserve(b[0].size());
    for (const Recommendation &r : b[0]) res.push_back(r.songIndex);
    return res;
}

int Recommender::findSongByTrackId(const std::string &trackId) const
{
    auto it = impl->by_id.find(trackId);
    return it == impl->by_id.end() ? -1 : it->second;
}

int Recommender::findSongByName(const std::string &trackName) const
{
    const std::string q = lower(trackName);
    auto it = impl->by_lower_name.find(q);  // first case-insensitive exact match
    if (it != impl->by_lower_name.end()) return it->second;
    // else the first song whose lower-cased name contains the query
    const std::string &blob = impl->lower_blob;
    const size_t n = impl->lower_off.size() - 1;
    if (q.empty()) return n ? 0 : -1;
    std::string needle = q;
    for (char &ch : needle)
        if (ch == '\0') ch = '\1';  // as the blob stores it
    size_t pos = blob.find(needle);
    if (pos == std::string::npos) return -1;
    // names are '\0'-separated and the needle has no '\0', so a hit lies inside one name
    size_t i = std::upper_bound(impl->lower_off.begin(), impl->lower_off.end(), pos) - impl->lower_off.begin() - 1;
    return i < n ? (int)i : -1;
}

std::vector<int> Recommender::recommend(const std::string &trackId, int topN)
{
    int index = findSongByTrackId(trackId);
    if (index == -1) {
        std::cerr << "Error: Song with track_id '" << trackId << "' not found" << std::endl;
        return {};
    }
    return recommendByIndex(index, topN);
}

std::vector<int> Recommender::recommendByName(const std::string &trackName, int topN)
{
    int index = findSongByName(trackName);
    if (index == -1) {
        std::cerr << "Error: Song with name '" << trackName << "' not found" << std::endl;
        return {};
    }
    return recommendByIndex(index, topN);
}

// ---- C API around the class (FFI users, tests) -------------------------------------------
extern "C" {

// ids / names may be NULL: songs are then called "id<i>" / "Track <i>".
void *sr_recommender_create(const float *features, int64_t n, const char *const *ids, const char *const *names)
{
    if (!features || n <= 0) return nullptr;
    std::vector<Song> songs((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        Song &s = songs[(size_t)i];
        s.track_id = ids ? ids[i] : "id" + std::to_string(i);
        s.track_name = names ? names[i] : "Track " + std::to_string(i);
        s.artists = "Artist";
        s.genre_id = 0;
        std::memcpy(s.features, features + i * FEATURE_COUNT, sizeof(s.features));
    }
    Recommender *r = new Recommender();
    std::streambuf *old = std::cout.rdbuf(nullptr);
    const bool ok = r->initialize(songs);
    std::cout.rdbuf(old);
    if (!ok) {
        delete r;
        return nullptr;
    }
    return r;
}

// Straight from a songs_data.bin written by the reference's --preprocess.
void *sr_recommender_create_from_file(const char *path)
{
    if (!path) return nullptr;
    Recommender *r = new Recommender();
    std::streambuf *old = std::cout.rdbuf(nullptr);
    const bool ok = r->initializeFromFile(path);
    std::cout.rdbuf(old);
    if (!ok) {
        delete r;
        return nullptr;
    }
    return r;
}

void sr_recommender_destroy(void *h) { delete static_cast<Recommender *>(h); }
int sr_recommender_song_count(void *h) { return static_cast<Recommender *>(h)->getSongCount(); }
int sr_recommender_gpu_enabled(void *h) { return static_cast<Recommender *>(h)->isGPUEnabled() ? 1 : 0; }

static int copy_out(const std::vector<int> &r, int32_t *out)
{
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}
int sr_recommender_by_index(void *h, int idx, int k, int32_t *out)
{
    return copy_out(static_cast<Recommender *>(h)->recommendByIndex(idx, k), out);
}
int sr_recommender_by_name(void *h, const char *name, int k, int32_t *out)
{
    return copy_out(static_cast<Recommender *>(h)->recommendByName(name, k), out);
}
int sr_recommender_by_id(void *h, const char *id, int k, int32_t *out)
{
    return copy_out(static_cast<Recommender *>(h)->recommend(id, k), out);
}
int sr_recommender_find_name(void *h, const char *name) { return static_cast<Recommender *>(h)->findSongByName(name); }
int sr_recommender_find_id(void *h, const char *id) { return static_cast<Recommender *>(h)->findSongByTrackId(id); }

}  // extern "C"
