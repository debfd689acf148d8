/*
 * sr_engine.h -- C ABI of the B200 (sm_100a) cosine-similarity + top-K scoring
 * engine: the drop-in boundary for the hot path of Iamdarika/Spotify_recommender.
 *
 * The reference has no FFI: its hot path is reached through the C++ class in
 * Recommender.h:28-82.  Each entry point below names the reference code it
 * replaces; sr_recommender.hpp (same directory) re-creates that class on top of
 * these calls so the reference's main.cpp keeps working unchanged.
 *
 * Conventions
 *   - plain C types only; every call returns 0 on success or an SR_E* code and
 *     leaves a message readable through sr_engine_last_error().
 *   - one engine = one CUDA device (one process per GPU); calls on one engine
 *     are serialised by the caller (the reference is single-threaded, §8b).
 *   - there is NO CPU fallback: without an sm_100 device create() fails.
 *   - song ids are 32-bit and GLOBAL: id = id_base + local row, so a row shard
 *     of a larger store answers with ids of the whole store.
 *   - results are ordered by (score descending, id ascending); rows shorter
 *     than k are padded with id -1 / score 0.  Scores are bit-identical to the
 *     reference's CPU arithmetic (Recommender.cu:256-273); the one exception is
 *     that a score of -0.0 (only non-finite inputs produce one) is returned as +0.0.
 *   - any k >= 1: like the reference (Recommender.cu:300-315) a query yields
 *     min(k, songs - 1) results.  Lists longer than 1024 are produced 1024 at a
 *     time (one more pass over the store each, every pass continuing below the
 *     last key of the one before).
 */
#ifndef SR_ENGINE_H
#define SR_ENGINE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SR_FEATURE_COUNT 12 /* reference Song.h:12 */

enum {
    SR_OK = 0,
    SR_EINVAL = 1,   /* bad argument (null, k <= 0, index out of range ...) */
    SR_ENODEVICE = 2,/* no usable sm_100 CUDA device                        */
    SR_ECUDA = 3,    /* a CUDA runtime call or kernel failed                 */
    SR_ENOMEM = 4,   /* host or device allocation failed                    */
    SR_ESTATE = 5    /* call made before load_features                      */
};

typedef struct sr_engine sr_engine;

#define SR_ENGINE_OWN_STREAM ((void *)(intptr_t)-1)

/* Replaces the device probing / cublasCreate part of Recommender::initialize
 * (Recommender.cu:117-149).  device < 0 selects the current device. */
int sr_engine_create(sr_engine **out, int device);

/* Replaces Recommender::~Recommender (Recommender.cu:86-98). */
void sr_engine_destroy(sr_engine *e);

/* Last error text of this engine ("" if none); e == NULL gives the text of the
 * last failed create(). */
const char *sr_engine_last_error(const sr_engine *e);

/* Replaces the pack + upload part of Recommender::initialize
 * (Recommender.cu:153-175): rows is a HOST dense row-major n x 12 FP32 matrix
 * (Song::features of every song, Song.h:26).  Builds the device-resident
 * stores (raw rows + exact norms, and the pre-normalised scan store).  May be
 * called again to replace the store.  id_base is the global id of row 0. */
int sr_engine_load_features(sr_engine *e, const float *rows, int64_t n, int64_t id_base);

/* Same, rows already on this engine's device (n x 12, 16-byte aligned). */
int sr_engine_load_features_device(sr_engine *e, const float *d_rows, int64_t n, int64_t id_base);

/* Recommender::getSongCount (Recommender.h:82). */
int64_t sr_engine_song_count(const sr_engine *e);

/* Replaces Recommender::recommendByIndex (Recommender.cu:275-318) for a BATCH
 * of in-store query songs: qidx are global ids owned by this engine; each
 * query excludes itself by id (Recommender.cu:296).  HOST buffers;
 * out_idx / out_score are nq x k.  out_score may be NULL. */
int sr_engine_query_by_index(sr_engine *e, const int32_t *qidx, int nq, int k,
                             int32_t *out_idx, float *out_score);

/* Same scoring for arbitrary query rows (nq x 12, HOST).  exclude is NULL or nq
 * global ids to skip (-1: none).  This is what a row shard of a multi-GPU store
 * is asked: the query row may live on another shard. */
int sr_engine_query_by_vector(sr_engine *e, const float *qrows, const int32_t *exclude,
                              int nq, int k, int32_t *out_idx, float *out_score);

/* Device-resident variants: all pointers are device memory of this engine's
 * device, work is enqueued on `stream` and NOT synchronised.  `stream` is a
 * cudaStream_t (NULL = CUDA's legacy default stream, as everywhere in CUDA) or
 * SR_ENGINE_OWN_STREAM for the engine's own non-blocking stream.  Used by the multi-GPU host (torch owns the
 * buffers and the NCCL exchange) and by bench.py's kernel-only timing.
 * A query id the store does not own cannot be refused before the work is enqueued: its result row is
 * -1 / 0 and the engine remembers it -- the next sr_engine_synchronize() (or any host-buffer call) returns
 * SR_EINVAL, and sr_engine_get_stat("bad_index") reads and clears the flag. */
int sr_engine_query_by_index_dev(sr_engine *e, const int32_t *d_qidx, int nq, int k,
                                 int32_t *d_out_idx, float *d_out_score, void *stream);
int sr_engine_query_by_vector_dev(sr_engine *e, const float *d_qrows, const int32_t *d_exclude,
                                  int nq, int k, int32_t *d_out_idx, float *d_out_score,
                                  void *stream);

/* Final step of the row-sharded multi-GPU path: merges `parts`
 * per-shard result lists (parts x nq x k, as gathered by NCCL all-gather; -1
 * padded) into one list per query in (score desc, id asc) order.  Device
 * pointers, stream-ordered. */
int sr_engine_merge_topk_dev(sr_engine *e, const int32_t *d_idx, const float *d_score,
                             int parts, int nq, int k, int32_t *d_out_idx, float *d_out_score,
                             void *stream);

/* The same local scoring with the result in the EXCHANGE FORMAT of the row-sharded path: one packed 64-bit key per
 * candidate (orderable score << 32 | ~id; 0 = none), nq x k, so that a step needs exactly one all-gather.
 * d_ceil is NULL or nq keys: only candidates ordered strictly after d_ceil[q] are admitted (how a list longer
 * than 1024 continues across shards).  d_blocks is NULL or the max-reduced block maxima of the SHARED bound
 * pass below: the shard then starts from thresholds that bound the k-th best of the whole store, and its list
 * may hold fewer than k keys (only those that can still make the merged top-k).  1 <= k <= 1024. */
int sr_engine_query_keys_by_vector_dev(sr_engine *e, const float *d_qrows, const int32_t *d_exclude,
                                       int nq, int k, const uint64_t *d_ceil, const float *d_blocks,
                                       uint64_t *d_out_keys, void *stream);

/* The threshold bound pass shared between the `shards` row shards of one store: this shard
 * samples 1 / shards of the tiles a single store would and writes, per query, the best filter score of each of
 * sr_engine_bound_block_count(k) disjoint blocks of its sample (nq x blocks floats, -inf = none).  The caller
 * max-reduces the arrays of all shards element-wise (one all-reduce) and hands the result to
 * sr_engine_query_keys_by_vector_dev.  Block b of every shard holds different songs, so the (k+1)-th largest
 * reduced maximum bounds the store-wide k-th best: every shard scans with the whole store's threshold at an
 * eighth of the sampling cost.  block_count is 0 when lists of k have no bound pass (k > 255). */
int sr_engine_bound_block_count(sr_engine *e, int k);
int sr_engine_bound_blocks_dev(sr_engine *e, const float *d_qrows, int nq, int k, int shards,
                               float *d_blocks, void *stream);

/* Merge of `parts` gathered key lists (parts x nq x k, as one NCCL all-gather of the buffers above delivers
 * them) into columns [col, col + k) of the nq x stride result rows; d_ceil_out (or NULL) receives each
 * query's k-th key (0 when fewer than k candidates exist).  stride <= 0 means k. */
int sr_engine_merge_keys_dev(sr_engine *e, const uint64_t *d_keys, int parts, int nq, int k,
                             int32_t *d_out_idx, float *d_out_score, int stride, int col,
                             uint64_t *d_ceil_out, void *stream);

/* First step of the row-sharded multi-GPU path: d_out (count x 12) receives the raw
 * feature rows (Song::features, Song.h:26) of the global ids this engine owns and
 * zeros for the others, so that one sum all-reduce over the shards hands every rank
 * the whole query matrix.  Device pointers, stream-ordered. */
int sr_engine_gather_rows_dev(sr_engine *e, const int32_t *d_ids, int count, float *d_out, void *stream);

/* All-pairs neighbour table (BASELINE config 5): for every owned song with
 * global id in [q_lo, q_hi) its top-k neighbours, streamed through the same
 * kernels in batches.  HOST outputs, (q_hi - q_lo) x k. */
int sr_engine_all_pairs_topk(sr_engine *e, int64_t q_lo, int64_t q_hi, int k,
                             int32_t *out_idx, float *out_score);

/* Tunables (DESIGN.md "knobs"):
 *   "variant"   scan kernel shape, index into the table sr_engine_variant_name() lists;
 *               -1 (default) picks by batch size
 *   "qt"        queries per shared-memory tile (1..256)
 *   "batch"     max queries per internal pass (workspace is sized for it)
 *   "sample"    threshold-bootstrap sample size per query (0 = off, else power of two <= 4096)
 *   "hit_cap"   hit-buffer entries per query per CTA (multiple of 32; 0 = sized from k)
 *   "list_ws"   1 (default): the CTAs' top-k lists may live in an L2-resident workspace instead of shared
 *               memory when that keeps the query tile at 256 queries (used for 16 < k <= "list_ws_kmax"); 0: never
 *   "small_max" batches of at most this many queries take the TMA-staged small-batch kernel shape
 *   "graphs"    1 (default): single-group passes of up to 256 queries are captured once per (buffers, nq, k) as a
 *               CUDA graph and replayed with one launch; 0: always launch kernel by kernel
 *   "bound_blocks", "prefetch", "refresh_every", "mid_max", "list_ws_kmax": development knobs
 *   "bound_tiles" layout tiles (2048 songs each) sampled by the threshold bound pass (0 = auto: 48 for k <= 16, else 128)
 *   "trigger_at" a settle phase starts when some hit buffer holds this many ids (0 = cap / 4)
 *   "settle_at" ... and scores and merges every buffer holding at least this many (0 = cap / 32)
 *   "bound"     0: skip the bound pass (threshold bootstrap at filter speed)
 *   "profile"   1: bracket every kernel with CUDA events (read with sr_engine_get_timing)
 *   "reset"     any value: zero the counters and timings below               */
int sr_engine_set_option(sr_engine *e, const char *key, int64_t value);

/* Counters since create()/reset (synchronises the engine's stream):
 * "kernel_launches", "queries", "filter_hits", "settles", "rescans", "refilters", "rescored",
 * "irregular_songs", "sm_count", "scan_grid", "scan_tile_songs", "device_bytes",
 * "variant" (the shape the last pass used), "qt", "lists_in_smem" (of the last pass), "graph_replays",
 * "bad_index" (reads and clears the flag described above). */
int sr_engine_get_stat(sr_engine *e, const char *key, int64_t *value);

/* With "profile" on: total device milliseconds and launch count of one kernel
 * ("prep", "sample", "bound", "scan", "finalize", "merge") since the last reset, measured
 * with CUDA events on the launching stream.  Synchronises that stream. */
int sr_engine_get_timing(sr_engine *e, const char *kernel, double *ms_total, int64_t *launches);

/* Name of scan-kernel shape i ("S8xT256x2" ...), NULL past the end. */
const char *sr_engine_variant_name(int i);

/* FP32 pipe microbenchmark on this engine's device, the measured FP32 roofline
 * denominator.  variant 0: FFMA, 1: packed FFMA2 (fma.rn.f32x2), 2: unfused
 * FMUL+FADD (the oracle's instruction mix).  Returns TFLOP/s (2 flop per FMA). */
int sr_engine_measure_fp32(sr_engine *e, int variant, double *tflops);

/* Self-test hook: out[i] = the engine's device-side IEEE division a[i] / b[i] (b > 0),
 * the one operation of the reference's scoring (Recommender.cu:271) that is not a
 * single hardware instruction on the GPU.  HOST buffers.  Used by the tests only. */
int sr_engine_selftest_div(sr_engine *e, const float *a, const float *b, int n, float *out);

/* SURVEY 8 f4 -- the min-max normalisation step of the reference's preprocessing
 * (DataManager.cpp:270-301) on the GPU, deterministic and bit-identical to the reference's
 * arithmetic: per column j < 11 of `raw11` (n x 11 row-major: danceability, energy, key,
 * loudness, mode, speechiness, acousticness, instrumentalness, liveness, valence, tempo --
 * DataManager.cpp:156-159) min and max over all rows (NaN never enters, as with std::min /
 * std::max); out[i][j] = range > 1e-4f ? (x - min) / range : 0.5f; out[i][11] =
 * (float)genre_id[i] / max(1, n_genres - 1).  `out` is n x 12, ready for
 * sr_engine_load_features[_device]; `minmax` (22 floats: minima then maxima) may be NULL.
 * A zero minimum / maximum counts as +0 whatever the order of signed zeros in the column.
 * The _dev form takes device pointers (16-byte aligned) and is ordered on `stream`; the host
 * form stages through temporary device buffers. */
int sr_engine_normalize_features(sr_engine *e, const float *raw11, const int32_t *genre_id, int64_t n, int32_t n_genres,
                                 float *out, float *minmax);
int sr_engine_normalize_features_dev(sr_engine *e, const float *d_raw11, const int32_t *d_genre_id, int64_t n,
                                     int32_t n_genres, float *d_out, float *d_minmax, void *stream);

/* SURVEY 8 f4 -- genre name -> genre id for `n` songs (host only, no engine needed).
 * mode 0: order of first appearance = what DataManager.cpp:244-250 yields with one OpenMP thread
 * (with more threads the reference's ids depend on scheduling); mode 1: rank of the name in
 * sorted (byte-wise) order -- deterministic whatever the order of the rows. */
int sr_genre_ids(const char *const *names, int64_t n, int mode, int32_t *ids, int32_t *n_genres);

/* Blocks until everything enqueued on the engine's own stream has finished; returns SR_EINVAL if a
 * device-pointer call since the last check met a query id the store does not own. */
int sr_engine_synchronize(sr_engine *e);

/* ---- several GPUs behind one handle, in ONE process (the reference itself is hard-wired to
 * device 0, Recommender.cu:124).  This is what include/sr_recommender.hpp uses when more than one sm_100
 * device is visible, so the reference's C++ host reaches the whole box without Python or MPI.
 *
 *   devices / n_devices   CUDA ordinals, one shard each (NULL / 0: every visible device).  An ordinal may
 *                         repeat -- several shards on one GPU -- which is how the path is tested on one GPU.
 *   load_features         replicate == 0: rows are split into contiguous row shards (shard s owns
 *                         [s * ceil(n / G), ...)), queries visit every shard, the shards' top-k key lists are
 *                         merged by one kernel that reads them over NVLink peer access (north_star (4));
 *                         replicate != 0: every shard holds all rows and serves a slice of the QUERIES
 *                         (BASELINE config 5: all-pairs).
 *   query_by_index        as sr_engine_query_by_index, ids global in [0, n); any k >= 1.
 *   all_pairs_topk        the n x k neighbour table (HOST), every song a query.
 * Results are identical, bit for bit, whatever the number of shards. */
typedef struct sr_sharded sr_sharded;
int sr_sharded_create(sr_sharded **out, const int *devices, int n_devices);
void sr_sharded_destroy(sr_sharded *s);
const char *sr_sharded_last_error(const sr_sharded *s);
int sr_sharded_load_features(sr_sharded *s, const float *rows, int64_t n, int replicate);
int64_t sr_sharded_song_count(const sr_sharded *s);
int sr_sharded_shard_count(const sr_sharded *s);
sr_engine *sr_sharded_engine(sr_sharded *s, int shard); /* for set_option / get_stat */
int sr_sharded_query_by_index(sr_sharded *s, const int32_t *qidx, int nq, int k,
                              int32_t *out_idx, float *out_score);
int sr_sharded_all_pairs_topk(sr_sharded *s, int k, int32_t *out_idx, float *out_score);

#ifdef __cplusplus
}
#endif
#endif /* SR_ENGINE_H */
