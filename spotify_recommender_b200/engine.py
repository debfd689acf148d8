"""Python binding of the C ABI (include/sr_engine.h) -- ctypes only, no torch types in
the signatures.  This module is the host-side mirror of the reference's scoring API
(`Recommender::recommendByIndex`, Recommender.cu:275-318) for whole query batches.

There is NO CPU fallback: if libsr_engine.so is missing or no sm_100 device is
present, construction raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ENGINE_SO = os.environ.get("SR_ENGINE_SO") or os.path.join(PKG, "libsr_engine.so")  # (the override: development builds)
FEATURE_COUNT = 12  # reference Song.h:12

SR_OK, SR_EINVAL, SR_ENODEVICE, SR_ECUDA, SR_ENOMEM, SR_ESTATE = range(6)
_ERRNAMES = {1: "SR_EINVAL", 2: "SR_ENODEVICE", 3: "SR_ECUDA", 4: "SR_ENOMEM", 5: "SR_ESTATE"}

EXPORTS = [
    "sr_engine_create", "sr_engine_destroy", "sr_engine_last_error", "sr_engine_load_features",
    "sr_engine_load_features_device", "sr_engine_song_count", "sr_engine_query_by_index",
    "sr_engine_query_by_vector", "sr_engine_query_by_index_dev", "sr_engine_query_by_vector_dev",
    "sr_engine_merge_topk_dev", "sr_engine_gather_rows_dev", "sr_engine_all_pairs_topk", "sr_engine_set_option", "sr_engine_get_stat",
    "sr_engine_get_timing", "sr_engine_variant_name", "sr_engine_measure_fp32", "sr_engine_selftest_div",
    "sr_engine_synchronize", "sr_engine_normalize_features", "sr_engine_normalize_features_dev", "sr_genre_ids",
    "sr_engine_query_keys_by_vector_dev", "sr_engine_merge_keys_dev", "sr_engine_bound_block_count", "sr_engine_bound_blocks_dev",
    "sr_sharded_create", "sr_sharded_destroy", "sr_sharded_last_error", "sr_sharded_load_features", "sr_sharded_song_count",
    "sr_sharded_shard_count", "sr_sharded_engine", "sr_sharded_query_by_index", "sr_sharded_all_pairs_topk",
]


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{_ERRNAMES.get(code, code)}: {msg}")
        self.code = code


_lib = None


def load_library() -> C.CDLL:
    """dlopen the in-tree engine; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(ENGINE_SO):
        raise FileNotFoundError(
            f"{ENGINE_SO} is missing: build it with `python -m spotify_recommender_b200.build` "
            "(the CUDA engine is the only implementation; there is no CPU fallback)")
    L = C.CDLL(ENGINE_SO)
    vp, i32, i64 = C.c_void_p, C.c_int, C.c_int64
    L.sr_engine_create.argtypes = [C.POINTER(vp), i32]
    L.sr_engine_destroy.argtypes = [vp]
    L.sr_engine_destroy.restype = None
    L.sr_engine_last_error.argtypes = [vp]
    L.sr_engine_last_error.restype = C.c_char_p
    L.sr_engine_load_features.argtypes = [vp, vp, i64, i64]
    L.sr_engine_load_features_device.argtypes = [vp, vp, i64, i64]
    L.sr_engine_song_count.argtypes = [vp]
    L.sr_engine_song_count.restype = i64
    L.sr_engine_query_by_index.argtypes = [vp, vp, i32, i32, vp, vp]
    L.sr_engine_query_by_vector.argtypes = [vp, vp, vp, i32, i32, vp, vp]
    L.sr_engine_query_by_index_dev.argtypes = [vp, vp, i32, i32, vp, vp, vp]
    L.sr_engine_query_by_vector_dev.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp]
    L.sr_engine_merge_topk_dev.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, vp]
    L.sr_engine_gather_rows_dev.argtypes = [vp, vp, i32, vp, vp]
    L.sr_engine_all_pairs_topk.argtypes = [vp, i64, i64, i32, vp, vp]
    L.sr_engine_set_option.argtypes = [vp, C.c_char_p, i64]
    L.sr_engine_get_stat.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    L.sr_engine_get_timing.argtypes = [vp, C.c_char_p, C.POINTER(C.c_double), C.POINTER(i64)]
    L.sr_engine_variant_name.argtypes = [i32]
    L.sr_engine_variant_name.restype = C.c_char_p
    L.sr_engine_measure_fp32.argtypes = [vp, i32, C.POINTER(C.c_double)]
    L.sr_engine_selftest_div.argtypes = [vp, vp, vp, i32, vp]
    L.sr_engine_synchronize.argtypes = [vp]
    L.sr_engine_normalize_features.argtypes = [vp, vp, vp, i64, i32, vp, vp]
    L.sr_engine_normalize_features_dev.argtypes = [vp, vp, vp, i64, i32, vp, vp, vp]
    L.sr_genre_ids.argtypes = [C.POINTER(C.c_char_p), i64, i32, vp, C.POINTER(C.c_int32)]
    L.sr_engine_query_keys_by_vector_dev.argtypes = [vp, vp, vp, i32, i32, vp, vp, vp, vp]
    L.sr_engine_bound_block_count.argtypes = [vp, i32]
    L.sr_engine_bound_blocks_dev.argtypes = [vp, vp, i32, i32, i32, vp, vp]
    L.sr_engine_merge_keys_dev.argtypes = [vp, vp, i32, i32, i32, vp, vp, i32, i32, vp, vp]
    L.sr_sharded_create.argtypes = [C.POINTER(vp), C.POINTER(C.c_int), i32]
    L.sr_sharded_destroy.argtypes = [vp]
    L.sr_sharded_destroy.restype = None
    L.sr_sharded_last_error.argtypes = [vp]
    L.sr_sharded_last_error.restype = C.c_char_p
    L.sr_sharded_load_features.argtypes = [vp, vp, i64, i32]
    L.sr_sharded_song_count.argtypes = [vp]
    L.sr_sharded_song_count.restype = i64
    L.sr_sharded_shard_count.argtypes = [vp]
    L.sr_sharded_engine.argtypes = [vp, i32]
    L.sr_sharded_engine.restype = vp
    L.sr_sharded_query_by_index.argtypes = [vp, vp, i32, i32, vp, vp]
    L.sr_sharded_all_pairs_topk.argtypes = [vp, i32, vp, vp]
    _lib = L
    return L


def genre_ids(names: list[str], sorted_ids: bool = True) -> tuple[np.ndarray, int]:
    """Genre name -> id (sr_genre_ids): first-appearance order (the reference with one thread) or sorted."""
    L = load_library()
    arr = (C.c_char_p * len(names))(*[s.encode() for s in names])
    ids = np.empty(len(names), np.int32)
    ng = C.c_int32(0)
    rc = L.sr_genre_ids(arr, len(names), 1 if sorted_ids else 0, ids.ctypes.data_as(C.c_void_p), C.byref(ng))
    if rc:
        raise EngineError(rc, "sr_genre_ids: bad arguments")
    return ids, int(ng.value)


def variant_names() -> list[str]:
    L = load_library()
    out, i = [], 0
    while True:
        n = L.sr_engine_variant_name(i)
        if n is None:
            return out
        out.append(n.decode())
        i += 1


def _ptr(a) -> C.c_void_p:
    """numpy array -> host pointer; torch CUDA tensor / int -> device pointer."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, np.ndarray):
        return C.c_void_p(a.ctypes.data)
    if isinstance(a, int):
        return C.c_void_p(a)
    return C.c_void_p(a.data_ptr())  # torch tensor


def _stream(stream) -> C.c_void_p:
    """None -> the engine's own stream (SR_ENGINE_OWN_STREAM); an int is a cudaStream_t
    handle (0 = CUDA's legacy default stream, torch's default)."""
    return C.c_void_p(-1 & (2 ** 64 - 1)) if stream is None else C.c_void_p(int(stream))


class Engine:
    """One scoring engine = one CUDA device (one process per GPU in multi-GPU runs)."""

    def __init__(self, device: int = -1, _borrowed=None):
        self.L = load_library()
        self._owned = _borrowed is None
        if _borrowed is not None:  # a shard of a ShardedEngine: options / stats only
            self.h = C.c_void_p(_borrowed)
            return
        h = C.c_void_p()
        rc = self.L.sr_engine_create(C.byref(h), device)
        if rc != SR_OK:
            raise EngineError(rc, self.L.sr_engine_last_error(None).decode())
        self.h = h

    # -- lifecycle -----------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "h", None):
            if self._owned:
                self.L.sr_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != SR_OK:
            raise EngineError(rc, self.L.sr_engine_last_error(self.h).decode())

    # -- store (replaces Recommender::initialize, Recommender.cu:100-182) -------
    def load_features(self, rows, id_base: int = 0) -> None:
        """rows: host float32 (n, 12) array, or a torch CUDA tensor of that shape."""
        if isinstance(rows, np.ndarray):
            rows = np.ascontiguousarray(rows, np.float32)
            if rows.ndim != 2 or rows.shape[1] != FEATURE_COUNT:
                raise ValueError("rows must be (n, 12) float32")
            self._check(self.L.sr_engine_load_features(self.h, _ptr(rows), rows.shape[0], id_base))
        else:
            if tuple(rows.shape[1:]) != (FEATURE_COUNT,) or not rows.is_contiguous():
                raise ValueError("rows must be a contiguous (n, 12) float32 CUDA tensor")
            self._check(self.L.sr_engine_load_features_device(self.h, _ptr(rows), rows.shape[0], id_base))

    @property
    def song_count(self) -> int:
        return int(self.L.sr_engine_song_count(self.h))

    # -- queries, HOST buffers (the public, end-to-end path) -------------------
    def query_by_index(self, qidx, k: int, scores: bool = True):
        """Batch form of Recommender::recommendByIndex (Recommender.cu:275-318):
        each query song excludes itself; returns (idx[nq,k] int32, score[nq,k] f32)."""
        qidx = np.ascontiguousarray(qidx, np.int32).ravel()
        out_i = np.empty((qidx.size, max(k, 0)), np.int32)
        out_s = np.empty((qidx.size, max(k, 0)), np.float32) if scores else None
        self._check(self.L.sr_engine_query_by_index(self.h, _ptr(qidx), qidx.size, k, _ptr(out_i), _ptr(out_s)))
        return out_i, out_s

    def query_by_vector(self, qrows, k: int, exclude=None, scores: bool = True):
        qrows = np.ascontiguousarray(qrows, np.float32).reshape(-1, FEATURE_COUNT)
        ex = None if exclude is None else np.ascontiguousarray(exclude, np.int32).ravel()
        out_i = np.empty((qrows.shape[0], max(k, 0)), np.int32)
        out_s = np.empty((qrows.shape[0], max(k, 0)), np.float32) if scores else None
        self._check(self.L.sr_engine_query_by_vector(self.h, _ptr(qrows), _ptr(ex), qrows.shape[0], k,
                                                     _ptr(out_i), _ptr(out_s)))
        return out_i, out_s

    def all_pairs_topk(self, q_lo: int, q_hi: int, k: int, scores: bool = True):
        out_i = np.empty((q_hi - q_lo, k), np.int32)
        out_s = np.empty((q_hi - q_lo, k), np.float32) if scores else None
        self._check(self.L.sr_engine_all_pairs_topk(self.h, q_lo, q_hi, k, _ptr(out_i), _ptr(out_s)))
        return out_i, out_s

    # -- preprocessing: min-max normalisation (DataManager.cpp:270-301) --------
    def normalize_features(self, raw11: np.ndarray, genre_id: np.ndarray, n_genres: int):
        """raw11 (n x 11 float32), genre ids -> (n x 12 normalised features, 22 minima/maxima); host buffers."""
        raw11 = np.ascontiguousarray(raw11, np.float32)
        genre_id = np.ascontiguousarray(genre_id, np.int32)
        n = raw11.shape[0]
        assert raw11.shape == (n, 11) and genre_id.shape == (n,)
        out = np.empty((n, 12), np.float32)
        mm = np.empty(22, np.float32)
        self._check(self.L.sr_engine_normalize_features(self.h, _ptr(raw11), _ptr(genre_id), n, n_genres, _ptr(out), _ptr(mm)))
        return out, mm

    def normalize_features_dev(self, d_raw11, d_genre_id, n: int, n_genres: int, d_out, d_minmax=None,
                               stream: int | None = None) -> None:
        self._check(self.L.sr_engine_normalize_features_dev(self.h, _ptr(d_raw11), _ptr(d_genre_id), n, n_genres,
                                                            _ptr(d_out), _ptr(d_minmax), _stream(stream)))

    # -- queries, DEVICE buffers (stream-ordered, not synchronised) ------------
    def query_by_index_dev(self, d_qidx, nq: int, k: int, d_out_idx, d_out_score=None, stream: int | None = None) -> None:
        self._check(self.L.sr_engine_query_by_index_dev(self.h, _ptr(d_qidx), nq, k, _ptr(d_out_idx),
                                                        _ptr(d_out_score), _stream(stream)))

    def query_by_vector_dev(self, d_qrows, d_exclude, nq: int, k: int, d_out_idx, d_out_score=None,
                            stream: int | None = None) -> None:
        self._check(self.L.sr_engine_query_by_vector_dev(self.h, _ptr(d_qrows), _ptr(d_exclude), nq, k,
                                                         _ptr(d_out_idx), _ptr(d_out_score), _stream(stream)))

    def merge_topk_dev(self, d_idx, d_score, parts: int, nq: int, k: int, d_out_idx, d_out_score=None,
                       stream: int | None = None) -> None:
        self._check(self.L.sr_engine_merge_topk_dev(self.h, _ptr(d_idx), _ptr(d_score), parts, nq, k,
                                                    _ptr(d_out_idx), _ptr(d_out_score), _stream(stream)))

    def query_keys_by_vector_dev(self, d_qrows, d_exclude, nq: int, k: int, d_out_keys, d_ceil=None,
                                 stream: int | None = None, d_blocks=None) -> None:
        """Local top-k in the exchange format of the row-sharded path: nq x k packed 64-bit keys (0 = none).
        d_blocks: the max-reduced block maxima of the shared bound pass (bound_blocks_dev), or None."""
        self._check(self.L.sr_engine_query_keys_by_vector_dev(self.h, _ptr(d_qrows), _ptr(d_exclude), nq, k, _ptr(d_ceil),
                                                              _ptr(d_blocks), _ptr(d_out_keys), _stream(stream)))

    def bound_block_count(self, k: int) -> int:
        """Blocks per query of the shared bound pass for lists of k (0: none -- k > 255)."""
        return int(self.L.sr_engine_bound_block_count(self.h, k))

    def bound_blocks_dev(self, d_qrows, nq: int, k: int, shards: int, d_blocks, stream: int | None = None) -> None:
        """This shard's part of the bound pass shared between `shards` row shards: nq x bound_block_count(k) float
        block maxima (-inf = none), to be max-reduced across the shards."""
        self._check(self.L.sr_engine_bound_blocks_dev(self.h, _ptr(d_qrows), nq, k, shards, _ptr(d_blocks), _stream(stream)))

    def merge_keys_dev(self, d_keys, parts: int, nq: int, k: int, d_out_idx, d_out_score=None, stride: int = 0, col: int = 0,
                       d_ceil_out=None, stream: int | None = None) -> None:
        self._check(self.L.sr_engine_merge_keys_dev(self.h, _ptr(d_keys), parts, nq, k, _ptr(d_out_idx), _ptr(d_out_score),
                                                    stride, col, _ptr(d_ceil_out), _stream(stream)))

    def gather_rows_dev(self, d_ids, count: int, d_out, stream: int | None = None) -> None:
        self._check(self.L.sr_engine_gather_rows_dev(self.h, _ptr(d_ids), count, _ptr(d_out), _stream(stream)))

    # -- knobs / introspection -----------------------------------------------
    def set_option(self, key: str, value: int) -> None:
        self._check(self.L.sr_engine_set_option(self.h, key.encode(), int(value)))

    def stat(self, key: str) -> int:
        v = C.c_int64()
        self._check(self.L.sr_engine_get_stat(self.h, key.encode(), C.byref(v)))
        return int(v.value)

    def timing(self, kernel: str):
        ms, n = C.c_double(), C.c_int64()
        self._check(self.L.sr_engine_get_timing(self.h, kernel.encode(), C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def measure_fp32(self, variant: int = 1) -> float:
        t = C.c_double()
        self._check(self.L.sr_engine_measure_fp32(self.h, variant, C.byref(t)))
        return float(t.value)

    def selftest_div(self, a, b):
        a = np.ascontiguousarray(a, np.float32).ravel()
        b = np.ascontiguousarray(b, np.float32).ravel()
        out = np.empty_like(a)
        self._check(self.L.sr_engine_selftest_div(self.h, _ptr(a), _ptr(b), a.size, _ptr(out)))
        return out

    def synchronize(self) -> None:
        self._check(self.L.sr_engine_synchronize(self.h))


class ShardedEngine:
    """Several GPUs behind one handle in ONE process (sr_sharded_* of include/sr_engine.h): a row-sharded store
    whose shards' top-k lists are merged over NVLink peer access, or a replicated store with sharded queries
    (all-pairs).  `devices` lists CUDA ordinals, one shard each; an ordinal may repeat (several shards on one
    GPU -- how the path is tested on a single-GPU box); None = every visible device."""

    def __init__(self, devices=None):
        self.L = load_library()
        h = C.c_void_p()
        if devices is None:
            rc = self.L.sr_sharded_create(C.byref(h), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.L.sr_sharded_create(C.byref(h), arr, len(devices))
        if rc != SR_OK:
            raise EngineError(rc, self.L.sr_sharded_last_error(None).decode())
        self.h = h

    def close(self) -> None:
        if getattr(self, "h", None):
            self.L.sr_sharded_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int) -> None:
        if rc != SR_OK:
            raise EngineError(rc, self.L.sr_sharded_last_error(self.h).decode())

    @property
    def shard_count(self) -> int:
        return int(self.L.sr_sharded_shard_count(self.h))

    @property
    def song_count(self) -> int:
        return int(self.L.sr_sharded_song_count(self.h))

    def shard(self, i: int) -> Engine:
        """Shard i's engine, for set_option / stat (owned by this object)."""
        p = self.L.sr_sharded_engine(self.h, i)
        if not p:
            raise IndexError(i)
        return Engine(_borrowed=p)

    def load_features(self, rows: np.ndarray, replicate: bool = False) -> None:
        rows = np.ascontiguousarray(rows, np.float32)
        if rows.ndim != 2 or rows.shape[1] != FEATURE_COUNT:
            raise ValueError("rows must be (n, 12) float32")
        self._check(self.L.sr_sharded_load_features(self.h, _ptr(rows), rows.shape[0], 1 if replicate else 0))

    def query_by_index(self, qidx, k: int, scores: bool = True):
        qidx = np.ascontiguousarray(qidx, np.int32).ravel()
        out_i = np.empty((qidx.size, max(k, 0)), np.int32)
        out_s = np.empty((qidx.size, max(k, 0)), np.float32) if scores else None
        self._check(self.L.sr_sharded_query_by_index(self.h, _ptr(qidx), qidx.size, k, _ptr(out_i), _ptr(out_s)))
        return out_i, out_s

    def all_pairs_topk(self, k: int, scores: bool = True):
        n = self.song_count
        out_i = np.empty((n, k), np.int32)
        out_s = np.empty((n, k), np.float32) if scores else None
        self._check(self.L.sr_sharded_all_pairs_topk(self.h, k, _ptr(out_i), _ptr(out_s)))
        return out_i, out_s
