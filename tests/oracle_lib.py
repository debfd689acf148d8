"""ctypes access to the CHECKERS: oracle/liboracle.so (C restatement) and, when
present, oracle/_ref/libref_{cpu,gpu}.so (the unmodified reference behind
oracle/ref_harness.cpp).  Test infrastructure only -- the product package never
imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def build_oracle() -> str:
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "cosine_topk_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "liboracle.so"])
    return so


class Oracle:
    def __init__(self) -> None:
        L = C.CDLL(build_oracle())
        L.sr_oracle_selfcheck_unfused.restype = C.c_int
        L.sr_oracle_query_norm.restype = C.c_float
        L.sr_oracle_query_norm.argtypes = [_f32p]
        L.sr_oracle_scores.argtypes = [_f32p, C.c_int64, _f32p, _f32p, C.c_int]
        L.sr_oracle_topk_canonical.restype = C.c_int
        L.sr_oracle_topk_canonical.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, C.c_int32, _i32p, _f32p]
        L.sr_oracle_topk_refheap.restype = C.c_int
        L.sr_oracle_topk_refheap.argtypes = [_f32p, C.c_int64, C.c_int64, C.c_int, _i32p]
        L.sr_oracle_query_rows.argtypes = [_f32p, C.c_int64, _f32p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_int32, _i32p, _f32p, C.c_int]
        L.sr_oracle_query_index.argtypes = [_f32p, C.c_int64, _i32p, C.c_int, C.c_int, _i32p, _f32p, C.c_int]
        L.sr_oracle_merge_parts.argtypes = [_i32p, _f32p, C.c_int, C.c_int, C.c_int, _i32p, _f32p]
        L.sr_oracle_minmax_normalize.argtypes = [_f32p, _i32p, C.c_int64, C.c_int32, _f32p, _f32p]
        L.sr_oracle_max_threads.restype = C.c_int
        self.L = L

    @property
    def max_threads(self) -> int:
        return int(self.L.sr_oracle_max_threads())

    def unfused(self) -> bool:
        return bool(self.L.sr_oracle_selfcheck_unfused())

    def scores(self, feats: np.ndarray, q: np.ndarray, threads: int = 1) -> np.ndarray:
        feats = np.ascontiguousarray(feats, np.float32)
        out = np.empty(feats.shape[0], np.float32)
        self.L.sr_oracle_scores(feats, feats.shape[0], np.ascontiguousarray(q, np.float32), out, threads)
        return out

    def topk_canonical(self, scores, exclude: int, k: int, id_base: int = 0):
        scores = np.ascontiguousarray(scores, np.float32)
        oi = np.empty(max(k, 1), np.int32)
        os_ = np.empty(max(k, 1), np.float32)
        n = self.L.sr_oracle_topk_canonical(scores, scores.size, exclude, k, id_base, oi, os_)
        return oi[:k], os_[:k], n

    def topk_refheap(self, scores, exclude: int, k: int) -> np.ndarray:
        scores = np.ascontiguousarray(scores, np.float32)
        oi = np.empty(max(1, min(max(k, 1), scores.size)), np.int32)
        n = self.L.sr_oracle_topk_refheap(scores, scores.size, exclude, k, oi)
        return oi[:n]

    def query_index(self, feats, qidx, k: int, threads: int = 1):
        feats = np.ascontiguousarray(feats, np.float32)
        qidx = np.ascontiguousarray(qidx, np.int32)
        oi = np.empty((qidx.size, k), np.int32)
        os_ = np.empty((qidx.size, k), np.float32)
        self.L.sr_oracle_query_index(feats, feats.shape[0], qidx, qidx.size, k, oi, os_, threads)
        return oi, os_

    def query_rows(self, feats, q, exclude, k: int, id_base: int = 0, threads: int = 1):
        feats = np.ascontiguousarray(feats, np.float32)
        q = np.ascontiguousarray(q, np.float32).reshape(-1, 12)
        oi = np.empty((q.shape[0], k), np.int32)
        os_ = np.empty((q.shape[0], k), np.float32)
        if exclude is None:
            ex_p = None
        else:
            ex = np.ascontiguousarray(exclude, np.int64)
            ex_p = ex.ctypes.data_as(C.c_void_p)
        self.L.sr_oracle_query_rows(feats, feats.shape[0], q, ex_p, q.shape[0], k, id_base, oi, os_, threads)
        return oi, os_

    def minmax_normalize(self, raw11, genre_id, n_genres: int):
        """DataManager.cpp:270-301 restated: (n x 12 features, 22 minima/maxima)."""
        raw11 = np.ascontiguousarray(raw11, np.float32)
        genre_id = np.ascontiguousarray(genre_id, np.int32)
        n = raw11.shape[0]
        out = np.empty((n, 12), np.float32)
        mm = np.empty(22, np.float32)
        self.L.sr_oracle_minmax_normalize(raw11, genre_id, n, n_genres, out, mm)
        return out, mm

    def merge_parts(self, idx, score):
        idx = np.ascontiguousarray(idx, np.int32)
        score = np.ascontiguousarray(score, np.float32)
        parts, nq, k = idx.shape
        oi = np.empty((nq, k), np.int32)
        os_ = np.empty((nq, k), np.float32)
        self.L.sr_oracle_merge_parts(idx, score, parts, nq, k, oi, os_)
        return oi, os_


class Reference:
    """The unmodified reference class (CPU build by default)."""

    def __init__(self, feats: np.ndarray, gpu: bool = False) -> None:
        so = os.path.join(ORACLE_DIR, "_ref", "libref_gpu.so" if gpu else "libref_cpu.so")
        if not os.path.exists(so):
            raise FileNotFoundError(so)
        L = C.CDLL(so)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [_f32p, C.c_int64]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_gpu_enabled.argtypes = [C.c_void_p]
        L.ref_recommend_by_index.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
        L.ref_recommend_by_name.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _i32p]
        L.ref_recommend_by_id.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _i32p]
        L.ref_scores.argtypes = [C.c_void_p, C.c_int, _f32p]
        L.ref_batch_by_index.argtypes = [C.c_void_p, _i32p, C.c_int, C.c_int, _i32p, C.c_int]
        L.ref_gen_mt19937_uniform.argtypes = [C.c_int64, C.c_uint32, _f32p]
        L.ref_max_threads.restype = C.c_int
        self.L = L
        feats = np.ascontiguousarray(feats, np.float32)
        self.n = feats.shape[0]
        self.h = L.ref_create(feats, self.n)
        if not self.h:
            raise RuntimeError("reference initialize() failed")

    @staticmethod
    def available(gpu: bool = False) -> bool:
        return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libref_gpu.so" if gpu else "libref_cpu.so"))

    def gpu_enabled(self) -> bool:
        return bool(self.L.ref_gpu_enabled(self.h))

    def by_index(self, idx: int, k: int) -> np.ndarray:
        out = np.empty(max(1, min(max(k, 1), self.n)), np.int32)
        n = self.L.ref_recommend_by_index(self.h, idx, k, out)
        return out[:n]

    def by_name(self, name: str, k: int) -> np.ndarray:
        out = np.empty(max(1, min(max(k, 1), self.n)), np.int32)
        n = self.L.ref_recommend_by_name(self.h, name.encode(), k, out)
        return out[:n]

    def by_id(self, tid: str, k: int) -> np.ndarray:
        out = np.empty(max(1, min(max(k, 1), self.n)), np.int32)
        n = self.L.ref_recommend_by_id(self.h, tid.encode(), k, out)
        return out[:n]

    def scores(self, idx: int) -> np.ndarray:
        out = np.empty(self.n, np.float32)
        self.L.ref_scores(self.h, idx, out)
        return out

    def batch(self, qidx, k: int, threads: int = 1) -> np.ndarray:
        qidx = np.ascontiguousarray(qidx, np.int32)
        out = np.empty((qidx.size, k), np.int32)
        self.L.ref_batch_by_index(self.h, qidx, qidx.size, k, out, threads)
        return out

    def mt19937_uniform(self, count: int, seed: int = 42) -> np.ndarray:
        out = np.empty(count, np.float32)
        self.L.ref_gen_mt19937_uniform(count, seed, out)
        return out

    def close(self) -> None:
        if self.h:
            self.L.ref_destroy(self.h)
            self.h = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass
