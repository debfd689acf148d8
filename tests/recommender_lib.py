"""ctypes access to libsr_recommender.so: the C API around the C++ `Recommender` class
of include/sr_recommender.hpp (the drop-in for reference Recommender.h:28-82)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "spotify_recommender_b200", "libsr_recommender.so")
EXPORTS = ["sr_recommender_create", "sr_recommender_create_from_file", "sr_recommender_destroy", "sr_recommender_song_count", "sr_recommender_gpu_enabled",
           "sr_recommender_by_index", "sr_recommender_by_name", "sr_recommender_by_id", "sr_recommender_find_name",
           "sr_recommender_find_id"]
_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def load():
    L = C.CDLL(SO)
    L.sr_recommender_create.restype = C.c_void_p
    L.sr_recommender_create.argtypes = [_f32p, C.c_int64, C.c_void_p, C.c_void_p]
    L.sr_recommender_create_from_file.restype = C.c_void_p
    L.sr_recommender_create_from_file.argtypes = [C.c_char_p]
    L.sr_recommender_destroy.argtypes = [C.c_void_p]
    L.sr_recommender_song_count.argtypes = [C.c_void_p]
    L.sr_recommender_gpu_enabled.argtypes = [C.c_void_p]
    L.sr_recommender_by_index.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p]
    L.sr_recommender_by_name.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _i32p]
    L.sr_recommender_by_id.argtypes = [C.c_void_p, C.c_char_p, C.c_int, _i32p]
    L.sr_recommender_find_name.argtypes = [C.c_void_p, C.c_char_p]
    L.sr_recommender_find_id.argtypes = [C.c_void_p, C.c_char_p]
    return L


class HostRecommender:
    @classmethod
    def from_file(cls, path: str):
        self = cls.__new__(cls)
        self.L = load()
        self.h = self.L.sr_recommender_create_from_file(path.encode())
        if not self.h:
            raise RuntimeError(f"initializeFromFile({path}) failed")
        self.n = int(self.L.sr_recommender_song_count(self.h))
        return self

    def __init__(self, feats, ids=None, names=None):
        self.L = load()
        feats = np.ascontiguousarray(feats, np.float32)
        self.n = feats.shape[0]

        def arr(strings):
            if strings is None:
                return None, None
            keep = [s.encode() for s in strings]
            a = (C.c_char_p * len(keep))(*keep)
            return a, keep
        ia, self._k1 = arr(ids)
        na, self._k2 = arr(names)
        self.h = self.L.sr_recommender_create(feats, self.n, C.cast(ia, C.c_void_p) if ia else None,
                                              C.cast(na, C.c_void_p) if na else None)
        if not self.h:
            raise RuntimeError("Recommender::initialize failed (no sm_100 device?)")

    def _call(self, fn, arg, k):
        out = np.empty(max(1, min(max(k, 1), self.n)), np.int32)
        n = fn(self.h, arg, k, out)
        return out[:n]

    def by_index(self, idx, k):
        return self._call(self.L.sr_recommender_by_index, idx, k)

    def by_name(self, name, k):
        return self._call(self.L.sr_recommender_by_name, name.encode(), k)

    def by_id(self, tid, k):
        return self._call(self.L.sr_recommender_by_id, tid.encode(), k)

    def find_name(self, name):
        return int(self.L.sr_recommender_find_name(self.h, name.encode()))

    def find_id(self, tid):
        return int(self.L.sr_recommender_find_id(self.h, tid.encode()))

    def gpu_enabled(self):
        return bool(self.L.sr_recommender_gpu_enabled(self.h))

    def close(self):
        if self.h:
            self.L.sr_recommender_destroy(self.h)
            self.h = None
