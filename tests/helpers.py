"""Shared parity helpers (test infrastructure)."""
from __future__ import annotations

import json
import os

import numpy as np

from spotify_recommender_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_kat.json")
_GENERATORS = {
    "gen_mt_uniform_114k": lambda n: synth.mt19937_uniform(n * 12, 42).reshape(n, 12),
    "gen_synth_spotify": lambda n: synth.features(n),
    "gen_adversarial": lambda n: synth.adversarial(n),
}


def load_golden():
    with open(GOLDEN) as fh:
        return json.load(fh)["cases"]


def case_features(case) -> np.ndarray:
    if "features_bits" in case:
        return np.array(case["features_bits"], np.uint32).view(np.float32).reshape(case["n"], 12).copy()
    return np.ascontiguousarray(_GENERATORS[case["generator"]](case["n"]), np.float32)


def from_bits(b) -> np.ndarray:
    return np.array(b, np.uint32).view(np.float32)


def assert_same_up_to_ties(got_idx, want_idx, scores, exclude=-1):
    """Two result lists agree "per tie group": identical length and score
    sequence, no duplicates, self excluded, and identical membership for every
    score strictly above the last listed score (members AT the cut may differ
    between the reference's heap artefact and canonical order: SURVEY App. A)."""
    got = np.asarray(got_idx)
    want = np.asarray(want_idx)
    got = got[got >= 0]
    want = want[want >= 0]
    assert got.size == want.size, (got, want)
    if got.size == 0:
        return
    assert len(set(got.tolist())) == got.size
    assert exclude not in got.tolist()
    sg, sw = scores[got], scores[want]
    assert np.array_equal(sg.view(np.uint32), sw.view(np.uint32)), (sg, sw)
    cut = sw[-1]
    assert set(got[sg > cut].tolist()) == set(want[sw > cut].tolist())


def canonical_from_scores(scores, exclude, k, id_base=0):
    """numpy statement of the canonical order: score desc, index asc."""
    n = scores.size
    idx = np.arange(n)
    keep = idx != exclude
    idx = idx[keep]
    order = np.lexsort((idx, -scores[idx].astype(np.float64)))
    top = idx[order][:max(k, 0)]
    return (top + id_base).astype(np.int32), scores[top]
