#!/usr/bin/env python
"""Regenerates tests/golden/reference_kat.json by RUNNING THE UNMODIFIED
REFERENCE (oracle/_ref/libref_cpu.so, built by oracle/Makefile from
/root/reference).  Run in the build container only:

    make -C oracle && python tests/golden/make_golden.py

Every case stores the inputs (or their seed), the reference's returned index
list (its heap order, ties included) and, where useful, raw score bit patterns
from the reference's calculateSimilaritiesCPU (Recommender.cu:256-273).
Cases A*/B*/C*/D* are SURVEY.md Appendix A.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle_lib import Reference  # noqa: E402
from spotify_recommender_b200 import synth  # noqa: E402


def bits(a):
    return [int(x) for x in np.asarray(a, np.float32).view(np.uint32)]


def small_case(name, feats, calls, with_scores=True):
    feats = np.ascontiguousarray(feats, np.float32)
    ref = Reference(feats)
    out = {"name": name, "n": int(feats.shape[0]), "features_bits": bits(feats.ravel()), "calls": []}
    for q, k in calls:
        c = {"q": int(q), "k": int(k), "ref_idx": [int(x) for x in ref.by_index(q, k)]}
        if with_scores:
            c["score_bits"] = bits(ref.scores(q))
        out["calls"].append(c)
    ref.close()
    return out


def seeded_case(name, gen, n, calls, n_score_probe=64):
    feats = gen(n)
    ref = Reference(feats)
    out = {"name": name, "n": int(n), "generator": gen.__name__, "calls": []}
    for q, k in calls:
        sc = ref.scores(q)
        idx = ref.by_index(q, k)
        probe = np.unique(np.concatenate([idx, np.linspace(0, n - 1, n_score_probe).astype(np.int64)]))
        out["calls"].append({"q": int(q), "k": int(k), "ref_idx": [int(x) for x in idx],
                             "probe_idx": [int(x) for x in probe], "probe_score_bits": bits(sc[probe])})
    ref.close()
    return out


def gen_mt_uniform_114k(n):
    return synth.mt19937_uniform(n * 12, 42).reshape(n, 12)


def gen_synth_spotify(n):
    return synth.features(n)


def gen_adversarial(n):
    return synth.adversarial(n)


def main():
    cases = []
    a = np.full((12, 12), 0.5, np.float32)
    cases.append(small_case("A_all_equal", a, [(0, 5), (0, 11), (6, 4)]))
    b = np.zeros((16, 12), np.float32)
    b[:, 0] = 1.0
    b[:, 1] = 1.0
    b[0] = 0.0; b[0, 0] = 1.0
    b[13] = 0.0; b[13, 0] = 2.0
    b[15] = 0.0; b[15, 0] = 3.0
    cases.append(small_case("B_ties_and_scaled", b, [(0, 4), (0, 6), (0, 20)]))
    c = np.zeros((5, 12), np.float32)
    for i in range(5):
        for j in range(12):
            c[i, j] = 0.0 if i == 2 else np.float32(0.1 * (i + 1) + 0.01 * j * (i % 2))
    cases.append(small_case("C_zero_vector", c, [(0, 4), (2, 4)]))
    rng = np.random.default_rng(123)
    r = (np.floor(rng.random((257, 12)) * 1000) / 1000).astype(np.float32)
    cases.append(small_case("R_random_257", r, [(0, 1), (3, 10), (200, 100), (256, 300)]))
    adv = synth.adversarial(512)
    cases.append(small_case("ADV_512", adv, [(3, 10), (17, 5), (31, 8), (99, 70), (200, 40), (40, 3)]))
    cases.append(seeded_case("D1_mt19937_114000", gen_mt_uniform_114k, 114000, [(0, 10), (13, 100)]))
    cases.append(seeded_case("D2_mt19937_1000000", gen_mt_uniform_114k, 1000000, [(0, 10)]))
    cases.append(seeded_case("S_synth_spotify_200000", gen_synth_spotify, 200000, [(13, 10), (7932, 100)]))
    cases.append(seeded_case("ADV_4096", gen_adversarial, 4096, [(3, 10), (99, 100), (205, 50)]))
    with open(os.path.join(HERE, "reference_kat.json"), "w") as fh:
        json.dump({"source": "unmodified reference CPU build (oracle/_ref/libref_cpu.so), g++ 13.3 -O3, x86-64",
                   "cases": cases}, fh, separators=(",", ":"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
