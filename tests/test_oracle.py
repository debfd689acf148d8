"""Pins the C oracle (oracle/cosine_topk_oracle.c) on the reference:
golden vectors captured from the unmodified reference CPU build (SURVEY.md
Appendix A + tests/golden/make_golden.py) and, when oracle/_ref is present, the
live reference class itself.  CPU only."""
import numpy as np
import pytest

from helpers import (assert_same_up_to_ties, canonical_from_scores, case_features, from_bits,
                     load_golden)
from oracle_lib import Reference
from spotify_recommender_b200 import synth

CASES = load_golden()


def test_oracle_built_without_fma(oracle):
    assert oracle.unfused()


def test_mt19937_generator_matches_libstdcxx():
    # first values of std::mt19937(42) + uniform_real_distribution<float>
    v = synth.mt19937_uniform(5, 42)
    assert np.allclose(v, [0.37454012, 0.796543, 0.9507143, 0.18343478, 0.7319939], atol=0, rtol=1e-7)
    if Reference.available():
        r = Reference(np.zeros((2, 12), np.float32))
        assert np.array_equal(r.mt19937_uniform(50000, 42), synth.mt19937_uniform(50000, 42))


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_scores_and_selection(oracle, case):
    feats = case_features(case)
    for call in case["calls"]:
        q, k = call["q"], call["k"]
        sc = oracle.scores(feats, feats[q])
        if "score_bits" in call:  # bit-exact against Recommender.cu:256-273
            assert np.array_equal(sc.view(np.uint32), np.array(call["score_bits"], np.uint32))
        else:
            probe = np.array(call["probe_idx"])
            assert np.array_equal(sc[probe].view(np.uint32), np.array(call["probe_score_bits"], np.uint32))
        ref_idx = np.array(call["ref_idx"], np.int32)
        # the reference's own order, heap artefact included (Recommender.cu:293-315)
        assert np.array_equal(oracle.topk_refheap(sc, q, k), ref_idx)
        # canonical order agrees per tie group, and is what numpy lexsort says
        ci, cs, n = oracle.topk_canonical(sc, q, k)
        assert n == min(k, feats.shape[0] - 1) == ref_idx.size
        assert_same_up_to_ties(ci[:n], ref_idx, sc, exclude=q)
        ni, ns = canonical_from_scores(sc, q, k)
        assert np.array_equal(ci[:n], ni) and np.array_equal(cs[:n], ns)
        assert np.all(ci[n:] == -1)


def test_appendix_a_literals(oracle):
    """SURVEY.md Appendix A rows, typed in by hand (independent of the JSON)."""
    a = np.full((12, 12), 0.5, np.float32)
    sc = oracle.scores(a, a[0])
    assert oracle.topk_refheap(sc, 0, 5).tolist() == [4, 2, 5, 3, 1]
    assert oracle.topk_canonical(sc, 0, 5)[0].tolist() == [1, 2, 3, 4, 5]
    sc6 = oracle.scores(a, a[6])
    assert oracle.topk_refheap(sc6, 6, 4).tolist() == [3, 1, 2, 0]
    b = np.zeros((16, 12), np.float32)
    b[:, :2] = 1.0
    b[0] = 0; b[0, 0] = 1
    b[13] = 0; b[13, 0] = 2
    b[15] = 0; b[15, 0] = 3
    sc = oracle.scores(b, b[0])
    assert oracle.topk_refheap(sc, 0, 4).tolist() == [13, 15, 2, 4]
    assert oracle.topk_canonical(sc, 0, 4)[0].tolist() == [13, 15, 1, 2]
    assert oracle.topk_canonical(sc, 0, 6)[0].tolist() == [13, 15, 1, 2, 3, 4]
    idx, _, n = oracle.topk_canonical(sc, 0, 20)
    assert n == 15 and idx[15:].tolist() == [-1] * 5
    f = synth.mt19937_uniform(114000 * 12, 42).reshape(-1, 12)
    oi, _ = oracle.query_index(f, [0], 10)
    assert oi[0].tolist() == [35723, 105294, 12393, 25136, 35929, 105510, 9246, 64817, 2013, 59538]


def test_edge_semantics(oracle):
    f = synth.adversarial(512)
    # zero query: every score is exactly 0 (Recommender.cu:271) => index order
    oi, os_ = oracle.query_index(f, [17], 6)
    assert oi[0].tolist() == [0, 1, 2, 3, 4, 5] and np.all(os_ == 0)
    # k <= 0 returns nothing; k > n-1 returns n-1
    sc = oracle.scores(f, f[3])
    assert oracle.topk_canonical(sc, 3, 0)[2] == 0
    idx, _, n = oracle.topk_canonical(sc, 3, 600)
    assert n == 511 and 3 not in idx[:n].tolist() and np.all(idx[n:] == -1)
    # duplicates of the query stay in (self excluded by index only), clamp => ties at 1.0
    top = oracle.topk_canonical(sc, 3, 6)
    assert set([5, 256]).issubset(set(top[0].tolist())) and top[1][0] == 1.0
    assert sc[40] == -1.0 and sc[17] == 0.0 and sc[31] == 0.0 and sc[30] != 0.0


def test_threads_do_not_change_results(oracle):
    f = synth.features(50000)
    q = synth.query_indices(9, 50000)
    a = oracle.query_index(f, q, 25, threads=1)
    b = oracle.query_index(f, q, 25, threads=4)
    c = oracle.query_index(f, q[:2], 25, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
    assert np.array_equal(a[0][:2], c[0])


def test_sharded_merge_equals_whole(oracle):
    """CPU statement of the row-shard + merge step: any split gives
    the identical list."""
    n, k = 30011, 40
    f = synth.adversarial(n)
    qi = np.array([3, 99, 205, 17, 29000], np.int32)
    whole = oracle.query_index(f, qi, k)
    for parts in (2, 3, 8):
        bounds = np.linspace(0, n, parts + 1).astype(np.int64)
        pi, ps = [], []
        for p in range(parts):
            lo, hi = bounds[p], bounds[p + 1]
            ex = np.where((qi >= lo) & (qi < hi), qi - lo, -1).astype(np.int64)
            i_, s_ = oracle.query_rows(f[lo:hi], f[qi], ex, k, id_base=int(lo))
            pi.append(i_); ps.append(s_)
        mi, ms = oracle.merge_parts(np.stack(pi), np.stack(ps))
        assert np.array_equal(mi, whole[0])
        assert np.array_equal(ms.view(np.uint32), whole[1].view(np.uint32))


@pytest.mark.skipif(not Reference.available(), reason="oracle/_ref not built (no /root/reference)")
def test_live_reference_agrees(oracle):
    rng = np.random.default_rng(5)
    for n, k in ((1, 3), (2, 1), (33, 7), (5000, 64)):
        f = (np.floor(rng.random((n, 12)) * 100) / 100).astype(np.float32)  # coarse => many ties
        ref = Reference(f)
        for q in sorted({0, n // 2, n - 1}):
            sc = oracle.scores(f, f[q])
            assert np.array_equal(ref.scores(q).view(np.uint32), sc.view(np.uint32))
            got = ref.by_index(q, k)
            assert np.array_equal(got, oracle.topk_refheap(sc, q, k))
            ci, _, cn = oracle.topk_canonical(sc, q, k)
            assert_same_up_to_ties(ci[:cn], got, sc, exclude=q)
        ref.close()


def test_native_libraries_export_every_declared_symbol():
    """The C-ABI library loads (no GPU needed) and exports every entry point include/sr_engine.h
    declares; same for the C API around the C++ Recommender class."""
    import ctypes, os, re
    from spotify_recommender_b200 import build, engine
    import recommender_lib
    build.build_all()
    hdr = open(os.path.join(build.ROOT, "include", "sr_engine.h")).read()
    declared = sorted(set(re.findall(r"\b(sr_(?:engine_)?[a-z0-9_]+)\s*\(", hdr)))
    assert set(declared) == set(engine.EXPORTS)
    lib = engine.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    rl = ctypes.CDLL(recommender_lib.SO)
    for name in recommender_lib.EXPORTS:
        assert hasattr(rl, name), name


def test_engine_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without an sm_100 device construction raises."""
    import torch
    from spotify_recommender_b200.engine import Engine, EngineError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(EngineError):
        Engine(0)
