"""Seeded random cases against the oracle: store size, batch size, k, data family, kernel shape, internal batch and
tile options all drawn at random, results compared bit for bit (index lists and score bits) -- through the single
engine and through the single-process sharded ABI with several shards on this GPU.  The fixed cases of the other
test files pin known edges; this one looks for the combinations nobody thought of."""
import numpy as np
import pytest

from spotify_recommender_b200 import synth

pytestmark = pytest.mark.gpu


def same_scores(got, want):
    """Bit-identical, except that a -0.0 of the reference comes back as +0.0 (keys fold signed zeros so that they
    tie, sr_device.cuh make_key; only hostile inputs -- an infinite denominator -- produce one)."""
    return bool(np.all((got.view(np.uint32) == want.view(np.uint32)) | ((got == 0) & (want == 0))))


def _data(rng, kind, n):
    if kind == "features":
        return synth.features(n)
    if kind == "uniform":
        return synth.uniform(n, seed=int(rng.integers(1, 1 << 30)))
    if kind == "hostile":
        # values no preprocessing would produce, sprinkled over ordinary rows: NaN, +-inf, huge, tiny, negative
        f = synth.uniform(n, seed=int(rng.integers(1, 1 << 30)))
        specials = np.array([np.nan, np.inf, -np.inf, 3e38, -3e38, 1e30, 1e-30, 1e-45, -0.0, -1.0], np.float32)
        m = max(1, n // 50)
        rows = rng.integers(0, n, m)
        cols = rng.integers(0, 12, m)
        f[rows, cols] = specials[rng.integers(0, specials.size, m)]
        return f
    if kind == "adversarial":
        return np.tile(synth.adversarial(max(256, min(n, 4096))), (n // 256 + 1, 1))[:n].copy() if n >= 256 else synth.uniform(n)
    # clustered: a few tight clusters in index order (genre-sorted stores look like this to a scan)
    c = max(1, int(rng.integers(2, 40)))
    centers = rng.random((c, 12), dtype=np.float32)
    f = centers[(np.arange(n) * c) // max(n, 1)] + np.float32(0.01) * rng.random((n, 12), dtype=np.float32)
    return np.ascontiguousarray(f, np.float32)


def _case(rng):
    n = int(rng.choice([1, 2, 7, 100, 1000, 4095, 4097, 20_000, 70_000, 150_000, 300_000]))
    n = max(1, n + int(rng.integers(-3, 4)) if n > 10 else n)
    nq = int(rng.choice([1, 2, 5, 16, 31, 33, 64, 100, 192, 193, 257, 700, 1281, 2600]))
    k = int(rng.choice([1, 2, 9, 10, 16, 17, 33, 64, 72, 73, 100, 128, 255, 256, 300, 1024, 1100]))
    kind = str(rng.choice(["features", "uniform", "adversarial", "clustered", "hostile"]))
    return n, nq, k, kind


@pytest.mark.parametrize("seed", range(96))
def test_random_case_equals_oracle(oracle, seed):
    from spotify_recommender_b200.engine import Engine, EngineError, variant_names
    rng = np.random.default_rng(1000 + seed)
    n, nq, k, kind = _case(rng)
    f = _data(rng, kind, n)
    q = rng.integers(0, n, nq).astype(np.int32)
    want = oracle.query_index(f, q, k, threads=8)
    with Engine(0) as e:
        e.load_features(f)
        opts = {}
        if rng.random() < 0.4:
            opts["variant"] = int(rng.integers(0, len(variant_names())))
        if rng.random() < 0.3:
            opts["batch"] = int(rng.choice([17, 300, 1000, 8192]))
        if rng.random() < 0.3:
            opts["qt"] = int(rng.choice([3, 16, 100, 256]))
        if rng.random() < 0.2:
            opts["bound"] = 0
        if rng.random() < 0.2:
            opts["graphs"] = 0
        for key, v in opts.items():
            e.set_option(key, v)
        try:
            got = e.query_by_index(q, k)
        except EngineError as exc:  # an explicit shape may not have the shared memory for this k: that must be said, not guessed
            assert "variant" in opts and ("does not fit" in str(exc) or "query tiles exceed" in str(exc)), (opts, str(exc))
            return
        assert np.array_equal(got[0], want[0]), (n, nq, k, kind, opts, np.argwhere(got[0] != want[0])[:3])
        assert same_scores(got[1], want[1]), (n, nq, k, kind, opts)
        again = e.query_by_index(q, k)  # and once more (graph replay / warm workspace)
        assert np.array_equal(again[0], want[0]) and same_scores(again[1], want[1])


@pytest.mark.parametrize("seed", range(24))
def test_random_sharded_case_equals_oracle(oracle, seed):
    from spotify_recommender_b200.engine import ShardedEngine
    rng = np.random.default_rng(5000 + seed)
    shards = int(rng.choice([2, 3, 5, 8]))
    n = int(rng.choice([shards * 3, 5000, 40_003, 200_000]))
    nq = int(rng.choice([1, 7, 64, 500, 8300]))
    k = int(rng.choice([1, 10, 50, 100, 255, 256, 1030]))
    kind = str(rng.choice(["features", "uniform", "adversarial", "clustered", "hostile"]))
    f = _data(rng, kind, n)
    q = rng.integers(0, n, nq).astype(np.int32)
    want = oracle.query_index(f, q, k, threads=8)
    with ShardedEngine([0] * shards) as se:
        se.load_features(f)
        got = se.query_by_index(q, k)
    assert np.array_equal(got[0], want[0]), (n, nq, k, kind, shards, np.argwhere(got[0] != want[0])[:3])
    assert same_scores(got[1], want[1]), (n, nq, k, kind, shards)


@pytest.mark.parametrize("seed", range(24))
def test_random_vector_queries_on_a_row_shard(oracle, seed):
    """query_by_vector on a shard with an id base: arbitrary (also hostile) query rows, exclusions inside and
    outside the shard, against the oracle's query_rows."""
    from spotify_recommender_b200.engine import Engine
    rng = np.random.default_rng(9000 + seed)
    n = int(rng.choice([50, 4096, 30_000, 120_000]))
    base = int(rng.choice([0, 1, 777_777, 2_000_000_000]))
    nq = int(rng.choice([1, 3, 40, 300, 1500]))
    k = int(rng.choice([1, 10, 40, 100, 500, 1030]))
    f = _data(rng, str(rng.choice(["features", "uniform", "clustered", "hostile"])), n)
    qv = _data(rng, str(rng.choice(["uniform", "hostile"])), nq)
    qv[rng.integers(0, nq)] = f[rng.integers(0, n)]          # one query IS a song of the store
    if rng.random() < 0.5:
        qv[rng.integers(0, nq)] = 0.0                          # a zero query: every score 0, order by id
    ex = rng.integers(base - 5, base + n + 5, nq).astype(np.int64)
    ex[rng.random(nq) < 0.3] = -1
    ex_local = np.where((ex >= base) & (ex < base + n), ex - base, -1)
    want = oracle.query_rows(f, qv, ex_local, k, id_base=base, threads=8)
    with Engine(0) as e:
        e.load_features(f, id_base=base)
        got = e.query_by_vector(qv, k, exclude=ex.astype(np.int32))
    assert np.array_equal(got[0], want[0]), (n, base, nq, k, np.argwhere(got[0] != want[0])[:3])
    assert same_scores(got[1], want[1]), (n, base, nq, k)
