"""GPU parity tests proper: the CUDA engine, called through the C ABI
(include/sr_engine.h via spotify_recommender_b200.engine), against the CPU oracle
(oracle/cosine_topk_oracle.c, pinned on the reference in test_oracle.py) and the
committed golden vectors.  Index lists must be bit-exact; scores must be
bit-identical to the oracle (the north_star tolerance is 1e-6 absolute, written
below where the reference's own GPU path is the comparison)."""
import numpy as np
import pytest

from helpers import case_features, from_bits, load_golden
from spotify_recommender_b200 import synth

pytestmark = pytest.mark.gpu
SCORE_TOL = 1e-6  # north_star: scores within 1e-6 absolute


@pytest.fixture()
def eng():
    """A fresh engine per test: every test starts from the default options (kernel shape chosen by the engine)."""
    from spotify_recommender_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


AUTO_SHAPES = (4, 5, 8)  # the shapes the engine picks by itself: TMA-staged small-batch, dynamic large-batch, two-CTA dynamic mid-batch


def assert_exact(got, want):
    gi, gs = got
    wi, ws = want
    assert np.array_equal(gi, wi), f"index lists differ at {np.argwhere(gi != wi)[:5]}"
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32)), "scores are not bit-identical to the oracle"


def test_division_matches_ieee(eng):
    """The device-side division (double Newton + one rounding, no subroutine call) is the
    host's IEEE divss bit for bit: random operands over the whole exponent range, cosine-
    like operands, subnormal quotients, and the special values."""
    rng = np.random.default_rng(11)
    n = 4_000_000
    a = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32)
    b = np.abs(rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32).view(np.float32))
    b[~(b > 0)] = 1.0
    a[:1_000_000] = (rng.random(1_000_000, dtype=np.float32) * 2 - 1) * b[:1_000_000]        # |q| <= 1
    a[1_000_000:1_200_000] = b[1_000_000:1_200_000] * np.float32(1e-38) * rng.random(200_000, dtype=np.float32)
    edge = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 1e-45, -1e-45, 3.4028235e38, 1.1754944e-38], np.float32)
    eb = np.array([1e-8, 1.0, 3.0, np.inf, 3.4028235e38, 1.1754944e-38, 1e-45, 7.0, 0.1], np.float32)
    a = np.concatenate([a, np.repeat(edge, eb.size)])
    b = np.concatenate([b, np.tile(eb, edge.size)])
    got = eng.selftest_div(a, b)
    with np.errstate(all="ignore"):
        want = a / b
    nan = np.isnan(want)
    assert np.array_equal(np.isnan(got), nan)
    assert np.array_equal(got[~nan].view(np.uint32), want[~nan].view(np.uint32))


CASES = load_golden()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_vectors(eng, oracle, case):
    """Every golden call: canonical list == oracle's, per-tie-group == the reference's."""
    from helpers import assert_same_up_to_ties
    feats = case_features(case)
    eng.load_features(feats)
    for call in case["calls"]:
        q, k = call["q"], call["k"]
        gi, gs = eng.query_by_index([q], k)
        wi, ws = oracle.query_index(feats, [q], k)
        assert_exact((gi, gs), (wi, ws))
        sc = oracle.scores(feats, feats[q])
        n = min(k, feats.shape[0] - 1)
        assert np.all(gi[0, n:] == -1)
        assert_same_up_to_ties(gi[0, :n], np.array(call["ref_idx"], np.int32), sc, exclude=q)
        if "score_bits" in call:  # raw reference scores, bit for bit
            ref_sc = from_bits(call["score_bits"])
            assert np.array_equal(gs[0, :n].view(np.uint32), ref_sc[gi[0, :n]].view(np.uint32))


@pytest.mark.parametrize("variant", range(10))
def test_every_kernel_shape(eng, oracle, variant):  # k = 1, 10, 100; 257 queries
    n = 150_000
    f = synth.features(n)
    eng.set_option("variant", variant)
    try:
        eng.load_features(f)
        q = synth.query_indices(257, n)
        for k in (1, 10, 100):
            assert_exact(eng.query_by_index(q, k), oracle.query_index(f, q, k, threads=8))
    finally:
        eng.set_option("variant", -1)


@pytest.mark.parametrize("n,nq,k", [(1, 1, 1), (2, 2, 1), (5, 5, 4), (33, 33, 7), (1000, 64, 10), (4097, 130, 100),
                                    (40_000, 1500, 10), (200_000, 300, 333), (100_000, 1, 1024), (60_000, 30, 100),
                                    (300_000, 600, 50), (150_000, 1300, 64)])  # last two: lists in the L2 workspace
def test_sizes_and_ragged_batches(eng, oracle, n, nq, k):
    f = synth.features(n)
    eng.load_features(f)
    q = synth.query_indices(nq, n)
    assert_exact(eng.query_by_index(q, k), oracle.query_index(f, q, k, threads=8))
    assert eng.stat("variant") in AUTO_SHAPES  # the shapes the engine selects by itself are the ones under test


@pytest.mark.parametrize("k", [40, 50, 64, 72])
def test_l2_list_workspace_path(eng, oracle, k):
    """16 < k <= 72 with full 256-query tiles: the CTAs' lists live in the L2 workspace (settles pull them into
    registers and write them back).  Uniform and clustered data, ties included."""
    n = 200_000
    for f in (synth.uniform(n), synth.features(n), np.tile(synth.adversarial(4096), (49, 1))[:n]):
        eng.load_features(f)
        q = synth.query_indices(1024, n)  # four full 256-query tiles
        assert_exact(eng.query_by_index(q, k), oracle.query_index(f, q, k, threads=8))
        assert eng.stat("variant") == 5 and eng.stat("lists_in_smem") == 0, "the L2-resident list workspace was not taken"


def test_adversarial_ties_zero_rows_irregular(eng, oracle):
    """Duplicates, scaled copies (clamped cos == 1 ties), an all-zero row and query,
    tiny-norm rows either side of the 1e-8 denominator cut, mass ties."""
    for n in (512, 4096, 30011):
        f = synth.adversarial(n)
        eng.load_features(f)
        q = np.array([3, 5, 17, 21, 30, 31, 32, 40, 41, 99, 100, 163, 200, 231, n - 1], np.int32)
        for k in (1, 6, 40, 70, 300):
            assert_exact(eng.query_by_index(q, k), oracle.query_index(f, q, k))
            assert eng.stat("variant") in AUTO_SHAPES
        # the same through the large-batch shape (a batch of > 40 queries), the queries repeated
        q2 = np.tile(q, 20)
        for k in (6, 50, 100):
            assert_exact(eng.query_by_index(q2, k), oracle.query_index(f, q2, k, threads=8))
            assert eng.stat("variant") == 5
    assert eng.stat("irregular_songs") >= 2


def test_mass_ties_whole_store(eng, oracle):
    """Every song identical: all scores tie, order is by index, k > n-1 pads with -1."""
    f = np.full((5000, 12), 0.5, np.float32)
    eng.load_features(f)
    q = np.array([0, 6, 4999], np.int32)
    assert_exact(eng.query_by_index(q, 11), oracle.query_index(f, q, 11))
    gi, _ = eng.query_by_index(q, 11)
    assert gi[0].tolist() == list(range(1, 12)) and gi[1].tolist() == [0, 1, 2, 3, 4, 5, 7, 8, 9, 10, 11]
    assert eng.stat("rescans") > 0  # the overflow path was exercised


def test_query_by_vector_and_row_shard_ids(eng, oracle):
    """A row shard answers with global ids and excludes by global id."""
    n, lo, hi, k = 90_000, 30_000, 61_234, 25
    f = synth.features(n)
    qi = np.array([5, 30_000, 45_678, 61_233, 89_999], np.int32)
    eng.load_features(f[lo:hi], id_base=lo)
    got = eng.query_by_vector(f[qi], k, exclude=qi)
    ex = np.where((qi >= lo) & (qi < hi), qi - lo, -1).astype(np.int64)
    want = oracle.query_rows(f[lo:hi], f[qi], ex, k, id_base=lo)
    assert_exact(got, want)
    # arbitrary (not in-store) query vectors, nothing excluded
    rng = np.random.default_rng(3)
    qv = rng.random((40, 12), dtype=np.float32)
    assert_exact(eng.query_by_vector(qv, k), oracle.query_rows(f[lo:hi], qv, None, k, id_base=lo))


def test_errors_are_loud(eng):
    from spotify_recommender_b200.engine import Engine, EngineError
    e2 = Engine(0)
    with pytest.raises(EngineError):  # no store yet
        e2.query_by_index([0], 5)
    e2.load_features(synth.features(100))
    for bad_k in (0, -3):
        with pytest.raises(EngineError):
            e2.query_by_index([0], bad_k)
    with pytest.raises(EngineError):  # not an owned song
        e2.query_by_index([100], 5)
    with pytest.raises(EngineError):
        e2.set_option("no_such_option", 1)
    e2.close()


def test_threshold_sharing_does_not_change_results(eng, oracle):
    """Bootstrap sample on/off, tiny query tiles, tiny batches: identical output."""
    n = 120_000
    f = synth.features(n)
    eng.load_features(f)
    q = synth.query_indices(200, n)
    want = oracle.query_index(f, q, 50, threads=8)
    try:
        for sample, qt, batch in ((0, 128, 8192), (4096, 7, 8192), (-1, 128, 33), (64, 1, 64)):
            eng.set_option("sample", sample); eng.set_option("qt", qt); eng.set_option("batch", batch)
            assert_exact(eng.query_by_index(q, 50), want)
    finally:
        eng.set_option("sample", -1); eng.set_option("qt", 256); eng.set_option("batch", 8192)


def test_full_size_properties(eng, oracle):
    """BASELINE config sizes (1M / batch 1024 / top-10) through size-independent
    properties plus an oracle spot check: sorted (score desc, id asc), self excluded,
    no duplicates, batch == single-query, and the returned scores recompute exactly."""
    n, nq, k = 1_000_000, 1024, 10
    f = synth.features(n)
    eng.load_features(f)
    q = synth.query_indices(nq, n)
    gi, gs = eng.query_by_index(q, k)
    assert gi.min() >= 0 and gi.max() < n
    assert np.all(gi != q[:, None])
    assert all(len(set(r.tolist())) == k for r in gi)
    ds = np.diff(gs.astype(np.float64), axis=1)
    assert np.all(ds <= 0)
    tie = ds == 0
    assert np.all(np.diff(gi.astype(np.int64), axis=1)[tie] > 0)
    for j in (0, 511, 1023):  # returned scores are the oracle's scores of those songs
        sc = oracle.scores(f, f[q[j]], threads=8)
        assert np.array_equal(sc[gi[j]].view(np.uint32), gs[j].view(np.uint32))
    spot = np.array([0, 1, 500, 1023])
    assert_exact(eng.query_by_index(q[spot], k), oracle.query_index(f, q[spot], k, threads=8))
    one_i, one_s = eng.query_by_index(q[777:778], k)
    assert np.array_equal(one_i[0], gi[777]) and np.array_equal(one_s[0], gs[777])


def test_sharded_host_logic_single_rank_on_torch_stream(eng, oracle):
    """ShardedRecommender with one rank, device buffers owned by torch, work enqueued on
    torch's current (legacy default) stream, several batches back to back without a
    synchronise in between: equals the oracle."""
    import torch
    from spotify_recommender_b200.sharded import ShardedRecommender
    n, k = 300_000, 10
    f = synth.features(n)
    sh = ShardedRecommender(eng, n, device=torch.device("cuda", 0))
    sh.load_shard(f)
    qs = [(((np.arange(700, dtype=np.int64) + b * 700) * 7919 + 13) % n).astype(np.int32) for b in range(4)]
    qd = [torch.from_numpy(q).cuda() for q in qs]
    outs = []
    for b in range(4):
        oi, os_ = sh.query_by_index_dev(qd[b], k)
        outs.append((oi.clone(), os_.clone()))
    torch.cuda.synchronize()
    for b in range(4):
        want = oracle.query_index(f, qs[b], k, threads=8)
        assert_exact((outs[b][0].cpu().numpy(), outs[b][1].cpu().numpy()), want)
        assert_exact(sh.query_by_index(qs[b], k), want)


def test_all_pairs_neighbour_table(eng, oracle):
    """BASELINE config 5 at test size: every song's top-K neighbours, streamed in batches."""
    n, k = 20_000, 10
    f = synth.features(n)
    eng.load_features(f)
    try:
        eng.set_option("batch", 3000)  # ragged last batch
        gi, gs = eng.all_pairs_topk(0, n, k)
        part_i, part_s = eng.all_pairs_topk(7000, 9000, k)  # a query shard of the table
    finally:
        eng.set_option("batch", 8192)
    want = oracle.query_index(f, np.arange(n, dtype=np.int32), k, threads=8)
    assert_exact((gi, gs), want)
    assert np.array_equal(part_i, want[0][7000:9000])
    # symmetry of the reference's arithmetic: s(i, j) == s(j, i) bit for bit
    j = gi[:, 0]
    back = np.array([oracle.scores(f[jj:jj + 1], f[i])[0] for i, jj in list(enumerate(j))[:200]], np.float32)
    assert np.array_equal(back.view(np.uint32), gs[:200, 0].view(np.uint32))


@pytest.mark.parametrize("k,nq", [(10, 64), (50, 600), (100, 300)])
def test_clustered_store_takes_the_refilter_path(eng, oracle, k, nq):
    """A genre-sorted store whose clusters the bound pass cannot see (bound off, tiny sample):
    tiles overflow their hit buffers and are re-filtered against the raised threshold.  (k = 50 keeps the
    CTA lists in the L2 workspace; re-filtered queries end their segment through the duplicate-checking
    settle, the others through the pending pass straight to the pool.)"""
    n = 300_000
    rng = np.random.default_rng(9)
    centers = rng.random((30, 12), dtype=np.float32)
    f = (centers[np.arange(n) // (n // 30 + 1)] + 0.02 * rng.random((n, 12), dtype=np.float32)).astype(np.float32)
    eng.load_features(f)
    q = synth.query_indices(nq, n)
    try:
        eng.set_option("bound", 0); eng.set_option("sample", 0); eng.set_option("reset", 1)
        got = eng.query_by_index(q, k)
        assert eng.stat("refilters") > 0
    finally:
        eng.set_option("bound", 1); eng.set_option("sample", -1)
    assert eng.stat("variant") in AUTO_SHAPES
    assert_exact(got, oracle.query_index(f, q, k, threads=8))
    assert_exact(eng.query_by_index(q, k), oracle.query_index(f, q, k, threads=8))


def test_foreign_query_id_on_the_device_api(eng, oracle):
    """sr_engine_query_by_index_dev cannot refuse an id before the work is enqueued: the row comes back -1 / 0,
    the other rows are untouched, and the next synchronising call reports SR_EINVAL (once)."""
    import torch
    from spotify_recommender_b200.engine import EngineError
    n, k = 50_000, 10
    f = synth.features(n)
    eng.load_features(f, id_base=1000)
    q = np.array([1000, 1000 + n - 1, 999, 1000 + n, 2000, -5], np.int32)
    dq = torch.from_numpy(q).cuda()
    oi = torch.full((q.size, k), 7, dtype=torch.int32, device="cuda")
    os_ = torch.full((q.size, k), 7.0, dtype=torch.float32, device="cuda")
    eng.query_by_index_dev(dq, q.size, k, oi, os_)
    with pytest.raises(EngineError) as exc:
        eng.synchronize()
    assert "not a song of this store" in str(exc.value)
    eng.synchronize()  # the flag was cleared by the report
    gi, gs = oi.cpu().numpy(), os_.cpu().numpy()
    good = np.array([0, 1, 4])
    wi, ws = oracle.query_index(f, q[good] - 1000, k)
    assert np.array_equal(gi[good], wi + 1000) and np.array_equal(gs[good].view(np.uint32), ws.view(np.uint32))
    bad = np.array([2, 3, 5])
    assert np.all(gi[bad] == -1) and np.all(gs[bad] == 0.0)
    eng.query_by_index_dev(dq, q.size, k, oi, os_)
    assert eng.stat("bad_index") == 1 and eng.stat("bad_index") == 0


@pytest.mark.parametrize("n,k", [(3000, 2999), (3000, 5000), (40_000, 1025), (40_000, 2500), (9000, 4096)])
def test_lists_longer_than_1024(eng, oracle, n, k):
    """topN beyond one pass's 1024 results: served 1024 at a time under a ceiling; min(k, n - 1) results like the
    reference (Recommender.cu:300-315), padded with -1."""
    f = synth.adversarial(n) if n == 9000 else synth.features(n)
    eng.load_features(f)
    q = np.array([3, 17, 100, 150, n - 1], np.int32)
    got = eng.query_by_index(q, k)
    want = oracle.query_index(f, q, k, threads=8)
    assert_exact(got, want)
    assert np.all(got[0][:, min(k, n - 1):] == -1)


@pytest.mark.parametrize("data", ["features", "uniform"])
@pytest.mark.parametrize("k", [10, 100])
def test_headline_size_parity(eng, oracle, data, k):
    """The sizes the headline is quoted on (BASELINE config 3): 10 M songs, a 4096-query batch, top-10 and
    top-100.  >= 48 sampled queries are compared bit-exactly (index lists and score bits) with the oracle run on
    all host threads; the batch runs three times and must be bit-identical every time (the scan's lock-free
    threshold sharing makes the WORK nondeterministic, never the result)."""
    n, nq = 10_000_000, 4096
    f = synth.features(n) if data == "features" else synth.uniform(n)
    eng.load_features(f)
    q = synth.query_indices(nq, n)
    first = eng.query_by_index(q, k)
    assert eng.stat("variant") == 5
    for _ in range(2):
        again = eng.query_by_index(q, k)
        assert np.array_equal(again[0], first[0]) and np.array_equal(again[1].view(np.uint32), first[1].view(np.uint32))
    sel = np.unique(np.concatenate([np.arange(0, nq, 86), [1, 2, nq - 1]]))
    assert sel.size >= 48
    want = oracle.query_index(f, q[sel], k, threads=oracle.max_threads)
    assert_exact((first[0][sel], first[1][sel]), want)
    # size-independent properties over the whole batch
    gi, gs = first
    assert gi.min() >= 0 and gi.max() < n and np.all(gi != q[:, None])
    ds = np.diff(gs.astype(np.float64), axis=1)
    assert np.all(ds <= 0) and np.all(np.diff(gi.astype(np.int64), axis=1)[ds == 0] > 0)


def test_small_passes_replayed_as_cuda_graphs(eng, oracle):
    """Single-query and small-batch calls go through one captured graph per (buffers, nq, k): the same buffers with
    new query ids replay it, a reloaded store or a changed option re-captures, and every answer equals the oracle."""
    import torch
    n, k = 300_000, 10
    f = synth.features(n)
    eng.load_features(f)
    before = eng.stat("graph_replays")
    for q in (5, 77_777, 299_999, 5):                      # host path: the engine's own staging buffers are the graph's
        assert_exact(eng.query_by_index([q], k), oracle.query_index(f, [q], k))
    assert eng.stat("graph_replays") >= before + 3
    dq = torch.zeros(8, dtype=torch.int32, device="cuda")
    oi = torch.empty((8, k), dtype=torch.int32, device="cuda")
    os_ = torch.empty((8, k), dtype=torch.float32, device="cuda")
    for rep in range(4):                                   # device path on torch's (legacy default) stream
        q = synth.query_indices(8, n) + rep * 1000
        dq.copy_(torch.from_numpy(q))
        eng.query_by_index_dev(dq, 8, k, oi, os_, stream=0)
        torch.cuda.synchronize()
        assert_exact((oi.cpu().numpy(), os_.cpu().numpy()), oracle.query_index(f, q, k))
    f2 = synth.uniform(200_000)
    eng.load_features(f2)                                  # new store: the cached graphs are dropped
    assert_exact(eng.query_by_index([5], k), oracle.query_index(f2, [5], k))
    eng.set_option("graphs", 0)
    r0 = eng.stat("graph_replays")
    assert_exact(eng.query_by_index([6], k), oracle.query_index(f2, [6], k))
    assert_exact(eng.query_by_index([7], k), oracle.query_index(f2, [7], k))
    assert eng.stat("graph_replays") == r0
