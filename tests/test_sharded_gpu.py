"""The row-sharded multi-GPU path exercised on ONE GPU: G engines on device 0, each owning a
contiguous row shard with global ids (id_base), driven exactly as the multi-process host drives them --
gather_rows (summed as the all-reduce would), per-shard top-K in the packed-key exchange format, the
concatenation an all-gather delivers, merge -- and compared with the single-store oracle.  The reference has
no multi-GPU path (Recommender.cu:124 hard-wires device 0), so the oracle over the whole store is the pin.
Also: the single-process C ABI (sr_sharded_*, several shards on one GPU), the replicated / query-sharded
all-pairs form (BASELINE config 5), and the C++ Recommender class on a sharded store."""
import os

import numpy as np
import pytest

from spotify_recommender_b200 import synth

pytestmark = pytest.mark.gpu


def assert_exact(got, want):
    gi, gs = got
    wi, ws = want
    assert np.array_equal(gi, wi), f"index lists differ at {np.argwhere(gi != wi)[:5]}"
    assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32)), "scores are not bit-identical to the oracle"


def _bounds(n, G):
    from spotify_recommender_b200.sharded import shard_bounds
    return [shard_bounds(n, G, r) for r in range(G)]


CASES = [("adversarial", 30_011), ("features", 400_003)]


@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("data,n", CASES)
def test_row_shards_on_one_gpu_equal_single_store(oracle, G, data, n):
    """gather_rows_kernel + scan (keys out) + merge_parts_kernel, as the NCCL host strings them together."""
    import torch
    from spotify_recommender_b200.engine import Engine
    f = synth.adversarial(n) if data == "adversarial" else synth.features(n)
    bounds = _bounds(n, G)
    engines = []
    for lo, hi in bounds:
        e = Engine(0)
        e.load_features(f[lo:hi], id_base=lo)
        engines.append(e)
    # ties across shard borders: the queries sit either side of every border; plus duplicates / zero rows
    q = [3, 5, 17, 21, 99, 100, 163, 200, 231, n // 2, n - 1]
    for lo, hi in bounds:
        q += [lo, hi - 1]
    q = np.array(sorted(set(q)), np.int32)
    nq = q.size
    dq = torch.from_numpy(q).cuda()
    shard_rows = max(hi - lo for lo, hi in bounds)
    for k in (1, 10, 100, min(1024, shard_rows + 50)):  # the last: k larger than a shard (its list is -1 padded)
        # 1. query rows: every shard contributes what it owns, the sum is the all-reduce
        parts = torch.zeros((G, nq, 12), dtype=torch.float32, device="cuda")
        for g, e in enumerate(engines):
            e.gather_rows_dev(dq, nq, parts[g], stream=0)
        qrows = parts.sum(0)
        torch.cuda.synchronize()
        assert np.array_equal(qrows.cpu().numpy().view(np.uint32), f[q].view(np.uint32))
        # 2. local top-K as packed keys; 3. the all-gather = concatenation in rank order
        all_keys = torch.zeros((G, nq, k), dtype=torch.int64, device="cuda")
        for g, e in enumerate(engines):
            e.query_keys_by_vector_dev(qrows, dq, nq, k, all_keys[g], None, stream=0)
        local_keys = all_keys.clone()
        # ... and with the bound pass shared between the shards (block maxima max-reduced as the all-reduce would):
        # the shards' lists may be shorter, the merged rows are the same
        nblk = engines[0].bound_block_count(k)
        if nblk:
            parts_b = torch.empty((G, nq, nblk), dtype=torch.float32, device="cuda")
            for g, e in enumerate(engines):
                e.bound_blocks_dev(qrows, nq, k, G, parts_b[g], stream=0)
            blocks = parts_b.max(0).values.contiguous()
            for g, e in enumerate(engines):
                e.query_keys_by_vector_dev(qrows, dq, nq, k, all_keys[g], None, stream=0, d_blocks=blocks)
            torch.cuda.synchronize()
            assert int((all_keys != 0).sum()) <= int((local_keys != 0).sum())
        # 4. merge
        oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        os_ = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        want = oracle.query_index(f, q, k, threads=8)
        for keys in (all_keys, local_keys):
            engines[0].merge_keys_dev(keys, G, nq, k, oi, os_, stream=0)
            torch.cuda.synchronize()
            assert_exact((oi.cpu().numpy(), os_.cpu().numpy()), want)
        # the (idx, score) form of the merge gives the same rows
        loc_i = torch.empty((G, nq, k), dtype=torch.int32, device="cuda")
        loc_s = torch.empty((G, nq, k), dtype=torch.float32, device="cuda")
        for g, e in enumerate(engines):
            e.query_by_vector_dev(qrows, dq, nq, k, loc_i[g], loc_s[g], stream=0)
        engines[0].merge_topk_dev(loc_i, loc_s, G, nq, k, oi, os_, stream=0)
        torch.cuda.synchronize()
        assert_exact((oi.cpu().numpy(), os_.cpu().numpy()), want)
    for e in engines:
        e.close()


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
def test_single_process_sharded_abi(oracle, shards):
    """sr_sharded_*: peer-access query gather, per-shard keys, fused gather + merge kernel, host rows out."""
    from spotify_recommender_b200.engine import EngineError, ShardedEngine
    n = 250_007
    f = synth.features(n)
    with ShardedEngine([0] * shards) as se:
        assert se.shard_count == shards
        se.load_features(f)
        assert se.song_count == n
        q = np.concatenate([synth.query_indices(300, n), [0, n - 1, n // (shards + 1), n // (shards + 1) - 1]]).astype(np.int32)
        for k in (10, 100):
            assert_exact(se.query_by_index(q, k), oracle.query_index(f, q, k, threads=8))
        # a big ragged batch (several internal batches of 8192)
        qb = synth.query_indices(9000, n)
        assert_exact(se.query_by_index(qb, 5), oracle.query_index(f, qb, 5, threads=8))
        with pytest.raises(EngineError):
            se.query_by_index([n], 5)
    # ties across shard borders, k beyond 1024 (ceilings travel between the shards)
    fa = synth.adversarial(9001)
    with ShardedEngine([0] * shards) as se:
        se.load_features(fa)
        qa = np.array([3, 5, 17, 99, 150, 205, 4500, 9000], np.int32)
        for k in (37, 1500, 9000, 9500):
            assert_exact(se.query_by_index(qa, k), oracle.query_index(fa, qa, k, threads=8))


@pytest.mark.parametrize("shards", [1, 3])
def test_query_sharded_all_pairs(oracle, shards):
    """BASELINE config 5 at test size: store replicated, the QUERIES split over the shards, no exchange."""
    from spotify_recommender_b200.engine import ShardedEngine
    n, k = 20_011, 10
    f = synth.features(n)
    want = oracle.query_index(f, np.arange(n, dtype=np.int32), k, threads=8)
    with ShardedEngine([0] * shards) as se:
        se.load_features(f, replicate=True)
        assert_exact(se.all_pairs_topk(k), want)
        q = synth.query_indices(500, n)
        assert_exact(se.query_by_index(q, k), (want[0][q], want[1][q]))
    with ShardedEngine([0] * shards) as se:  # the row-sharded store serves the same table (more exchange)
        se.load_features(f)
        assert_exact(se.all_pairs_topk(k), want)


def test_all_pairs_multi_process_host_single_rank(oracle):
    """QueryShardedAllPairs with world size 1 on the GPU (the world_size-2 form runs under gloo on the CPU)."""
    import torch
    from spotify_recommender_b200.engine import Engine
    from spotify_recommender_b200.sharded import QueryShardedAllPairs
    n, k = 12_345, 10
    f = synth.features(n)
    with Engine(0) as e:
        ap = QueryShardedAllPairs(e, device=torch.device("cuda", 0))
        ap.load_replicated(f)
        assert_exact(ap.all_pairs_topk(k), oracle.query_index(f, np.arange(n, dtype=np.int32), k, threads=8))


def test_recommender_class_on_a_sharded_store(oracle, monkeypatch):
    """The C++ drop-in class (include/sr_recommender.hpp) returns identical lists on 1 and several shards."""
    from recommender_lib import HostRecommender
    n, k = 60_000, 25
    f = synth.features(n)
    q = [0, 777, 30_000, 59_999]
    want = oracle.query_index(f, np.array(q, np.int32), k, threads=8)[0]
    for devices in ("0", "0,0,0"):
        monkeypatch.setenv("SR_DEVICES", devices)
        r = HostRecommender(f)
        for j, qi in enumerate(q):
            assert np.array_equal(r.by_index(qi, k), want[j])
        big = r.by_index(5, 3000)  # topN beyond one pass
        assert np.array_equal(big, oracle.query_index(f, np.array([5], np.int32), 3000, threads=8)[0][0])
        r.close()
