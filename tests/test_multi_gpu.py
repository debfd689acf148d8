"""Row-sharded path on real GPUs (NCCL): runs only where >= 2 GPUs are visible."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, n, nq, k, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.engine import Engine
    from spotify_recommender_b200.sharded import ShardedRecommender, shard_bounds
    eng = Engine(rank)
    sh = ShardedRecommender(eng, n, device=torch.device("cuda", rank))
    lo, hi = shard_bounds(n, world, rank)
    sh.load_shard(synth.features(n, lo, hi))
    q = synth.query_indices(nq, n)
    gi, gs = sh.query_by_index(q, k)
    np.save(os.path.join(out_dir, f"idx{rank}.npy"), gi)
    np.save(os.path.join(out_dir, f"score{rank}.npy"), gs)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def test_two_gpu_shards_equal_single_store(tmp_path, oracle):
    import torch
    import torch.multiprocessing as mp
    from spotify_recommender_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    n, nq, k = 400_003, 300, 20
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, n, nq, k, str(tmp_path)), nprocs=world, join=True)
    f = synth.features(n)
    q = synth.query_indices(nq, n)
    wi, ws = oracle.query_index(f, q, k, threads=8)
    for rank in range(world):
        gi = np.load(tmp_path / f"idx{rank}.npy")
        gs = np.load(tmp_path / f"score{rank}.npy")
        assert np.array_equal(gi, wi)
        assert np.array_equal(gs.view(np.uint32), ws.view(np.uint32))


def test_single_process_sharded_on_two_devices(oracle):
    """sr_sharded_* over two REAL devices: the query rows and the shards' key lists cross NVLink by peer access."""
    import torch
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.engine import ShardedEngine
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    devices = list(range(min(torch.cuda.device_count(), 4)))
    n = 400_003
    f = synth.features(n)
    q = np.concatenate([synth.query_indices(500, n), [0, n - 1, n // 2, n // 2 - 1]]).astype(np.int32)
    with ShardedEngine(devices) as se:
        se.load_features(f)
        for k in (10, 100, 1500):
            wi, ws = oracle.query_index(f, q, k, threads=8)
            gi, gs = se.query_by_index(q, k)
            assert np.array_equal(gi, wi) and np.array_equal(gs.view(np.uint32), ws.view(np.uint32))
    fa = synth.features(30_000)
    want = oracle.query_index(fa, np.arange(30_000, dtype=np.int32), 10, threads=8)
    with ShardedEngine(devices) as se:  # replicated store, queries sharded (BASELINE config 5)
        se.load_features(fa, replicate=True)
        gi, gs = se.all_pairs_topk(10)
        assert np.array_equal(gi, want[0]) and np.array_equal(gs.view(np.uint32), want[1].view(np.uint32))


def _all_pairs_worker(rank, world, port, n, k, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.engine import Engine
    from spotify_recommender_b200.sharded import QueryShardedAllPairs
    eng = Engine(rank)
    ap = QueryShardedAllPairs(eng, device=torch.device("cuda", rank))
    ap.load_replicated(synth.features(n))
    gi, gs = ap.all_pairs_topk(k)
    np.save(os.path.join(out_dir, f"ap_idx{rank}.npy"), gi)
    np.save(os.path.join(out_dir, f"ap_score{rank}.npy"), gs)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()


def test_query_sharded_all_pairs_over_nccl(tmp_path, oracle):
    """BASELINE config 5 on several GPUs: store replicated, queries sharded, one NCCL gather of the table."""
    import torch
    import torch.multiprocessing as mp
    from spotify_recommender_b200 import synth
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = min(torch.cuda.device_count(), 4)
    n, k = 50_003, 10
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_all_pairs_worker, args=(world, port, n, k, str(tmp_path)), nprocs=world, join=True)
    wi, ws = oracle.query_index(synth.features(n), np.arange(n, dtype=np.int32), k, threads=8)
    for rank in range(world):
        assert np.array_equal(np.load(tmp_path / f"ap_idx{rank}.npy"), wi)
        assert np.array_equal(np.load(tmp_path / f"ap_score{rank}.npy").view(np.uint32), ws.view(np.uint32))
