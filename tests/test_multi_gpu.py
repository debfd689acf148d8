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
