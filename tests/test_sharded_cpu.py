"""world_size-2 gloo test of the row-sharded host logic (spotify_recommender_b200.sharded):
query-row exchange, per-shard top-K with global ids, all-gather, merge.  The GPU engine is
replaced by a checker-backed stand-in (the oracle) -- tests only; the product path always
uses spotify_recommender_b200.engine.Engine."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


class CheckerEngine:
    """CPU stand-in with the Engine methods ShardedRecommender calls."""

    def __init__(self):
        from oracle_lib import Oracle
        self.o = Oracle()

    def load_features(self, rows, id_base=0):
        self.rows = np.ascontiguousarray(rows, np.float32)
        self.base = int(id_base)

    def gather_rows_dev(self, d_ids, count, d_out, stream=0):
        ids = d_ids.numpy().astype(np.int64) - self.base
        own = (ids >= 0) & (ids < self.rows.shape[0])
        out = np.zeros((count, 12), np.float32)
        out[own] = self.rows[ids[own]]
        d_out.copy_(torch.from_numpy(out))

    def query_by_vector_dev(self, d_qrows, d_excl, nq, k, d_out_idx, d_out_score=None, stream=0):
        ex = d_excl.numpy().astype(np.int64) - self.base
        ex[(ex < 0) | (ex >= self.rows.shape[0])] = -1
        oi, os_ = self.o.query_rows(self.rows, d_qrows.numpy(), ex, k, id_base=self.base)
        d_out_idx.copy_(torch.from_numpy(oi))
        d_out_score.copy_(torch.from_numpy(os_))

    def merge_topk_dev(self, d_idx, d_score, parts, nq, k, d_out_idx, d_out_score=None, stream=0):
        mi, ms = self.o.merge_parts(d_idx.numpy(), d_score.numpy())
        d_out_idx.copy_(torch.from_numpy(mi))
        d_out_score.copy_(torch.from_numpy(ms))

    # the exchange format of the row-sharded path: orderable(score + 0.0) << 32 | (0xFFFFFFFF - id), 0 = none
    def bound_block_count(self, k):
        return 64

    def bound_blocks_dev(self, d_qrows, nq, k, shards, d_blocks, stream=0):
        # the checker has no bound pass: "no information" (-inf) is a valid contribution to the max-reduce
        d_blocks.fill_(float("-inf"))
        d_blocks[:, self.base % 64] = float(self.base)  # (a marker per shard, checked after the all-reduce)
        self.shards = shards

    def query_keys_by_vector_dev(self, d_qrows, d_excl, nq, k, d_out_keys, d_ceil=None, stream=0, d_blocks=None):
        if d_blocks is not None:  # the reduced array carries every shard's marker
            assert int(torch.isfinite(d_blocks[0]).sum()) == self.shards
        ex = d_excl.numpy().astype(np.int64) - self.base
        ex[(ex < 0) | (ex >= self.rows.shape[0])] = -1
        oi, os_ = self.o.query_rows(self.rows, d_qrows.numpy(), ex, k, id_base=self.base)
        d_out_keys.copy_(torch.from_numpy(pack_keys(oi, os_)))

    def merge_keys_dev(self, d_keys, parts, nq, k, d_out_idx, d_out_score=None, stride=0, col=0, d_ceil_out=None, stream=0):
        idx, score = unpack_keys(d_keys.numpy().reshape(parts, nq, k))
        mi, ms = self.o.merge_parts(idx, score)
        d_out_idx.copy_(torch.from_numpy(mi))
        d_out_score.copy_(torch.from_numpy(ms))

    def all_pairs_topk(self, lo, hi, k, scores=True):
        return self.o.query_index(self.rows, np.arange(lo, hi, dtype=np.int32), k)


def pack_keys(idx, score):
    u = (score + np.float32(0.0)).view(np.uint32).astype(np.uint64)
    o = np.where(u & 0x80000000, ~u & 0xFFFFFFFF, u | 0x80000000).astype(np.uint64)
    key = (o << np.uint64(32)) | (np.uint64(0xFFFFFFFF) - idx.astype(np.int64).astype(np.uint64) & np.uint64(0xFFFFFFFF))
    key[idx < 0] = 0
    return key.view(np.int64)


def unpack_keys(key):
    key = np.ascontiguousarray(key).view(np.uint64)
    o = (key >> np.uint64(32)).astype(np.uint32)
    u = np.where(o & 0x80000000, o ^ 0x80000000, ~o).astype(np.uint32)
    idx = (np.uint64(0xFFFFFFFF) - (key & np.uint64(0xFFFFFFFF))).astype(np.int64).astype(np.int32)
    score = u.view(np.float32).copy()
    idx[key == 0] = -1
    score[key == 0] = 0.0
    return idx, score


def _worker(rank, world, port, n, k, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.sharded import ShardedRecommender, shard_bounds
    full = synth.adversarial(n)
    sh = ShardedRecommender(CheckerEngine(), n, device=torch.device("cpu"))
    lo, hi = shard_bounds(n, world, rank)
    assert (sh.lo, sh.hi) == (lo, hi)
    sh.load_shard(full[lo:hi])
    q = np.array([3, 5, 17, 99, 205, n // 2, n - 1, lo, hi - 1], np.int32)
    gi, gs = sh.query_by_index(q, k)
    np.save(os.path.join(out_dir, f"idx{rank}.npy"), gi)
    np.save(os.path.join(out_dir, f"score{rank}.npy"), gs)
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_equals_single_store(tmp_path, oracle, world):
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.sharded import shard_bounds
    n, k = 9001, 37
    mp.spawn(_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    full = synth.adversarial(n)
    lo, hi = shard_bounds(n, world, world - 1)
    for rank in range(world):
        lo_r, hi_r = shard_bounds(n, world, rank)
        q = np.array([3, 5, 17, 99, 205, n // 2, n - 1, lo_r, hi_r - 1], np.int32)
        wi, ws = oracle.query_index(full, q, k)
        gi = np.load(tmp_path / f"idx{rank}.npy")
        gs = np.load(tmp_path / f"score{rank}.npy")
        # every rank holds the final lists of ITS query batch... all ranks used their own q (lo/hi differ)
        assert np.array_equal(gi[:7], wi[:7])
        assert np.array_equal(gs[:7].view(np.uint32), ws[:7].view(np.uint32))


def test_shard_bounds_cover_everything():
    from spotify_recommender_b200.sharded import shard_bounds
    for n in (1, 7, 8, 100, 10_000_001):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert all(hi - lo <= -(-n // w) for lo, hi in b)


def _all_pairs_worker(rank, world, port, n, k, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spotify_recommender_b200 import synth
    from spotify_recommender_b200.sharded import QueryShardedAllPairs
    ap = QueryShardedAllPairs(CheckerEngine(), device=torch.device("cpu"))
    ap.load_replicated(synth.adversarial(n))
    gi, gs = ap.all_pairs_topk(k)
    np.save(os.path.join(out_dir, f"ap_idx{rank}.npy"), gi)
    np.save(os.path.join(out_dir, f"ap_score{rank}.npy"), gs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_query_sharded_all_pairs_equals_single_store(tmp_path, oracle, world):
    """BASELINE config 5 host logic (SURVEY 8e "All-pairs"): store replicated, queries sharded, one final gather
    -- every rank ends up with the whole n x k table, ragged last slice included."""
    from spotify_recommender_b200 import synth
    n, k = 1003, 10
    mp.spawn(_all_pairs_worker, args=(world, _free_port(), n, k, str(tmp_path)), nprocs=world, join=True)
    wi, ws = oracle.query_index(synth.adversarial(n), np.arange(n, dtype=np.int32), k)
    for rank in range(world):
        assert np.array_equal(np.load(tmp_path / f"ap_idx{rank}.npy"), wi)
        assert np.array_equal(np.load(tmp_path / f"ap_score{rank}.npy").view(np.uint32), ws.view(np.uint32))
