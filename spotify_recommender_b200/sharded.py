"""Row-sharded multi-GPU scoring (north_star (4)): one process per GPU, the
songs split into contiguous row shards, every rank answers the whole query batch over
its shard with GLOBAL song ids, ONE all-gather of packed 64-bit keys moves the K candidates
per query, and a merge kernel produces the final (score desc, id asc) lists -- bit-identical to a single
store holding all rows, because scores do not depend on the sharding and the order is total.

torch / torch.distributed are plumbing here (device buffers, the NCCL collectives over
NVLink); all scoring and merging is done by the engine behind the C ABI."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> tuple[int, int]:
    """Rows [lo, hi) owned by `rank`: ceil(n/world) rows each, the last shard shorter."""
    per = -(-n_total // world)
    lo = min(n_total, rank * per)
    return lo, min(n_total, lo + per)


class ShardedRecommender:
    """`engine` needs load_features / gather_rows_dev / query_by_vector_dev / bound_block_count / bound_blocks_dev /
    query_keys_by_vector_dev / merge_keys_dev (spotify_recommender_b200.engine.Engine).  `group` is a
    torch.distributed process group (NCCL on GPUs; the gloo tests drive the same
    host logic on CPU tensors with a checker-backed engine)."""

    def __init__(self, engine, n_total: int, group=None, device=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = int(n_total)
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._buf = {}

    # -- store ----------------------------------------------------------------
    def load_shard(self, rows) -> None:
        """rows: this rank's rows [lo, hi) (host array or device tensor)."""
        if rows.shape[0] != self.hi - self.lo:
            raise ValueError(f"rank {self.rank} owns rows [{self.lo}, {self.hi}), got {rows.shape[0]}")
        self.engine.load_features(rows, id_base=self.lo)

    def _scratch(self, name, shape, dtype):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._buf[name] = t
        return t

    def _stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream if self.device.type == "cuda" else 0

    # -- queries -----------------------------------------------------------------
    def query_by_index_dev(self, d_qidx: torch.Tensor, k: int):
        """d_qidx: int32 tensor of GLOBAL song ids, identical on every rank.  Returns
        (idx[nq,k] int32, score[nq,k] f32) device tensors, identical on every rank.
        Stream-ordered on the current torch stream; not synchronised.  1 <= k <= 1024."""
        nq = int(d_qidx.numel())
        st = self._stream()
        qrows = self._scratch("qrows", (nq, 12), torch.float32)
        # 1. the query vectors: each shard contributes the rows it owns, zeros elsewhere
        self.engine.gather_rows_dev(d_qidx, nq, qrows, st)
        if self.world > 1:
            dist.all_reduce(qrows, op=dist.ReduceOp.SUM, group=self.group)
        if self.world == 1:
            loc_i = self._scratch("loc_i", (nq, k), torch.int32)
            loc_s = self._scratch("loc_s", (nq, k), torch.float32)
            self.engine.query_by_vector_dev(qrows, d_qidx, nq, k, loc_i, loc_s, st)
            return loc_i, loc_s
        # 2. the threshold bound pass, shared: every shard samples 1 / world of the tiles a single store would, the
        #    block maxima are max-reduced (block b holds different songs on every shard), and every shard starts its
        #    scan from thresholds that bound the K-th best of the WHOLE store -- an eighth of the sampling cost per
        #    GPU at 8 GPUs, and no shard wastes time on candidates that cannot make the merged list
        nblk = self.engine.bound_block_count(k)
        blocks = None
        if nblk:
            blocks = self._scratch("blocks", (nq, nblk), torch.float32)
            self.engine.bound_blocks_dev(qrows, nq, k, self.world, blocks, st)
            dist.all_reduce(blocks, op=dist.ReduceOp.MAX, group=self.group)
        # 3. local exact top-K over this shard as packed 64-bit keys (orderable score << 32 | ~global id),
        #    self excluded by global id
        keys = self._scratch("keys", (nq, k), torch.int64)
        self.engine.query_keys_by_vector_dev(qrows, d_qidx, nq, k, keys, None, st, blocks)
        # 4. the ONE exchange of results: K keys per query from every shard
        all_keys = self._scratch("all_keys", (self.world, nq, k), torch.int64)
        dist.all_gather_into_tensor(all_keys.view(self.world * nq, k), keys, group=self.group)
        # 5. merge (every rank ends up with the final lists)
        out_i = self._scratch("out_i", (nq, k), torch.int32)
        out_s = self._scratch("out_s", (nq, k), torch.float32)
        self.engine.merge_keys_dev(all_keys, self.world, nq, k, out_i, out_s, 0, 0, None, st)
        return out_i, out_s

    def query_by_index(self, qidx, k: int):
        """Host in, host out (the end-to-end path): pinned staging, H2D of the ids,
        the device path above, D2H of the merged lists."""
        qidx = np.ascontiguousarray(qidx, np.int32).ravel()
        if qidx.size and (qidx.min() < 0 or qidx.max() >= self.n_total):
            raise ValueError(f"query ids must be in [0, {self.n_total})")
        nq = qidx.size
        pin = self._buf.get("pin_q")
        if pin is None or pin.numel() != nq:
            pin = torch.empty(nq, dtype=torch.int32, pin_memory=self.device.type == "cuda")
            self._buf["pin_q"] = pin
        pin.numpy()[:] = qidx
        d_q = self._scratch("d_q", (nq,), torch.int32)
        d_q.copy_(pin, non_blocking=True)
        out_i, out_s = self.query_by_index_dev(d_q, k)
        h_i = self._buf.get("pin_i")
        if h_i is None or tuple(h_i.shape) != (nq, k):
            h_i = torch.empty((nq, k), dtype=torch.int32, pin_memory=self.device.type == "cuda")
            h_s = torch.empty((nq, k), dtype=torch.float32, pin_memory=self.device.type == "cuda")
            self._buf["pin_i"], self._buf["pin_s"] = h_i, h_s
        h_s = self._buf["pin_s"]
        h_i.copy_(out_i, non_blocking=True)
        h_s.copy_(out_s, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        return h_i.numpy().copy(), h_s.numpy().copy()


class QueryShardedAllPairs:
    """BASELINE config 5 on several GPUs (SURVEY 8e "All-pairs"): the store is REPLICATED (48 B/song: 48 MB at
    1 M songs), rank r computes the neighbour lists of the query songs [r * ceil(N/G), ...) with the same
    kernels and no exchange at all; one all-gather at the very end assembles the N x K table on every rank."""

    def __init__(self, engine, group=None, device=None):
        self.engine = engine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.n = 0

    def load_replicated(self, rows) -> None:
        """rows: ALL rows, the same on every rank."""
        self.engine.load_features(rows, id_base=0)
        self.n = int(rows.shape[0])

    def local_range(self) -> tuple[int, int]:
        return shard_bounds(self.n, self.world, self.rank)

    def local_topk(self, k: int):
        """This rank's slice of the table (host arrays); empty slices give (0, k) arrays."""
        lo, hi = self.local_range()
        if hi <= lo:
            return np.empty((0, k), np.int32), np.empty((0, k), np.float32)
        return self.engine.all_pairs_topk(lo, hi, k)

    def gather_table(self, li: np.ndarray, ls: np.ndarray):
        """The one exchange of the path: every rank's slice -> the whole N x k table on every rank.
        (Indices and scores stay two contiguous arrays: packing them into one for a single all-gather was measured
        slower -- the strided host copies of an 80 MB table cost more than the second collective.)"""
        if self.world == 1:
            return li, ls
        k = li.shape[1]
        per = -(-self.n // self.world)
        pad_i = torch.full((per, k), -1, dtype=torch.int32)
        pad_s = torch.zeros((per, k), dtype=torch.float32)
        pad_i[:li.shape[0]] = torch.from_numpy(li)
        pad_s[:ls.shape[0]] = torch.from_numpy(ls)
        d_i, d_s = pad_i.to(self.device), pad_s.to(self.device)
        all_i = torch.empty((self.world * per, k), dtype=torch.int32, device=self.device)
        all_s = torch.empty((self.world * per, k), dtype=torch.float32, device=self.device)
        dist.all_gather_into_tensor(all_i, d_i, group=self.group)
        dist.all_gather_into_tensor(all_s, d_s, group=self.group)
        return all_i[:self.n].cpu().numpy(), all_s[:self.n].cpu().numpy()

    def all_pairs_topk(self, k: int):
        """(idx[N,k] int32, score[N,k] f32) host arrays, identical on every rank."""
        li, ls = self.local_topk(k)
        return self.gather_table(li, ls)
