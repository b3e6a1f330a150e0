"""Row-sharded context matching across GPUs (BASELINE config 5).

The character feature DB is split by rows over the ranks of a torch.distributed group (one process
per GPU, NCCL over NVLink); queries are replicated. Every rank finds its local exact top-k
(distance fp64 + GLOBAL row index int64), the [nq,k] candidate lists are all-gathered — a
latency-bound exchange of nq*k*16 bytes per rank — and merged by (distance, index) with a small
CUDA kernel (mocha_topk_merge). The exact fp64 re-rank happens on the shard that owns the row, so
only final distances travel."""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_rows: int, world: int, rank: int) -> "tuple[int, int]":
    """Contiguous, balanced row range [lo, hi) of `rank` (first n_rows % world ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_topk_cuda(all_dist: torch.Tensor, all_idx: torch.Tensor, k: int):
    """all_dist [S,nq,k] fp64, all_idx [S,nq,k] int64 (-1 = empty) on CUDA -> merged ([nq,k], [nq,k])."""
    S, nq, kk = all_dist.shape
    out_d = torch.empty((nq, k), dtype=torch.float64, device=all_dist.device)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=all_dist.device)
    if kk != k:
        raise _lib.MochaError("merge_topk_cuda: candidate lists must have length k")
    _lib.check(_lib.load().mocha_topk_merge(_lib.ptr(all_dist.contiguous()), _lib.ptr(all_idx.contiguous()), S, nq, k,
                                            _lib.ptr(out_d), _lib.ptr(out_i), _lib.stream_ptr()), "mocha_topk_merge")
    return out_d, out_i


class PeerExchange:
    """Peer-memory replacement of all-gather + merge (one fused kernel per rank, include/mocha_b200.h).

    Every rank allocates one exchange buffer with CUDA IPC (mocha_peer_alloc), the 64-byte handles are
    all-gathered through the process group (host side, once), and each rank maps its peers' buffers. A call then
    costs ONE kernel: P2P stores of the local lists into every peer's slot over NVLink, a system-scope
    arrival counter, and the merge. Needs one process per GPU on a single node with P2P access."""

    def __init__(self, nq: int, k: int, group=None):
        import ctypes as C
        self.lib = _lib.load()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.nq, self.k = nq, k
        self.epoch = 0
        nbytes = self.lib.mocha_topk_exchange_bytes(self.world, nq, k)
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        _lib.check(self.lib.mocha_peer_alloc(nbytes, C.byref(own), handle), "mocha_peer_alloc")
        self._own = own
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self._opened = []
        ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(own.value)
                continue
            p = C.c_void_p()
            _lib.check(self.lib.mocha_peer_open(C.create_string_buffer(h, 64), C.byref(p)), "mocha_peer_open")
            self._opened.append(p)
            ptrs.append(p.value)
        self._ptrs = (C.c_void_p * self.world)(*ptrs)
        dist.barrier(group=group)   # every buffer is mapped (and zeroed) before the first exchange

    def exchange_merge(self, d_loc: torch.Tensor, i_loc: torch.Tensor):
        nq, k = d_loc.shape
        if (nq, k) != (self.nq, self.k):
            raise _lib.MochaError("PeerExchange: list shape differs from the one the buffers were sized for")
        out_d = torch.empty_like(d_loc)
        out_i = torch.empty_like(i_loc)
        _lib.check(self.lib.mocha_topk_exchange_merge(_lib.ptr(d_loc.contiguous()), _lib.ptr(i_loc.contiguous()), nq, k,
                                                      self.rank, self.world, self._ptrs, self.epoch, _lib.ptr(out_d),
                                                      _lib.ptr(out_i), _lib.stream_ptr()), "mocha_topk_exchange_merge")
        self.epoch += 1
        return out_d, out_i

    def close(self):
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for p in self._opened:
            self.lib.mocha_peer_close(p)
        self._opened = []
        if self._own is not None:
            self.lib.mocha_peer_free(self._own)
            self._own = None


class ShardedMatcher:
    """k-NN over a DB whose rows live on different ranks.

    local_query(q, k) -> (dist [nq,k'] fp64, idx [nq,k'] int64 LOCAL row ids) with k' = min(k, rows);
    defaults to this rank's `BallTree.query_device`. merge(all_dist, all_idx, k) defaults to the CUDA
    merge kernel. Both are injectable so the host-side protocol is testable on CPU with gloo.
    """

    def __init__(self, n_rows_total: int, local_query, group=None, merge=merge_topk_cuda, exchange="nccl"):
        if exchange not in ("nccl", "peer"):
            raise ValueError("exchange must be 'nccl' (all-gather + merge kernel) or 'peer' (fused P2P kernel)")
        self.exchange = exchange
        self._peer = None
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.n_total = n_rows_total
        self.lo, self.hi = shard_bounds(n_rows_total, self.world, self.rank)
        self.local_query = local_query
        self.merge = merge

    def query(self, q: torch.Tensor, k: int = 1):
        n_local = self.hi - self.lo
        kk = min(k, n_local)
        dev = q.device
        d_loc = torch.full((q.shape[0], k), float("inf"), dtype=torch.float64, device=dev)
        i_loc = torch.full((q.shape[0], k), -1, dtype=torch.int64, device=dev)
        if kk > 0:
            d, i = self.local_query(q, kk)
            d_loc[:, :kk] = d
            i_loc[:, :kk] = i + self.lo          # global row index = local index + shard offset
        if self.world == 1:
            return d_loc, i_loc
        nq = d_loc.shape[0]
        if self.exchange == "peer":
            if self._peer is None or (self._peer.nq, self._peer.k) != (nq, k):
                if self._peer is not None:
                    self._peer.close()
                self._peer = PeerExchange(nq, k, group=self.group)
            return self._peer.exchange_merge(d_loc, i_loc)
        all_d = torch.empty((self.world * nq, k), dtype=torch.float64, device=dev)
        all_i = torch.empty((self.world * nq, k), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_d, d_loc.contiguous(), group=self.group)
        dist.all_gather_into_tensor(all_i, i_loc.contiguous(), group=self.group)
        return self.merge(all_d.view(self.world, nq, k), all_i.view(self.world, nq, k), k)


def sharded_balltree(db_rows_local: torch.Tensor, n_rows_total: int, group=None, **tree_kwargs) -> ShardedMatcher:
    """Build the per-rank BallTree over this rank's rows and wrap it in a ShardedMatcher."""
    from .balltree import BallTree
    tree = BallTree(db_rows_local, **tree_kwargs)
    m = ShardedMatcher(n_rows_total, lambda q, k: tree.query_device(q, k=k, return_distance=True), group=group)
    if m.hi - m.lo != tree.N:
        raise _lib.MochaError(f"rank {m.rank} holds {tree.N} rows but its shard is [{m.lo},{m.hi})")
    m.tree = tree
    return m
