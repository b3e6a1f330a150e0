"""Host-side protocol of the row-sharded matcher on CPU: world_size=2, gloo backend. The local
search and the merge are injected NumPy/torch stand-ins (the CUDA kernels are covered by -m gpu)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _merge_cpu(all_d, all_i, k):
    S, nq, kk = all_d.shape
    d = all_d.permute(1, 0, 2).reshape(nq, S * kk).numpy()
    i = all_i.permute(1, 0, 2).reshape(nq, S * kk).numpy()
    d = np.where(i < 0, np.inf, d)
    order = np.lexsort((i, d), axis=-1)[:, :k]
    return torch.from_numpy(np.take_along_axis(d, order, 1)), torch.from_numpy(np.take_along_axis(i, order, 1))


def _worker(rank, world, port, n_rows, k, q_out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mocha_oracle import matching
    from mocha_sigasia2023_b200.sharded import ShardedMatcher, shard_bounds
    rng = np.random.default_rng(0)
    db = rng.standard_normal((n_rows, 96)).astype(np.float32)
    q = rng.standard_normal((7, 96)).astype(np.float32)
    lo, hi = shard_bounds(n_rows, world, rank)

    def local_query(qt, kk):
        d, i = matching.knn(db[lo:hi], qt.numpy(), kk)
        return torch.from_numpy(d), torch.from_numpy(i)

    m = ShardedMatcher(n_rows, local_query, merge=_merge_cpu)
    d, i = m.query(torch.from_numpy(q), k=k)
    wd, wi = matching.knn(db, q, k)
    ok = bool((i.numpy() == wi).all() and np.allclose(d.numpy(), wd, rtol=1e-12))
    res = torch.tensor([int(ok)])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        q_out.put(int(res.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows,k", [(101, 3), (2, 1), (5, 4)])
def test_sharded_matcher_world2_gloo(n_rows, k):
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29500 + (os.getpid() + n_rows * 7 + k) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_rows, k, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q_out.get(timeout=5) == 1


def test_shard_bounds_cover_and_balance():
    from mocha_sigasia2023_b200.sharded import shard_bounds
    for n in (0, 1, 7, 16_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)
