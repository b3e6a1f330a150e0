"""GPU parity of context matching vs sklearn BallTree goldens and the float64 oracle."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from mocha_oracle import matching
from mocha_sigasia2023_b200.balltree import BallTree

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(gi.MATCH_CASES))
def test_exact_matcher_vs_balltree(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "match.npz"))
    db, q = gi.match_inputs(name)
    k = gi.MATCH_CASES[name][3]
    tree = BallTree(db, use_tensor_cores=False)
    dist, idx = tree.query(q, k=k, return_distance=True)
    assert idx.dtype == np.int64 and idx.shape == (q.shape[0], k)
    np.testing.assert_array_equal(idx, g[name + "_idx"])          # bit-exact indices
    np.testing.assert_allclose(dist, g[name + "_dist"], rtol=1e-12, atol=1e-12)
    only = tree.query(q[:1], k=1, return_distance=False)
    assert only.shape == (1, 1) and only[0, 0] == g[name + "_idx"][0, 0]


def test_reference_call_pattern():
    # test_fullframework.py:294-296: tree.query(x.reshape(n,-1), k=1, return_distance=False)[:,0][0]
    db, q = gi.match_inputs("small")
    tree = BallTree(db.reshape(db.shape[0], -1))
    fi = tree.query(q[0:1].reshape(1, -1), k=1, return_distance=False)[:, 0][0]
    want = matching.knn(db, q[0:1], 1)[1][0, 0]
    assert int(fi) == int(want)


# the 23040-d cases with few rows take the split-K coarse pass (tc_match_coarse_splitk)
@pytest.mark.parametrize("shape", [(4096, 512, 300, 2), (1000, 23040, 130, 1), (777, 192, 5, 3), (385, 23040, 128, 2),
                                   (385, 23040, 1, 1)])
def test_tensor_core_matcher(shape):
    N, D, nq, k = shape
    rng = np.random.default_rng(N + D)
    db = rng.standard_normal((N, D)).astype(np.float32)
    pick = rng.integers(0, N, size=nq)
    q = (db[pick] + 0.05 * rng.standard_normal((nq, D))).astype(np.float32)
    tree = BallTree(db, use_tensor_cores=True, kc=8)
    dist, idx = tree.query(q, k=k, return_distance=True)
    wd, wi = matching.knn_gemm(db, q, k)
    np.testing.assert_array_equal(idx[:, 0], pick)                 # planted neighbour found
    margins_ok = np.ones(nq, dtype=bool)
    if k > 1:
        margins_ok = (wd[:, 1:] - wd[:, :-1]).min(axis=1) > 1e-5
    np.testing.assert_array_equal(idx[margins_ok], wi[margins_ok])
    np.testing.assert_allclose(dist[margins_ok], wd[margins_ok], rtol=1e-9)


def test_tensor_core_matcher_iid_recall():
    # distance-concentrated worst case: report agreement on the margin-filtered subset
    rng = np.random.default_rng(11)
    N, D, nq = 20000, 1024, 256
    db = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    tree = BallTree(db, use_tensor_cores=True, kc=16)
    idx = tree.query(q, k=1, return_distance=False)[:, 0]
    wd, wi = matching.knn_gemm(db, q, 2)
    ok = (wd[:, 1] - wd[:, 0]) > 1e-5
    agree = (idx[ok] == wi[ok, 0]).mean()
    assert agree >= 0.98, f"coarse top-16 recall too low: {agree}"


def test_query_validation():
    db, q = gi.match_inputs("ragged")
    tree = BallTree(db)
    with pytest.raises(ValueError):
        tree.query(q[:, :10], k=1)
    with pytest.raises(ValueError):
        tree.query(q, k=db.shape[0] + 1)


def test_topk_merge_logical_shards():
    """Row-sharded matching on ONE device: G logical shards, CUDA merge kernel (SURVEY §4 tier note)."""
    from mocha_sigasia2023_b200.sharded import merge_topk_cuda, shard_bounds
    rng = np.random.default_rng(5)
    N, D, nq, k, G = 1003, 160, 33, 3, 4
    db = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    qd = torch.from_numpy(q).cuda()
    ds, ids = [], []
    for r in range(G):
        lo, hi = shard_bounds(N, G, r)
        d, i = BallTree(db[lo:hi], use_tensor_cores=False).query_device(qd, k=k)
        ds.append(d); ids.append(i + lo)
    d, i = merge_topk_cuda(torch.stack(ds), torch.stack(ids), k)
    wd, wi = matching.knn(db, q, k)
    np.testing.assert_array_equal(i.cpu().numpy(), wi)
    np.testing.assert_allclose(d.cpu().numpy(), wd, rtol=1e-12)


@pytest.mark.parametrize("N,D,nq,k", [(385, 23040, 128, 1), (131, 203, 37, 3), (64, 64, 16, 2)])
def test_exact_matcher_batched_tiles(N, D, nq, k):
    """nq >= 16 takes the 8x8-tiled fp64 kernel; ragged N, nq and D exercise its clamped edges."""
    rng = np.random.default_rng(N * 3 + nq)
    db = rng.standard_normal((N, D)).astype(np.float32)
    q = rng.standard_normal((nq, D)).astype(np.float32)
    dist, idx = BallTree(db, use_tensor_cores=False).query(q, k=k, return_distance=True)
    wd, wi = matching.knn(db, q, k)
    np.testing.assert_array_equal(idx, wi)
    np.testing.assert_allclose(dist, wd, rtol=1e-12)


@pytest.mark.parametrize("shape", [(4096, 512, 300, 2), (1000, 23040, 130, 1)])
def test_tf32_fp32_storage_matcher(shape):
    """BASELINE config 3 'fp32' leg: fp32 rows fed to tcgen05 as TF32, exact fp64 re-rank on the fp32 rows."""
    N, D, nq, k = shape
    rng = np.random.default_rng(N + D + 1)
    db = rng.standard_normal((N, D)).astype(np.float32)
    pick = rng.integers(0, N, size=nq)
    q = (db[pick] + 0.05 * rng.standard_normal((nq, D))).astype(np.float32)
    tree = BallTree(db, use_tensor_cores=True, kc=8, tc_storage="fp32")
    dist, idx = tree.query(q, k=k, return_distance=True)
    wd, wi = matching.knn_gemm(db, q, k)
    np.testing.assert_array_equal(idx[:, 0], pick)
    np.testing.assert_array_equal(idx[:, 0], wi[:, 0])
    np.testing.assert_allclose(dist[:, 0], wd[:, 0], rtol=1e-9)


@pytest.mark.parametrize("nq,k", [(300, 3), (301, 3), (3, 1), (77, 1)])     # odd nq*k: slots stay 16 B aligned (round-1 advisor finding)
def test_peer_exchange_two_logical_ranks_on_one_gpu(nq, k):
    """mocha_topk_exchange_merge (fused P2P store / signal / merge kernel of the DB-sharded matcher): two
    logical ranks on ONE GPU, their kernels on two streams so that both are resident and signal each other
    through their exchange buffers, three consecutive epochs; result = merge of the two shards' lists."""
    import ctypes as C
    from mocha_sigasia2023_b200 import _lib
    lib = _lib.load()
    world = 2
    nbytes = lib.mocha_topk_exchange_bytes(world, nq, k)
    bufs, handles = [], []
    for r in range(world):
        p = C.c_void_p(); h = C.create_string_buffer(64)
        _lib.check(lib.mocha_peer_alloc(nbytes, C.byref(p), h), "mocha_peer_alloc")
        bufs.append(p); handles.append(h)
    ptrs = (C.c_void_p * world)(*[b.value for b in bufs])
    streams = [torch.cuda.Stream() for _ in range(world)]
    rng = np.random.default_rng(5)
    try:
        for epoch in range(3):
            d_np = np.sort(rng.random((world, nq, k)), axis=2)
            i_np = rng.integers(0, 10_000, size=(world, nq, k)).astype(np.int64)
            i_np[1, ::7, -1] = -1                      # a shard with fewer than k rows for some queries
            d_np[1, ::7, -1] = np.inf
            d = [torch.from_numpy(d_np[r].copy()).cuda() for r in range(world)]     # own allocations: 16 B aligned lists
            i = [torch.from_numpy(i_np[r].copy()).cuda() for r in range(world)]
            outs = []
            torch.cuda.synchronize()
            for r in range(world):
                od = torch.empty((nq, k), dtype=torch.float64, device="cuda")
                oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
                with torch.cuda.stream(streams[r]):
                    _lib.check(lib.mocha_topk_exchange_merge(_lib.ptr(d[r]), _lib.ptr(i[r]), nq, k, r, world, ptrs, epoch,
                                                             _lib.ptr(od), _lib.ptr(oi), C.c_void_p(streams[r].cuda_stream)),
                               "mocha_topk_exchange_merge")
                outs.append((od, oi))
            torch.cuda.synchronize()
            # oracle: merge by (distance, index)
            cd = d_np.transpose(1, 0, 2).reshape(nq, world * k)
            ci = i_np.transpose(1, 0, 2).reshape(nq, world * k)
            cd = np.where(ci >= 0, cd, np.inf)
            order = np.lexsort((ci, cd), axis=1)[:, :k]
            want_d = np.take_along_axis(cd, order, axis=1)
            want_i = np.take_along_axis(ci, order, axis=1)
            for od, oi in outs:
                np.testing.assert_array_equal(oi.cpu().numpy(), want_i)
                np.testing.assert_array_equal(od.cpu().numpy(), want_d)
    finally:
        for b in bufs:
            lib.mocha_peer_free(b)
