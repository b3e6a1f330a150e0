"""Context matching oracle: exact Euclidean k-NN in float64 — the arithmetic of
sklearn.neighbors.BallTree (BallTree64, p=2) as called at test_fullframework.py:293-296,:440-443.
sklearn is a third-party dependency of the reference (environment.yml:15, unpinned; 1.9.0 in the
build container); its published algorithm is an exact search, so a float64 brute force returns the
same neighbours wherever distances are distinct. gen_golden.py pins this against BallTree itself."""
from __future__ import annotations

import numpy as np


def knn(db, q, k=1, chunk=2048):
    """db [N,D], q [nq,D] (any float dtype) -> (dist [nq,k] float64, idx [nq,k] int64),
    ascending by (distance, index). Difference form, float64 accumulation."""
    db = np.asarray(db)
    q = np.asarray(q)
    nq, N = q.shape[0], db.shape[0]
    d2 = np.empty((nq, N), dtype=np.float64)
    q64 = q.astype(np.float64)
    for s in range(0, N, chunk):
        blk = db[s:s + chunk].astype(np.float64)
        for i in range(nq):
            diff = blk - q64[i]
            d2[i, s:s + chunk] = np.einsum("nd,nd->n", diff, diff)
    order = np.lexsort((np.broadcast_to(np.arange(N), d2.shape), d2), axis=-1)[:, :k]
    return np.sqrt(np.take_along_axis(d2, order, axis=1)), order.astype(np.int64)


def knn_gemm(db, q, k=1):
    """Faster float64 variant (||x||^2 - 2 q.x + ||q||^2 via BLAS) for large CPU baselines; ranking
    is refined in difference form on the best 4k candidates."""
    db64, q64 = np.asarray(db, dtype=np.float64), np.asarray(q, dtype=np.float64)
    d2 = (db64 * db64).sum(1)[None, :] - 2.0 * (q64 @ db64.T) + (q64 * q64).sum(1)[:, None]
    kk = min(db64.shape[0], max(4 * k, 8))
    cand = np.argpartition(d2, kk - 1, axis=1)[:, :kk]
    out_d, out_i = np.empty((q64.shape[0], k)), np.empty((q64.shape[0], k), dtype=np.int64)
    for i in range(q64.shape[0]):
        diff = db64[cand[i]] - q64[i]
        e = np.einsum("nd,nd->n", diff, diff)
        o = np.lexsort((cand[i], e))[:k]
        out_d[i], out_i[i] = np.sqrt(e[o]), cand[i][o]
    return out_d, out_i
