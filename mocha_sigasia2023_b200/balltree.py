"""Drop-in for `sklearn.neighbors.BallTree` as used by the reference's context matching
(test_fullframework.py:293-296, :440-443; train_CVAE.py:207-211): exact Euclidean k-NN with float64
arithmetic on the float32 feature rows. No tree is built — the DB is streamed from HBM (fp64 brute
force, the default: exact like sklearn) or, opt-in for large DBs / query batches (use_tensor_cores=True /
"auto"), ranked on tcgen05 tensor cores by a bf16 / TF32 coarse score whose kc best rows are re-ranked
exactly in fp64 (approximate only in that a true neighbour must survive the coarse cut)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


class BallTree:
    # above this many (query x row) pairs the tensor-core path is used
    TC_THRESHOLD_PAIRS = 1 << 22

    def __init__(self, X, leaf_size=40, metric="minkowski", device=None, use_tensor_cores=None, kc=8,
                 tc_storage="bf16", **kwargs):
        if metric not in ("minkowski", "euclidean", "l2"):
            raise _lib.MochaError(f"BallTree metric {metric!r} unsupported (Euclidean only)")
        if kwargs.get("p", 2) != 2:
            raise _lib.MochaError("BallTree: only p=2 is supported")
        if isinstance(X, torch.Tensor):
            db = X
        else:
            db = torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float32)))
        if db.dim() != 2 or db.shape[0] == 0:
            raise ValueError("BallTree expects a non-empty 2-D array")
        if device is None:
            device = db.device if db.is_cuda else torch.device("cuda")
        self.data = db.to(device=device, dtype=torch.float32).contiguous()
        self.N, self.D = self.data.shape
        self.use_tensor_cores = use_tensor_cores
        self.kc = kc
        if tc_storage not in ("bf16", "fp32"):
            raise _lib.MochaError("tc_storage must be 'bf16' or 'fp32'")
        self.tc_storage = tc_storage   # operand format of the tensor-core coarse pass (fp32 -> TF32 MMA)
        self._norm32 = None
        self.center = None       # origin of the bf16 operands (set by _ensure_bf16 / from_feature_db)
        self._db16 = None
        self._norm = None
        self._ws = None

    @classmethod
    def from_feature_db(cls, fdb, **kwargs):
        """Tree over the local rows of a feature_db.FeatureDB without re-packing: the builder already wrote the
        fp32 rows, the bf16 rows and their norms in the matcher's layout."""
        if fdb.rows32 is None:
            raise _lib.MochaError("BallTree.from_feature_db needs the fp32 rows (keep_fp32=True)")
        t = cls(fdb.rows32, **kwargs)
        if fdb.rows16 is not None and fdb.norms is not None:
            t._db16, t._norm, t.center = fdb.rows16, fdb.norms, fdb.center
        return t

    def _scratch(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(max(nbytes, 1 << 16), dtype=torch.uint8, device=self.data.device)
        return self._ws

    def _ensure_bf16(self):
        """bf16 rows + norms for the tensor-core coarse pass, packed around the DB mean (`self.center`): feature
        rows share a large common component (||x|| ~ 370 vs ||x - x'|| ~ 2 on the synthetic character DB), and
        rounding x - mean instead of x keeps the bf16 error on the part that decides the ranking."""
        if self._db16 is None:
            lib = _lib.load()
            dev = self.data.device
            self.center = self.data.mean(dim=0, dtype=torch.float32).contiguous() if self.N > 1 else None
            self._db16 = torch.empty((self.N, self.D), dtype=torch.bfloat16, device=dev)
            self._norm = torch.empty((self.N,), dtype=torch.float32, device=dev)
            step = max(1, (256 << 20) // (4 * self.D))        # centred fp32 scratch of <= 256 MB at a time
            for s in range(0, self.N, step):
                rows = self.data[s:s + step]
                if self.center is not None:
                    rows = (rows - self.center).contiguous()
                _lib.check(lib.mocha_db_pack_bf16(_lib.ptr(rows), rows.shape[0], self.D, _lib.ptr(self._db16[s:s + step]),
                                                  _lib.ptr(self._norm[s:s + step]), _lib.stream_ptr()), "mocha_db_pack_bf16")

    def query_device(self, q: torch.Tensor, k: int = 1, return_distance: bool = True):
        """Same as query() but takes/returns CUDA tensors (no host round trip)."""
        lib = _lib.load()
        q = q.to(device=self.data.device, dtype=torch.float32).contiguous()
        if q.dim() != 2 or q.shape[1] != self.D:
            raise ValueError(f"query has dimension {tuple(q.shape)}, tree was built with D={self.D}")
        if k > self.N:
            raise ValueError("k must be less than or equal to the number of training points")
        nq = q.shape[0]
        idx = torch.empty((nq, k), dtype=torch.int64, device=q.device)
        dist = torch.empty((nq, k), dtype=torch.float64, device=q.device)
        use_tc = self.use_tensor_cores
        if use_tc is None:
            # The drop-in stays EXACT by default, like sklearn's BallTree: the tensor-core path keeps the kc best rows
            # of a bf16 / TF32 coarse score and re-ranks those exactly, which finds the true neighbours only when
            # they survive the coarse cut (measured: tests/test_gpu_match_large.py, bench.py match_sweep) - it is
            # opt-in (use_tensor_cores=True, or "auto" for the size-based switch of the throughput path).
            use_tc = False
        elif use_tc == "auto":
            use_tc = nq * self.N >= self.TC_THRESHOLD_PAIRS and self.D % 8 == 0 and self.D >= 64 and k <= self.kc
        if use_tc and self.tc_storage == "fp32":
            if self._norm32 is None:
                self._norm32 = torch.empty((self.N,), dtype=torch.float32, device=self.data.device)
                _lib.check(lib.mocha_db_norms_f32(_lib.ptr(self.data), self.N, self.D, _lib.ptr(self._norm32),
                                                  _lib.stream_ptr()), "mocha_db_norms_f32")
            ws = self._scratch(lib.mocha_match_tc_workspace_bytes(nq, self.N, self.D, self.kc))
            _lib.check(lib.mocha_match_tc(_lib.ptr(q), None, nq, None, _lib.ptr(self.data), _lib.ptr(self._norm32),
                                          self.N, self.D, k, self.kc, 0, _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(ws),
                                          ws.numel(), _lib.stream_ptr()), "mocha_match_tc(tf32)")
        elif use_tc:
            self._ensure_bf16()
            q16 = (q if self.center is None else q - self.center).to(torch.bfloat16)
            ws = self._scratch(lib.mocha_match_tc_workspace_bytes(nq, self.N, self.D, self.kc))
            _lib.check(lib.mocha_match_tc(_lib.ptr(q), _lib.ptr(q16), nq, _lib.ptr(self._db16), _lib.ptr(self.data),
                                          _lib.ptr(self._norm), self.N, self.D, k, self.kc, 0, _lib.ptr(idx),
                                          _lib.ptr(dist), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                       "mocha_match_tc")
        else:
            ws = self._scratch(lib.mocha_match_exact_workspace_bytes(nq, self.N, k))
            _lib.check(lib.mocha_match_exact(_lib.ptr(q), nq, _lib.ptr(self.data), self.N, self.D, k, 0,
                                             _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(ws), ws.numel(),
                                             _lib.stream_ptr()), "mocha_match_exact")
        return (dist, idx) if return_distance else idx

    def query(self, X, k=1, return_distance=True, **kwargs):
        if isinstance(X, torch.Tensor):
            q = X
        else:
            q = torch.from_numpy(np.ascontiguousarray(np.asarray(X, dtype=np.float32)))
        if q.dim() == 1:
            q = q[None]
        res = self.query_device(q, k=k, return_distance=return_distance)
        if return_distance:
            return res[0].cpu().numpy(), res[1].cpu().numpy()
        return res.cpu().numpy()
