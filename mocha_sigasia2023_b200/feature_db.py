"""Feature-DB builder: batched encode of whole character databases into the matcher's
layout, one shard per rank.

What the reference does in collect_CVAE_feature_action.py:167-190 and compute_cnt_norm.py:157-179 -
`mot_embedding -> + pos_emb -> encoder -> mean_variance_norm` over the windows in batches of 32, results
concatenated on the host - and what test_fullframework.py:266-272,:293 does for the target clip, done on the
GPU with the same CUDA stages the per-frame path uses, written straight into

  * `encoded`  [n, 90, 256]  fp32   (what the CVAE path / NN seed reads, test_fullframework.py:298)
  * `cnt`      [n, 90, 256]  fp32   (optional; the raw context feature, for `cnt_norm.npz` statistics)
  * `rows32`   [n, 23040]    fp32   `(cnt - cnt_mean) / cnt_std` flattened: the exact matcher / re-rank operand
  * `rows16`   [n, 23040]    bf16 + `norms` [n] = ||row16||^2: the tensor-core matcher's operand (row-major,
                                    K contiguous - what the TMA descriptors of mocha_match_tc expect), packed
                                    around `center` = the mean row when the fp32 rows are kept (see
                                    include/mocha_b200.h, mocha_match_tc: any common origin ranks identically in
                                    exact arithmetic, the mean minimises the bf16 rounding error of the ranking)

Rows are split over ranks with `sharded.shard_bounds` (contiguous, balanced), so `build_feature_db(...,
shard=(rank, world))` on every rank yields exactly the DB-sharded matcher's layout (config 5)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, packing
from .sharded import shard_bounds


@dataclass
class FeatureDB:
    lo: int                      # global index of the first local row
    hi: int
    n_total: int
    encoded: torch.Tensor | None
    cnt: torch.Tensor | None
    rows32: torch.Tensor | None
    rows16: torch.Tensor | None
    norms: torch.Tensor | None
    center: torch.Tensor | None = None   # origin the bf16 rows are expressed around (None: the true origin)

    @property
    def n_local(self) -> int:
        return self.hi - self.lo

    def cnt_statistics(self):
        """Per-(token, channel) mean / std of the raw context feature over the local rows, the quantities
        compute_cnt_norm.py:175-179 saves as cnt_norm.npz (population std, like np.std)."""
        if self.cnt is None:
            raise _lib.MochaError("FeatureDB built without keep_cnt=True")
        return self.cnt.mean(dim=0), self.cnt.std(dim=0, unbiased=False)


class FeatureEncoder:
    """mot_embedding -> + pos_emb -> encoder -> context feature for batches of pose windows (the encode
    stages of CharacterizationSession without the per-frame state)."""

    def __init__(self, gen_sd, cfg, cnt_mean, cnt_std, batch: int, device="cuda", precision: str = "fp32"):
        self.lib = _lib.load()
        _lib.check(self.lib.mocha_check_device(), "mocha_check_device")
        self.dev = torch.device(device)
        self.B = batch
        self.prec = _lib.precision_code(precision)
        self.gen = packing.PackedGenerator(gen_sd, cfg, self.dev)
        d = self.gen.dims
        self.n, self.D = self.gen.ntok, d.D
        f32 = dict(dtype=torch.float32, device=self.dev)
        self.cnt_mean = torch.as_tensor(np.ascontiguousarray(cnt_mean), **f32).contiguous()
        self.cnt_std = torch.as_tensor(np.ascontiguousarray(cnt_std), **f32).contiguous()
        self.X = torch.zeros((batch, d.T, d.V, d.Cin), **f32)
        self.tokens = torch.empty((batch, self.n, self.D), **f32)
        dims = C.byref(self.gen.struct.dims)
        with _lib.workspace_precision(self.prec):
            nbytes = max(self.lib.mocha_embed_workspace_bytes(dims, batch), self.lib.mocha_encoder_workspace_bytes(dims, batch))
        self.ws = torch.empty(nbytes + 4096, dtype=torch.uint8, device=self.dev)

    def encode_into(self, X: torch.Tensor, encoded: torch.Tensor, cnt, rows32, rows16):
        """X [m <= batch, T, V, Cin] (CUDA fp32) -> slices of the caller's output tensors (first dim m)."""
        m = X.shape[0]
        if m > self.B:
            raise _lib.MochaError("batch larger than the encoder was built for")
        self.X[:m].copy_(X)
        if m < self.B:
            self.X[m:].zero_()
        lib, g, s = self.lib, C.byref(self.gen.struct), _lib.stream_ptr()
        wp, wn = _lib.ptr(self.ws), self.ws.numel()
        full = m == self.B
        enc = encoded if full else torch.empty((self.B, self.n, self.D), dtype=torch.float32, device=self.dev)
        _lib.check(lib.mocha_embed_fwd(g, _lib.ptr(self.X), self.B, _lib.ptr(self.tokens), 1, self.prec, wp, wn, s), "embed")
        _lib.check(lib.mocha_encoder_fwd(g, _lib.ptr(self.tokens), self.B, _lib.ptr(enc), self.prec, wp, wn, s), "encoder")
        # a ragged last batch goes through scratch outputs of full batch size
        def out(t, shape, dtype):
            if t is None:
                return None, None
            if full:
                return t, t
            return torch.empty((self.B,) + shape, dtype=dtype, device=self.dev), t
        c_buf, c_dst = out(cnt, (self.n, self.D), torch.float32)
        r32_buf, r32_dst = out(rows32, (self.n * self.D,), torch.float32)
        r16_buf, r16_dst = out(rows16, (self.n * self.D,), torch.bfloat16)
        _lib.check(lib.mocha_cnt_features(_lib.ptr(enc), self.B, self.n, self.D, 1e-5,
                                          None if c_buf is None else _lib.ptr(c_buf), _lib.ptr(self.cnt_mean),
                                          _lib.ptr(self.cnt_std), None if r32_buf is None else _lib.ptr(r32_buf),
                                          None if r16_buf is None else _lib.ptr(r16_buf), None, s), "cnt_features")
        if not full:
            encoded.copy_(enc[:m])
            for buf, dst in ((c_buf, c_dst), (r32_buf, r32_dst), (r16_buf, r16_dst)):
                if dst is not None:
                    dst.copy_(buf[:m])


def build_feature_db(gen_sd, cfg, cnt_mean, cnt_std, windows, batch: int = 256, precision: str = "fp32",
                     shard=(0, 1), device="cuda", keep_encoded: bool = True, keep_cnt: bool = False,
                     keep_fp32: bool = True, keep_bf16: bool = True) -> FeatureDB:
    """windows: [N, T, V, Cin] normalised pose windows (NumPy or tensor, host or device) of the WHOLE database;
    this rank encodes rows shard_bounds(N, world, rank). cnt_mean / cnt_std: the [90,256] tables the driver
    applies (:293; std already divided by std_weight, :89)."""
    rank, world = shard
    N = int(windows.shape[0])
    lo, hi = shard_bounds(N, world, rank)
    n = hi - lo
    dev = torch.device(device)
    encd = FeatureEncoder(gen_sd, cfg, cnt_mean, cnt_std, min(batch, max(n, 1)), dev, precision)
    nt, D = encd.n, encd.D
    f32 = dict(dtype=torch.float32, device=dev)
    encoded = torch.empty((n, nt, D), **f32)          # always produced (scratch when not kept)
    cnt = torch.empty((n, nt, D), **f32) if keep_cnt else None
    rows32 = torch.empty((n, nt * D), **f32) if keep_fp32 else None
    rows16 = torch.empty((n, nt * D), dtype=torch.bfloat16, device=dev) if keep_bf16 else None
    center_rows = keep_fp32 and keep_bf16 and n > 1
    for s in range(0, n, encd.B):
        m = min(encd.B, n - s)
        X = windows[lo + s:lo + s + m]
        X = torch.as_tensor(np.ascontiguousarray(X) if isinstance(X, np.ndarray) else X, dtype=torch.float32).to(dev)
        encd.encode_into(X, encoded[s:s + m], None if cnt is None else cnt[s:s + m],
                         None if rows32 is None else rows32[s:s + m],
                         None if (rows16 is None or center_rows) else rows16[s:s + m])
    norms = center = None
    if rows16 is not None and n > 0:
        norms = torch.empty((n,), **f32)
        if center_rows:
            center = rows32.mean(dim=0, dtype=torch.float32).contiguous()
        for s in range(0, n, 8192):
            if center_rows:
                rows16[s:s + 8192] = (rows32[s:s + 8192] - center).to(torch.bfloat16)
            # ||row16||^2 of the ROUNDED rows, accumulated in fp32 like mocha_db_pack_bf16 does
            r = rows16[s:s + 8192].float()
            norms[s:s + 8192] = (r * r).sum(dim=1)
    return FeatureDB(lo, hi, N, encoded if keep_encoded else None, cnt, rows32, rows16, norms, center)
