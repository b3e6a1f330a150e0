"""Drop-in `Generator` for the reference's model.py (model.py:15-106), running on libmocha_b200.

Same constructor (`Generator(config['model'])`), same `forward(src_X, cha_X, extract_feature=False)`,
same attribute surface the inference driver touches (`mot_embedding`, `pos_emb`, `encoder`,
`decoder`, `to_mot`, `eval()`; test_fullframework.py:49,190-192,301-302,455-456) and the same
state_dict keys/shapes, so released or random-init reference checkpoints load with strict=True.
All math runs in the hand-written CUDA kernels behind the C ABI; inputs must be CUDA tensors.
"""
from __future__ import annotations

import ctypes as C

import torch
from torch import nn

from . import _lib, packing, weights


class _Holder(nn.Module):
    """Parameter/buffer container reproducing a slice of the reference's module tree."""


def _attach(root: nn.Module, dotted: str, tensor: torch.Tensor, buffer: bool) -> None:
    parts = dotted.split(".")
    mod = root
    for p in parts[:-1]:
        if p not in mod._modules:
            mod.add_module(p, _Holder())
        mod = mod._modules[p]
    if buffer:
        mod.register_buffer(parts[-1], tensor.clone())
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor.clone(), requires_grad=False))


_BUFFER_KEYS = ("A_j", "A_b", "mot_embedding.3.weight", "to_mot.3.weight", ".pe")


def _is_buffer(key: str) -> bool:
    return key.endswith("A_j") or key.endswith("A_b") or key in ("mot_embedding.3.weight", "to_mot.3.weight") \
        or key.endswith(".pe")


class _Workspace:
    """Grow-only device scratch shared by the stage calls of one module."""

    def __init__(self):
        self.buf = None

    def get(self, nbytes: int, device) -> torch.Tensor:
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        return self.buf


class _Stage(nn.Module):
    """Callable sub-module (`model.mot_embedding(X)` etc.) that owns part of the parameter tree."""

    def __init__(self, owner, kind):
        super().__init__()
        object.__setattr__(self, "_owner_ref", owner)
        self._kind = kind

    def forward(self, *args):
        owner = object.__getattribute__(self, "_owner_ref")
        return getattr(owner, "_run_" + self._kind)(*args)


class Generator(nn.Module):
    def __init__(self, config, precision: str = "fp32"):
        super().__init__()
        self.config = dict(config)
        self.precision = precision
        sd = weights.generator_state_dict(seed=1777, cfg=self.config)
        self.pos_emb = nn.Parameter(sd["pos_emb"].clone(), requires_grad=False)
        self.mot_embedding = _Stage(self, "mot_embedding")
        self.encoder = _Stage(self, "encoder")
        self.decoder = _Stage(self, "decoder")
        self.to_mot = _Stage(self, "to_mot")
        for k, v in sd.items():
            if k == "pos_emb":
                continue
            _attach(self, k, v, _is_buffer(k))
        self._packed = None
        self._packed_key = None
        self._ws = _Workspace()

    # -- weight packing (lazy; invalidated when parameters are replaced or moved) -----------------
    def _pack(self):
        params = list(self.state_dict().values())
        key = tuple((p.data_ptr(), p._version) for p in params)
        dev = self.pos_emb.device
        if self._packed is None or key != self._packed_key:
            if dev.type != "cuda":
                raise _lib.MochaError("Generator must live on a CUDA device: call .to('cuda') (no CPU fallback)")
            self._packed = packing.PackedGenerator(self.state_dict(), self.config, dev)
            self._packed_key = key
        return self._packed

    def _prec(self) -> int:
        return _lib.precision_code(self.precision)

    def _check_in(self, x, shape_tail):
        _lib.require_cuda(x)
        if x.dtype != torch.float32:
            raise _lib.MochaError("expected float32 input")
        if tuple(x.shape[1:]) != tuple(shape_tail):
            raise _lib.MochaError(f"expected input [B,{','.join(map(str, shape_tail))}], got {tuple(x.shape)}")

    # -- stages ------------------------------------------------------------------------------------
    def _run_mot_embedding(self, X, add_pos_emb: bool = False):
        pk = self._pack()
        d = pk.dims
        X = X.contiguous()
        self._check_in(X, (d.T, d.V, d.Cin))
        B = X.shape[0]
        lib = _lib.load()
        out = torch.empty((B, pk.ntok, d.D), dtype=torch.float32, device=X.device)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_embed_workspace_bytes(C.byref(pk.struct.dims), B)
        ws = self._ws.get(nbytes, X.device)
        _lib.check(lib.mocha_embed_fwd(C.byref(pk.struct), _lib.ptr(X), B, _lib.ptr(out), int(add_pos_emb),
                                       self._prec(), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "mocha_embed_fwd")
        return out

    def _run_encoder(self, tokens, sty=None):
        pk = self._pack()
        d = pk.dims
        tokens = tokens.contiguous()
        self._check_in(tokens, (pk.ntok, d.D))
        B = tokens.shape[0]
        lib = _lib.load()
        out = torch.empty_like(tokens)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_encoder_workspace_bytes(C.byref(pk.struct.dims), B)
        ws = self._ws.get(nbytes, tokens.device)
        _lib.check(lib.mocha_encoder_fwd(C.byref(pk.struct), _lib.ptr(tokens), B, _lib.ptr(out), self._prec(),
                                         _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "mocha_encoder_fwd")
        return out

    def _run_decoder(self, src, sty):
        pk = self._pack()
        d = pk.dims
        src, sty = src.contiguous(), sty.contiguous()
        self._check_in(src, (pk.ntok, d.D))
        self._check_in(sty, (pk.ntok, d.D))
        if sty.shape[0] != src.shape[0]:
            raise _lib.MochaError("decoder: source and style batch sizes differ")
        B = src.shape[0]
        lib = _lib.load()
        out = torch.empty_like(src)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_decoder_workspace_bytes(C.byref(pk.struct.dims), B)
        ws = self._ws.get(nbytes, src.device)
        _lib.check(lib.mocha_decoder_fwd(C.byref(pk.struct), _lib.ptr(src), _lib.ptr(sty), B, _lib.ptr(out),
                                         self._prec(), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "mocha_decoder_fwd")
        return out

    def _run_to_mot(self, tokens):
        pk = self._pack()
        d = pk.dims
        tokens = tokens.contiguous()
        self._check_in(tokens, (pk.ntok, d.D))
        B = tokens.shape[0]
        lib = _lib.load()
        out = torch.empty((B, d.T, d.V, d.Cin), dtype=torch.float32, device=tokens.device)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_to_mot_workspace_bytes(C.byref(pk.struct.dims), B)
        ws = self._ws.get(nbytes, tokens.device)
        _lib.check(lib.mocha_to_mot_fwd(C.byref(pk.struct), _lib.ptr(tokens), B, _lib.ptr(out), None, None, None,
                                        self._prec(), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()),
                   "mocha_to_mot_fwd")
        return out

    # -- reference forward (model.py:82-106) ---------------------------------------------------------
    def forward(self, src_X, cha_X, extract_feature=False):
        from .transformer import mean_variance_norm
        src_tokens = self._run_mot_embedding(src_X, add_pos_emb=True)
        cha_tokens = self._run_mot_embedding(cha_X, add_pos_emb=True)
        src_encoded = self._run_encoder(src_tokens)
        cha_encoded = self._run_encoder(cha_tokens)
        if extract_feature:
            src_cnt = mean_variance_norm(src_encoded.permute(0, 2, 1))
            cha_cnt = mean_variance_norm(cha_encoded.permute(0, 2, 1))
            return src_encoded, cha_encoded, src_cnt.permute(0, 2, 1), cha_cnt.permute(0, 2, 1)
        trans_decoded = self._run_decoder(src_encoded, cha_encoded)
        return self._run_to_mot(trans_decoded)
