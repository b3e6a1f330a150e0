"""Drop-in `CVAE` for the reference's model_CVAE.py (model_CVAE.py:8-46) on libmocha_b200.

Constructor and method signatures are the reference's; state_dict keys/shapes match so
`network_cvae.load_state_dict(torch.load(...))` (test_fullframework.py:56-57) works unchanged.
`sample()` (PriorNet + Decoder) is the inference hot path and runs in CUDA kernels. The posterior
`Encoder` is training-only: its parameters are kept for checkpoint compatibility
and `forward`/`encode` raise.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib, packing, weights
from .model import _attach, _is_buffer, _Workspace


class CVAE(nn.Module):
    def __init__(self, output_seq, latent_dim=256, depth=2, nheads=4, feedforward_dim=512, dropout=0.1,
                 activation=F.relu, precision: str = "fp32"):
        super().__init__()
        if activation is not F.relu:
            raise _lib.MochaError("CVAE kernels implement activation=F.relu only (test_fullframework.py:55)")
        self.output_seq, self.latent_dim, self.depth = output_seq, latent_dim, depth
        self.nheads, self.feedforward_dim = nheads, feedforward_dim
        self.precision = precision
        sd = weights.cvae_state_dict(seed=1778, latent_dim=latent_dim, depth=depth, dff=feedforward_dim)
        for k, v in sd.items():
            _attach(self, k, v, _is_buffer(k))
        self._packed = None
        self._packed_key = None
        self._packed_post = None
        self._packed_post_key = None
        self._ws = _Workspace()
        # Source of the reparameterisation noise. Default mirrors the reference
        # (torch.randn_like on the tensor's device, model_CVAE.py:83); parity runs inject host draws.
        self.eps_fn = None

    def _pack(self):
        sd = self.state_dict()
        key = tuple((p.data_ptr(), p._version) for p in sd.values())
        dev = next(iter(sd.values())).device
        if self._packed is None or key != self._packed_key:
            if dev.type != "cuda":
                raise _lib.MochaError("CVAE must live on a CUDA device: call .to('cuda') (no CPU fallback)")
            self._packed = packing.PackedCVAE(sd, self.output_seq, self.latent_dim, self.depth, self.nheads,
                                              self.feedforward_dim, dev)
            self._packed_key = key
        return self._packed

    def _prec(self):
        return _lib.precision_code(self.precision)

    def _pack_posterior(self):
        """The posterior Encoder's weights in the token-network slots of a second ABI struct (training-side surface:
        built on first use of encode / forward)."""
        sd = self.state_dict()
        key = tuple((p.data_ptr(), p._version) for p in sd.values())
        dev = next(iter(sd.values())).device
        if self._packed_post is None or key != self._packed_post_key:
            if dev.type != "cuda":
                raise _lib.MochaError("CVAE must live on a CUDA device: call .to('cuda') (no CPU fallback)")
            self._packed_post = packing.PackedCVAE(sd, self.output_seq, self.latent_dim, self.depth, self.nheads,
                                                   self.feedforward_dim, dev, token_net="encoder")
            self._packed_post_key = key
        return self._packed_post

    def _run_tokens(self, pk, tokens):
        """mu, logvar of a token network (prior or posterior) over `tokens` [B, n, D]: mocha_cvae_sample without outputs."""
        tokens = tokens.contiguous()
        _lib.require_cuda(tokens)
        if tokens.dtype != torch.float32 or tokens.dim() != 3 or tokens.shape[2] != self.latent_dim:
            raise _lib.MochaError(f"tokens must be float32 [B, n, {self.latent_dim}]")
        B, n = tokens.shape[0], tokens.shape[1]
        lib = _lib.load()
        mu = torch.empty((B, self.latent_dim), dtype=torch.float32, device=tokens.device)
        logvar = torch.empty_like(mu)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_cvae_workspace_bytes(C.byref(pk.struct), B, n)
        ws = self._ws.get(nbytes, tokens.device)
        _lib.check(lib.mocha_cvae_sample(C.byref(pk.struct), _lib.ptr(tokens), B, n, None, None, _lib.ptr(mu),
                                         _lib.ptr(logvar), None, None, None, self._prec(), _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr()), "mocha_cvae_sample(tokens)")
        return mu, logvar

    def _run(self, c, eps):
        pk = self._pack()
        c = c.contiguous()
        _lib.require_cuda(c)
        if c.dtype != torch.float32 or c.dim() != 3 or c.shape[2] != self.latent_dim:
            raise _lib.MochaError(f"condition must be float32 [B, n, {self.latent_dim}]")
        B, ncond = c.shape[0], c.shape[1]
        lib = _lib.load()
        out = torch.empty((B, self.output_seq, self.latent_dim), dtype=torch.float32, device=c.device)
        mu = torch.empty((B, self.latent_dim), dtype=torch.float32, device=c.device)
        logvar = torch.empty_like(mu)
        with _lib.workspace_precision(self._prec()):
            nbytes = lib.mocha_cvae_workspace_bytes(C.byref(pk.struct), B, ncond)
        ws = self._ws.get(nbytes, c.device)
        _lib.check(lib.mocha_cvae_sample(C.byref(pk.struct), _lib.ptr(c), B, ncond, _lib.ptr(eps), _lib.ptr(out),
                                         _lib.ptr(mu), _lib.ptr(logvar), None, None, None, self._prec(),
                                         _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "mocha_cvae_sample")
        return out, mu, logvar

    def _draw_eps(self, c):
        shape = (c.shape[0], self.latent_dim)
        if self.eps_fn is not None:
            e = self.eps_fn(shape)
            return e.to(device=c.device, dtype=torch.float32).contiguous()
        return torch.randn(shape, dtype=torch.float32, device=c.device)

    def prior(self, c):
        _, mu, logvar = self._run(c, None)
        return mu, logvar

    def sample(self, c, deterministic=False):
        # the reference draws eps even when deterministic (model_CVAE.py:83); keep the RNG in step
        eps = self._draw_eps(c)
        out, _, _ = self._run(c, None if deterministic else eps)
        return out

    def encode(self, x, c):
        """(model_CVAE.py:33-35) posterior mu, logvar: the Encoder is the prior's network over [mu, logvar, c, x] (:116-126).
        Forward values only - no autograd (training stays out of scope)."""
        if x.dim() != 3 or c.dim() != 3 or x.shape[0] != c.shape[0]:
            raise _lib.MochaError("encode(x, c): expected [B, n_x, D] and [B, n_c, D]")
        return self._run_tokens(self._pack_posterior(), torch.cat((c, x), dim=1))

    def forward(self, x, c):
        """(model_CVAE.py:37-42) out, (mu_po, logvar_po), (mu_pr, logvar_pr) with z_po = mu_po + eps * exp(logvar_po / 2).
        The decoder kernel path takes its latent as `mu_pr + eps' * std_pr`, so z_po is handed over as the equivalent eps'.
        Like the reference, two noise tensors are drawn (posterior, then prior: :128, :83); only the first is used."""
        mu_po, logvar_po = self.encode(x, c)
        eps_po = self._draw_eps(c)
        self._draw_eps(c)                       # the prior's draw (PriorNet.forward), unused by the output
        z_po = mu_po + eps_po * torch.exp(0.5 * logvar_po)
        mu_pr, logvar_pr = self.prior(c)
        eps_equiv = ((z_po - mu_pr) * torch.exp(-0.5 * logvar_pr)).contiguous()
        out, _, _ = self._run(c, eps_equiv)
        return out, (mu_po, logvar_po), (mu_pr, logvar_pr)
