"""Drop-in for the reference's net/transformer.py surface used by the inference driver:
`mean_variance_norm(input, eps=1e-5)` (net/transformer.py:13-20; call site test_fullframework.py:193)."""
from __future__ import annotations

import torch

from . import _lib


def mean_variance_norm(input, eps=1e-5):
    """(x - mean) / (std_unbiased + eps) over all trailing dims per (batch, channel).

    `input` is (B, C, *) as in the reference. The kernel wants token-major [B, n, C]; the driver
    always passes `encoded.permute(0, 2, 1)`, i.e. a transposed view of such a tensor, which is
    consumed in place without a copy.
    """
    if not isinstance(input, torch.Tensor) or not input.is_cuda:
        raise _lib.MochaError("mean_variance_norm needs a CUDA tensor (no CPU fallback)")
    if input.dtype != torch.float32:
        raise _lib.MochaError("mean_variance_norm expects float32")
    size = input.size()
    B, Cc = size[0], size[1]
    x = input.reshape(B, Cc, -1)
    n = x.shape[2]
    if n < 2:
        raise _lib.MochaError("mean_variance_norm needs at least 2 elements per channel")
    tok = x.permute(0, 2, 1)            # [B, n, C]
    if not tok.is_contiguous():
        tok = tok.contiguous()
    out = torch.empty_like(tok)
    lib = _lib.load()
    _lib.check(lib.mocha_cnt_features(_lib.ptr(tok), B, n, Cc, float(eps), _lib.ptr(out), None, None, None, None, None,
                                      _lib.stream_ptr()), "mocha_cnt_features")
    return out.permute(0, 2, 1).reshape(size)
