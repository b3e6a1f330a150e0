"""Fused tcgen05 kernels of the transformer layers vs a plain PyTorch fp32 reference of the same op (operands rounded to
bf16 exactly where the kernel rounds them: the GEMM inputs, the bf16 copy of Y and the hidden activations)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from mocha_sigasia2023_b200 import _lib

pytestmark = pytest.mark.gpu


def _tail_ref(A0, W0, b0, R0, ln1, Hd, act, W1, b1, W2, b2, ln2, eps):
    y = A0.float() @ W0.float().T
    if b0 is not None:
        y = y + b0
    if R0 is not None:
        y = y + R0
    if ln1 is not None:
        y = F.layer_norm(y, (256,), ln1[0], ln1[1], eps)
    if Hd == 0:
        return y
    x = y.to(torch.bfloat16).float()
    h = x @ W1.float().T + b1
    h = {0: lambda t: t, 1: torch.relu, 2: lambda t: F.gelu(t), 3: lambda t: F.leaky_relu(t, 0.2)}[act](h)
    z = y + h.to(torch.bfloat16).float() @ W2.float().T + b2
    if ln2 is not None:
        z = F.layer_norm(z, (256,), ln2[0], ln2[1], eps)
    return z


def _run_tail(M, K0, Hd, act, ln1, ln2, res=True, bias=True, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    rn = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g, device="cuda")).contiguous()
    A0 = rn(M, K0).to(torch.bfloat16)
    W0 = rn(256, K0, sc=K0 ** -0.5).to(torch.bfloat16)
    b0 = rn(256, sc=0.1) if bias else None
    R0 = rn(M, 256) if res else None
    l1 = (1.0 + 0.1 * rn(256), 0.1 * rn(256)) if ln1 else None
    l2 = (1.0 + 0.1 * rn(256), 0.1 * rn(256)) if ln2 else None
    W1 = rn(max(Hd, 1), 256, sc=1 / 16).to(torch.bfloat16)
    b1 = rn(max(Hd, 1), sc=0.1)
    W2 = rn(256, max(Hd, 1), sc=max(Hd, 1) ** -0.5).to(torch.bfloat16)
    b2 = rn(256, sc=0.1)
    eps = 1e-5
    O32 = torch.full((M, 256), float("nan"), device="cuda")
    O16 = torch.full((M, 256), float("nan"), device="cuda", dtype=torch.bfloat16)
    P = lambda t: None if t is None else _lib.ptr(t)
    lib = _lib.load()
    _lib.check(lib.mocha_block_tail(P(A0), K0, K0, P(W0), P(b0), P(R0), P(l1[0]) if l1 else None, P(l1[1]) if l1 else None,
                                    Hd, act, P(W1), P(b1), P(W2), P(b2), P(l2[0]) if l2 else None, P(l2[1]) if l2 else None,
                                    eps, P(O32), P(O16), M, _lib.stream_ptr()), "mocha_block_tail")
    torch.cuda.synchronize()
    want = _tail_ref(A0, W0, b0, R0, l1, Hd, act, W1, b1, W2, b2, l2, eps)
    scale = want.abs().max().item()
    err32 = (O32 - want).abs().max().item() / scale
    err16 = (O16.float() - want).abs().max().item() / scale
    return err32, err16


@pytest.mark.parametrize("M,K0,Hd,act,ln1,ln2", [
    (11520, 512, 512, 2, False, False),      # Generator encoder layer: out-proj + GELU FFN, 128 clips x 90 tokens
    (11520, 1024, 512, 2, False, False),     # Generator decoder layer (4 heads x 256)
    (23296, 256, 512, 1, True, True),        # CVAE prior layer: 128 x 182 tokens, post-LN, ReLU
    (300, 256, 512, 1, True, True),          # ragged last tile
    (256, 256, 0, 0, True, False),           # out-proj + LayerNorm only (CVAE decoder self-attention block)
    (90, 512, 512, 2, False, False),         # batch-1 streaming: a single partial tile
    (1000, 256, 512, 1, False, True),
])
def test_block_tail_vs_torch(M, K0, Hd, act, ln1, ln2):
    err32, err16 = _run_tail(M, K0, Hd, act, ln1, ln2)
    # fp32 output: accumulation order + the GELU polynomial (3e-7) only; bf16 output: + one rounding (2^-9 relative)
    assert err32 < 2e-3, f"fp32 output error {err32:.3e} of range"
    assert err16 < 8e-3, f"bf16 output error {err16:.3e} of range"


def test_block_tail_no_bias_no_residual():
    err32, err16 = _run_tail(640, 512, 512, 1, False, False, res=False, bias=False, seed=3)
    assert err32 < 2e-3 and err16 < 8e-3


# ---------------------------------------------------------------------------------------------------
# fused attention core
# ---------------------------------------------------------------------------------------------------
def _attn_ref(q, k, v, B, H, nq, nkv, dh):
    qf = q.float().view(B, nq, H, dh).permute(0, 2, 1, 3)
    kf = k.float().view(B, nkv, H, dh).permute(0, 2, 1, 3)
    vf = v.float().view(B, nkv, H, dh).permute(0, 2, 1, 3)
    s = (qf @ kf.transpose(-1, -2)) / dh ** 0.5
    p = torch.exp(s - s.amax(-1, keepdim=True)).to(torch.bfloat16).float()     # the kernel rounds P to bf16 ...
    o = (p @ vf) / p.sum(-1, keepdim=True)                                     # ... and normalises by the rounded sum
    return o.permute(0, 2, 1, 3).reshape(B, nq, H * dh)


@pytest.mark.parametrize("B,H,nq,nkv,dh", [
    (5, 4, 90, 90, 128),       # Generator encoder self-attention
    (3, 4, 90, 90, 256),       # Generator decoder cross-attention
    (3, 4, 182, 182, 64),      # CVAE prior: two query tiles per (clip, head), keys padded to 192
    (4, 4, 90, 181, 64),       # CVAE decoder cross-attention
    (6, 4, 2, 182, 64),        # CVAE prior last layer: only the mu / logvar rows query
    (128, 4, 90, 90, 128),     # the benchmarked batch: 512 units over 148 persistent CTAs
    (2, 2, 33, 100, 64),
])
def test_attention_core_vs_torch(B, H, nq, nkv, dh):
    g = torch.Generator(device="cuda").manual_seed(B * 7 + nq)
    inner = H * dh
    # q / k / v live side by side in one projection buffer, like the QKV GEMM leaves them
    qkv_q = torch.randn((B * nq, inner), generator=g, device="cuda").to(torch.bfloat16)
    kv = torch.randn((B * nkv, 2 * inner), generator=g, device="cuda").to(torch.bfloat16)
    k, v = kv[:, :inner], kv[:, inner:]
    out = torch.full((B, nq, inner), float("nan"), device="cuda", dtype=torch.bfloat16)
    lib = _lib.load()
    _lib.check(lib.mocha_attention_core(_lib.ptr(qkv_q), inner, C.c_void_p(k.data_ptr()), 2 * inner, C.c_void_p(v.data_ptr()),
                                        2 * inner, B, H, nq, nkv, dh, _lib.ptr(out), inner, _lib.stream_ptr()), "attention_core")
    torch.cuda.synchronize()
    want = _attn_ref(qkv_q, k.contiguous(), v.contiguous(), B, H, nq, nkv, dh)
    err = (out.float() - want).abs().max().item() / want.abs().max().item()
    assert err < 1e-2, f"attention core error {err:.3e} of range"
