"""tcgen05 GEMM (TMA + TMEM) vs torch on the same bf16-rounded operands."""
import ctypes as C

import numpy as np
import pytest
import torch

from mocha_sigasia2023_b200 import _lib

pytestmark = pytest.mark.gpu


def run_linear(A, W, bias, res, act, precision):
    lib = _lib.load()
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    nbytes = lib.mocha_linear_workspace_bytes(M, N, K, precision)
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda")
    _lib.check(lib.mocha_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), M, N, K, act,
                                precision, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "mocha_linear")
    torch.cuda.synchronize()
    return out


def torch_ref(A, W, bias, res, act):
    y = A.double() @ W.double().T
    if bias is not None:
        y = y + bias.double()
    if act == 1:
        y = torch.relu(y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    elif act == 3:
        y = torch.nn.functional.leaky_relu(y, 0.2)
    if res is not None:
        y = y + res.double()
    return y


SHAPES = [(90, 256, 256), (90, 1536, 256), (128, 128, 64), (257, 512, 1280), (11520, 256, 512), (1, 512, 256),
          (180, 768, 256), (1440, 64, 320), (300, 40, 128)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 2])
def test_fp32_linear(M, N, K, act):
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N)
    A = torch.randn((M, K), generator=g, device="cuda")
    W = torch.randn((N, K), generator=g, device="cuda") / K ** 0.5
    b = torch.randn((N,), generator=g, device="cuda")
    r = torch.randn((M, N), generator=g, device="cuda")
    out = run_linear(A, W, b, r, act, _lib.MOCHA_FP32)
    want = torch_ref(A, W, b, r, act)
    err = (out.double() - want).abs().max() / want.abs().max()
    assert err < 1e-5, err


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [0, 1])
def test_tcgen05_linear(M, N, K, act):
    lib = _lib.load()
    g = torch.Generator(device="cuda").manual_seed(M * 3 + K)
    A = torch.randn((M, K), generator=g, device="cuda")
    W = (torch.randn((N, K), generator=g, device="cuda") / K ** 0.5).contiguous()
    b = torch.randn((N,), generator=g, device="cuda")
    r = torch.randn((M, N), generator=g, device="cuda")
    W16 = W.to(torch.bfloat16)
    _lib.check(lib.mocha_register_bf16_blob(_lib.ptr(W), _lib.ptr(W16), W.numel()))
    out = run_linear(A, W, b, r, act, _lib.MOCHA_BF16)
    # exact reference on the bf16-rounded operands: only accumulation order differs
    want = torch_ref(A.to(torch.bfloat16).float(), W16.float(), b, r, act)
    err = (out.double() - want).abs().max() / want.abs().max()
    assert err < 2e-5, err


@pytest.mark.parametrize("B", [2, 16])
def test_tconv_pair_kernel_vs_fp32_path(B):
    """JointBlock temporal conv through mocha_bench_tconv: B=16 gives 96 pair tiles, enough for the 2-CTA
    (cta_group::2) kernel; B=2 stays on the 1-CTA kernel. Both must agree with the fp32 SIMT path on
    bf16-rounded inputs (weights are bf16-rounded by the tensor-core path: tolerance 2e-2 of the range)."""
    from mocha_sigasia2023_b200 import packing, weights
    lib = _lib.load()
    pk = packing.PackedGenerator(weights.generator_state_dict(1777), weights.DEFAULT_MODEL_CFG, torch.device("cuda"))
    d = pk.struct.dims
    rows = B * d.T * d.V
    g = torch.Generator(device="cuda").manual_seed(B)
    x = torch.randn((rows, d.D), generator=g, device="cuda").to(torch.bfloat16).float()
    ws = torch.empty(lib.mocha_embed_workspace_bytes(C.byref(d), B), dtype=torch.uint8, device="cuda")
    outs = []
    for prec in (_lib.MOCHA_FP32, _lib.MOCHA_BF16):
        out = torch.full((rows, d.D), float("nan"), device="cuda")
        _lib.check(lib.mocha_bench_tconv(C.byref(pk.struct), _lib.ptr(x), B, _lib.ptr(out), prec, 1, _lib.ptr(ws), ws.numel(),
                                         _lib.stream_ptr()), "mocha_bench_tconv")
        torch.cuda.synchronize()
        outs.append(out)
    assert torch.isfinite(outs[1]).all()
    err = (outs[1] - outs[0]).abs().max() / outs[0].abs().max()
    assert err < 2e-2, err
