"""GPU parity: the CUDA Generator / CVAE behind the C ABI vs the reference goldens and the oracle."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from mocha_oracle import nets
from mocha_sigasia2023_b200 import weights
from mocha_sigasia2023_b200.model import Generator
from mocha_sigasia2023_b200.model_CVAE import CVAE
from mocha_sigasia2023_b200.transformer import mean_variance_norm

pytestmark = pytest.mark.gpu

RTOL_FP32 = 1e-4   # north_star: 1e-4 relative in fp32
RTOL_BF16 = 2e-2   # 2e-2 in bf16


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "nets.npz"))


@pytest.fixture(scope="module")
def gen():
    g = Generator(weights.DEFAULT_MODEL_CFG)
    g.load_state_dict(weights.generator_state_dict(1777), strict=True)
    return g.to("cuda").eval()


@pytest.fixture(scope="module")
def cvae():
    c = CVAE(output_seq=90)
    c.load_state_dict(weights.cvae_state_dict(1778), strict=True)
    return c.to("cuda").eval()


def cu(x):
    return torch.from_numpy(np.ascontiguousarray(x)).cuda()


def test_mot_embedding(gen, gold):
    src, _ = gi.pose_windows()
    tok = gen.mot_embedding(cu(src)).cpu().numpy()
    assert rel_err(tok, gold["tokens"]) < RTOL_FP32


def test_encoder_and_cnt(gen, gold):
    src, cha = gi.pose_windows()
    tok = gen.mot_embedding(cu(src))
    tok = tok + gen.pos_emb[:, :tok.shape[1]]
    enc = gen.encoder(tok)
    assert rel_err(enc.cpu().numpy(), gold["src_encoded"]) < RTOL_FP32
    cnt = mean_variance_norm(enc.permute(0, 2, 1)).permute(0, 2, 1)
    assert rel_err(cnt.cpu().numpy(), gold["src_cnt"]) < RTOL_FP32


def test_decoder_and_to_mot(gen, gold):
    dec = gen.decoder(cu(gold["src_encoded"]), cu(gold["cha_encoded"]))
    assert rel_err(dec.cpu().numpy(), gold["decoded"]) < RTOL_FP32
    y = gen.to_mot(cu(gold["decoded"]))
    assert rel_err(y.cpu().numpy(), gold["Ytil"]) < RTOL_FP32


def test_forward_signature(gen, gold):
    src, cha = gi.pose_windows()
    y = gen(cu(src), cu(cha))
    assert tuple(y.shape) == (2, 60, 24, 15)
    assert rel_err(y.cpu().numpy(), gold["forward"]) < 2 * RTOL_FP32
    feats = gen(cu(src), cu(cha), extract_feature=True)
    assert len(feats) == 4
    assert rel_err(feats[3].cpu().numpy(), gold["feat_cha_cnt"]) < 2 * RTOL_FP32


def test_batch_sizes_against_oracle(gen):
    sd = {k: v.numpy() for k, v in weights.generator_state_dict(1777).items()}
    for B, seed in ((1, 5), (3, 6)):
        src, cha = gi.pose_windows(B=B, seed=seed)
        want = nets.generator_forward(sd, src, cha)
        got = gen(cu(src), cu(cha)).cpu().numpy()
        assert rel_err(got, want) < 2 * RTOL_FP32


def test_cvae_sample(cvae, gold):
    cond, eps = gi.cvae_inputs()
    out = cvae.sample(cu(cond), deterministic=True).cpu().numpy()
    assert rel_err(out, gold["cvae_det"]) < RTOL_FP32
    mu, logvar = cvae.prior(cu(cond))
    assert rel_err(mu.cpu().numpy(), gold["cvae_mu"]) < RTOL_FP32
    assert rel_err(logvar.cpu().numpy(), gold["cvae_logvar"]) < RTOL_FP32
    cvae.eps_fn = lambda shape: torch.from_numpy(eps)
    try:
        out_e = cvae.sample(cu(cond), deterministic=False).cpu().numpy()
    finally:
        cvae.eps_fn = None
    assert rel_err(out_e, gold["cvae_eps"]) < RTOL_FP32


def test_bf16_tensor_core_mode(gold):
    g = Generator(weights.DEFAULT_MODEL_CFG, precision="bf16")
    g.load_state_dict(weights.generator_state_dict(1777), strict=True)
    g = g.to("cuda").eval()
    src, cha = gi.pose_windows()
    tok = g.mot_embedding(cu(src)).cpu().numpy()
    assert rel_err(tok, gold["tokens"]) < RTOL_BF16
    enc = g.encoder(cu(gold["tokens"]) + g.pos_emb[:, :90]).cpu().numpy()
    assert rel_err(enc, gold["src_encoded"]) < RTOL_BF16
    dec = g.decoder(cu(gold["src_encoded"]), cu(gold["cha_encoded"])).cpu().numpy()
    assert rel_err(dec, gold["decoded"]) < RTOL_BF16
    y = g.to_mot(cu(gold["decoded"])).cpu().numpy()
    assert rel_err(y, gold["Ytil"]) < RTOL_BF16
    c = CVAE(output_seq=90, precision="bf16")
    c.load_state_dict(weights.cvae_state_dict(1778), strict=True)
    c = c.to("cuda").eval()
    cond, _ = gi.cvae_inputs()
    out = c.sample(cu(cond), deterministic=True).cpu().numpy()
    assert rel_err(out, gold["cvae_det"]) < RTOL_BF16
    # the last prior layer runs on its two read rows in one SIMT kernel (cvae_prior_last): mu / logvar directly
    mu, logvar = c.prior(cu(cond))
    assert rel_err(mu.cpu().numpy(), gold["cvae_mu"]) < RTOL_BF16
    assert rel_err(logvar.cpu().numpy(), gold["cvae_logvar"]) < RTOL_BF16


def test_bf16_mode_at_bench_batch_sizes():
    """The batched step takes kernel variants small batches never reach (CTA-pair temporal conv from 13 clips,
    64-column bf16 TMA boxes, split-K matcher from 16): the whole Generator at B = 16 in bf16 tensor-core mode
    must agree with the fp32 parity mode on the same inputs (2e-2 of the range, north_star's bf16 tolerance)."""
    B = 16
    src, cha = gi.pose_windows()
    rng = np.random.default_rng(16)
    x = np.concatenate([src, cha] * (B // 2), axis=0)[:B].astype(np.float32)
    x = x + 0.01 * rng.standard_normal(x.shape).astype(np.float32)
    outs = {}
    for prec in ("fp32", "bf16"):
        g = Generator(weights.DEFAULT_MODEL_CFG, precision=prec)
        g.load_state_dict(weights.generator_state_dict(1777), strict=True)
        g = g.to("cuda").eval()
        tok = g.mot_embedding(cu(x))
        enc = g.encoder(tok + g.pos_emb[:, :90])
        dec = g.decoder(enc, enc.flip(0))
        y = g.to_mot(dec)
        outs[prec] = [t.cpu().numpy() for t in (tok, enc, dec, y)]
    for name, a, b in zip(("tokens", "encoded", "decoded", "Ytil"), outs["bf16"], outs["fp32"]):
        assert np.isfinite(a).all(), name
        assert rel_err(a, b) < RTOL_BF16, (name, rel_err(a, b))


def test_tf32x3_tensor_core_parity_mode(gold):
    """MOCHA_TF32X3: every linear layer / temporal convolution as a split-fp32 (3xTF32) tcgen05 GEMM must meet the fp32
    tolerance (1e-4 relative, north_star) against the reference goldens - the parity mode on the tensor cores."""
    g = Generator(weights.DEFAULT_MODEL_CFG, precision="tf32x3")
    g.load_state_dict(weights.generator_state_dict(1777), strict=True)
    g = g.to("cuda").eval()
    src, cha = gi.pose_windows()
    tok = g.mot_embedding(cu(src)).cpu().numpy()
    assert rel_err(tok, gold["tokens"]) < RTOL_FP32
    enc = g.encoder(cu(gold["tokens"]) + g.pos_emb[:, :90]).cpu().numpy()
    assert rel_err(enc, gold["src_encoded"]) < RTOL_FP32
    dec = g.decoder(cu(gold["src_encoded"]), cu(gold["cha_encoded"])).cpu().numpy()
    assert rel_err(dec, gold["decoded"]) < RTOL_FP32
    y = g.to_mot(cu(gold["decoded"])).cpu().numpy()
    assert rel_err(y, gold["Ytil"]) < RTOL_FP32
    c = CVAE(output_seq=90, precision="tf32x3")
    c.load_state_dict(weights.cvae_state_dict(1778), strict=True)
    c = c.to("cuda").eval()
    cond, _ = gi.cvae_inputs()
    out = c.sample(cu(cond), deterministic=True).cpu().numpy()
    assert rel_err(out, gold["cvae_det"]) < RTOL_FP32


def test_tf32x3_linear_chunked_and_periodic_bias():
    """mocha_linear in MOCHA_TF32X3 mode vs float64: a workspace that only fits a fraction of the split A operand forces the
    row-chunked path; error must stay at fp32 level (far below bf16's 4e-3)."""
    import ctypes as C
    from mocha_sigasia2023_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    M, N, K = 1000, 264, 192
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    R = rng.standard_normal((M, N)).astype(np.float32)
    want = np.maximum(A.astype(np.float64) @ W.astype(np.float64).T + b, 0) + R
    dA, dW, db, dR = cu(A), cu(W), cu(b), cu(R)
    for ws_rows in (M + 8, 300):     # whole operand / 256-row chunks
        nbytes = (N + ws_rows) * 3 * K * 4 + 8192
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        out = torch.empty((M, N), dtype=torch.float32, device="cuda")
        _lib.check(lib.mocha_linear(_lib.ptr(dA), _lib.ptr(dW), _lib.ptr(db), _lib.ptr(dR), _lib.ptr(out), M, N, K, 1,
                                    _lib.MOCHA_TF32X3, _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "mocha_linear")
        got = out.cpu().numpy()
        assert np.abs(got - want).max() < 2e-5, (ws_rows, np.abs(got - want).max())


@pytest.mark.parametrize("precision,tol", [("fp32", RTOL_FP32), ("tf32x3", RTOL_FP32), ("bf16", RTOL_BF16)])
def test_cvae_posterior_surface(gold, precision, tol):
    """CVAE.encode / CVAE.forward (model_CVAE.py:33-42): the posterior Encoder is the prior's network over 272 tokens
    [mu, logvar, c, x]; forward decodes z_po. Noise replayed from the golden run (posterior draw first, prior draw second)."""
    c = CVAE(output_seq=90, precision=precision)
    c.load_state_dict(weights.cvae_state_dict(1778), strict=True)
    c = c.to("cuda").eval()
    cond, eps = gi.cvae_inputs()
    x = gi.cvae_posterior_inputs()
    mu, logvar = c.encode(cu(x), cu(cond))
    assert rel_err(mu.cpu().numpy(), gold["cvae_enc_mu"]) < tol
    assert rel_err(logvar.cpu().numpy(), gold["cvae_enc_logvar"]) < tol
    draws = iter([torch.from_numpy(eps), torch.zeros(eps.shape)])
    c.eps_fn = lambda shape: next(draws)
    out, (mu_po, lv_po), (mu_pr, lv_pr) = c(cu(x), cu(cond))
    assert rel_err(out.cpu().numpy(), gold["cvae_fwd"]) < 2 * tol
    assert rel_err(mu_po.cpu().numpy(), gold["cvae_enc_mu"]) < tol
    assert rel_err(mu_pr.cpu().numpy(), gold["cvae_mu"]) < tol
    assert rel_err(lv_pr.cpu().numpy(), gold["cvae_logvar"]) < tol


def test_tf32x3_chunked_when_workspace_is_short(gold):
    """MOCHA_TF32X3 with a workspace sized for the OTHER modes (no room for whole split operands): the linear layers run
    as row chunks and the temporal convolutions as image chunks; results must not change."""
    import ctypes as C
    from mocha_sigasia2023_b200 import _lib
    lib = _lib.load()
    g = Generator(weights.DEFAULT_MODEL_CFG, precision="tf32x3")
    g.load_state_dict(weights.generator_state_dict(1777), strict=True)
    g = g.to("cuda").eval()
    pk = g._pack()
    src, cha = gi.pose_windows()
    X = cu(np.concatenate([src, cha, src, cha], axis=0))           # 8 clips: several image chunks
    B = X.shape[0]
    nbytes = lib.mocha_embed_workspace_bytes(C.byref(pk.struct.dims), B)      # default (fp32 / bf16) sizing
    with _lib.workspace_precision(_lib.MOCHA_TF32X3):
        assert lib.mocha_embed_workspace_bytes(C.byref(pk.struct.dims), B) > 2 * nbytes
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty((B, pk.ntok, pk.dims.D), dtype=torch.float32, device="cuda")
    _lib.check(lib.mocha_embed_fwd(C.byref(pk.struct), _lib.ptr(X), B, _lib.ptr(out), 0, _lib.MOCHA_TF32X3, _lib.ptr(ws),
                                   ws.numel(), _lib.stream_ptr()), "mocha_embed_fwd")
    tok = out.cpu().numpy()
    assert rel_err(tok[:2], gold["tokens"]) < RTOL_FP32
    assert rel_err(tok[4:6], gold["tokens"]) < RTOL_FP32
