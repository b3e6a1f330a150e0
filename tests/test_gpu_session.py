"""GPU parity of the whole per-frame loop (session) vs the CPU oracle pipeline, frame by frame."""
import numpy as np
import pytest
import torch

from mocha_oracle.pipeline import OraclePipeline
from mocha_sigasia2023_b200 import skeleton, workload

pytestmark = pytest.mark.gpu

KEYS = ("pos", "rot", "vel", "ang", "blend_pos", "ik_pos", "src_root_pos", "src_root_rot")


def _np_sd(sd):
    return {k: v.numpy() for k, v in sd.items()}


def _run(precision, frames, B, use_graph, tol):
    sess, gen_sd, cvae_sd, stats = workload.build_session(B, n_db=48, precision=precision)
    ora = OraclePipeline(_np_sd(gen_sd), _np_sd(cvae_sd), workload.stats_as_dict(stats),
                         sess.cha_encoded.cpu().numpy(), sess.tree.data.cpu().numpy(), B, skeleton.BONE_PARENTS)
    for f in range(frames):
        inp = workload.step_inputs(B, seed=100 + f)
        if use_graph and f == 2:
            sess.capture()
        got = sess.step_host(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        want = ora.step(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        if precision in ("fp32", "tf32x3"):
            np.testing.assert_array_equal(sess.match_idx[:, 0].cpu().numpy(), ora.last["match_idx"])
        Y = sess.Y.cpu().numpy()
        err = np.abs(Y - ora.last["Y"]).max() / np.abs(ora.last["Y"]).max()
        assert err < tol, f"frame {f}: decoder output rel err {err}"
        for k in KEYS:
            np.testing.assert_allclose(got[k], want[k], rtol=10 * tol, atol=10 * tol, err_msg=f"frame {f} {k}")
        dots = np.abs((got["ik_rot"] * want["ik_rot"]).sum(-1))
        assert dots.min() > 1 - 100 * tol, f"frame {f} ik_rot"


def test_session_fp32_eager():
    _run("fp32", frames=4, B=3, use_graph=False, tol=2e-4)


def test_session_fp32_cuda_graph():
    _run("fp32", frames=5, B=2, use_graph=True, tol=2e-4)


def test_session_tf32x3_cuda_graph():
    """the parity mode on the tensor cores (3xTF32 GEMMs): same tolerance and exact match indices as fp32"""
    _run("tf32x3", frames=4, B=3, use_graph=True, tol=2e-4)


def test_session_bf16():
    _run("bf16", frames=3, B=2, use_graph=False, tol=3e-2)


def test_session_bf16_batched_kernel_variants():
    """16 clips reach the kernel variants of the batched step that small batches never take (split-K
    tensor-core matcher, CTA-pair temporal conv, 64-column TMA boxes); one frame against the CPU oracle."""
    _run("bf16", frames=1, B=16, use_graph=False, tol=3e-2)
