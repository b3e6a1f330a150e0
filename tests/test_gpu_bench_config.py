"""Parity at the configuration bench.py measures: 128 clips per GPU, bf16 tensor-core mode, 385-row character
DB, steady-state frames replayed from ONE CUDA graph - against the CPU oracle pipeline on the same inputs.
B = 128 takes kernel variants no smaller batch reaches (256-wide tiles, CTA-pair temporal conv from 13 clips,
split-K matcher with 128 queries, fused transformer kernels at 90 m-tiles)."""
import numpy as np
import pytest

from mocha_oracle import matching
from mocha_oracle.pipeline import OraclePipeline
from mocha_sigasia2023_b200 import skeleton, workload

pytestmark = pytest.mark.gpu

B, N_DB, FRAMES = 128, 385, 4
KEYS = ("pos", "rot", "vel", "ang", "blend_pos", "ik_pos", "src_root_pos", "src_root_rot")


@pytest.fixture(scope="module")
def run():
    sess, gen_sd, cvae_sd, stats = workload.build_session(B, n_db=N_DB, precision="bf16")
    ora = OraclePipeline({k: v.numpy() for k, v in gen_sd.items()}, {k: v.numpy() for k, v in cvae_sd.items()},
                         workload.stats_as_dict(stats), sess.cha_encoded.cpu().numpy(), sess.tree.data.cpu().numpy(), B,
                         skeleton.BONE_PARENTS)
    db = sess.tree.data.cpu().numpy()
    frames = []
    for f in range(FRAMES):
        inp = workload.step_inputs(B, seed=300 + f)
        if f == 2:
            sess.capture()                      # frames 2.. replay the captured graph, exactly like bench.py
        got = sess.step_host(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        want = ora.step(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        frames.append({"got": got, "want": want, "Y": sess.Y.cpu().numpy(), "Y_ora": ora.last["Y"],
                       "idx": sess.match_idx[:, 0].cpu().numpy().copy(), "idx_ora": ora.last["match_idx"].copy(),
                       "q_gpu": sess.cnt_nm.cpu().numpy().copy(),
                       "q_ora": ((ora.last["cnt"] - ora.st["cnt_mean"][None]) / ora.st["cnt_std"][None]).reshape(B, -1),
                       "graph": sess.graph is not None and f >= 2})
    return sess, db, frames


def test_graph_was_replayed(run):
    _, _, frames = run
    assert [f["graph"] for f in frames] == [False, False, True, True]


def test_decoder_output_vs_oracle(run):
    """bf16 mode: 2e-2 of the tensor's range (max-abs error over max-abs value), every frame."""
    _, _, frames = run
    for i, f in enumerate(frames):
        err = np.abs(f["Y"] - f["Y_ora"]).max() / np.abs(f["Y_ora"]).max()
        assert err < 2e-2, f"frame {i}: decoder output error {err:.3e} of range"
        # per-clip bound as well, so one bad clip cannot hide in the batch maximum
        per = np.abs(f["Y"] - f["Y_ora"]).reshape(B, -1).max(1) / np.abs(f["Y_ora"]).reshape(B, -1).max(1)
        assert per.max() < 3e-2, f"frame {i}: worst clip {per.argmax()} error {per.max():.3e}"


def test_pose_outputs_vs_oracle(run):
    _, _, frames = run
    for i, f in enumerate(frames):
        worst = {k: float(np.abs(f["got"][k] - f["want"][k]).max()) for k in KEYS}
        print(f"[bench-config frame {i}] max abs pose errors {worst}")
        for k in KEYS:
            np.testing.assert_allclose(f["got"][k], f["want"][k], rtol=0.3, atol=0.3, err_msg=f"frame {i} {k}")
        dots = np.abs((f["got"]["ik_rot"] * f["want"]["ik_rot"]).sum(-1))
        assert dots.min() > 0.97, f"frame {i} ik_rot {dots.min()}"


def test_matched_indices_bf16_mode(run):
    """The matcher is exact for the query it is given: against the float64 oracle k-NN of the GPU's OWN query
    (the context feature of the bf16 encoder) indices are bit-exact wherever the margin exceeds 1e-5. Against the
    fp32 oracle's query the feature itself differs by e = ||q_gpu - q_ora||, which moves every distance by at most
    e: indices must agree wherever the oracle's top-1 / top-2 margin exceeds 2e; the overall agreement is reported."""
    _, db, frames = run
    tot = agree_all = 0
    for i, f in enumerate(frames):
        wd, wi = matching.knn_gemm(db, f["q_gpu"], 2)
        ok = (wd[:, 1] - wd[:, 0]) > 1e-5
        np.testing.assert_array_equal(f["idx"][ok], wi[ok, 0], err_msg=f"frame {i}: matcher vs fp64 k-NN of its own query")
        od, oi = matching.knn_gemm(db, f["q_ora"], 2)
        assert (oi[:, 0] == f["idx_ora"]).all()
        e = np.linalg.norm(f["q_gpu"].astype(np.float64) - f["q_ora"].astype(np.float64), axis=1)
        must = (od[:, 1] - od[:, 0]) > 2.0 * e + 1e-5
        np.testing.assert_array_equal(f["idx"][must], f["idx_ora"][must], err_msg=f"frame {i}: bf16 vs fp32-oracle indices")
        tot += B
        agree_all += int((f["idx"] == f["idx_ora"]).sum())
        print(f"[bench-config frame {i}] query rel err {np.median(e / np.linalg.norm(f['q_ora'], axis=1)):.3e}, "
              f"margin-certified {int(must.sum())}/{B}, index agreement {(f['idx'] == f['idx_ora']).mean():.4f}")
    assert agree_all / tot >= 0.9, f"bf16-mode matched-index agreement with the fp32 oracle: {agree_all / tot:.4f}"
