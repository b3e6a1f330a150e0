"""Sub-batch lanes (CharacterizationSession(lanes=k)): the clips of a session cut into k groups whose frames run on separate
streams inside the captured graph must produce exactly what one lane produces (clips are independent)."""
import numpy as np
import pytest

from mocha_sigasia2023_b200 import workload

pytestmark = pytest.mark.gpu


def _run(lanes, B=12, frames=4):
    sess, *_ = workload.build_session(B, n_db=48, precision="bf16", lanes=lanes)
    out = []
    for f in range(frames):
        inp = workload.step_inputs(B, seed=700 + f)
        if f == 2:
            sess.capture()
        got = sess.step_host(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        out.append((sess.Y.cpu().numpy().copy(), got["ik_pos"].copy(), got["ik_rot"].copy(), sess.match_idx.cpu().numpy().copy()))
    return out


def test_lanes_are_bit_identical():
    one, three = _run(1), _run(3)
    for f, (a, b) in enumerate(zip(one, three)):
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y, err_msg=f"frame {f}")
