"""The oracle's whole-clip driver (oracle/mocha_oracle/clip.py: window features -> encode -> OraclePipeline
frame loop -> final payload) pinned to tests/golden/e2e.npz, the recording of the UNMODIFIED reference
`test_fullframework.main()` (oracle/ref_harness.py). CPU only: this is what entitles smoke(), the session
tests and bench.py's cpu_baseline to use OraclePipeline as the checker."""
import os

import numpy as np
import pytest

from mocha_oracle import clip
from mocha_sigasia2023_b200 import preprocess, skeleton, synthetic, weights

FRAMES = 48   # frames of the 225-frame loop replayed here (CPU budget); the GPU e2e tests cover all 225


@pytest.fixture(scope="module")
def run(golden_dir):
    g = np.load(os.path.join(golden_dir, "e2e.npz"))
    gen_sd = {k: v.numpy() for k, v in weights.generator_state_dict(1777).items()}
    cvae_sd = {k: v.numpy() for k, v in weights.cvae_state_dict(1778).items()}
    out = clip.run_clip(preprocess.process_clip(synthetic.make_clip(240, 0)), preprocess.process_clip(synthetic.make_clip(400, 1)),
                        synthetic.make_norm_stats(), gen_sd, cvae_sd, skeleton.BONE_PARENTS, eps_seq=g["eps"], frames=FRAMES)
    return g, out


def _angle_err_deg(a, b):
    d = np.abs(a - b) % 360.0
    return np.minimum(d, 360.0 - d)


def test_db_and_matched_indices(run):
    g, out = run
    assert out["n_db"] == int(g["db_shape"][0])
    np.testing.assert_array_equal(out["match"], g["match"][:FRAMES])


def test_network_output_rows(run):
    g, out = run
    ref = g["Ytil_last_rows"][0::2][:FRAMES]     # the 'trans' decode of each frame (even entries)
    rel = np.abs(out["Ytil_last_rows"] - ref).reshape(FRAMES, -1).max(axis=1) / np.abs(ref).max()
    assert rel[0] < 1e-5 and rel.max() < 1e-3, (rel.argmax(), rel.max())


def test_payloads(run):
    g, out = run
    np.testing.assert_allclose(out["src_positions"], g["src_positions"][:FRAMES], rtol=1e-4, atol=1e-4)
    assert _angle_err_deg(out["src_rotations"], g["src_rotations"][:FRAMES]).max() < 0.05
    np.testing.assert_allclose(out["ours_positions"], g["ours_positions"][:FRAMES], rtol=1e-3, atol=1e-3)
    err = _angle_err_deg(out["ours_rotations"], g["ours_rotations"][:FRAMES])
    assert np.median(err) < 0.01 and (err < 0.5).mean() > 0.995
