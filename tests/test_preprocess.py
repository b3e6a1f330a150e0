"""Host-side clip ingestion (preprocess.process_clip) vs the reference's process_data golden."""
import os

import numpy as np

from mocha_sigasia2023_b200 import preprocess, skeleton, synthetic


def test_process_clip_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "prep.npz"))
    out = preprocess.process_clip(synthetic.make_clip(240, 0))
    assert out["pos"].shape[0] == int(g["nwin"]) == 225
    np.testing.assert_array_equal(out["parents"], g["parents"])
    assert list(out["parents"]) == skeleton.BONE_PARENTS
    for k in ("pos", "vel", "rot", "ang"):
        np.testing.assert_allclose(out[k][g["pick"]], g[k], rtol=1e-6, atol=1e-7)
        assert abs(out[k].sum(dtype=np.float64) - float(g[k + "_sum"])) < 1e-3
    np.testing.assert_array_equal(out["contacts"], g["contacts"])
    assert 0.02 < out["contacts"].mean() < 0.6      # the synthetic clip does exercise foot contacts


def test_edge_windows_are_padded():
    x = np.arange(100, dtype=np.float64)[:, None]
    w = preprocess.sliding_windows(x, 60, 1)
    assert w.shape == (85, 60, 1)                      # range(0, len - window//4)
    assert (w[-1][:8] == w[-1][8]).all() or w[-1][0, 0] == 84   # left padding repeats the first pose
    z = preprocess.sliding_windows(x, 60, 1, zero_pad=True)
    assert z[-1][0, 0] == 0.0 and z[-1][-1, 0] == 0.0
