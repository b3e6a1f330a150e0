"""Window feature extraction on the GPU (test_fullframework.py:135-186) vs the CPU oracle restatement
(oracle/mocha_oracle/clip.window_features, itself pinned end to end by tests/test_oracle_e2e.py), and the training-side
twins (trainer.py:249-374, train_CVAE.py:196-211) vs goldens recorded from the reference functions."""
import os

import numpy as np
import pytest
import torch

from mocha_oracle import clip, matching
from mocha_sigasia2023_b200 import features, preprocess, skeleton, synthetic, training_twins

pytestmark = pytest.mark.gpu


def test_window_features_vs_oracle():
    raw = synthetic.make_norm_stats()
    win = preprocess.process_clip(synthetic.make_clip(240, 0))
    want = clip.window_features(win, raw["norm"]["X_mean"], raw["norm"]["X_std"], skeleton.BONE_PARENTS)
    got = features.extract(win, raw["norm"]["X_mean"], raw["norm"]["X_std"])
    assert tuple(got["X"].shape) == want["X"].shape == (225, 60, 24, 15)
    for k in ("X", "Yrvel", "Yrang", "Ypos", "Yvel"):
        g = got[k].cpu().numpy()
        assert np.abs(g - want[k]).max() <= 2e-4 * max(1.0, np.abs(want[k]).max()), k
    dots = np.abs((got["Yrot"].cpu().numpy() * want["Yrot"]).sum(-1))
    assert dots.min() > 1 - 1e-5
    np.testing.assert_array_equal(got["contacts"].cpu().numpy(), want["contacts"])


def test_convert_YtilToX_and_recon_criterion_vs_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "train.npz"))
    Ytil, Ygt = torch.from_numpy(g["Ytil"]).cuda(), torch.from_numpy(g["Ygt"]).cuda()
    X = training_twins.convert_YtilToX(Ytil, Ygt[:, :, 0:1], skeleton.BONE_PARENTS).cpu().numpy()
    assert X.shape == g["X"].shape
    np.testing.assert_allclose(X, g["X"], rtol=2e-4, atol=2e-4 * np.abs(g["X"]).max())
    loss = float(training_twins.recon_criterion(Ytil, Ygt, skeleton.BONE_PARENTS))
    assert abs(loss - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))


def test_action_matcher_vs_oracle():
    rng = np.random.default_rng(9)
    rows = rng.standard_normal((300, 90, 256)).astype(np.float32)[:, 0]     # first context token rows, [n, 256]
    labels = rng.integers(0, 4, size=300)
    m = training_twins.ActionMatcher(rows, labels)
    q = rng.standard_normal((16, 256)).astype(np.float32)
    for lab in range(4):
        idx = np.where(labels == lab)[0]
        want = matching.knn(rows[idx], q, 1)[1][:, 0]
        np.testing.assert_array_equal(m.nearest(lab, q), want)
    assert m.nearest(7, q).shape == (0,)
