"""Synthetic motion clips and normalisation statistics (BASELINE.json configs: "synthetic clip,
skeleton from configs/dataset.yaml, random-init weights"). The released BVH files, norm.npz,
cnt_norm.npz and cvae_norm.npz are not available offline (reference download.sh:1-30), so every run
in this repository — parity, smoke and bench — uses these seeded generators.

A clip has the dict layout `motion/bvh.py:load` returns (rotations in degrees, Euler order 'zyx',
positions in cm, offsets, parents, names)."""
from __future__ import annotations

import numpy as np

from . import skeleton

# plausible T-pose offsets in cm for the 24 'mocha' joints (configs/dataset.yaml:54-61)
_OFFSETS_CM = np.array([
    [0, 90, 0],
    [9, -5, 0], [0, -42, 0], [0, -42, 0], [0, -6, 14],
    [0, 8, 0], [0, 10, 0], [0, 10, 0], [0, 10, 0],
    [4, 8, 0], [12, 0, 0], [27, 0, 0], [25, 0, 0],
    [0, 12, 0], [0, 6, 0], [0, 8, 0],
    [-4, 8, 0], [-12, 0, 0], [-27, 0, 0], [-25, 0, 0],
    [-9, -5, 0], [0, -42, 0], [0, -42, 0], [0, -6, 14]], dtype=np.float64)


def make_clip(n_frames: int = 240, seed: int = 0, fps: float = 60.0) -> dict:
    """Smooth synthetic locomotion-like clip: per-joint Euler curves amp*sin(2 pi f t + phi) with
    amp in [5,25] deg and f in [0.5,2] Hz, root drifting forward."""
    rng = np.random.default_rng(seed)
    J = len(skeleton.JOINT_NAMES)
    t = np.arange(n_frames) / fps
    amp = rng.uniform(5.0, 25.0, size=(J, 3))
    freq = rng.uniform(0.5, 2.0, size=(J, 3))
    phase = rng.uniform(0.0, 2 * np.pi, size=(J, 3))
    rotations = amp[None] * np.sin(2 * np.pi * freq[None] * t[:, None, None] + phase[None])
    rotations[:, 0, :] *= 0.3                                   # keep the pelvis mostly upright
    legs = [1, 2, 3, 4, 20, 21, 22, 23]
    rotations[:, legs, :] *= 0.3                                # slow feet -> stance phases (foot contacts) occur
    rotations[:, 0, 1] += 20.0 * np.sin(2 * np.pi * 0.1 * t + rng.uniform(0, 2 * np.pi))  # slow heading change
    offsets = _OFFSETS_CM + rng.normal(0.0, 1.0, size=_OFFSETS_CM.shape)
    positions = np.repeat(offsets[None], n_frames, axis=0)
    speed = rng.uniform(16.0, 32.0)                             # cm/s (slow walk so toes drop below 0.5 m/s)
    positions[:, 0, 0] = 10.0 * np.sin(2 * np.pi * 0.25 * t)
    positions[:, 0, 1] = offsets[0, 1] + 3.0 * np.sin(2 * np.pi * 1.8 * t)
    positions[:, 0, 2] = speed * t
    return {
        "rotations": rotations.astype(np.float64),
        "positions": positions.astype(np.float64),
        "offsets": offsets.astype(np.float64),
        "parents": np.array(skeleton.JOINT_PARENTS),
        "names": list(skeleton.JOINT_NAMES),
        "order": "zyx",
    }


def make_norm_stats(seed: int = 7) -> dict:
    """Stand-ins for norm.npz / cnt_norm.npz / cvae_norm.npz with the shapes the driver expects
    (test_fullframework.py:64-92): mean ~ N(0, 0.1), std ~ U(0.5, 1.5), std_weight = linspace(1,3,15)
    repeated over the 6 body parts (train_CVAE.py:64-66)."""
    rng = np.random.default_rng(seed)

    def mean(shape):
        return rng.normal(0.0, 0.1, size=shape).astype(np.float32)

    def std(shape):
        return rng.uniform(0.5, 1.5, size=shape).astype(np.float32)

    w = np.repeat(np.linspace(1.0, 3.0, 15, dtype=np.float32), 6)[:, None] * np.ones((1, 256), dtype=np.float32)
    return {
        "norm": {"X_mean": mean((25, 15)), "X_std": std((25, 15)), "Y_mean": mean((25, 15)), "Y_std": std((25, 15))},
        "cnt_norm": {"mean": mean((90, 256)), "std": std((90, 256))},
        "cvae_norm": {"std_weight": w, "src_cnt_mean": mean((90, 256)), "src_cnt_std": std((90, 256)),
                      "cha_cnt_mean": mean((90, 256)), "cha_cnt_std": std((90, 256)),
                      "cha_encoded_mean": mean((90, 256)), "cha_encoded_std": std((90, 256))},
    }
