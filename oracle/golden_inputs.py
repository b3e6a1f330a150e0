"""Seeded inputs shared by oracle/gen_golden.py (which feeds them to the reference) and tests/
(which feed them to the oracle and to the CUDA path). TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import numpy as np


def pose_windows(B=2, seed=0):
    rng = np.random.default_rng(seed)
    src = rng.standard_normal((B, 60, 24, 15)).astype(np.float32)
    cha = rng.standard_normal((B, 60, 24, 15)).astype(np.float32)
    return src, cha


def cvae_inputs(B=2, seed=1):
    rng = np.random.default_rng(seed)
    cond = rng.standard_normal((B, 180, 256)).astype(np.float32)
    eps = rng.standard_normal((B, 256)).astype(np.float32)
    return cond, eps


def cvae_posterior_inputs(B=2, seed=2):
    """x [B, 90, 256] for CVAE.encode / CVAE.forward (the condition and eps are cvae_inputs()'s)."""
    return np.random.default_rng(seed).standard_normal((B, 90, 256)).astype(np.float32)


def _rand_quat(rng, shape, dtype):
    q = rng.standard_normal(tuple(shape) + (4,))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    return q.astype(dtype)


def kin_inputs(seed=2):
    rng = np.random.default_rng(seed)
    F, J = 37, 25
    d = {
        "lrot": _rand_quat(rng, (F, J), np.float32),
        "lpos": (0.3 * rng.standard_normal((F, J, 3))).astype(np.float32),
        "lvel": rng.standard_normal((F, J, 3)).astype(np.float32),
        "lang": rng.standard_normal((F, J, 3)).astype(np.float32),
        "xy": rng.standard_normal((200, 3, 2)).astype(np.float32),
        "vec3": np.concatenate([rng.standard_normal((30, 3)), 1e-7 * rng.standard_normal((4, 3))]).astype(np.float32),
    }
    n = 48
    root = rng.standard_normal((n, 3))
    mid = root + np.array([0.0, -0.45, 0.05]) + 0.05 * rng.standard_normal((n, 3))
    end = mid + np.array([0.0, -0.45, -0.05]) + 0.05 * rng.standard_normal((n, 3))
    target = end + 0.15 * rng.standard_normal((n, 3))
    target[::6] = root[::6] + 3.0 * (end[::6] - root[::6])   # out of reach -> clamped
    d["ik2"] = {"root": root, "mid": mid, "end": end, "target": target, "fwd": rng.standard_normal((n, 3)),
                "root_gr": _rand_quat(rng, (n,), np.float64), "mid_gr": _rand_quat(rng, (n,), np.float64),
                "par_gr": _rand_quat(rng, (n,), np.float64)}
    # contact trajectories: a foot moving forward with stance phases (flag on) and a few far jumps
    S, L = 6, 90
    t = np.arange(L) / 60.0
    pos = np.zeros((S, L, 3))
    flag = np.zeros((S, L), dtype=np.uint8)
    for s in range(S):
        phase = rng.uniform(0, 2 * np.pi)
        speed = rng.uniform(0.5, 2.0)
        pos[s, :, 0] = 0.1 * np.sin(2 * np.pi * 0.7 * t + phase)
        pos[s, :, 1] = 0.02 + 0.08 * np.maximum(0.0, np.sin(2 * np.pi * 1.3 * t + phase))
        pos[s, :, 2] = speed * t + 0.15 * np.sin(2 * np.pi * 1.3 * t + phase)
        flag[s] = (np.sin(2 * np.pi * 1.3 * t + phase) < 0.1).astype(np.uint8)
        if s % 2 == 1:
            pos[s, 40:, 2] += 0.6          # jump beyond the unlock radius while locked
    d["contact"] = {"pos": pos, "flag": flag}
    n = 5
    d["pose"] = {
        "root_pos": rng.standard_normal((n, 3)), "root_vel": rng.standard_normal((n, 3)),
        "root_rot": _rand_quat(rng, (n,), np.float64), "root_ang": rng.standard_normal((n, 3)),
        "src_pos": rng.standard_normal((n, J, 3)), "src_vel": rng.standard_normal((n, J, 3)),
        "src_rot": _rand_quat(rng, (n, J), np.float64), "src_ang": rng.standard_normal((n, J, 3)),
        "dst_pos": rng.standard_normal((n, J, 3)), "dst_vel": rng.standard_normal((n, J, 3)),
        "dst_rot": _rand_quat(rng, (n, J), np.float64), "dst_ang": rng.standard_normal((n, J, 3)),
    }
    return d


# name -> (N, D, nq, k)
MATCH_CASES = {
    "small": (300, 23040, 5, 3),      # the reference's regime: a few hundred rows of 90*256 features
    "ragged": (77, 200, 9, 4),        # D not a multiple of 4/64, N not a multiple of anything
    "planted": (2048, 512, 64, 2),    # queries = DB rows + small noise (large margins)
    "single": (1, 64, 3, 1),          # one-row DB
}


def match_inputs(name):
    N, D, nq, k = MATCH_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    db = rng.standard_normal((N, D)).astype(np.float32)
    if name == "planted":
        pick = rng.integers(0, N, size=nq)
        q = (db[pick] + 0.05 * rng.standard_normal((nq, D))).astype(np.float32)
    else:
        q = rng.standard_normal((nq, D)).astype(np.float32)
    return db, q
