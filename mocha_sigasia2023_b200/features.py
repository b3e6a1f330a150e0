"""Window feature extraction on the GPU: the driver's set-up math between `process_data` and
`mot_embedding` (test_fullframework.py:135-186) — FK with velocities, re-rooting every window on
its last frame's simulation root, inverse kinematics back to local space, 6-D rotation encoding,
central-difference velocities and normalisation. SURVEY §8(f) "next" row 2.

Everything runs in the library's kernels: `mocha_fk_vel` (world space), `mocha_window_features` (re-rooting on
the window's last root + 6-D rotation encoding + normalisation, one pass), `mocha_ik` (back to local space),
`mocha_quat_op` (body-frame root velocities); only the 3-point central differences are torch expressions."""
from __future__ import annotations

import numpy as np
import torch

from . import kinematics as kin
from . import _lib, skeleton


def _central_diff_time(x):
    """(:164-169): central differences along the window axis, linear extrapolation at both ends."""
    v = torch.empty_like(x)
    v[:, 1:-1] = 0.5 * (x[:, 2:] - x[:, 1:-1]) * 60.0 + 0.5 * (x[:, 1:-1] - x[:, :-2]) * 60.0
    v[:, 0] = v[:, 1] - (v[:, 3] - v[:, 2])
    v[:, -1] = v[:, -2] + (v[:, -2] - v[:, -3])
    return v


def extract(windows: dict, X_mean: np.ndarray, X_std: np.ndarray, device="cuda") -> dict:
    """windows: output of preprocess.process_clip. X_mean/X_std: [25,15] tables of norm.npz.
    Returns CUDA tensors: X [nwin,60,24,15], Yrvel/Yrang [nwin,60,3], Ypos [nwin,60,25,3],
    Yrot [nwin,60,25,4], Yvel [nwin,60,25,3] and contacts [nwin,60,2] (uint8)."""
    f32 = dict(dtype=torch.float32, device=device)
    Ypos = torch.as_tensor(windows["pos"], **f32).contiguous()
    Yvel = torch.as_tensor(windows["vel"], **f32).contiguous()
    Yrot = torch.as_tensor(windows["rot"], **f32).contiguous()
    Yang = torch.as_tensor(windows["ang"], **f32).contiguous()
    par = kin.parents_tensor(windows.get("parents", skeleton.BONE_PARENTS), device)
    window = Ypos.shape[1]
    # local root velocities in the body frame (:142-143)
    nwin, J = Ypos.shape[0], Ypos.shape[2]
    r0 = Yrot[:, :, 0].reshape(-1, 4).contiguous()
    Yrvel = kin.quat_op("inv_mul_vec", r0, Yvel[:, :, 0].reshape(-1, 3).contiguous()).reshape(nwin, window, 3)
    Yrang = kin.quat_op("inv_mul_vec", r0, Yang[:, :, 0].reshape(-1, 3).contiguous()).reshape(nwin, window, 3)
    # world space (:146), then every window relative to its last frame's root, encoded and normalised (:148-158,:180-186)
    Grot, Gpos, Gvel, Gang = kin.fk_vel(Yrot, Ypos, Yvel, Yang, par)
    xm = torch.as_tensor(np.ascontiguousarray(np.asarray(X_mean, dtype=np.float32)), **f32).contiguous()
    xs = torch.as_tensor(np.ascontiguousarray(np.asarray(X_std, dtype=np.float32)), **f32).contiguous()
    X = torch.empty((nwin, window, J - 1, 15), **f32)
    Xrot = torch.empty_like(Grot)
    Xpos = torch.empty_like(Gpos)
    _lib.check(_lib.load().mocha_window_features(_lib.ptr(Grot), _lib.ptr(Gpos), _lib.ptr(Gvel), _lib.ptr(Gang), nwin, window, J,
                                                 _lib.ptr(xm), _lib.ptr(xs), _lib.ptr(X), _lib.ptr(Xrot), _lib.ptr(Xpos),
                                                 _lib.stream_ptr()), "mocha_window_features")
    Yrot2, Ypos2 = kin.ik(Xrot, Xpos, par)                             # (:160)
    Yvel2 = _central_diff_time(Ypos2)                                  # (:164-169)
    return {"X": X, "Yrvel": Yrvel.contiguous(), "Yrang": Yrang.contiguous(), "Ypos": Ypos2, "Yrot": Yrot2,
            "Yvel": Yvel2, "contacts": torch.as_tensor(windows["contacts"], dtype=torch.uint8, device=device)}
