"""Window feature extraction on the GPU: the driver's set-up math between `process_data` and
`mot_embedding` (test_fullframework.py:135-186) — FK with velocities, re-rooting every window on
its last frame's simulation root, inverse kinematics back to local space, 6-D rotation encoding,
central-difference velocities and normalisation. SURVEY §8(f) "next" row 2.

Heavy parts (fk_vel, ik, to_xform_xy over nwin*60 skeletons) run in the library's kernels; the
re-rooting algebra is element-wise torch on CUDA tensors."""
from __future__ import annotations

import numpy as np
import torch

from . import kinematics as kin
from . import skeleton, tq


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
    Yrvel = tq.inv_mul_vec(Yrot[:, :, 0], Yvel[:, :, 0])
    Yrang = tq.inv_mul_vec(Yrot[:, :, 0], Yang[:, :, 0])
    # world space (:146), then every frame's root replaced by the window's last root (:148-151)
    Grot, Gpos, Gvel, Gang = kin.fk_vel(Yrot, Ypos, Yvel, Yang, par)
    for G in (Gpos, Grot, Gvel, Gang):
        G[:, :, 0:1] = G[:, -1:, 0:1].expand(-1, window, -1, -1).clone()
    R0 = Grot[:, :, 0:1]
    Xpos = tq.inv_mul_vec(R0, Gpos - Gpos[:, :, 0:1])                 # (:154-158)
    Xrot = tq.inv_mul(R0, Grot)
    Xtxy = kin.quat_to_xy(Xrot.contiguous())
    Xvel = tq.inv_mul_vec(R0, Gvel)
    Xang = tq.inv_mul_vec(R0, Gang)
    Yrot2, Ypos2 = kin.ik(Xrot.contiguous(), Xpos.contiguous(), par)   # (:160)
    Yvel2 = _central_diff_time(Ypos2)                                  # (:164-169)
    nwin, ns, nj = Xtxy.shape[:3]
    X = torch.cat([Xpos, Xtxy.reshape(nwin, ns, nj, 6), Xvel, Xang], dim=-1)   # (:180-185)
    xm = torch.as_tensor(np.asarray(X_mean, dtype=np.float32)[1:], **f32)
    xs = torch.as_tensor(np.asarray(X_std, dtype=np.float32)[1:], **f32)
    X = ((X[:, :, 1:] - xm) / xs).contiguous()                         # (:186)
    return {"X": X, "Yrvel": Yrvel.contiguous(), "Yrang": Yrang.contiguous(), "Ypos": Ypos2, "Yrot": Yrot2,
            "Yvel": Yvel2, "contacts": torch.as_tensor(windows["contacts"], dtype=torch.uint8, device=device)}
