"""Training-side twins of the hot-path kernels, forward only.

The reference's training code re-uses the inference math: `convert_YtilToX` (trainer.py:337-374) turns a decoded pose
window back into a character-space input window (from_xform_xy -> fk_vel -> per-frame re-rooting -> to_xform_xy),
`recon_criterion` (trainer.py:249-335) evaluates its FK losses, and train_CVAE.py:207-211 runs a BallTree query per
training batch over the rows of one action label. Here they run on the CUDA kernels of the inference path
(`mocha_xy_to_quat`, `mocha_fk_vel`, `mocha_window_features`, `mocha_quat_op`, the matcher). They are FORWARD values
only - evaluation metrics and data preparation; no autograd is attached (training itself is out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, kinematics as kin
from .balltree import BallTree


def _split(Y):
    lead = Y.shape[:-1]
    return Y[..., :3], Y[..., 3:9].reshape(*lead, 3, 2), Y[..., 9:12], Y[..., 12:15]


def _character_space(pos, txy, vel, ang, parents):
    """local pose streams [B,T,J,*] -> (Gpos, X [B,T,J,15] per-frame character space, Xrot [B,T,J,4])."""
    B, T, J = pos.shape[:3]
    par = kin.parents_tensor(parents, pos.device)
    rot = kin.xy_to_quat(txy.contiguous())
    grot, gpos, gvel, gang = kin.fk_vel(rot, pos.contiguous(), vel.contiguous(), ang.contiguous(), par)
    X = torch.empty((B, T, J, 15), dtype=torch.float32, device=pos.device)
    xrot = torch.empty_like(grot)
    _lib.check(_lib.load().mocha_window_features(_lib.ptr(grot), _lib.ptr(gpos), _lib.ptr(gvel), _lib.ptr(gang), B, T, J, None, None,
                                                 _lib.ptr(X), _lib.ptr(xrot), None, _lib.stream_ptr()), "mocha_window_features")
    return X, xrot


def convert_YtilToX(Ytil: torch.Tensor, Ygrd: torch.Tensor, parents) -> torch.Tensor:
    """trainer.py:337-374. Ytil [B,T,V,15] decoded window, Ygrd [B,T,1,15] root stream -> X [B,T,V+1,15]."""
    for t in (Ytil, Ygrd):
        if not t.is_cuda:
            raise _lib.MochaError("convert_YtilToX needs CUDA tensors (no CPU fallback)")
    pos, txy, vel, ang = (torch.cat([g, y], dim=2) for g, y in zip(_split(Ygrd.float()), _split(Ytil.float())))
    X, _ = _character_space(pos, txy, vel, ang, parents)
    return X


def recon_criterion(Ytil: torch.Tensor, Ygt: torch.Tensor, parents) -> torch.Tensor:
    """Forward value of trainer.py:249-335 (no gradient). Ytil [B,T,V,15], Ygt [B,T,V+1,15].
    The FK runs on quaternions (mocha_xy_to_quat orthonormalises the 6-D columns); the reference's matrix FK keeps the
    first column un-normalised (txform.py:22-33), so the two agree when the 6-D columns are orthonormal - exactly for
    ground-truth windows, up to the network's rotation noise for decoded ones."""
    dt = 1.0 / 60.0
    g_pos, g_txy, g_vel, g_ang = _split(Ygt.float())
    t_pos, t_txy, t_vel, t_ang = _split(Ytil.float())
    t_pos = torch.cat([g_pos[:, :, 0:1], t_pos], dim=2)
    t_txy = torch.cat([g_txy[:, :, 0:1], t_txy], dim=2)
    t_vel = torch.cat([g_vel[:, :, 0:1], t_vel], dim=2)
    t_ang = torch.cat([g_ang[:, :, 0:1], t_ang], dim=2)
    Xg, rg = _character_space(g_pos, g_txy, g_vel, g_ang, parents)
    Xt, rt = _character_space(t_pos, t_txy, t_vel, t_ang, parents)
    # character-space rotation matrices (txform.inv_mul of the FK'd transforms == to_xform of the re-rooted quaternion)
    Qg = kin.quat_op("to_xform", rg.reshape(-1, 4).contiguous()).reshape(*rg.shape[:-1], 3, 3)
    Qt = kin.quat_op("to_xform", rt.reshape(-1, 4).contiguous()).reshape(*rt.shape[:-1], 3, 3)
    m = lambda w, a, b: torch.mean(w * torch.abs(a - b))
    d = lambda x: (x[:, 1:] - x[:, :-1]) / dt
    loss = (m(75.0, g_pos, t_pos) + m(10.0, g_txy, t_txy) + m(10.0, g_vel, t_vel) + m(1.25, g_ang, t_ang)
            + m(15.0, Xg[..., :3], Xt[..., :3]) + m(5.0, Qg, Qt) + m(2.0, Xg[..., 9:12], Xt[..., 9:12])
            + m(0.75, Xg[..., 12:15], Xt[..., 12:15])
            + m(10.0, d(g_pos), d(t_pos)) + m(1.75, d(g_txy), d(t_txy)) + m(2.0, d(Xg[..., :3]), d(Xt[..., :3]))
            + m(0.75, d(Qg), d(Qt)))
    return loss


class ActionMatcher:
    """train_CVAE.py:196-211: one exact nearest-neighbour index per action label over the first context token rows
    `(cnt[:,0] - cnt_mean)/cnt_std` of that label's character windows; `nearest(label, queries)` returns indices into
    the label's subset, like `tree.query(..., k=1, return_distance=False)[:,0]`."""

    def __init__(self, rows_nm, action_labels, device="cuda"):
        rows_nm = torch.as_tensor(rows_nm, dtype=torch.float32)
        labels = np.asarray(action_labels)
        self.members, self.trees = {}, {}
        for lab in np.unique(labels):
            idx = np.where(labels == lab)[0]
            self.members[int(lab)] = idx
            self.trees[int(lab)] = BallTree(rows_nm[idx].reshape(len(idx), -1), device=device)

    def nearest(self, label: int, queries) -> np.ndarray:
        if int(label) not in self.trees:
            return np.empty((0,), dtype=np.int64)
        q = torch.as_tensor(np.asarray(queries), dtype=torch.float32)
        return self.trees[int(label)].query(q.reshape(q.shape[0], -1), k=1, return_distance=False)[:, 0]
