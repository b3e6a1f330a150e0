"""Drop-in for the reference's `motion/quat.py` (NumPy in, NumPy out, same positional signatures,
quaternions [w,x,y,z]) computing on the GPU through libmocha_b200 / CUDA tensors.

Batch operators of the hot path run in the library's kernels:
from_xform_xy, to_xform_xy, fk, fk_vel, ik, ik_two_bone, fk_partial / fk_vel_bone (via the FK
kernels). Small element-wise helpers run as CUDA tensor expressions (tq.py). Kernels compute in
float32 (float64 for ik_two_bone); results come back in the input's dtype."""
from __future__ import annotations

import numpy as np
import torch

from . import kinematics as kin
from . import tq


def _cu(a, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a)), dtype=dtype).cuda()


def _back(t, like):
    out = t.cpu().numpy()
    like = np.asarray(like)
    return out.astype(like.dtype if like.dtype.kind == "f" else np.float64)


def _bin(fn, a, b):
    ta, tb = _cu(a), _cu(b)
    shape = np.broadcast_shapes(ta.shape[:-1], tb.shape[:-1])
    ta, tb = ta.expand(shape + ta.shape[-1:]), tb.expand(shape + tb.shape[-1:])
    return _back(fn(ta, tb), np.zeros(1, dtype=np.result_type(np.asarray(a).dtype, np.asarray(b).dtype, np.float32)))


def eye(shape, dtype=np.float32):
    return np.ones(list(shape) + [4], dtype=dtype) * np.asarray([1, 0, 0, 0], dtype=dtype)


def length(x):
    return _back(torch.sqrt((_cu(x) ** 2).sum(-1)), x)


def normalize(x, eps=1e-8):
    t = _cu(x)
    return _back(t / (torch.sqrt((t * t).sum(-1, keepdim=True)) + eps), x)


def abs(x):
    t = _cu(x)
    return _back(torch.where(t[..., 0:1] > 0.0, t, -t), x)


def inv(q):
    return _back(tq.inv(_cu(q)), q)


def mul(x, y):
    return _bin(tq.mul, x, y)


def inv_mul(x, y):
    return _bin(lambda a, b: tq.mul(tq.inv(a), b), x, y)


def mul_inv(x, y):
    return _bin(lambda a, b: tq.mul(a, tq.inv(b)), x, y)


def mul_vec(q, x):
    return _bin(tq.mul_vec, q, x)


def inv_mul_vec(q, x):
    return _bin(lambda a, b: tq.mul_vec(tq.inv(a), b), q, x)


def from_angle_axis(angle, axis):
    a = _cu(angle)
    ax = _cu(axis).expand(a.shape + (3,))
    return _back(torch.cat([torch.cos(a / 2.0)[..., None], torch.sin(a / 2.0)[..., None] * ax], dim=-1), angle)


def exp(x, eps=1e-5):
    t = _cu(x)
    h = torch.sqrt((t * t).sum(-1, keepdim=True))
    c = torch.where(h < eps, torch.ones_like(h), torch.cos(h))
    s = torch.where(h < eps, torch.ones_like(h), torch.sin(h) / torch.where(h < eps, torch.ones_like(h), h))
    return _back(torch.cat([c, s * t], dim=-1), x)


def log(x, eps=1e-5):
    t = _cu(x)
    ln = torch.sqrt((t[..., 1:] ** 2).sum(-1, keepdim=True))
    safe = torch.where(ln < eps, torch.ones_like(ln), ln)
    half = torch.where(ln < eps, torch.ones_like(ln), torch.atan2(ln, t[..., 0:1]) / safe)
    return _back(half * t[..., 1:], x)


def to_scaled_angle_axis(x, eps=1e-5):
    return 2.0 * log(x, eps)


def from_scaled_angle_axis(x, eps=1e-5):
    return exp(np.asarray(x) / 2.0, eps)


def between(x, y):
    a, b = _cu(x), _cu(y)
    shape = np.broadcast_shapes(a.shape, b.shape)
    a, b = a.expand(shape), b.expand(shape)
    w = torch.sqrt((a * a).sum(-1) * (b * b).sum(-1))[..., None] + (a * b).sum(-1)[..., None]
    return _back(torch.cat([w, tq.cross(a, b)], dim=-1), x)


def to_xform_xy(x):
    return _back(kin.quat_to_xy(_cu(x)), x)


def from_xform_xy(x):
    return _back(kin.xy_to_quat(_cu(x)), x)


def fk(lrot, lpos, parents):
    par = kin.parents_tensor(parents, "cuda")
    gr, gp = kin.fk(_cu(lrot), _cu(lpos), par)
    return _back(gr, lrot), _back(gp, lpos)


def ik(grot, gpos, parents):
    par = kin.parents_tensor(parents, "cuda")
    lr, lp = kin.ik(_cu(grot), _cu(gpos), par)
    return _back(lr, grot), _back(lp, gpos)


def fk_vel(lrot, lpos, lvel, lang, parents):
    par = kin.parents_tensor(parents, "cuda")
    out = kin.fk_vel(_cu(lrot), _cu(lpos), _cu(lvel), _cu(lang), par)
    return tuple(_back(t, lrot) for t in out)


def fk_vel_bone(bone_positions, bone_velocities, bone_rotations, bone_angular_velocities, bone_parents, bone):
    gr, gp, gv, ga = fk_vel(np.asarray(bone_rotations)[None], np.asarray(bone_positions)[None],
                            np.asarray(bone_velocities)[None], np.asarray(bone_angular_velocities)[None], bone_parents)
    return gp[0, bone], gv[0, bone], gr[0, bone], ga[0, bone]


def fk_partial(global_bone_positions, global_bone_rotations, global_bone_computed, local_bone_positions,
               local_bone_rotations, bone_parents, bone):
    """Fills (in place, like the reference) the global transforms of `bone` and of every ancestor not
    yet marked computed."""
    gr, gp = fk(np.asarray(local_bone_rotations)[None], np.asarray(local_bone_positions)[None], bone_parents)
    j = int(bone)
    chain = []
    while j != -1:
        chain.append(j)
        j = int(bone_parents[j])
    for idx, j in enumerate(chain):
        if idx == 0 or not global_bone_computed[j]:
            global_bone_positions[j] = gp[0, j]
            global_bone_rotations[j] = gr[0, j]
            global_bone_computed[j] = True
    return global_bone_positions, global_bone_rotations, global_bone_computed


def ik_two_bone(bone_root_lr, bone_mid_lr, bone_root, bone_mid, bone_end, target, fwd, bone_root_gr, bone_mid_gr,
                bone_par_gr, max_length_buffer):
    f64 = torch.float64
    args = [_cu(np.asarray(a, dtype=np.float64)[None], f64) for a in
            (bone_root, bone_mid, bone_end, target, fwd, bone_root_gr, bone_mid_gr, bone_par_gr)]
    a, b = kin.ik_two_bone(*args, float(max_length_buffer))
    return a[0].cpu().numpy(), b[0].cpu().numpy()


def from_euler(e, order="zyx"):
    axis = {"x": np.asarray([1, 0, 0], dtype=np.float32), "y": np.asarray([0, 1, 0], dtype=np.float32),
            "z": np.asarray([0, 0, 1], dtype=np.float32)}
    e = np.asarray(e)
    q0 = from_angle_axis(e[..., 0], axis[order[0]])
    q1 = from_angle_axis(e[..., 1], axis[order[1]])
    q2 = from_angle_axis(e[..., 2], axis[order[2]])
    return mul(q0, mul(q1, q2))


def unroll(x):
    t = _cu(x)
    y = t.clone()
    for i in range(1, y.shape[0]):
        flip = (y[i] * y[i - 1]).sum(-1) < 0.0
        y[i][flip] = -y[i][flip]
    return _back(y, x)


def to_euler(x, order="xyz"):
    t = _cu(x)
    q0, q1, q2, q3 = t[..., 0:1], t[..., 1:2], t[..., 2:3], t[..., 3:4]
    if order == "xyz":
        out = torch.cat([torch.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)),
                         torch.asin((2 * (q0 * q2 - q3 * q1)).clamp(-1, 1)),
                         torch.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))], dim=-1)
    elif order == "yzx":
        out = torch.cat([torch.atan2(2 * (q1 * q0 - q2 * q3), -q1 * q1 + q2 * q2 - q3 * q3 + q0 * q0),
                         torch.atan2(2 * (q2 * q0 - q1 * q3), q1 * q1 - q2 * q2 - q3 * q3 + q0 * q0),
                         torch.asin((2 * (q1 * q2 + q3 * q0)).clamp(-1, 1))], dim=-1)
    else:
        raise NotImplementedError("Cannot convert from ordering %s" % order)
    return _back(out, x)
