"""Drop-in for the reference's `motion/quat.py` (NumPy in, NumPy out, same positional signatures,
quaternions [w,x,y,z]) computing on the GPU through libmocha_b200.

Every function computes in the precision NumPy would have used for the same call: float32 operands stay
float32, anything else (float64 arrays, integer constants such as `np.array([1, 0, 0, 0])`) is float64 -
the reference driver's per-frame loop runs on float64 state (test_fullframework.py:321-353, :476-509) and a
drop-in must not narrow it. Element-wise algebra goes through `mocha_quat_op` (one templated kernel family,
float32 / float64), the batch operators of the hot path through
`mocha_fk*`, `mocha_ik*`, `mocha_fk_chain`, `mocha_ik_two_bone`, `mocha_xy_to_quat`, `mocha_quat_to_xy`."""
from __future__ import annotations

import builtins

import numpy as np
import torch

from . import kinematics as kin


# ---------------------------------------------------------------------------------------------------
# marshalling
# ---------------------------------------------------------------------------------------------------
def _wdt(*arrays):
    """NumPy's result dtype for these operands, restricted to the two precisions the kernels have."""
    dt = np.result_type(*[np.asarray(a).dtype for a in arrays])
    return np.float32 if dt == np.float32 else np.float64


def _cu(a, dt):
    return torch.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=dt)).cuda()


def _unary(name, x, wa, wo, param=0.0, dt=None):
    x = np.asarray(x)
    dt = dt or _wdt(x)
    lead = x.shape[:-1]
    if x.shape[-1] != wa:
        raise ValueError(f"quat.{name}: last axis must have {wa} elements, got {x.shape}")
    out = kin.quat_op(name, _cu(x.reshape(-1, wa), dt), None, param)
    return out.cpu().numpy().reshape(lead + (wo,))


def _binary(name, a, b, wa, wb, wo, dt=None):
    a, b = np.asarray(a), np.asarray(b)
    dt = dt or _wdt(a, b)
    if a.shape[-1] != wa or b.shape[-1] != wb:
        raise ValueError(f"quat.{name}: operand shapes {a.shape} / {b.shape} need last axes {wa} / {wb}")
    lead = np.broadcast_shapes(a.shape[:-1], b.shape[:-1])
    a = np.broadcast_to(a, lead + (wa,)).reshape(-1, wa)
    b = np.broadcast_to(b, lead + (wb,)).reshape(-1, wb)
    out = kin.quat_op(name, _cu(a, dt), _cu(b, dt))
    return out.cpu().numpy().reshape(lead + (wo,))


# ---------------------------------------------------------------------------------------------------
# element-wise algebra (motion/quat.py:3-164, :346-368)
# ---------------------------------------------------------------------------------------------------
def _fast_cross(a, b):
    return _binary("cross", a, b, 3, 3, 3)


def eye(shape, dtype=np.float32):
    return np.ones(list(shape) + [4], dtype=dtype) * np.asarray([1, 0, 0, 0], dtype=dtype)


def length(x):
    x = np.asarray(x)
    if x.shape[-1] not in (3, 4):
        raise ValueError("quat.length: last axis must have 3 or 4 elements")
    return _unary("length%d" % x.shape[-1], x, x.shape[-1], 1)[..., 0]


def normalize(x, eps=1e-8):
    x = np.asarray(x)
    if x.shape[-1] not in (3, 4):
        raise ValueError("quat.normalize: last axis must have 3 or 4 elements")
    return _unary("normalize%d" % x.shape[-1], x, x.shape[-1], x.shape[-1], eps)


def abs(x):
    return _unary("abs", x, 4, 4)


def inv(q):
    return _unary("inv", q, 4, 4)


def mul(x, y):
    return _binary("mul", x, y, 4, 4, 4)


def inv_mul(x, y):
    return _binary("inv_mul", x, y, 4, 4, 4)


def mul_inv(x, y):
    return _binary("mul_inv", x, y, 4, 4, 4)


def mul_vec(q, x):
    return _binary("mul_vec", q, x, 4, 3, 3)


def inv_mul_vec(q, x):
    return _binary("inv_mul_vec", q, x, 4, 3, 3)


def from_angle_axis(angle, axis):
    angle, axis = np.asarray(angle), np.asarray(axis)
    return _binary("from_angle_axis", angle[..., np.newaxis], axis, 1, 3, 4)


def exp(x, eps=1e-5):
    return _unary("exp", x, 3, 4, eps)


def log(x, eps=1e-5):
    return _unary("log", x, 4, 3, eps)


def to_scaled_angle_axis(x, eps=1e-5):
    return 2.0 * log(x, eps)


def from_scaled_angle_axis(x, eps=1e-5):
    return exp(np.asarray(x) / 2.0, eps)


def between(x, y):
    return _binary("between", x, y, 3, 3, 4)


def to_xform(x):
    x = np.asarray(x)
    return _unary("to_xform", x, 4, 9).reshape(x.shape[:-1] + (3, 3))


def from_xform(ts):
    ts = np.asarray(ts)
    if ts.shape[-2:] != (3, 3):
        raise ValueError("quat.from_xform expects [..., 3, 3] rotation matrices")
    return _unary("from_xform", ts.reshape(ts.shape[:-2] + (9,)), 9, 4)


def to_xform_xy(x):
    x = np.asarray(x)
    if _wdt(x) == np.float32:       # batch path of the driver's set-up (:156, :161): dedicated float4 kernel
        return kin.quat_to_xy(_cu(x, np.float32)).cpu().numpy()
    return _unary("to_xform_xy", x, 4, 6).reshape(x.shape[:-1] + (3, 2))


def from_xform_xy(x):
    x = np.asarray(x)
    if x.shape[-2:] != (3, 2):
        raise ValueError("quat.from_xform_xy expects [..., 3, 2]")
    if _wdt(x) == np.float32:       # per-frame path (:461): dedicated kernel
        return kin.xy_to_quat(_cu(x, np.float32)).cpu().numpy()
    return _unary("from_xform_xy", x.reshape(x.shape[:-2] + (6,)), 6, 4)


def from_euler(e, order="zyx"):
    axis = {"x": np.asarray([1, 0, 0], dtype=np.float32), "y": np.asarray([0, 1, 0], dtype=np.float32),
            "z": np.asarray([0, 0, 1], dtype=np.float32)}
    e = np.asarray(e)
    q0 = from_angle_axis(e[..., 0], axis[order[0]])
    q1 = from_angle_axis(e[..., 1], axis[order[1]])
    q2 = from_angle_axis(e[..., 2], axis[order[2]])
    return mul(q0, mul(q1, q2))


def unroll(x):
    """Keep consecutive frames on one hemisphere (:135-141): frame i is negated when it opposes the
    already-unrolled frame i-1, i.e. sign_i = sign_{i-1} * sign(<x_i, x_{i-1}>) - a cumulative product."""
    x = np.asarray(x)
    t = _cu(x, _wdt(x))
    d = (t[1:] * t[:-1]).sum(-1)
    sgn = torch.where(d < 0, -torch.ones_like(d), torch.ones_like(d))
    s = torch.cat([torch.ones_like(sgn[:1]), torch.cumprod(sgn, dim=0)], dim=0)
    return (t * s[..., None]).cpu().numpy()


def to_euler(x, order="xyz"):
    if order not in ("xyz", "yzx"):
        raise NotImplementedError("Cannot convert from ordering %s" % order)
    return _unary("to_euler_" + order, x, 4, 3)


# ---------------------------------------------------------------------------------------------------
# kinematics (motion/quat.py:166-343)
# ---------------------------------------------------------------------------------------------------
def fk(lrot, lpos, parents):
    dt = _wdt(lrot, lpos)
    par = kin.parents_tensor(parents, "cuda")
    gr, gp = kin.fk(_cu(lrot, dt), _cu(lpos, dt), par)
    return gr.cpu().numpy(), gp.cpu().numpy()


def ik(grot, gpos, parents):
    dt = _wdt(grot, gpos)
    par = kin.parents_tensor(parents, "cuda")
    lr, lp = kin.ik(_cu(grot, dt), _cu(gpos, dt), par)
    return lr.cpu().numpy(), lp.cpu().numpy()


def fk_vel(lrot, lpos, lvel, lang, parents):
    dt = _wdt(lrot, lpos, lvel, lang)
    par = kin.parents_tensor(parents, "cuda")
    out = kin.fk_vel(_cu(lrot, dt), _cu(lpos, dt), _cu(lvel, dt), _cu(lang, dt), par)
    return tuple(t.cpu().numpy() for t in out)


def fk_vel_bone(bone_positions, bone_velocities, bone_rotations, bone_angular_velocities, bone_parents, bone):
    gr, gp, gv, ga = fk_vel(np.asarray(bone_rotations)[None], np.asarray(bone_positions)[None],
                            np.asarray(bone_velocities)[None], np.asarray(bone_angular_velocities)[None], bone_parents)
    return gp[0, bone], gv[0, bone], gr[0, bone], ga[0, bone]


def fk_partial(global_bone_positions, global_bone_rotations, global_bone_computed, local_bone_positions,
               local_bone_rotations, bone_parents, bone):
    """(:241-272) In place, like the reference: `bone` is always recomputed; the walk towards the root stops at
    the first ancestor already marked computed and continues from that ancestor's STORED global transform."""
    chain = [int(bone)]
    while bone_parents[chain[-1]] != -1 and not global_bone_computed[bone_parents[chain[-1]]]:
        chain.append(int(bone_parents[chain[-1]]))
    chain.reverse()                                   # top of the chain first
    top_parent = int(bone_parents[chain[0]])
    dt = _wdt(global_bone_positions, global_bone_rotations, local_bone_positions, local_bone_rotations)
    lpos = _cu(np.asarray(local_bone_positions)[chain][None], dt)
    lrot = _cu(np.asarray(local_bone_rotations)[chain][None], dt)
    if top_parent == -1:
        gpos, grot = kin.fk_chain(lpos, lrot)
    else:
        gpos, grot = kin.fk_chain(lpos, lrot, _cu(np.asarray(global_bone_positions)[top_parent][None], dt),
                                  _cu(np.asarray(global_bone_rotations)[top_parent][None], dt))
    gpos, grot = gpos.cpu().numpy()[0], grot.cpu().numpy()[0]
    for c, j in enumerate(chain):
        global_bone_positions[j] = gpos[c]
        global_bone_rotations[j] = grot[c]
        global_bone_computed[j] = True
    return global_bone_positions, global_bone_rotations, global_bone_computed


def ik_look_at(bone_rotation, global_parent_rotation, global_rotation, global_position, child_position,
               target_position, eps=1e-5):
    """(:276-290)"""
    curr_dir = normalize(np.asarray(child_position) - np.asarray(global_position))
    targ_dir = normalize(np.asarray(target_position) - np.asarray(global_position))
    if builtins.abs(1.0 - float(np.dot(curr_dir, targ_dir))) > eps:
        bone_rotation = inv_mul(global_parent_rotation, mul(between(curr_dir, targ_dir), global_rotation))
    return bone_rotation


def ik_two_bone(bone_root_lr, bone_mid_lr, bone_root, bone_mid, bone_end, target, fwd, bone_root_gr, bone_mid_gr,
                bone_par_gr, max_length_buffer):
    """(:295-343) float64 kernel; the reference's inputs here are float64 state."""
    args = [_cu(np.asarray(a, dtype=np.float64)[None], np.float64) for a in
            (bone_root, bone_mid, bone_end, target, fwd, bone_root_gr, bone_mid_gr, bone_par_gr)]
    a, b = kin.ik_two_bone(*args, float(max_length_buffer))
    return a[0].cpu().numpy(), b[0].cpu().numpy()
