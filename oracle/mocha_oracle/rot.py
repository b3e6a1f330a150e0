"""NumPy restatement of the reference's quaternion / kinematics helpers (motion/quat.py).

Quaternions are [w, x, y, z] on the last axis. Functions keep the dtype of their inputs the way the
reference does (float32 arrays stay float32; Python-float constants do not promote)."""
from __future__ import annotations

import numpy as np


def _c(a, i):
    return a[..., i:i + 1]


def cross3(a, b):
    """_fast_cross (motion/quat.py:3-7)."""
    return np.concatenate([
        _c(a, 1) * _c(b, 2) - _c(a, 2) * _c(b, 1),
        _c(a, 2) * _c(b, 0) - _c(a, 0) * _c(b, 2),
        _c(a, 0) * _c(b, 1) - _c(a, 1) * _c(b, 0)], axis=-1)


def norm(x):
    """length (motion/quat.py:12-13)."""
    return np.sqrt((x * x).sum(axis=-1))


def normalize(x, eps=1e-8):
    """motion/quat.py:15-16."""
    return x / (norm(x)[..., None] + eps)


def q_abs(q):
    """motion/quat.py:18-19."""
    return np.where(_c(q, 0) > 0.0, q, -q)


def q_inv(q):
    """motion/quat.py:109-110 (the sign vector is float32 there)."""
    return np.asarray([1, -1, -1, -1], dtype=np.float32) * q


def q_mul(a, b):
    """motion/quat.py:112-120."""
    a0, a1, a2, a3 = _c(a, 0), _c(a, 1), _c(a, 2), _c(a, 3)
    b0, b1, b2, b3 = _c(b, 0), _c(b, 1), _c(b, 2), _c(b, 3)
    return np.concatenate([
        b0 * a0 - b1 * a1 - b2 * a2 - b3 * a3,
        b0 * a1 + b1 * a0 - b2 * a3 + b3 * a2,
        b0 * a2 + b1 * a3 + b2 * a0 - b3 * a1,
        b0 * a3 - b1 * a2 + b2 * a1 + b3 * a0], axis=-1)


def q_inv_mul(a, b):
    return q_mul(q_inv(a), b)


def q_mul_inv(a, b):
    return q_mul(a, q_inv(b))


def q_rotate(q, v):
    """mul_vec (motion/quat.py:128-130)."""
    t = 2.0 * cross3(q[..., 1:], v)
    return v + _c(q, 0) * t + cross3(q[..., 1:], t)


def q_inv_rotate(q, v):
    return q_rotate(q_inv(q), v)


def q_from_angle_axis(angle, axis):
    """motion/quat.py:21-25."""
    half = angle / 2.0
    return np.concatenate([np.cos(half)[..., None], np.sin(half)[..., None] * axis], axis=-1)


def q_exp(x, eps=1e-5):
    """motion/quat.py:154-158."""
    h = np.sqrt(np.square(x).sum(axis=-1))[..., None]
    c = np.where(h < eps, np.ones_like(h), np.cos(h))
    s = np.where(h < eps, np.ones_like(h), np.sinc(h / np.pi))
    return np.concatenate([c, s * x], axis=-1)


def q_log(q, eps=1e-5):
    """motion/quat.py:149-152."""
    ln = np.sqrt(np.square(q[..., 1:]).sum(axis=-1))[..., None]
    half = np.where(ln < eps, np.ones_like(ln), np.arctan2(ln, _c(q, 0)) / ln)
    return half * q[..., 1:]


def q_to_scaled_angle_axis(q, eps=1e-5):
    return 2.0 * q_log(q, eps)


def q_from_scaled_angle_axis(x, eps=1e-5):
    return q_exp(x / 2.0, eps)


def q_between(a, b):
    """motion/quat.py:143-147."""
    return np.concatenate([
        np.sqrt((a * a).sum(axis=-1) * (b * b).sum(axis=-1))[..., None] + (a * b).sum(axis=-1)[..., None],
        cross3(a, b)], axis=-1)


def q_to_xy(q):
    """to_xform_xy (motion/quat.py:42-55): first two columns of the rotation matrix, [...,3,2]."""
    w, x, y, z = _c(q, 0), _c(q, 1), _c(q, 2), _c(q, 3)
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, wx = x * x2, y * y2, w * x2
    xy, yz, wy = x * y2, y * z2, w * y2
    xz, zz, wz = x * z2, z * z2, w * z2
    rows = [np.concatenate([1.0 - (yy + zz), xy - wz], axis=-1),
            np.concatenate([xy + wz, 1.0 - (xx + zz)], axis=-1),
            np.concatenate([xz - wy, yz + wx], axis=-1)]
    return np.stack(rows, axis=-2)


def q_to_matrix(q):
    """to_xform (motion/quat.py:27-40)."""
    w, x, y, z = _c(q, 0), _c(q, 1), _c(q, 2), _c(q, 3)
    x2, y2, z2 = x + x, y + y, z + z
    xx, yy, wx = x * x2, y * y2, w * x2
    xy, yz, wy = x * y2, y * z2, w * y2
    xz, zz, wz = x * z2, z * z2, w * z2
    rows = [np.concatenate([1.0 - (yy + zz), xy - wz, xz + wy], axis=-1),
            np.concatenate([xy + wz, 1.0 - (xx + zz), yz - wx], axis=-1),
            np.concatenate([xz - wy, yz + wx, 1.0 - (xx + yy)], axis=-1)]
    return np.stack(rows, axis=-2)


def q_from_matrix(m):
    """from_xform (motion/quat.py:69-94): four-branch extraction then normalize."""
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]

    def pack(a, b, c, d):
        return np.stack([a, b, c, d], axis=-1)

    neg_a = pack(m[..., 2, 1] - m[..., 1, 2], 1.0 + m00 - m11 - m22, m[..., 1, 0] + m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0])
    neg_b = pack(m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] + m[..., 0, 1], 1.0 - m00 + m11 - m22, m[..., 2, 1] + m[..., 1, 2])
    pos_a = pack(m[..., 1, 0] - m[..., 0, 1], m[..., 0, 2] + m[..., 2, 0], m[..., 2, 1] + m[..., 1, 2], 1.0 - m00 - m11 + m22)
    pos_b = pack(1.0 + m00 + m11 + m22, m[..., 2, 1] - m[..., 1, 2], m[..., 0, 2] - m[..., 2, 0], m[..., 1, 0] - m[..., 0, 1])
    q = np.where((m22 < 0.0)[..., None],
                 np.where((m00 > m11)[..., None], neg_a, neg_b),
                 np.where((m00 < -m11)[..., None], pos_a, pos_b))
    return normalize(q)


def q_from_xy(xy):
    """from_xform_xy (motion/quat.py:96-107): [...,3,2] -> quaternion."""
    c0_in, c1_in = xy[..., 0], xy[..., 1]
    c2 = cross3(c0_in, c1_in)
    c2 = c2 / np.sqrt(np.square(c2).sum(axis=-1))[..., None]
    c1 = cross3(c2, c0_in)
    c1 = c1 / np.sqrt(np.square(c1).sum(axis=-1))[..., None]
    return q_from_matrix(np.stack([c0_in, c1, c2], axis=-1))


def fk(lrot, lpos, parents):
    """motion/quat.py:166-173."""
    gp, gr = [lpos[..., :1, :]], [lrot[..., :1, :]]
    for j in range(1, len(parents)):
        p = parents[j]
        gp.append(q_rotate(gr[p], lpos[..., j:j + 1, :]) + gp[p])
        gr.append(q_mul(gr[p], lrot[..., j:j + 1, :]))
    return np.concatenate(gr, axis=-2), np.concatenate(gp, axis=-2)


def fk_vel(lrot, lpos, lvel, lang, parents):
    """motion/quat.py:189-204."""
    gp, gr, gv, ga = [lpos[..., :1, :]], [lrot[..., :1, :]], [lvel[..., :1, :]], [lang[..., :1, :]]
    for j in range(1, len(parents)):
        p = parents[j]
        rp = q_rotate(gr[p], lpos[..., j:j + 1, :])
        gp.append(rp + gp[p])
        gr.append(q_mul(gr[p], lrot[..., j:j + 1, :]))
        gv.append(q_rotate(gr[p], lvel[..., j:j + 1, :]) + cross3(ga[p], rp) + gv[p])
        ga.append(q_rotate(gr[p], lang[..., j:j + 1, :]) + ga[p])
    return (np.concatenate(gr, axis=-2), np.concatenate(gp, axis=-2),
            np.concatenate(gv, axis=-2), np.concatenate(ga, axis=-2))


def ik(grot, gpos, parents):
    """motion/quat.py:175-187: global -> local."""
    par = np.asarray(parents[1:])
    lrot = np.concatenate([grot[..., :1, :], q_mul(q_inv(grot[..., par, :]), grot[..., 1:, :])], axis=-2)
    lpos = np.concatenate([gpos[..., :1, :],
                           q_rotate(q_inv(grot[..., par, :]), gpos[..., 1:, :] - gpos[..., par, :])], axis=-2)
    return lrot, lpos


def chain_to_root(parents, bone):
    out = []
    while bone != -1:
        out.append(int(bone))
        bone = parents[bone]
    return out[::-1]


def fk_vel_bone(pos, vel, rot, ang, parents, bone):
    """motion/quat.py:207-237, unrolled along the ancestor chain instead of recursing."""
    chain = chain_to_root(parents, bone)
    gp, gv, gr, ga = pos[chain[0]], vel[chain[0]], rot[chain[0]], ang[chain[0]]
    for j in chain[1:]:
        rp = q_rotate(gr, pos[j])
        nv = gv + q_rotate(gr, vel[j]) + cross3(ga, rp)
        na = q_rotate(gr, ang[j]) + ga
        gp, gv, ga = rp + gp, nv, na
        gr = q_mul(gr, rot[j])
    return gp, gv, gr, ga


def fk_chain(pos, rot, parents, bone):
    """Global transforms of every bone on the ancestor chain of `bone` — what repeated
    fk_partial calls (motion/quat.py:241-272) leave in the global arrays."""
    chain = chain_to_root(parents, bone)
    gpos, grot = {}, {}
    gpos[chain[0]], grot[chain[0]] = pos[chain[0]], rot[chain[0]]
    for a, j in zip(chain[:-1], chain[1:]):
        gpos[j] = q_rotate(grot[a], pos[j]) + gpos[a]
        grot[j] = q_mul(grot[a], rot[j])
    return gpos, grot


def ik_two_bone(bone_root, bone_mid, bone_end, target, fwd, root_gr, mid_gr, par_gr, max_length_buffer):
    """motion/quat.py:295-343 (the two incoming local rotations are overwritten there, so they are
    not parameters here)."""
    max_ext = norm(bone_root - bone_mid) + norm(bone_mid - bone_end) - max_length_buffer
    t = target
    if norm(target - bone_root) > max_ext:
        t = bone_root + max_ext * normalize(target - bone_root)
    axis_dwn = normalize(bone_end - bone_root)
    axis_rot = normalize(np.cross(axis_dwn, fwd))
    a, b, c = bone_root, bone_mid, bone_end
    lab, lcb, lat = norm(b - a), norm(b - c), norm(t - a)
    ac_ab_0 = np.arccos(np.clip(np.dot(normalize(c - a), normalize(b - a)), -1.0, 1.0))
    ba_bc_0 = np.arccos(np.clip(np.dot(normalize(a - b), normalize(c - b)), -1.0, 1.0))
    ac_ab_1 = np.arccos(np.clip((lab * lab + lat * lat - lcb * lcb) / (2.0 * lab * lat), -1.0, 1.0))
    ba_bc_1 = np.arccos(np.clip((lab * lab + lcb * lcb - lat * lat) / (2.0 * lab * lcb), -1.0, 1.0))
    r0 = q_from_angle_axis(ac_ab_1 - ac_ab_0, axis_rot)
    r1 = q_from_angle_axis(ba_bc_1 - ba_bc_0, axis_rot)
    c_a, t_a = normalize(bone_end - bone_root), normalize(t - bone_root)
    r2 = q_from_angle_axis(np.arccos(np.clip(np.dot(c_a, t_a), -1.0, 1.0)), normalize(np.cross(c_a, t_a)))
    root_lr = q_inv_mul(par_gr, q_mul(r2, q_mul(r0, root_gr)))
    mid_lr = q_inv_mul(root_gr, q_mul(r1, mid_gr))
    return root_lr, mid_lr


def q_from_euler(e, order="zyx"):
    """motion/quat.py:57-67."""
    axis = {"x": np.asarray([1, 0, 0], dtype=np.float32), "y": np.asarray([0, 1, 0], dtype=np.float32),
            "z": np.asarray([0, 0, 1], dtype=np.float32)}
    q0 = q_from_angle_axis(e[..., 0], axis[order[0]])
    q1 = q_from_angle_axis(e[..., 1], axis[order[1]])
    q2 = q_from_angle_axis(e[..., 2], axis[order[2]])
    return q_mul(q0, q_mul(q1, q2))


def q_unroll(x):
    """motion/quat.py:135-141: flip signs so consecutive frames stay on one hemisphere."""
    y = x.copy()
    for i in range(1, len(x)):
        flip = (y[i] * y[i - 1]).sum(axis=-1) < (-y[i] * y[i - 1]).sum(axis=-1)
        y[i][flip] = -y[i][flip]
    return y


def q_to_euler(q, order="xyz"):
    """motion/quat.py:346-368."""
    q0, q1, q2, q3 = _c(q, 0), _c(q, 1), _c(q, 2), _c(q, 3)
    if order == "xyz":
        return np.concatenate([
            np.arctan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 * q1 + q2 * q2)),
            np.arcsin((2 * (q0 * q2 - q3 * q1)).clip(-1, 1)),
            np.arctan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 * q2 + q3 * q3))], axis=-1)
    if order == "yzx":
        return np.concatenate([
            np.arctan2(2 * (q1 * q0 - q2 * q3), -q1 * q1 + q2 * q2 - q3 * q3 + q0 * q0),
            np.arctan2(2 * (q2 * q0 - q1 * q3), q1 * q1 - q2 * q2 - q3 * q3 + q0 * q0),
            np.arcsin((2 * (q1 * q2 + q3 * q0)).clip(-1, 1))], axis=-1)
    raise NotImplementedError(order)
