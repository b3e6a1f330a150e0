"""NumPy restatement of the reference's inertialization helpers (motion/Inertialization.py)."""
from __future__ import annotations

import numpy as np

from . import rot


def fast_negexp(x):
    """Inertialization.py:10-11."""
    return 1.0 / (1.0 + x + 0.48 * x * x + 0.235 * x * x * x)


def halflife_to_damping(halflife, eps=1e-5):
    """Inertialization.py:13-14."""
    return (4.0 * np.log(2.0)) / (halflife + eps)


def decay_pos(x, v, halflife, dt):
    """decay_spring_damper_exact vec3 branch / _pos (Inertialization.py:18-26, :50-54)."""
    y = halflife_to_damping(halflife) / 2.0
    j1 = v + x * y
    e = fast_negexp(y * dt)
    return e * (x + j1 * dt), e * (v - j1 * y * dt)


def decay_rot(x, v, halflife, dt):
    """decay_spring_damper_exact_rot (Inertialization.py:28-37)."""
    y = halflife_to_damping(halflife) / 2.0
    j0 = rot.q_to_scaled_angle_axis(x)
    j1 = v + j0 * y
    e = fast_negexp(y * dt)
    return rot.q_from_scaled_angle_axis(e * (j0 + j1 * dt)), e * (v - j1 * y * dt)


def contact_update(state, lock, position, velocity, point, target, off_pos, off_vel, in_pos, in_state,
                   unlock_radius, foot_height, halflife, dt, eps=1e-8):
    """Inertialization.py:300-377. Returns the 8 updated state entries in the reference's order."""
    in_vel = (in_pos - target) / (dt + eps)
    target = in_pos
    off_pos, off_vel = decay_pos(off_pos, off_vel, halflife, dt)          # inertialize_update :110-127
    fed_x = point if lock else in_pos
    fed_v = np.zeros(3) if lock else in_vel
    position, velocity = fed_x + off_pos, fed_v + off_vel
    unlock = lock and (rot.norm(point - in_pos) > unlock_radius)
    if (not state) and in_state:
        lock = True
        point = position.copy()
        point[1] = foot_height
        off_pos, off_vel = (in_pos + off_pos) - point, (in_vel + off_vel) - np.zeros(3)   # transition :93-108
    elif (lock and state and not in_state) or unlock:
        lock = False
        off_pos, off_vel = (point + off_pos) - in_pos, (np.zeros(3) + off_vel) - in_vel
    state = in_state
    return state, lock, position, velocity, point, target, off_pos, off_vel


def pose_transition(off_pos, off_vel, off_rot, off_ang, root_pos, root_vel, root_rot, root_ang,
                    src_pos, src_vel, src_rot, src_ang, dst_pos, dst_vel, dst_rot, dst_ang):
    """Inertialization.py:136-209 (arrays are copied, not mutated)."""
    off_pos, off_vel, off_rot, off_ang = off_pos.copy(), off_vel.copy(), off_rot.copy(), off_ang.copy()
    tr_dst_pos, tr_dst_rot = root_pos, root_rot
    tr_src_pos, tr_src_rot = dst_pos[0], dst_rot[0]
    ws_vel = rot.q_rotate(tr_dst_rot, rot.q_rotate(tr_src_rot, dst_vel[0]))
    ws_ang = rot.q_rotate(tr_dst_rot, rot.q_rotate(tr_src_rot, dst_ang[0]))
    off_pos[0], off_vel[0] = (root_pos + off_pos[0]) - root_pos, (root_vel + off_vel[0]) - ws_vel
    off_rot[0] = rot.q_abs(rot.q_mul(rot.q_mul(off_rot[0], root_rot), rot.q_inv(root_rot)))
    off_ang[0] = (off_ang[0] + root_ang) - ws_ang
    for i in range(1, len(off_pos)):
        off_pos[i], off_vel[i] = (src_pos[i] + off_pos[i]) - dst_pos[i], (src_vel[i] + off_vel[i]) - dst_vel[i]
        off_rot[i] = rot.q_abs(rot.q_mul(rot.q_mul(off_rot[i], src_rot[i]), rot.q_inv(dst_rot[i])))
        off_ang[i] = (off_ang[i] + src_ang[i]) - dst_ang[i]
    return off_pos, off_vel, off_rot, off_ang, tr_src_pos, tr_src_rot, tr_dst_pos, tr_dst_rot


def pose_update(off_pos, off_vel, off_rot, off_ang, in_pos, in_vel, in_rot, in_ang,
                tr_src_pos, tr_src_rot, tr_dst_pos, tr_dst_rot, halflife, dt):
    """Inertialization.py:217-297. Returns (pos, vel, rot, ang, off_pos, off_vel, off_rot, off_ang)."""
    n = len(in_pos)
    pos, vel, ang = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
    rt = np.zeros((n, 4))
    off_pos, off_vel, off_rot, off_ang = off_pos.copy(), off_vel.copy(), off_rot.copy(), off_ang.copy()
    for i in range(n):
        ix, iv, ir, ia = in_pos[i], in_vel[i], in_rot[i], in_ang[i]
        if i == 0:
            ix = rot.q_rotate(tr_dst_rot, rot.q_inv_rotate(tr_src_rot, in_pos[0] - tr_src_pos)) + tr_dst_pos
            iv = rot.q_rotate(tr_dst_rot, rot.q_inv_rotate(tr_src_rot, in_vel[0]))
            ir = rot.normalize(rot.q_mul(tr_dst_rot, rot.q_inv_mul(tr_src_rot, in_rot[0])))
            ia = rot.q_rotate(tr_dst_rot, rot.q_inv_rotate(tr_src_rot, in_ang[0]))
        off_pos[i], off_vel[i] = decay_pos(off_pos[i], off_vel[i], halflife, dt)
        pos[i], vel[i] = ix + off_pos[i], iv + off_vel[i]
        off_rot[i], off_ang[i] = decay_rot(off_rot[i], off_ang[i], halflife, dt)
        rt[i], ang[i] = rot.q_mul(off_rot[i], ir), off_ang[i] + ia
    return pos, vel, rt, ang, off_pos, off_vel, off_rot, off_ang
