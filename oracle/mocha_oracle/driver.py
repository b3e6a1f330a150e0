"""NumPy restatement of the per-frame post-process of the reference driver
(test_fullframework.py:303-308, :321-434 for frame 0 and :457-462, :476-641 for frames i >= 1):
de-normalised decoder output -> pose, root integration, position blending, partial FK,
contact inertializer, two-bone IK. dtype behaviour follows the reference: joint quantities are
float32 values, the root / contact state is float64, the source root stays float32."""
from __future__ import annotations

import numpy as np

from . import inertial, rot

IDENT = np.array([1, 0, 0, 0])  # int64 on purpose, as in the reference (:321)


class PostParams:
    def __init__(self, parents, contact_bones=(5, 24), dt=1.0 / 60.0, ik_enabled=True, ik_max_length_buffer=0.015,
                 ik_foot_height=0.02, ik_unlock_radius=0.2, ik_blending_halflife=0.1):
        self.parents = np.asarray(parents)
        self.contact_bones = np.asarray(contact_bones)
        self.dt = dt
        self.ik_enabled = ik_enabled
        self.ik_max_length_buffer = ik_max_length_buffer
        self.ik_foot_height = ik_foot_height
        self.ik_unlock_radius = ik_unlock_radius
        self.ik_blending_halflife = ik_blending_halflife


def pose_from_output(Y):
    """Y [T,V,15] de-normalised decoder output -> last-frame pos/rot/ang and all-frame vel
    (test_fullframework.py:458-462)."""
    pos = Y[-1, :, :3]
    txy = Y[-1, :, 3:9].reshape(pos.shape[0], 3, 2)
    vel = Y[..., 9:12]
    ang = Y[-1, :, 12:15]
    return pos, rot.q_from_xy(txy), vel, ang


def speed_ratio(vel_all, src_hips_vel):
    """test_fullframework.py:492-496."""
    r = np.linalg.norm(vel_all[:, 0], axis=1).mean() / np.linalg.norm(src_hips_vel, axis=1).mean()
    if r > 3.0 or r < 0.33:
        r = 1.0
    return r


class ClipPost:
    """State of one clip across frames (the reference's python lists, reduced to what is read)."""

    def __init__(self, params: PostParams):
        self.P = params
        self.root_pos = None      # trans_Ypos_list[-1][0]
        self.root_rot = None      # trans_Yrot_list[-1][0]
        self.prev_pos = None      # trans_Ypos_list[-1]
        self.prev_ik_pos = None   # ik_trans_Ypos_list[-1]
        self.src_root_pos = None  # float32
        self.src_root_rot = None
        n = len(self.P.contact_bones)
        self.c_state = np.zeros(n, dtype=bool)
        self.c_lock = np.zeros(n, dtype=bool)
        self.c_pos = np.zeros((n, 3))
        self.c_vel = np.zeros((n, 3))
        self.c_point = np.zeros((n, 3))
        self.c_target = np.zeros((n, 3))
        self.c_off_pos = np.zeros((n, 3))
        self.c_off_vel = np.zeros((n, 3))

    def _src_root(self, src_rvel, src_rang, init):
        dt = self.P.dt
        if init:
            vel = rot.q_rotate(IDENT, src_rvel)
            ang = rot.q_rotate(IDENT, src_rang)
            pos = np.array([0, 0, 0]) + vel * dt
            rt = rot.q_mul(IDENT, rot.q_from_scaled_angle_axis(ang * dt))
        else:
            vel = rot.q_rotate(self.src_root_rot, src_rvel)
            ang = rot.q_rotate(self.src_root_rot, src_rang)
            pos = self.src_root_pos + vel * dt
            rt = rot.q_mul(self.src_root_rot, rot.q_from_scaled_angle_axis(ang * dt))
        # stored into float32 arrays (src_Ypos[i,-1,0] = ..., :480-483)
        self.src_root_pos = pos.astype(np.float32)
        self.src_root_rot = rt.astype(np.float32)
        return {"src_root_pos": self.src_root_pos.astype(np.float64), "src_root_rot": self.src_root_rot.astype(np.float64),
                "src_root_vel": np.asarray(vel, dtype=np.float32).astype(np.float64),
                "src_root_ang": np.asarray(ang, dtype=np.float32).astype(np.float64)}

    def frame(self, Y, src_hips_vel, src_rvel, src_rang, contacts, init=False):
        """One frame. Y [T,V,15] float32; src_hips_vel [T,3]; src_rvel/src_rang [3]; contacts [2]."""
        P, dt = self.P, self.P.dt
        out = self._src_root(src_rvel, src_rang, init)
        jpos, jrot, vel_all, jang = pose_from_output(Y)
        ratio = speed_ratio(vel_all, src_hips_vel)
        yrvel = src_rvel * ratio
        yrang = src_rang
        prev_rot = IDENT if init else self.root_rot
        prev_pos = np.array([0, 0, 0]) if init else self.root_pos
        rootvel = rot.q_rotate(prev_rot, yrvel)
        rootang = rot.q_rotate(prev_rot, yrang)
        rootpos = prev_pos + rootvel * dt
        rootrot = rot.q_mul(prev_rot, rot.q_from_scaled_angle_axis(rootang * dt))
        pos = np.concatenate([rootpos[None], jpos], axis=0)
        vel = np.concatenate([rootvel[None], vel_all[-1]], axis=0)
        rt = np.concatenate([rootrot[None], jrot], axis=0)
        ang = np.concatenate([rootang[None], jang], axis=0)
        out.update(pos=pos, vel=vel, rot=rt, ang=ang)
        if init:
            for f, bone in enumerate(P.contact_bones):
                bp, bv, _, _ = rot.fk_vel_bone(pos, vel, rt, ang, P.parents, bone)
                self.c_state[f] = False
                self.c_lock[f] = False
                self.c_pos[f], self.c_vel[f], self.c_point[f], self.c_target[f] = bp, bv, bp, bp
                self.c_off_pos[f], self.c_off_vel[f] = 0.0, 0.0
            self.root_pos, self.root_rot = rootpos, rootrot
            self.prev_pos, self.prev_ik_pos = pos.copy(), pos.copy()
            out.update(blend_pos=pos.copy(), ik_pos=pos.copy(), ik_rot=rt.copy())
            return out

        bone_pos = ((self.prev_ik_pos + vel * dt) * 0.5 + pos * 0.5).copy()     # :532-535
        adj_rot = rt.copy()
        cflags = np.asarray(contacts).astype(bool)
        if P.ik_enabled:
            for f, toe in enumerate(P.contact_bones):
                heel = P.parents[toe]
                knee = P.parents[heel]
                hip = P.parents[knee]
                rootb = P.parents[hip]
                gpos, grot = rot.fk_chain(bone_pos, rt, P.parents, toe)           # :549-557
                (self.c_state[f], self.c_lock[f], self.c_pos[f], self.c_vel[f], self.c_point[f], self.c_target[f],
                 self.c_off_pos[f], self.c_off_vel[f]) = inertial.contact_update(
                    self.c_state[f], self.c_lock[f], self.c_pos[f], self.c_vel[f], self.c_point[f], self.c_target[f],
                    self.c_off_pos[f], self.c_off_vel[f], gpos[toe], cflags[f], P.ik_unlock_radius, P.ik_foot_height,
                    P.ik_blending_halflife, dt)
                self.c_pos[f][1] = max(self.c_pos[f][1], P.ik_foot_height)        # aliasing clamp :581-582
                target = self.c_pos[f] + (gpos[heel] - gpos[toe])
                fwd = rot.q_rotate(grot[knee], np.array([0.0, 1.0, 0.0], dtype=np.float32))
                adj_rot[hip], adj_rot[knee] = rot.ik_two_bone(
                    gpos[hip], gpos[knee], gpos[heel], target, fwd, grot[hip], grot[knee], grot[rootb],
                    P.ik_max_length_buffer)
        blend = (self.prev_pos + vel * dt) * 0.5 + pos * 0.5                    # :626
        self.root_pos, self.root_rot = blend[0], rootrot
        self.prev_pos, self.prev_ik_pos = blend, bone_pos
        out.update(blend_pos=blend, ik_pos=bone_pos, ik_rot=adj_rot)
        return out
