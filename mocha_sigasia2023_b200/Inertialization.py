"""Drop-in for the reference's `motion/Inertialization.py` entry points on the hot path, same
positional signatures (NumPy in/out), computing in the library's fp64 kernels:
`contact_update` (Inertialization.py:300-377), `pose_transition` (:136-209), `pose_update` (:217-297).
Batched device-resident variants live in kinematics.py (contact_update / pose_* on CUDA tensors)."""
from __future__ import annotations

import numpy as np
import torch

from . import kinematics as kin


def _d(a, shape=None):
    t = torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).cuda()
    return t.reshape(shape) if shape is not None else t


def contact_update(contact_state, contact_lock, contact_position, contact_velocity, contact_point, contact_target,
                   contact_offset_position, contact_offset_velocity, input_contact_position, input_contact_state,
                   unlock_radius, foot_height, halflife, dt, eps=1e-8):
    st = torch.tensor([int(bool(contact_state))], dtype=torch.int32, device="cuda")
    lk = torch.tensor([int(bool(contact_lock))], dtype=torch.int32, device="cuda")
    vecs = [_d(a, (1, 3)) for a in (contact_position, contact_velocity, contact_point, contact_target,
                                    contact_offset_position, contact_offset_velocity)]
    kin.contact_update(st, lk, *vecs, _d(input_contact_position, (1, 3)),
                       torch.tensor([int(bool(input_contact_state))], dtype=torch.int32, device="cuda"),
                       float(unlock_radius), float(foot_height), float(halflife), float(dt))
    out = [v[0].cpu().numpy() for v in vecs]
    return (bool(st.item()), bool(lk.item()), out[0], out[1], out[2], out[3], out[4], out[5])


def pose_transition(bone_offset_positions, bone_offset_velocities, bone_offset_rotations,
                    bone_offset_angular_velocities, root_position, root_velocity, root_rotation,
                    root_angular_velocity, bone_src_positions, bone_src_velocities, bone_src_rotations,
                    bone_src_angular_velocities, bone_dst_positions, bone_dst_velocities, bone_dst_rotations,
                    bone_dst_angular_velocities):
    J = len(bone_offset_positions)
    off = tuple(_d(a, (1, J, -1)) for a in (bone_offset_positions, bone_offset_velocities, bone_offset_rotations,
                                            bone_offset_angular_velocities))
    root = tuple(_d(a, (1, -1)) for a in (root_position, root_velocity, root_rotation, root_angular_velocity))
    src = tuple(_d(a, (1, J, -1)) for a in (bone_src_positions, bone_src_velocities, bone_src_rotations,
                                            bone_src_angular_velocities))
    dst = tuple(_d(a, (1, J, -1)) for a in (bone_dst_positions, bone_dst_velocities, bone_dst_rotations,
                                            bone_dst_angular_velocities))
    tr = kin.pose_transition(off, root, src, dst)
    o = [t[0].cpu().numpy() for t in off]
    t4 = [t[0].cpu().numpy() for t in tr]
    return (o[0], o[1], o[2], o[3], t4[0], t4[1], t4[2], t4[3])


def pose_update(bone_positions, bone_velocities, bone_rotations, bone_angular_velocities, bone_offset_positions,
                bone_offset_velocities, bone_offset_rotations, bone_offset_angular_velocities, bone_input_positions,
                bone_input_velocities, bone_input_rotations, bone_input_angular_velocities, transition_src_position,
                transition_src_rotation, transition_dst_position, transition_dst_rotation, halflife, dt):
    J = len(bone_input_positions)
    off = tuple(_d(a, (1, J, -1)) for a in (bone_offset_positions, bone_offset_velocities, bone_offset_rotations,
                                            bone_offset_angular_velocities))
    inp = tuple(_d(a, (1, J, -1)) for a in (bone_input_positions, bone_input_velocities, bone_input_rotations,
                                            bone_input_angular_velocities))
    tr = tuple(_d(a, (1, -1)) for a in (transition_src_position, transition_src_rotation, transition_dst_position,
                                        transition_dst_rotation))
    out = kin.pose_update(off, inp, tr, float(halflife), float(dt))
    return tuple(t[0].cpu().numpy() for t in (*out, *off))
