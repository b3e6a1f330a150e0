"""GPU parity of the quaternion / FK / IK / inertialization kernels vs reference goldens + oracle."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from mocha_oracle import driver as odriver
from mocha_sigasia2023_b200 import kinematics as kin
from mocha_sigasia2023_b200 import skeleton

pytestmark = pytest.mark.gpu


def cu(x, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


@pytest.fixture(scope="module")
def gkin(golden_dir):
    return np.load(os.path.join(golden_dir, "kin.npz"))


def close(a, b, rtol=1e-4, atol=1e-5):
    np.testing.assert_allclose(a.cpu().numpy() if isinstance(a, torch.Tensor) else a, b, rtol=rtol, atol=atol)


def test_rotation_formats(gkin):
    d = gi.kin_inputs()
    got = kin.xy_to_quat(cu(d["xy"])).cpu().numpy()
    want = gkin["from_xform_xy"]
    # quaternion sign-invariant comparison near branch boundaries (SURVEY §7 hard parts)
    dots = np.abs((got * want).sum(-1))
    assert dots.min() > 1 - 1e-5
    close(kin.quat_to_xy(cu(d["lrot"])), gkin["to_xform_xy"])


def test_fk_family(gkin):
    d = gi.kin_inputs()
    par = kin.parents_tensor(skeleton.BONE_PARENTS, "cuda")
    gr, gp = kin.fk(cu(d["lrot"]), cu(d["lpos"]), par)
    close(gr, gkin["fk_grot"]); close(gp, gkin["fk_gpos"])
    g4 = kin.fk_vel(cu(d["lrot"]), cu(d["lpos"]), cu(d["lvel"]), cu(d["lang"]), par)
    for a, k in zip(g4, ("fkv_grot", "fkv_gpos", "fkv_gvel", "fkv_gang")):
        close(a, gkin[k], 1e-4, 2e-5)
    lr, lp = kin.ik(cu(gkin["fk_grot"]), cu(gkin["fk_gpos"]), par)
    close(lr, gkin["ik_lrot"]); close(lp, gkin["ik_lpos"])
    # leading batch dims and a big ragged batch
    big = np.tile(d["lrot"], (31, 1, 1))[None]
    bigp = np.tile(d["lpos"], (31, 1, 1))[None]
    gr2, gp2 = kin.fk(cu(big), cu(bigp), par)
    close(gr2[0, -37:], gkin["fk_grot"]); close(gp2[0, :37], gkin["fk_gpos"])


def test_fk_streaming_kernel_large_ragged_batch():
    """Batches of >= 4096 skeletons take the thread-per-skeleton streaming kernel (fk_rows_kernel): a ragged size (last warp
    group partly filled) against the oracle's fk / fk_vel (motion/quat.py:166-204) in float64."""
    from mocha_oracle import rot as orot
    rng = np.random.default_rng(5)
    F, J = 4096 + 32 * 7 + 21, len(skeleton.BONE_PARENTS)
    lrot = rng.standard_normal((F, J, 4)); lrot /= np.linalg.norm(lrot, axis=-1, keepdims=True)
    lpos = rng.standard_normal((F, J, 3)); lvel = rng.standard_normal((F, J, 3)); lang = rng.standard_normal((F, J, 3))
    par = kin.parents_tensor(skeleton.BONE_PARENTS, "cuda")
    f32 = lambda x: cu(x.astype(np.float32))
    want = orot.fk_vel(lrot.astype(np.float32).astype(np.float64), lpos.astype(np.float32).astype(np.float64),
                       lvel.astype(np.float32).astype(np.float64), lang.astype(np.float32).astype(np.float64),
                       list(skeleton.BONE_PARENTS))
    gr, gp = kin.fk(f32(lrot), f32(lpos), par)
    close(gr, want[0], 1e-4, 5e-5); close(gp, want[1], 1e-4, 5e-5)
    g4 = kin.fk_vel(f32(lrot), f32(lpos), f32(lvel), f32(lang), par)
    for a, w in zip(g4, want):
        close(a, w, 1e-4, 2e-4)
    # and the two kernels agree on a prefix small enough for the warp-per-skeleton kernel
    gr_s, gp_s = kin.fk(f32(lrot[:1000]), f32(lpos[:1000]), par)
    close(gr[:1000], gr_s.cpu().numpy(), 1e-5, 1e-5); close(gp[:1000], gp_s.cpu().numpy(), 1e-5, 1e-5)


def test_two_bone_ik(gkin):
    t = gi.kin_inputs()["ik2"]
    a, b = kin.ik_two_bone(*[cu(t[k]) for k in ("root", "mid", "end", "target", "fwd", "root_gr", "mid_gr", "par_gr")],
                           0.015)
    got = torch.cat([a, b], dim=1).cpu().numpy()
    np.testing.assert_allclose(got, gkin["ik_two_bone"], rtol=1e-9, atol=1e-10)


def test_contact_update_trajectories(gkin):
    d = gi.kin_inputs()["contact"]
    S, L = d["flag"].shape
    state = torch.zeros(S, dtype=torch.int32, device="cuda")
    lock = torch.zeros(S, dtype=torch.int32, device="cuda")
    p0 = cu(d["pos"][:, 0])
    position, point, target = p0.clone(), p0.clone(), p0.clone()
    velocity, off_pos, off_vel = torch.zeros_like(p0), torch.zeros_like(p0), torch.zeros_like(p0)
    rec = np.zeros((S, L - 1, 20))
    for f in range(1, L):
        kin.contact_update(state, lock, position, velocity, point, target, off_pos, off_vel, cu(d["pos"][:, f]),
                           cu(d["flag"][:, f].astype(np.int32)), 0.2, 0.02, 0.1, 1.0 / 60.0)
        position[:, 1] = torch.clamp(position[:, 1], min=0.02)
        rec[:, f - 1] = torch.cat([state[:, None].double(), lock[:, None].double(), position, velocity, point, target,
                                   off_pos, off_vel], dim=1).cpu().numpy()
    got = rec.reshape(-1, 20)
    np.testing.assert_array_equal(got[:, :2], gkin["contact_traj"][:, :2])    # state machine bits exact
    np.testing.assert_allclose(got, gkin["contact_traj"], rtol=1e-9, atol=1e-11)


def test_pose_inertialization(gkin):
    pz = gi.kin_inputs()["pose"]
    n = pz["src_pos"].shape[0]
    off = (torch.zeros((n, 25, 3), dtype=torch.float64, device="cuda"),
           torch.zeros((n, 25, 3), dtype=torch.float64, device="cuda"),
           cu(np.tile(np.array([1.0, 0, 0, 0]), (n, 25, 1))),
           torch.zeros((n, 25, 3), dtype=torch.float64, device="cuda"))
    root = tuple(cu(pz[k]) for k in ("root_pos", "root_vel", "root_rot", "root_ang"))
    src = tuple(cu(pz[k]) for k in ("src_pos", "src_vel", "src_rot", "src_ang"))
    dst = tuple(cu(pz[k]) for k in ("dst_pos", "dst_vel", "dst_rot", "dst_ang"))
    tr = kin.pose_transition(off, root, src, dst)
    got = torch.cat([t.reshape(n, -1) for t in (*off, *tr)], dim=1).cpu().numpy()
    np.testing.assert_allclose(got, gkin["pose_transition"], rtol=1e-9, atol=1e-11)
    out = kin.pose_update(off, dst, tr, 0.1, 1.0 / 60.0)
    got = torch.cat([t.reshape(n, -1) for t in (*out, *off)], dim=1).cpu().numpy()
    np.testing.assert_allclose(got, gkin["pose_update"], rtol=1e-8, atol=1e-10)


def test_post_frame_against_oracle():
    rng = np.random.default_rng(3)
    B, T, V = 5, 60, 24
    post = kin.PostProcessor(B, "cuda")
    oracles = [odriver.ClipPost(odriver.PostParams(skeleton.BONE_PARENTS)) for _ in range(B)]
    t = np.arange(12)[:, None, None, None] / 60.0
    for f in range(12):
        Y = (0.3 * rng.standard_normal((B, T, V, 15))).astype(np.float32)
        # plausible leg geometry so the two-bone IK is well conditioned
        Y[..., 1] -= 0.35
        hips = rng.standard_normal((B, T, 3)).astype(np.float32)
        rvel = rng.standard_normal((B, 3)).astype(np.float32)
        rang = (0.5 * rng.standard_normal((B, 3))).astype(np.float32)
        contacts = (rng.random((B, 2)) < 0.5).astype(np.uint8)
        post.step(cu(Y), cu(hips), cu(rvel), cu(rang), cu(contacts))
        got = post.read()
        for b in range(B):
            want = oracles[b].frame(Y[b], hips[b], rvel[b], rang[b], contacts[b], init=(f == 0))
            for k in ("pos", "rot", "vel", "ang", "blend_pos", "ik_pos", "src_root_pos", "src_root_rot"):
                np.testing.assert_allclose(got[k][b], want[k], rtol=1e-4, atol=2e-5, err_msg=f"frame {f} clip {b} {k}")
            dots = np.abs((got["ik_rot"][b] * want["ik_rot"]).sum(-1))
            assert dots.min() > 1 - 1e-5, f"frame {f} clip {b} ik_rot"
