"""The NumPy-signature drop-ins (quat.py, Inertialization.py, Trainer) vs reference goldens."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as gi
from mocha_sigasia2023_b200 import Inertialization as inert
from mocha_sigasia2023_b200 import quat, skeleton, weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gkin(golden_dir):
    return np.load(os.path.join(golden_dir, "kin.npz"))


def test_quat_module_surface(gkin):
    d = gi.kin_inputs()
    par = np.array(skeleton.BONE_PARENTS)
    gr, gp = quat.fk(d["lrot"], d["lpos"], par)
    np.testing.assert_allclose(gr, gkin["fk_grot"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gp, gkin["fk_gpos"], rtol=1e-4, atol=1e-5)
    g4 = quat.fk_vel(d["lrot"], d["lpos"], d["lvel"], d["lang"], par)
    np.testing.assert_allclose(g4[2], gkin["fkv_gvel"], rtol=1e-4, atol=2e-5)
    lr, lp = quat.ik(gkin["fk_grot"], gkin["fk_gpos"], par)
    np.testing.assert_allclose(lr, gkin["ik_lrot"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(quat.exp(d["vec3"]), gkin["exp"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(quat.log(d["lrot"][0]), gkin["log"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(quat.mul_vec(d["lrot"][0], d["lpos"][0]), gkin["mul_vec"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(quat.to_xform_xy(d["lrot"]), gkin["to_xform_xy"], rtol=1e-5, atol=1e-6)
    q = quat.from_xform_xy(d["xy"])
    assert np.abs((q * gkin["from_xform_xy"]).sum(-1)).min() > 1 - 1e-5
    # fk_vel_bone / fk_partial on single poses
    i = 0
    r = quat.fk_vel_bone(d["lpos"][i], d["lvel"][i], d["lrot"][i], d["lang"][i], par, 5)
    np.testing.assert_allclose(np.concatenate(r), gkin["fk_vel_bone"][0], rtol=1e-4, atol=2e-5)
    gpos, grot, done = np.zeros((25, 3)), np.zeros((25, 4)), np.zeros(25, dtype=bool)
    quat.fk_partial(gpos, grot, done, d["lpos"][i], d["lrot"][i], par, 24)
    assert done[[0, 1, 21, 22, 23, 24]].all() and done.sum() == 6
    np.testing.assert_allclose(gpos[24], gkin["fk_gpos"][i, 24], rtol=1e-4, atol=1e-5)
    t = d["ik2"]
    a, b = quat.ik_two_bone(t["root_gr"][3], t["mid_gr"][3], t["root"][3], t["mid"][3], t["end"][3], t["target"][3],
                            t["fwd"][3], t["root_gr"][3], t["mid_gr"][3], t["par_gr"][3], 0.015)
    np.testing.assert_allclose(np.concatenate([a, b]), gkin["ik_two_bone"][3], rtol=1e-9, atol=1e-10)


def test_contact_update_dropin(gkin):
    d = gi.kin_inputs()["contact"]
    s = 1
    st = [False, False, d["pos"][s, 0].copy(), np.zeros(3), d["pos"][s, 0].copy(), d["pos"][s, 0].copy(), np.zeros(3),
          np.zeros(3)]
    L = d["pos"].shape[1]
    for f in range(1, L):
        st = list(inert.contact_update(*st, d["pos"][s, f], bool(d["flag"][s, f]), 0.2, 0.02, 0.1, 1.0 / 60.0))
        st[2][1] = max(st[2][1], 0.02)
        want = gkin["contact_traj"][s * (L - 1) + f - 1]
        got = np.concatenate([[float(st[0]), float(st[1])], *st[2:]])
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)


def test_trainer_shim_loads_checkpoint(tmp_path):
    from mocha_sigasia2023_b200.trainer import Trainer
    cfg = {"model": weights.DEFAULT_MODEL_CFG, "model_dir": str(tmp_path),
           "dataset": {"mocha": {"parents": skeleton.JOINT_PARENTS}}}
    sd = weights.generator_state_dict(4242)
    path = os.path.join(tmp_path, "gen_125.pt")
    torch.save({"gen": sd, "gen_ema": sd, "gen_opt": {}}, path)
    tr = Trainer(cfg)
    assert tr.load_checkpoint(path) == 125
    model = tr.gen_ema.eval()
    x = torch.randn(1, 60, 24, 15, device="cuda")
    tokens = model.mot_embedding(x)
    tokens = tokens + model.pos_emb[:, :tokens.shape[1]]
    assert tuple(model.encoder(tokens).shape) == (1, 90, 256)
    assert list(tr.parents) == skeleton.BONE_PARENTS
