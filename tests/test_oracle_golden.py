"""Pin the CPU oracle (oracle/mocha_oracle) against golden vectors produced by the live reference
(oracle/gen_golden.py). CPU-only; the oracle is test infrastructure, never a product path."""
import os

import numpy as np
import pytest

import golden_inputs as gi
from mocha_oracle import inertial, matching, nets, rot
from mocha_sigasia2023_b200 import skeleton, weights


def _close(a, b, rtol, atol):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


@pytest.fixture(scope="module")
def gnets(golden_dir):
    return np.load(os.path.join(golden_dir, "nets.npz"))


@pytest.fixture(scope="module")
def gkin(golden_dir):
    return np.load(os.path.join(golden_dir, "kin.npz"))


@pytest.fixture(scope="module")
def gen_sd():
    return {k: v.numpy() for k, v in weights.generator_state_dict(1777).items()}


@pytest.fixture(scope="module")
def cvae_sd():
    return {k: v.numpy() for k, v in weights.cvae_state_dict(1778).items()}


def test_inputs_reproducible(gnets):
    src, cha = gi.pose_windows()
    cond, eps = gi.cvae_inputs()
    got = np.array([src.sum(dtype=np.float64), cha.sum(dtype=np.float64), cond.sum(dtype=np.float64),
                    eps.sum(dtype=np.float64)])
    np.testing.assert_array_equal(got, gnets["input_checksum"])


def test_graph_buffers_match_reference_layout():
    # column-normalised 'distance' partition: every column of sum_k A[k] sums to 1
    A_j, A_b = skeleton.joint_adjacency(), skeleton.body_adjacency()
    assert A_j.shape == (3, 24, 24) and A_b.shape == (2, 6, 6)
    np.testing.assert_allclose(A_j.sum(axis=(0, 1)), 1.0, rtol=1e-6)
    np.testing.assert_allclose(A_b.sum(axis=(0, 1)), 1.0, rtol=1e-6)
    assert [int((A_j[k] != 0).sum()) for k in range(3)] == [24, 46, 52]   # SURVEY §2.2 K3 probe
    assert skeleton.BONE_PARENTS == [-1, 0, 1, 2, 3, 4, 1, 6, 7, 8, 9, 10, 11, 12, 9, 14, 15, 9, 17, 18, 19, 1, 21, 22, 23]


def test_generator_stages(gnets, gen_sd):
    src, cha = gi.pose_windows()
    tok = nets.mot_embedding(gen_sd, src)
    _close(tok, gnets["tokens"], 1e-4, 1e-5)
    enc_s = nets.encoder(gen_sd, tok + gen_sd["pos_emb"][:, :90])
    _close(enc_s, gnets["src_encoded"], 1e-4, 2e-5)
    enc_c = nets.encoder(gen_sd, nets.mot_embedding(gen_sd, cha) + gen_sd["pos_emb"][:, :90])
    _close(enc_c, gnets["cha_encoded"], 1e-4, 2e-5)
    cnt = np.transpose(nets.mean_variance_norm(np.transpose(enc_s, (0, 2, 1))), (0, 2, 1))
    _close(cnt, gnets["src_cnt"], 1e-4, 2e-5)
    dec = nets.decoder(gen_sd, gnets["src_encoded"], gnets["cha_encoded"])
    _close(dec, gnets["decoded"], 1e-4, 5e-5)
    _close(nets.to_mot(gen_sd, gnets["decoded"]), gnets["Ytil"], 1e-4, 2e-5)


def test_generator_forward(gnets, gen_sd):
    src, cha = gi.pose_windows()
    _close(nets.generator_forward(gen_sd, src, cha), gnets["forward"], 2e-4, 1e-4)
    feats = nets.generator_forward(gen_sd, src, cha, extract_feature=True)
    _close(feats[3], gnets["feat_cha_cnt"], 2e-4, 5e-5)


def test_cvae(gnets, cvae_sd):
    cond, eps = gi.cvae_inputs()
    out, mu, logvar = nets.cvae_sample(cvae_sd, cond, None)
    _close(mu, gnets["cvae_mu"], 1e-4, 2e-5)
    _close(logvar, gnets["cvae_logvar"], 1e-4, 2e-5)
    _close(out, gnets["cvae_det"], 1e-4, 2e-5)
    out_e, _, _ = nets.cvae_sample(cvae_sd, cond, eps)
    _close(out_e, gnets["cvae_eps"], 1e-4, 2e-5)


def test_cvae_posterior(gnets, cvae_sd):
    """CVAE.encode / CVAE.forward (training-side surface, model_CVAE.py:33-42) with the posterior noise replayed."""
    cond, eps = gi.cvae_inputs()
    x = gi.cvae_posterior_inputs()
    mu, logvar = nets.cvae_encode(cvae_sd, x, cond)
    _close(mu, gnets["cvae_enc_mu"], 1e-4, 2e-5)
    _close(logvar, gnets["cvae_enc_logvar"], 1e-4, 2e-5)
    out, _, (mu_pr, _) = nets.cvae_forward(cvae_sd, x, cond, eps)
    _close(out, gnets["cvae_fwd"], 1e-4, 2e-5)
    _close(mu_pr, gnets["cvae_mu"], 1e-4, 2e-5)


def test_rotation_helpers(gkin):
    d = gi.kin_inputs()
    par = skeleton.BONE_PARENTS
    _close(rot.q_from_xy(d["xy"]), gkin["from_xform_xy"], 1e-6, 1e-6)
    _close(rot.q_to_xy(d["lrot"]), gkin["to_xform_xy"], 1e-6, 1e-6)
    gr, gp = rot.fk(d["lrot"], d["lpos"], par)
    _close(gr, gkin["fk_grot"], 1e-6, 1e-6)
    _close(gp, gkin["fk_gpos"], 1e-6, 1e-6)
    g4 = rot.fk_vel(d["lrot"], d["lpos"], d["lvel"], d["lang"], par)
    for a, k in zip(g4, ("fkv_grot", "fkv_gpos", "fkv_gvel", "fkv_gang")):
        _close(a, gkin[k], 1e-6, 1e-6)
    lr, lp = rot.ik(gkin["fk_grot"], gkin["fk_gpos"], par)
    _close(lr, gkin["ik_lrot"], 1e-6, 1e-6)
    _close(lp, gkin["ik_lpos"], 1e-6, 1e-6)
    _close(rot.q_exp(d["vec3"]), gkin["exp"], 1e-6, 1e-7)
    _close(rot.q_log(d["lrot"][0]), gkin["log"], 1e-6, 1e-7)
    _close(rot.q_rotate(d["lrot"][0], d["lpos"][0]), gkin["mul_vec"], 1e-6, 1e-7)


def test_fk_vel_bone_and_two_bone_ik(gkin):
    d = gi.kin_inputs()
    par = np.array(skeleton.BONE_PARENTS)
    rows = []
    for i in range(4):
        for bone in (5, 24):
            r = rot.fk_vel_bone(d["lpos"][i].astype(np.float64), d["lvel"][i].astype(np.float64),
                                d["lrot"][i].astype(np.float64), d["lang"][i].astype(np.float64), par, bone)
            rows.append(np.concatenate([r[0], r[1], r[2], r[3]]))
    _close(np.stack(rows), gkin["fk_vel_bone"], 1e-12, 1e-12)
    t = d["ik2"]
    rows = []
    for i in range(t["root"].shape[0]):
        a, b = rot.ik_two_bone(t["root"][i], t["mid"][i], t["end"][i], t["target"][i], t["fwd"][i], t["root_gr"][i],
                               t["mid_gr"][i], t["par_gr"][i], 0.015)
        rows.append(np.concatenate([a, b]))
    _close(np.stack(rows), gkin["ik_two_bone"], 1e-10, 1e-12)


def test_contact_state_machine(gkin):
    d = gi.kin_inputs()["contact"]
    rec = []
    locks = 0
    for s in range(d["pos"].shape[0]):
        st = [False, False, d["pos"][s, 0].copy(), np.zeros(3), d["pos"][s, 0].copy(), d["pos"][s, 0].copy(),
              np.zeros(3), np.zeros(3)]
        for f in range(1, d["pos"].shape[1]):
            st = list(inertial.contact_update(*st, d["pos"][s, f], bool(d["flag"][s, f]), 0.2, 0.02, 0.1, 1.0 / 60.0))
            st[2] = st[2].copy()
            st[2][1] = max(st[2][1], 0.02)
            locks += int(st[1])
            rec.append(np.concatenate([[float(st[0]), float(st[1])], *st[2:]]))
    got = np.stack(rec)
    assert 0 < locks < got.shape[0]          # both locked and unlocked phases are exercised
    _close(got, gkin["contact_traj"], 1e-12, 1e-12)


def test_pose_inertialization(gkin):
    pz = gi.kin_inputs()["pose"]
    for s in range(pz["src_pos"].shape[0]):
        off = [np.zeros((25, 3)), np.zeros((25, 3)), np.tile(np.array([1.0, 0, 0, 0]), (25, 1)), np.zeros((25, 3))]
        r = inertial.pose_transition(*off, pz["root_pos"][s], pz["root_vel"][s], pz["root_rot"][s], pz["root_ang"][s],
                                     pz["src_pos"][s], pz["src_vel"][s], pz["src_rot"][s], pz["src_ang"][s],
                                     pz["dst_pos"][s], pz["dst_vel"][s], pz["dst_rot"][s], pz["dst_ang"][s])
        _close(np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in r]), gkin["pose_transition"][s],
               1e-12, 1e-12)
        u = inertial.pose_update(r[0], r[1], r[2], r[3], pz["dst_pos"][s], pz["dst_vel"][s], pz["dst_rot"][s],
                                 pz["dst_ang"][s], r[4], r[5], r[6], r[7], 0.1, 1.0 / 60.0)
        _close(np.concatenate([np.asarray(x, dtype=np.float64).reshape(-1) for x in u]), gkin["pose_update"][s],
               1e-10, 1e-12)


@pytest.mark.parametrize("name", list(gi.MATCH_CASES))
def test_matching_against_balltree(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "match.npz"))
    db, q = gi.match_inputs(name)
    k = gi.MATCH_CASES[name][3]
    dist, idx = matching.knn(db, q, k)
    np.testing.assert_array_equal(idx, g[name + "_idx"])
    _close(dist, g[name + "_dist"], 1e-10, 1e-10)
    d2, i2 = matching.knn_gemm(db, q, k)
    ok = g[name + "_margin"] > 1e-5
    np.testing.assert_array_equal(i2[ok], g[name + "_idx"][ok])
