"""CPU restatement of the reference driver around the per-frame loop, for ONE clip pair: the window
feature extraction of test_fullframework.py:135-186 (source) / :213-264 (character), the batch
encode (:188-194, :266-272), the frame loop (OraclePipeline, :288-641) and the final payload handed
to bvh.save (:672-721). TEST INFRASTRUCTURE ONLY (see package docstring).

Pinned by tests/test_oracle_e2e.py against tests/golden/e2e.npz, the recording of the UNMODIFIED
reference `main()` (oracle/ref_harness.py)."""
from __future__ import annotations

import numpy as np

from . import nets, rot
from .pipeline import OraclePipeline

F32 = np.float32


def _central_diff(x):
    """:164-169 - central differences along the window axis, linear extrapolation at both ends."""
    v = np.empty_like(x)
    v[:, 1:-1] = 0.5 * (x[:, 2:] - x[:, 1:-1]) * 60.0 + 0.5 * (x[:, 1:-1] - x[:, :-2]) * 60.0
    v[:, 0] = v[:, 1] - (v[:, 3] - v[:, 2])
    v[:, -1] = v[:, -2] + (v[:, -2] - v[:, -3])
    return v


def window_features(win: dict, X_mean, X_std, parents) -> dict:
    """win: process_data's window arrays (pos/vel/rot/ang [nwin,60,25,*], contacts [nwin,60,2]).
    Returns the arrays the loop reads (float32 like the reference's np.array(..., dtype=np.float32))."""
    Ypos = np.array(win["pos"], dtype=F32)
    Yvel = np.array(win["vel"], dtype=F32)
    Yrot = np.array(win["rot"], dtype=F32)
    Yang = np.array(win["ang"], dtype=F32)
    window = Ypos.shape[1]
    Yrvel = rot.q_inv_rotate(Yrot[:, :, 0], Yvel[:, :, 0])                   # :142
    Yrang = rot.q_inv_rotate(Yrot[:, :, 0], Yang[:, :, 0])                   # :143
    Grot, Gpos, Gvel, Gang = rot.fk_vel(Yrot, Ypos, Yvel, Yang, parents)     # :146
    for G in (Gpos, Grot, Gvel, Gang):                                       # :148-151
        G[:, :, 0:1] = np.repeat(G[:, -1:, 0:1], window, axis=1)
    R0 = Grot[:, :, 0:1]
    Xpos = rot.q_inv_rotate(R0, Gpos - Gpos[:, :, 0:1])                      # :154
    Xrot = rot.q_inv_mul(R0, Grot)                                           # :155
    Xtxy = rot.q_to_xy(Xrot).astype(F32)                                     # :156
    Xvel = rot.q_inv_rotate(R0, Gvel)                                        # :157
    Xang = rot.q_inv_rotate(R0, Gang)                                        # :158
    Yrot2, Ypos2 = rot.ik(Xrot, Xpos, parents)                               # :160
    Yvel2 = _central_diff(Ypos2)                                             # :164-169
    b, ns, nj = Xtxy.shape[:3]
    X = np.concatenate([Xpos, Xtxy.reshape(b, ns, nj, -1), Xvel, Xang], axis=-1)   # :180-185
    X = (X[:, :, 1:] - np.asarray(X_mean)[None, None, 1:]) / np.asarray(X_std)[None, None, 1:]   # :186
    return {"X": X.astype(F32), "Yrvel": Yrvel, "Yrang": Yrang, "Ypos": Ypos2, "Yrot": Yrot2, "Yvel": Yvel2,
            "contacts": np.array(win["contacts"], dtype=np.uint8)}


def encode_windows(gen_sd, X, batch=32):
    """:188-194 in chunks: encoded [n,90,256] and the context feature cnt [n,90,256]."""
    encs, cnts = [], []
    for s in range(0, X.shape[0], batch):
        tok = nets.mot_embedding(gen_sd, X[s:s + batch]) + gen_sd["pos_emb"][:, :90]
        enc = nets.encoder(gen_sd, tok)
        cnt = np.transpose(nets.mean_variance_norm(np.transpose(enc, (0, 2, 1))), (0, 2, 1))
        encs.append(enc.astype(F32))
        cnts.append(cnt.astype(F32))
    return np.concatenate(encs), np.concatenate(cnts)


def final_payload(rot_seq, pos_seq, parents):
    """:672-694 + :704-721: global FK of the sequence, the simulation root is dropped and the hips take
    their global transform; rotations as Euler degrees ('xyz' formulas of quat.to_euler)."""
    grot, gpos = rot.fk(rot_seq, pos_seq, parents)
    p = np.array(pos_seq[:, 1:])
    r = np.array(rot_seq[:, 1:])
    p[:, 0] = gpos[:, 1]
    r[:, 0] = grot[:, 1]
    return np.degrees(rot.q_to_euler(r)), p


def stats_for_loop(raw: dict) -> dict:
    """The driver's std /= std_weight (:89-92) and the joint rows of Y_mean / Y_std (:457)."""
    w = raw["cvae_norm"]["std_weight"]
    return {"Y_mean": np.asarray(raw["norm"]["Y_mean"][1:], dtype=F32), "Y_std": np.asarray(raw["norm"]["Y_std"][1:], dtype=F32),
            "cnt_mean": np.asarray(raw["cnt_norm"]["mean"], dtype=F32), "cnt_std": (raw["cnt_norm"]["std"] / w).astype(F32),
            "src_cnt_mean": np.asarray(raw["cvae_norm"]["src_cnt_mean"], dtype=F32),
            "src_cnt_std": (raw["cvae_norm"]["src_cnt_std"] / w).astype(F32),
            "cha_encoded_mean": np.asarray(raw["cvae_norm"]["cha_encoded_mean"], dtype=F32),
            "cha_encoded_std": (raw["cvae_norm"]["cha_encoded_std"] / w).astype(F32)}


def run_clip(src_win: dict, cha_win: dict, raw_stats: dict, gen_sd: dict, cvae_sd: dict, parents, eps_seq=None,
             frames=None) -> dict:
    """Everything main() computes between process_data and bvh.save, for one clip pair."""
    Xm, Xs = raw_stats["norm"]["X_mean"], raw_stats["norm"]["X_std"]
    src = window_features(src_win, Xm, Xs, parents)
    cha = window_features(cha_win, Xm, Xs, parents)
    st = stats_for_loop(raw_stats)
    src_enc, src_cnt = encode_windows(gen_sd, src["X"])
    cha_enc, cha_cnt = encode_windows(gen_sd, cha["X"])
    db = ((cha_cnt - st["cnt_mean"][None]) / st["cnt_std"][None]).reshape(cha_cnt.shape[0], -1)   # :293
    pipe = OraclePipeline(gen_sd, cvae_sd, st, cha_enc, db, 1, parents, deterministic=eps_seq is None)
    n = src["X"].shape[0] if frames is None else min(frames, src["X"].shape[0])
    outs, match, ytil = [], [], []
    for i in range(n):
        eps = None if (i == 0 or eps_seq is None) else np.asarray(eps_seq[i - 1], dtype=F32)[None]
        o = pipe.step(src["X"][i:i + 1], src["Yvel"][i:i + 1, :, 1], src["Yrvel"][i:i + 1, -1], src["Yrang"][i:i + 1, -1],
                      src["contacts"][i:i + 1, -1], eps)
        outs.append({k: v[0] for k, v in o.items()})
        match.append(int(pipe.last["match_idx"][0]))
        ytil.append(((pipe.last["Y"][0, -1] - st["Y_mean"]) / st["Y_std"]).astype(F32))
    src_pos = src["Ypos"][:n, -1].astype(F32)
    src_rot = src["Yrot"][:n, -1].astype(F32)
    src_pos[:, 0] = np.stack([o["src_root_pos"] for o in outs]).astype(F32)      # :478-483
    src_rot[:, 0] = np.stack([o["src_root_rot"] for o in outs]).astype(F32)
    ik_pos = np.stack([o["ik_pos"] for o in outs])
    ik_rot = np.stack([o["ik_rot"] for o in outs])
    src_eul, src_p = final_payload(src_rot, src_pos, parents)
    our_eul, our_p = final_payload(ik_rot, ik_pos, parents)
    return {"src_rotations": src_eul, "src_positions": src_p, "ours_rotations": our_eul, "ours_positions": our_p,
            "match": np.array(match), "Ytil_last_rows": np.stack(ytil), "n_db": int(cha_enc.shape[0])}
