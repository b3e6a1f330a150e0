"""CPU restatement of the whole per-frame loop of the reference driver
(test_fullframework.py:288-641) for a batch of independent clips: encode -> context feature ->
BallTree match -> CVAE sample -> AdaIN decode -> to_mot -> de-normalise -> post-process.
TEST INFRASTRUCTURE / CPU BASELINE ONLY (see package docstring)."""
from __future__ import annotations

import numpy as np

from . import driver, matching, nets

F32 = np.float32


class OraclePipeline:
    def __init__(self, gen_sd, cvae_sd, stats, cha_encoded, cha_cnt_nm, batch, parents, deterministic=False):
        """stats: dict with Y_mean/Y_std [24,15], cnt_mean/cnt_std, src_cnt_mean/std, cha_encoded_mean/std
        [90,256] (already divided by std_weight). cha_encoded [N,90,256], cha_cnt_nm [N,23040]."""
        self.g, self.c, self.st = gen_sd, cvae_sd, stats
        self.cha_encoded = np.asarray(cha_encoded, dtype=F32)
        self.db = np.asarray(cha_cnt_nm, dtype=F32)
        self.B = batch
        self.deterministic = deterministic
        self.posts = [driver.ClipPost(driver.PostParams(parents)) for _ in range(batch)]
        self.prev_cha = None
        self.frame = 0
        self.last = {}

    def encode(self, X):
        tok = nets.mot_embedding(self.g, X) + self.g["pos_emb"][:, :90]                  # :190-191
        enc = nets.encoder(self.g, tok)                                                   # :192
        cnt = np.transpose(nets.mean_variance_norm(np.transpose(enc, (0, 2, 1))), (0, 2, 1))  # :193
        return enc, cnt

    def step(self, X, src_hips_vel, src_rvel, src_rang, contacts, eps=None):
        st = self.st
        enc, cnt = self.encode(X)
        q = ((cnt - st["cnt_mean"][None]) / st["cnt_std"][None]).reshape(self.B, -1)      # :442
        _, idx = matching.knn(self.db, q, 1)                                              # :443
        idx = idx[:, 0]
        init = self.frame == 0
        if init:
            cur = self.cha_encoded[idx]                                                   # :298
        else:
            cond = np.concatenate([(cnt - st["src_cnt_mean"][None]) / st["src_cnt_std"][None],
                                   (self.prev_cha - st["cha_encoded_mean"][None]) / st["cha_encoded_std"][None]],
                                  axis=1).astype(F32)                                     # :446-447
            out, _, _ = nets.cvae_sample(self.c, cond, None if self.deterministic else eps)   # :448
            cur = (out * st["cha_encoded_std"][None] + st["cha_encoded_mean"][None]).astype(F32)  # :449
        self.prev_cha = cur
        dec = nets.decoder(self.g, enc, cur)                                              # :455
        Ytil = nets.to_mot(self.g, dec)                                                   # :456
        Y = (Ytil * st["Y_std"][None, None] + st["Y_mean"][None, None]).astype(F32)      # :457
        outs = [self.posts[b].frame(Y[b], src_hips_vel[b], src_rvel[b], src_rang[b], contacts[b], init=init)
                for b in range(self.B)]
        self.frame += 1
        self.last = {"encoded": enc, "cnt": cnt, "match_idx": idx, "cha": cur, "decoded": dec, "Y": Y}
        return {k: np.stack([o[k] for o in outs]) for k in outs[0]}
