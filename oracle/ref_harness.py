"""Run the UNMODIFIED reference driver (test_fullframework.main) on CPU under a stub harness and
record its outputs as the end-to-end golden (tests/golden/e2e.npz). Build container only.

Nothing in the reference is edited. What the harness supplies:
  * a scratch working directory with symlinks to the reference tree, holding the files main() opens:
    model_ours/pth/gen_125.pt, <cvae dir>/cvae_020000.pt (deterministic random-init weights from
    mocha_sigasia2023_b200.weights), datasets/mocha60/{norm,cnt_norm}.npz, <cvae dir>/cvae_norm.npz
    (synthetic tables from mocha_sigasia2023_b200.synthetic)
  * sys.modules['viz_motion'] with a no-op animation_plot (matplotlib is absent)
  * bvh.load -> seeded synthetic clips, bvh.save -> captured
  * recording spies (call-through wrappers) on torch.randn_like, BallTree.query, CVAE.sample and
    Generator.to_mot so intermediate values can be compared frame by frame
"""
from __future__ import annotations

import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MOCHA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

SRC_FRAMES, CHA_FRAMES = 240, 400
SRC_SEED, CHA_SEED = 0, 1
CVAE_DIR = "Neutral_AverageJoe2Neutral_Princess"


def _scratch():
    from mocha_sigasia2023_b200 import synthetic, weights
    d = tempfile.mkdtemp(prefix="mocha_ref_")
    for name in os.listdir(REF):
        os.symlink(os.path.join(REF, name), os.path.join(d, name))
    os.makedirs(os.path.join(d, "model_ours", "pth"))
    os.makedirs(os.path.join(d, CVAE_DIR))
    os.makedirs(os.path.join(d, "datasets_mocha60"))
    gsd = weights.generator_state_dict(1777)
    torch.save({"gen": gsd, "gen_ema": gsd, "gen_opt": {}}, os.path.join(d, "model_ours", "pth", "gen_125.pt"))
    torch.save(weights.cvae_state_dict(1778), os.path.join(d, CVAE_DIR, "cvae_020000.pt"))
    st = synthetic.make_norm_stats()
    np.savez(os.path.join(d, "datasets_mocha60", "norm.npz"), **st["norm"])
    np.savez(os.path.join(d, "datasets_mocha60", "cnt_norm.npz"), **st["cnt_norm"])
    np.savez(os.path.join(d, CVAE_DIR, "cvae_norm.npz"), **st["cvae_norm"])
    # config whose data_dir points at the scratch dataset dir (datasets/ is a symlink-free name here)
    import yaml
    cfg = yaml.load(open(os.path.join(REF, "configs", "config.yaml")), Loader=yaml.FullLoader)
    cfg["data_dir"] = "./datasets_mocha60/"
    with open(os.path.join(d, "harness_config.yaml"), "w") as f:
        yaml.dump(cfg, f)
    return d


def run_reference_main():
    from mocha_sigasia2023_b200 import synthetic
    d = _scratch()
    os.chdir(d)
    for p in ("", "etc", "motion", "preprocess", "net"):
        sys.path.insert(0, os.path.join(d, p))
    viz = types.ModuleType("viz_motion")
    viz.animation_plot = lambda *a, **k: None
    sys.modules["viz_motion"] = viz

    rec = {"eps": [], "match": [], "cvae_cond": [], "cvae_out": [], "Ytil": [], "saves": []}

    real_randn_like = torch.randn_like

    def spy_randn_like(t, *a, **k):
        out = real_randn_like(t, *a, **k)
        rec["eps"].append(out.detach().clone().numpy())
        return out

    torch.randn_like = spy_randn_like
    sys.argv = ["test_fullframework.py", "--config", "harness_config.yaml"]
    import test_fullframework as tf
    import bvh

    clips = {"src": synthetic.make_clip(SRC_FRAMES, SRC_SEED), "cha": synthetic.make_clip(CHA_FRAMES, CHA_SEED)}

    def fake_load(path, *a, **k):
        key = "cha" if "Princess" in path else "src"
        c = clips[key]
        return {k2: (v.copy() if isinstance(v, np.ndarray) else list(v) if isinstance(v, list) else v)
                for k2, v in c.items()}

    def fake_save(path, data, *a, **k):
        rec["saves"].append({"path": os.path.basename(path), "rotations": np.array(data["rotations"]),
                             "positions": np.array(data["positions"])})

    bvh.load, bvh.save = fake_load, fake_save
    tf.bvh.load, tf.bvh.save = fake_load, fake_save

    RealTree = tf.BallTree

    class SpyTree:
        def __init__(self, X, *a, **k):
            self._t = RealTree(X, *a, **k)
            rec["db_shape"] = np.array(X.shape)

        def query(self, q, *a, **k):
            r = self._t.query(q, *a, **k)
            rec["match"].append(np.array(r).reshape(-1)[0])
            return r

    tf.BallTree = SpyTree

    real_sample = tf.CVAE.sample

    def spy_sample(self, c, deterministic=False):
        out = real_sample(self, c, deterministic)
        if len(rec["cvae_out"]) < 6:
            rec["cvae_cond"].append(c.detach().clone().numpy())
            rec["cvae_out"].append(out.detach().clone().numpy())
        return out

    tf.CVAE.sample = spy_sample

    real_trainer_init = tf.Trainer.__init__

    def spy_trainer_init(self, cfg):
        real_trainer_init(self, cfg)
        real_to_mot = self.gen_ema.to_mot.forward

        def spy_to_mot(x):
            y = real_to_mot(x)
            rec["Ytil"].append(y.detach().clone().numpy()[0, -1])     # last frame row only (24,15)
            return y

        self.gen_ema.to_mot.forward = spy_to_mot

    tf.Trainer.__init__ = spy_trainer_init
    torch.set_num_threads(8)
    tf.main()
    torch.randn_like = real_randn_like
    return rec


def generate(out_path):
    rec = run_reference_main()
    saves = {s["path"].split("_")[0]: s for s in rec["saves"]}
    out = {
        "src_rotations": saves["Src"]["rotations"], "src_positions": saves["Src"]["positions"],
        "ours_rotations": saves["Ours"]["rotations"], "ours_positions": saves["Ours"]["positions"],
        "eps": np.stack([e.reshape(-1) for e in rec["eps"]]).astype(np.float32),
        "match": np.array(rec["match"], dtype=np.int64),
        "db_shape": rec["db_shape"],
        "cvae_cond0": rec["cvae_cond"][0].astype(np.float32), "cvae_out0": rec["cvae_out"][0].astype(np.float32),
        "Ytil_last_rows": np.stack(rec["Ytil"]).astype(np.float32),
    }
    np.savez_compressed(out_path, **out)
    print("e2e.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    generate(os.path.join(ROOT, "tests", "golden", "e2e.npz"))
