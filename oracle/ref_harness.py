"""Run the UNMODIFIED reference driver (test_fullframework.main) under a stub harness.

Three uses:
  * `python oracle/ref_harness.py` (build container): record the end-to-end golden tests/golden/e2e.npz from
    the reference on CPU;
  * `--patched --out x.npz`: Level-1 drop-in proof - the same unmodified `main()` with BallTree, quat,
    Inertialization, Trainer, CVAE and mean_variance_norm of its namespace replaced by this package
    (tests/test_gpu_level1.py compares the result with e2e.npz; the reparameterisation noise is replayed);
  * `--time --out x.json`: time the reference's own per-frame loop on the host cores (bench.py reference arm).
The tree is /root/reference in the build container, else the staged copy oracle/_ref/reference
(oracle/stage_reference.py).

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
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stage_reference  # noqa: E402

REF = stage_reference.reference_root()

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


def run_reference_main(patched: bool = False, eps_replay=None, timing: dict | None = None, threads: int = 8):
    """patched: replace the hot-path symbols of the driver's namespace by this package (GPU).
    eps_replay [frames, 256]: noise handed to the patched CVAE instead of fresh draws.
    timing: dict that receives perf_counter stamps of every BallTree.query call (loop timing)."""
    import time
    from mocha_sigasia2023_b200 import synthetic
    if REF is None:
        raise RuntimeError("reference tree unavailable: neither /root/reference nor oracle/_ref/reference exists")
    d = _scratch()
    os.chdir(d)
    for p in ("", "etc", "motion", "preprocess", "net"):
        sys.path.insert(0, os.path.join(d, p))
    viz = types.ModuleType("viz_motion")
    viz.animation_plot = lambda *a, **k: None
    sys.modules["viz_motion"] = viz

    rec = {"eps": [], "match": [], "cvae_cond": [], "cvae_out": [], "Ytil": [], "saves": []}

    # per-stage perf_counter wrappers, active in timing mode only: call-through, nothing edited
    stages = timing.setdefault("stages", {}) if timing is not None else None

    def staged(name, fn):
        if stages is None:
            return fn

        def wrapped(*a, **k):
            t0 = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                acc = stages.setdefault(name, [0.0, 0])
                acc[0] += time.perf_counter() - t0
                acc[1] += 1
        return wrapped

    real_randn_like = torch.randn_like

    def spy_randn_like(t, *a, **k):
        out = real_randn_like(t, *a, **k)
        rec["eps"].append(out.detach().cpu().clone().numpy())
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

    if patched:
        import mocha_sigasia2023_b200 as pkg
        from mocha_sigasia2023_b200 import Inertialization as our_inert, quat as our_quat
        from mocha_sigasia2023_b200.balltree import BallTree as OurTree
        from mocha_sigasia2023_b200.model_CVAE import CVAE as OurCVAE
        from mocha_sigasia2023_b200.trainer import Trainer as OurTrainer
        from mocha_sigasia2023_b200.transformer import mean_variance_norm as our_mvn
        replay = {"i": 0}

        tf.BallTree, tf.quat, tf.inert = OurTree, our_quat, our_inert
        tf.Trainer, tf.CVAE, tf.mean_variance_norm = OurTrainer, OurCVAE, our_mvn
        rec["patched_package"] = pkg.__name__
        _orig_cvae_init = OurCVAE.__init__

        def _init_with_replay(self, *a, **k):
            _orig_cvae_init(self, *a, **k)
            if eps_replay is not None:
                def eps_fn(shape):
                    e = torch.as_tensor(eps_replay[replay["i"]], dtype=torch.float32).reshape(shape)
                    replay["i"] += 1
                    rec["eps"].append(e.numpy().copy())
                    return e
                self.eps_fn = eps_fn

        OurCVAE.__init__ = _init_with_replay
        # the reference picks its device from torch.cuda.is_available(): make sure the patched run is on the GPU
        assert torch.cuda.is_available(), "the patched (Level-1) run needs a GPU: the package has no CPU fallback"
        tf.device = torch.device("cuda")

    RealTree = tf.BallTree

    class SpyTree:
        def __init__(self, X, *a, **k):
            self._t = RealTree(X, *a, **k)
            rec["db_shape"] = np.array(X.shape)

        def query(self, q, *a, **k):
            if timing is not None:
                timing.setdefault("query_t", []).append(time.perf_counter())
            r = staged("BallTree.query", self._t.query)(q, *a, **k)
            rec["match"].append(np.array(r).reshape(-1)[0])
            return r

    tf.BallTree = SpyTree

    real_sample = tf.CVAE.sample

    def spy_sample(self, c, deterministic=False):
        out = staged("CVAE.sample", real_sample)(self, c, deterministic)
        if len(rec["cvae_out"]) < 6:
            rec["cvae_cond"].append(c.detach().cpu().clone().numpy())
            rec["cvae_out"].append(out.detach().cpu().clone().numpy())
        return out

    tf.CVAE.sample = spy_sample

    real_trainer_init = tf.Trainer.__init__

    def spy_trainer_init(self, cfg):
        real_trainer_init(self, cfg)
        real_to_mot = self.gen_ema.to_mot.forward

        def spy_to_mot(x):
            y = staged("Generator.to_mot", real_to_mot)(x)
            rec["Ytil"].append(y.detach().cpu().clone().numpy()[0, -1])     # last frame row only (24,15)
            return y

        self.gen_ema.to_mot.forward = spy_to_mot
        self.gen_ema.decoder.forward = staged("Generator.decoder (Transformer.forward)", self.gen_ema.decoder.forward)
        self.gen_ema.encoder.forward = staged("Generator.encoder (Transformer.forward)", self.gen_ema.encoder.forward)
        self.gen_ema.mot_embedding.forward = staged("Generator.mot_embedding", self.gen_ema.mot_embedding.forward)

    tf.Trainer.__init__ = spy_trainer_init
    if stages is not None and not patched:
        for name in ("fk_partial", "ik_two_bone", "from_xform_xy", "fk", "fk_vel"):
            if hasattr(tf.quat, name):
                setattr(tf.quat, name, staged("quat." + name, getattr(tf.quat, name)))
        if hasattr(tf.inert, "contact_update"):
            tf.inert.contact_update = staged("inert.contact_update", tf.inert.contact_update)
    torch.set_num_threads(threads)
    if timing is not None:
        timing["t_main0"] = time.perf_counter()
    tf.main()
    if timing is not None:
        timing["t_main1"] = time.perf_counter()
        timing["threads"] = torch.get_num_threads()
    torch.randn_like = real_randn_like
    tf.Trainer.__init__ = real_trainer_init
    if patched:
        OurCVAE.__init__ = _orig_cvae_init
    return rec


def generate(out_path, **kw):
    rec = run_reference_main(**kw)
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


def time_reference_loop(threads: int) -> dict:
    """frames/s of the reference's own per-frame loop (test_fullframework.py:438-641) on the host cores: the
    loop starts every frame with `tree_cnt.query(...)` (:443), so the interval between the first and the last
    in-loop query call covers frames 1..223 exactly, set-up and BVH export excluded. Nothing is modified."""
    timing = {}
    run_reference_main(timing=timing, threads=threads)
    q = timing["query_t"]            # q[0]: frame-0 initialisation (:296); q[1:]: one per loop iteration
    frames = len(q) - 2
    loop_s = q[-1] - q[1]
    per_frame = np.diff(np.asarray(q[1:]))
    stages = {k: {"calls": c, "total_ms": 1e3 * t, "ms_per_call": 1e3 * t / max(c, 1)} for k, (t, c) in timing["stages"].items()}
    return {"frames": frames, "loop_s": loop_s, "frames_per_s": frames / loop_s, "ms_per_frame": 1e3 * loop_s / frames,
            "p50_ms_per_frame": float(np.median(per_frame) * 1e3), "p99_ms_per_frame": float(np.percentile(per_frame, 99) * 1e3),
            "main_s": timing["t_main1"] - timing["t_main0"], "threads": timing["threads"], "stages": stages}


if __name__ == "__main__":
    import argparse
    import json
    ap = argparse.ArgumentParser()
    ap.add_argument("--patched", action="store_true")
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--threads", type=int, default=8)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if a.time:
        res = time_reference_loop(a.threads)
        with open(a.out, "w") as f:
            json.dump(res, f)
    elif a.patched:
        g = np.load(os.path.join(ROOT, "tests", "golden", "e2e.npz"))
        generate(a.out, patched=True, eps_replay=g["eps"])
    else:
        generate(a.out or os.path.join(ROOT, "tests", "golden", "e2e.npz"))
