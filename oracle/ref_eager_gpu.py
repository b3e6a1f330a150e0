"""Same-GPU library baseline: the reference's own `Generator` / `CVAE` modules, UNMODIFIED, moved to
the B200 and run with stock PyTorch eager kernels (cuBLAS / cuDNN / ATen), batched over clips.

This is checker-side infrastructure like the rest of oracle/: bench.py runs it in a subprocess for the reported
`torch_eager_gpu` key ("hand-written kernels vs library kernels" on the same GPU); nothing in the product imports it.

One step = the network portion of one frame of test_fullframework.py's loop (:438-460) for B clips at once:
  mot_embedding + pos_emb + encoder (model.py:43-49) -> mean_variance_norm -> normalised query -> nearest DB row
  (torch.cdist + argmin on the GPU instead of sklearn's BallTree) -> CVAE condition (:446-447) -> CVAE.sample (:448) ->
  de-normalise (:449) -> decoder (:455) -> to_mot (:456) -> de-normalised Y copied to the host (:457).
The reference's NumPy kinematics after that point (FK / IK / inertialization, :462-641) are NOT included, nor is the
second "cm_trans" decode: the number is an upper bound for a GPU port of the reference that keeps its Python loop.

    python oracle/ref_eager_gpu.py --clips 128 --steps 20 --warmup 5 --out x.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import stage_reference  # noqa: E402


def build(dev):
    ref = stage_reference.reference_root()
    if ref is None:
        raise RuntimeError("reference tree unavailable: neither /root/reference nor oracle/_ref/reference exists")
    os.chdir(ref)
    for p in ("", "net", "motion", "etc", "preprocess"):
        sys.path.insert(0, os.path.join(ref, p))
    from model import Generator
    from model_CVAE import CVAE
    from transformer import mean_variance_norm
    from mocha_sigasia2023_b200 import weights
    gen = Generator(weights.DEFAULT_MODEL_CFG)
    gen.load_state_dict(weights.generator_state_dict(1777), strict=True)
    cvae = CVAE(output_seq=90)
    cvae.load_state_dict(weights.cvae_state_dict(1778), strict=True)
    return gen.to(dev).eval(), cvae.to(dev).eval(), mean_variance_norm


def run(clips: int, steps: int, warmup: int, db_rows: int, mode: str, device: str = "cuda") -> dict:
    import time
    from mocha_sigasia2023_b200 import workload
    dev = torch.device(device)
    cuda = dev.type == "cuda"
    torch.set_grad_enabled(False)
    torch.backends.cuda.matmul.allow_tf32 = mode == "tf32"
    torch.backends.cudnn.allow_tf32 = mode == "tf32"
    gen, cvae, mvn = build(dev)
    st = workload.stats_as_dict(workload.driver_stats())
    f32 = dict(dtype=torch.float32, device=dev)
    cnt_mean, cnt_std = torch.as_tensor(st["cnt_mean"], **f32), torch.as_tensor(st["cnt_std"], **f32)
    m0, s0 = torch.as_tensor(st["src_cnt_mean"], **f32), torch.as_tensor(st["src_cnt_std"], **f32)
    m1, s1 = torch.as_tensor(st["cha_encoded_mean"], **f32), torch.as_tensor(st["cha_encoded_std"], **f32)
    Y_mean, Y_std = torch.as_tensor(st["Y_mean"], **f32), torch.as_tensor(st["Y_std"], **f32)
    n = 90

    def encode(X):
        tok = gen.mot_embedding(X)
        enc = gen.encoder(tok + gen.pos_emb[:, :tok.shape[1]])
        cnt = mvn(enc.permute(0, 2, 1)).permute(0, 2, 1)
        return enc, cnt

    # character DB: encoded windows + normalised context rows, built with the same modules
    wins = torch.as_tensor(workload.pose_windows(db_rows, 5000), **f32)
    cha_enc, cha_cnt = [], []
    for i in range(0, db_rows, 64):
        e, c = encode(wins[i:i + 64])
        cha_enc.append(e)
        cha_cnt.append(c)
    cha_enc, cha_cnt = torch.cat(cha_enc), torch.cat(cha_cnt)
    db = ((cha_cnt - cnt_mean[None]) / cnt_std[None]).reshape(db_rows, -1).contiguous()
    prev = cha_enc[:1].expand(clips, -1, -1).contiguous()

    pool = [torch.as_tensor(workload.step_inputs(clips, seed=f)["X"]) for f in range(4)]
    y_host = torch.empty((clips, 60, 24, 15), dtype=torch.float32)
    if cuda:
        pool, y_host = [p.pin_memory() for p in pool], y_host.pin_memory()
    autocast = torch.autocast(dev.type, dtype=torch.bfloat16, enabled=mode == "bf16_autocast")

    def step(i):
        nonlocal prev
        X = pool[i % len(pool)].to(dev, non_blocking=True)
        with autocast:
            enc, cnt = encode(X)
            q = ((cnt.float() - cnt_mean[None]) / cnt_std[None]).reshape(clips, -1)
            idx = torch.cdist(q, db).argmin(dim=1)                       # noqa: F841 (the cm_trans decode would use it)
            cond = torch.cat([(cnt.float() - m0[None]) / s0[None], (prev - m1[None]) / s1[None]], dim=1)
            out = cvae.sample(cond, deterministic=False)
            cur = out.float() * s1[None] + m1[None]
            prev = cur
            dec = gen.decoder(enc, cur)
            ytil = gen.to_mot(dec)
            y = ytil.float() * Y_std + Y_mean            # [24, 15] tables broadcast over [B, 60, 24, 15]
        y_host.copy_(y.reshape(y_host.shape), non_blocking=True)

    for i in range(warmup):
        step(i)
    if cuda:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
    else:   # CPU dry run of the script itself (build container)
        t0 = time.perf_counter()
        for i in range(steps):
            step(i)
        ms = (time.perf_counter() - t0) * 1e3 / steps
    return {"mode": mode, "clips": clips, "steps": steps, "ms_per_step": ms, "value": clips / ms * 1e3, "unit": "frames/s",
            "db_rows": db_rows, "torch": torch.__version__}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=128)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--db-rows", type=int, default=385)
    ap.add_argument("--modes", default="fp32,tf32,bf16_autocast")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    res = []
    for m in a.modes.split(","):
        try:
            res.append(run(a.clips, a.steps, a.warmup, a.db_rows, m, a.device))
        except Exception as e:  # noqa: BLE001 - a mode the reference modules cannot run is reported, not fatal
            res.append({"mode": m, "error": f"{type(e).__name__}: {e}"[:300]})
    with open(a.out, "w") as f:
        json.dump(res, f)
