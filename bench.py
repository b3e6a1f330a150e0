#!/usr/bin/env python
"""bench.py — characterised frames/s of MOCHA's per-frame hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips C] [--precision bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU port on the host cores

A step = one pass of the hot path over one batch: every clip on this GPU advances one frame
(encode its new 60-frame window -> context feature -> nearest-neighbour match -> CVAE sample ->
AdaIN decode -> to_mot -> root integration / blending / foot-lock IK). Clips are independent, so
N GPUs run N x C clips with no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "characterized_frames_per_s"
UNIT = "frames/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=128, help="clips per GPU (config 4: 1024 clips / 8 GPUs)")
    ap.add_argument("--db-rows", type=int, default=385, help="character DB rows (400-frame character clip)")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--lanes", type=int, default=1, help="sub-batch lanes: the clips of a GPU are cut into this many groups "
                    "whose frames run concurrently on separate streams inside the captured graph (measured on B200 at 128 "
                    "clips: 1.37 / 1.49 / 1.63 / 1.83 ms per step for 1 / 2 / 3 / 4 lanes - one lane is fastest)")
    ap.add_argument("--latency-frames", type=int, default=200)
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args):
    return {
        "workload": "batched characterization, clips sharded data-parallel (BASELINE config 4 per-GPU share): "
                    f"{args.clips} clips/GPU advance one frame per step, 60x24x15 pose window per clip",
        "clips_per_gpu": args.clips,
        "window": [60, 24, 15],
        "db_rows": args.db_rows,
        "feature_dim": 23040,
        "cvae": "stochastic (eps injected)",
        "second_decode_cm_trans": False,
        "l2": "per-step working set (activations + weights + DB) exceeds the 126 MB L2; no explicit flush",
        "parallelism": f"dp{args.gpus} over clips, no collective on the data path",
    }


# -------------------------------------------------------------------------------------------------
# clocks
# -------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout_s=3.0):
        """nvidia-smi needs ~0.1-0.5 s to emit its first line; block until it has, then mark the offset so
        that only samples taken after this point (the timed region) are reported."""
        self.skip = 0
        if self.p is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout_s:
            try:
                if os.path.getsize(self.f.name) > 0:
                    break
            except OSError:
                pass
            time.sleep(0.02)
        try:
            self.skip = os.path.getsize(self.f.name)
        except OSError:
            self.skip = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.03)  # let the sample in flight land
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(getattr(self, "skip", 0))
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(power))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm's NumPy port (oracle/) on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_port_rate(args, clips, steps, warmup, db_rows=None):
    """frames/s of the oracle port for `clips` clips advancing `steps` frames (after `warmup`)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    from mocha_oracle import nets
    from mocha_oracle.pipeline import OraclePipeline
    from mocha_sigasia2023_b200 import skeleton, synthetic, weights, workload
    gen_sd = {k: v.numpy() for k, v in weights.generator_state_dict(1777).items()}
    cvae_sd = {k: v.numpy() for k, v in weights.cvae_state_dict(1778).items()}
    stats = workload.stats_as_dict(workload.driver_stats())
    n_db = db_rows or args.db_rows
    # character DB through the port's own encoder (set-up, untimed)
    encs, nms = [], []
    cha_X = workload.pose_windows(n_db, 5000)
    for s in range(0, n_db, 32):
        tok = nets.mot_embedding(gen_sd, cha_X[s:s + 32]) + gen_sd["pos_emb"][:, :90]
        enc = nets.encoder(gen_sd, tok)
        cnt = np.transpose(nets.mean_variance_norm(np.transpose(enc, (0, 2, 1))), (0, 2, 1))
        encs.append(enc)
        nms.append(((cnt - stats["cnt_mean"][None]) / stats["cnt_std"][None]).reshape(enc.shape[0], -1))
    pipe = OraclePipeline(gen_sd, cvae_sd, stats, np.concatenate(encs), np.concatenate(nms), clips,
                          skeleton.BONE_PARENTS)
    times = []
    for f in range(warmup + steps):
        inp = workload.step_inputs(clips, seed=f)
        t0 = time.perf_counter()
        pipe.step(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
        dt = time.perf_counter() - t0
        if f >= warmup:
            times.append(dt)
    total = sum(times)
    return clips * len(times) / total, total / len(times)


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy  # noqa: F401  (loads BLAS so threadpoolctl sees it)
    sample_clips = 4
    steps = max(1, min(args.steps, 6))
    warmup = max(1, min(args.warmup, 2))
    rate, sec_per_step = cpu_port_rate(args, sample_clips, steps, warmup)
    cores = host_threads()
    cfg = workload_config(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_clips} clips x {steps} frames of the same per-frame path "
                                   f"(NumPy port of the reference, oracle/mocha_oracle); a full step is "
                                   f"{args.clips} clips, rate is per frame so it is independent of the sample size"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
# B200 arm
# -------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mocha_sigasia2023_b200 import _lib, workload

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the single JSON line: with NCCL_DEBUG set (VERSION / WARN / INFO) NCCL prints its
        # version banner to stdout, so the variable is dropped for this process and anything else goes to a file
        os.environ.pop("NCCL_DEBUG", None)
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/mocha_bench_nccl_%h_%p.log")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    _lib.check(lib.mocha_check_device(), "mocha_check_device")
    dev = torch.device("cuda", local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.clips, args.steps, max(args.warmup, 3)
    sess, *_ = workload.build_session(B, n_db=args.db_rows, precision=args.precision, device=dev,
                                      seed=rank, match_tensor_cores=None, lanes=args.lanes)
    P = 4  # distinct input sets cycled through
    host_pool = [workload.step_inputs(B, seed=1000 * rank + i) for i in range(P)]
    dev_pool = []
    for h in host_pool:
        side = np.concatenate([h["src_hips_vel"].reshape(B, -1), h["src_rvel"], h["src_rang"]], axis=1)
        dev_pool.append({"X": torch.from_numpy(h["X"]).to(dev), "side": torch.from_numpy(side).to(dev),
                         "contacts": torch.from_numpy(h["contacts"]).to(dev), "eps": torch.from_numpy(h["eps"]).to(dev)})

    def load_dev(i):
        d = dev_pool[i % P]
        sess.X.copy_(d["X"]); sess.side.copy_(d["side"]); sess.contacts.copy_(d["contacts"]); sess.eps.copy_(d["eps"])

    # init frame + one eager steady-state frame to count launches, then capture the CUDA graph
    load_dev(0); sess.step_device()
    lib.mocha_reset_launch_count()
    load_dev(1); sess.step_device()
    launches_per_step = int(lib.mocha_launch_count())
    sess.capture()
    for i in range(W):
        load_dev(i); sess.step_device()

    # ---- device-resident throughput ("value") ----
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        sampler.wait_first_sample()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        load_dev(i); sess.step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * B * K / (ms_total * 1e-3)

    # ---- end to end through the public API with HOST buffers: every step copies its inputs from pinned
    # host memory, runs the frame and copies the pose structs back; the H2D of step i+1 overlaps step i ----
    pinned_pool = []
    for h in host_pool:
        pin = sess.pinned_inputs()
        pin["X"].numpy()[...] = h["X"]
        pin["side"].numpy()[...] = np.concatenate([h["src_hips_vel"].reshape(B, -1), h["src_rvel"], h["src_rang"]], axis=1)
        pin["contacts"].numpy()[...] = h["contacts"]
        pin["eps"].numpy()[...] = h["eps"]
        pinned_pool.append(pin)
    for i in range(3):
        sess.collect(sess.submit(pinned_pool[i % P]))
    barrier()
    checksum = 0.0
    e0.record()
    prev = None
    for i in range(K):
        t = sess.submit(pinned_pool[i % P])
        if prev is not None:
            checksum += float(sess.collect(prev)["ik_pos"][0, 0, 0])
        prev = t
    checksum += float(sess.collect(prev)["ik_pos"][0, 0, 0])
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t.item()) * 1e-3)
    assert checksum == checksum, "end-to-end outputs contain NaN"

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": sess.h2d_bytes(),
                "d2h_bytes_per_step": sess.d2h_bytes()},
        "gpu_launches": launches_per_step * K,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clocks,
    }

    if rank == 0:
        line["roofline"] = roofline_pass(args, sess, torch, lib, _lib)
        if not args.no_latency:
            line["latency_batch1"] = latency_pass(args, dev, torch, workload)
            line["latency_batch1_bf16"] = latency_pass(args, dev, torch, workload, precision="bf16")
        if world == 1 and not args.no_cpu_baseline:
            rate, _ = cpu_port_rate(args, 2, 3, 1)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": host_threads(), "kind": "port",
                                    "sample": "2 clips x 3 frames of the same per-frame path on the host "
                                              "(NumPy port of the reference, oracle/mocha_oracle)"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def roofline_pass(args, sess, torch, lib, _lib):
    """Dominant kernel of the step, timed alone with CUDA events on its stream: the reflect-padded
    temporal convolution of mot_embedding's JointBlock (5 taps, 256->256 over B*1440 rows = 78 % of
    the embedding FLOPs, 31 % of the step's), i.e. the tcgen05 GEMM kernel the step launches ~75 times
    with smaller shapes. FLOPs are algorithmic: 2 * rows * (5*256) * 256."""
    import ctypes as C
    peaks, src = load_peaks()
    B = sess.B
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=sess.dev)
    fn = getattr(lib, "mocha_bench_tconv", None)
    if fn is None:
        return None
    rows = B * 60 * 24
    flops = 2.0 * rows * 1280 * 256
    x = torch.randn((rows, 256), device=sess.dev)
    out = torch.empty((rows, 256), device=sess.dev)
    wp, wn = _lib.ptr(sess.ws), sess.ws.numel()
    # Operands (x 189 MB fp32 -> 106 MB bf16 padded copy, out 189 MB) exceed the 126 MB L2, so launches
    # run back to back without a flush; R launches between two events amortise the launch gap.
    R = 10
    times = []
    for i in range(5):
        flush.zero_()
        # stage once + 1 warm launch, then time R launches of the GEMM kernel alone
        _lib.check(fn(C.byref(sess.gen.struct), _lib.ptr(x), B, _lib.ptr(out), sess.prec, 1, wp, wn, _lib.stream_ptr()),
                   "mocha_bench_tconv")
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        _lib.check(fn(C.byref(sess.gen.struct), _lib.ptr(x), B, _lib.ptr(out), sess.prec, 1, wp, wn, _lib.stream_ptr()),
                   "mocha_bench_tconv")
        e1.record()
        _lib.check(fn(C.byref(sess.gen.struct), _lib.ptr(x), B, _lib.ptr(out), sess.prec, 1 + R, wp, wn, _lib.stream_ptr()),
                   "mocha_bench_tconv")
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            # (stage + (1+R) gemm) - (stage + 1 gemm) = R gemm launches
            times.append((e1.elapsed_time(e2) - e0.elapsed_time(e1)) / R)
    ms = sum(times) / len(times)
    achieved = flops / (ms * 1e-3) / 1e12
    if sess.prec == _lib.MOCHA_BF16:
        peak = peaks.get("bf16_tflops", 1590.0)
        return {"kernel": "tc_gemm2_kernel<LinearEpiT<1>> (cta_group::2, 256x256 pair tiles, TMA-store epilogue): mot_embedding "
                          "JointBlock temporal conv (5 taps, 256->256) as a TMA-shifted implicit GEMM on tcgen05 "
                          "(largest launch of the step, 31 % of its FLOPs)",
                "bound": "tensor", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of this launch at 128 clips from one
                # `ncu --set full` capture (profiles/r01_ncu_tconv_pair_tc_gemm2_LinearEpiT1.txt); algorithmic
                # bytes are 106 MB (padded bf16 operand) + 189 MB (fp32 output of the stand-alone entry)
                "traffic": (101463040 + 135415040) * (B / 128.0), "traffic_unit": "B/launch",
                "peak_source": src + " (burst)",
                "ms_per_launch": ms, "flops_per_launch": flops}
    return {"kernel": "sgemm_kernel<128,128,8,8> (temporal conv as implicit GEMM, fp32 FFMA)", "bound": "fp32-simt",
            "achieved": achieved, "peak": 80.0, "unit": "TFLOP/s", "frac": achieved / 80.0, "traffic": None,
            "peak_source": "nominal fp32 FFMA", "ms_per_launch": ms, "flops_per_launch": flops}


def latency_pass(args, dev, torch, workload, precision="fp32"):
    """BASELINE config 2: batch-1 streaming, one CUDA graph per frame; per-frame latency with CUDA
    events (H2D of the new window + frame + D2H of the pose inside the timed region)."""
    sess, *_ = workload.build_session(1, n_db=args.db_rows, precision=precision, device=dev, seed=99)
    pool = [workload.step_inputs(1, seed=7000 + i) for i in range(8)]
    h = pool[0]
    sess.step_host(h["X"], h["src_hips_vel"], h["src_rvel"], h["src_rang"], h["contacts"], h["eps"])
    h = pool[1]
    sess.step_host(h["X"], h["src_hips_vel"], h["src_rvel"], h["src_rang"], h["contacts"], h["eps"])
    sess.capture()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run(n, do_flush):
        ts = []
        for i in range(n + 10):
            h = pool[i % len(pool)]
            if do_flush:
                flush.zero_()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sess.step_host(h["X"], h["src_hips_vel"], h["src_rvel"], h["src_rang"], h["contacts"], h["eps"])
            e1.record()
            torch.cuda.synchronize()
            if i >= 10:
                ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2], ts[min(len(ts) - 1, int(len(ts) * 0.99))]

    p50, p99 = run(args.latency_frames, True)
    w50, w99 = run(args.latency_frames, False)
    return {"p50_ms": p50, "p99_ms": p99, "p50_ms_l2_warm": w50, "p99_ms_l2_warm": w99, "frames": args.latency_frames,
            "precision": precision, "db_rows": args.db_rows, "budget_ms": 33.3,
            "note": "L2 flushed (256 MB write) before every timed frame for p50_ms; *_l2_warm without flush"}


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
