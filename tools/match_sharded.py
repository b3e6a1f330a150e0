"""BASELINE config 5: DB-sharded matching. Launch with torchrun (one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29511 \
        tools/match_sharded.py --rows-per-gpu 2000000 --q 4096
Each rank holds rows_per_gpu rows of a bf16 DB (weak scaling in N, see SURVEY §7 footprint note);
queries are replicated; local tcgen05 match + exact re-rank, NCCL all-gather of the [Q,k] lists, CUDA merge."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from mocha_sigasia2023_b200 import _lib
from mocha_sigasia2023_b200.sharded import ShardedMatcher

ap = argparse.ArgumentParser()
ap.add_argument("--rows-per-gpu", type=int, default=2_000_000)
ap.add_argument("--q", type=int, default=4096)
ap.add_argument("--d", type=int, default=23040)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--exchange", default="nccl", choices=["nccl", "peer", "both"],
                help="nccl: all-gather + merge kernel; peer: one fused P2P store/signal/merge kernel; both: run and compare")
a = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = _lib.load()
dev = torch.device("cuda", local)
Nl, Q, D = a.rows_per_gpu, a.q, a.d
N = Nl * world
g = torch.Generator(device=dev).manual_seed(1234 + rank)
db16 = torch.empty((Nl, D), dtype=torch.bfloat16, device=dev)
norm = torch.empty((Nl,), dtype=torch.float32, device=dev)
for s in range(0, Nl, 16384):
    m = min(16384, Nl - s)
    rows = torch.randn((m, D), generator=g, device=dev)
    _lib.check(lib.mocha_db_pack_bf16(_lib.ptr(rows), m, D, _lib.ptr(db16[s:s + m]), _lib.ptr(norm[s:s + m]), _lib.stream_ptr()))
# planted queries: query j is a noisy copy of row (j mod Nl) of rank (j mod world); built on its owner, then shared
gq = torch.Generator(device=dev).manual_seed(99)
owner = torch.arange(Q, device=dev) % world
row = (torch.arange(Q, device=dev) * 7919) % Nl
q = torch.zeros((Q, D), device=dev)
mine = owner == rank
q[mine] = db16[row[mine]].float() + 0.05 * torch.randn((int(mine.sum()), D), generator=gq, device=dev)
if world > 1:
    dist.all_reduce(q)
q16 = q.to(torch.bfloat16)
ws = torch.empty(lib.mocha_match_tc_workspace_bytes(Q, Nl, D, 8) + 1024, dtype=torch.uint8, device=dev)


def local_query(qt, k):
    idx = torch.empty((Q, k), dtype=torch.int64, device=dev)
    dd = torch.empty((Q, k), dtype=torch.float64, device=dev)
    _lib.check(lib.mocha_match_tc(_lib.ptr(qt), _lib.ptr(q16), Q, _lib.ptr(db16), None, _lib.ptr(norm), Nl, D, k, 8, 0,
                                  _lib.ptr(idx), _lib.ptr(dd), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    return dd, idx


modes = ["nccl", "peer"] if a.exchange == "both" else [a.exchange]
results = {}
for mode in modes:
  m = ShardedMatcher(N, local_query, exchange=mode if world > 1 else "nccl")
  times = []
  xt = []
  for it in range(a.iters + 1):
      if world > 1:
          dist.barrier()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      d, i = m.query(q, k=2)
      e1.record()
      torch.cuda.synchronize()
      t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
      if world > 1:
          dist.all_reduce(t, op=dist.ReduceOp.MAX)
      if it > 0:
          times.append(float(t.item()))
  results[mode] = (sum(times) / len(times), d.clone(), i.clone())
  # the exchange alone (local lists already computed): time 20 back-to-back calls
  if world > 1:
      dl, il = local_query(q, 2)
      il = il + m.lo
      torch.cuda.synchronize(); dist.barrier()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(20):
          if mode == "peer":
              m._peer.exchange_merge(dl, il)
          else:
              all_d = torch.empty((world * Q, 2), dtype=torch.float64, device=dev)
              all_i = torch.empty((world * Q, 2), dtype=torch.int64, device=dev)
              dist.all_gather_into_tensor(all_d, dl); dist.all_gather_into_tensor(all_i, il)
              m.merge(all_d.view(world, Q, 2), all_i.view(world, Q, 2), 2)
      e1.record(); torch.cuda.synchronize()
      xt = e0.elapsed_time(e1) / 20 * 1e3
      results[mode] = results[mode] + (xt,)
  if getattr(m, "_peer", None) is not None:
      m._peer.close()
ms, d, i = results[modes[0]][:3]
want = owner * Nl + row      # shard_bounds is contiguous and equal-sized here
found = float((i[:, 0] == want).float().mean())
same = True
if len(modes) == 2:
    same = bool((results["nccl"][2] == results["peer"][2]).all()) and bool((results["nccl"][1] == results["peer"][1]).all())
if rank == 0:
    print(json.dumps({"workload": f"DB-sharded match {Q}q x {N} rows ({Nl}/GPU) x {D} bf16, k=2, exchange={a.exchange}",
                      "exchange_us": {mo: (results[mo][3] if len(results[mo]) > 3 else None) for mo in modes},
                      "ms_by_mode": {mo: results[mo][0] for mo in modes}, "peer_equals_nccl": same,
                      "n_gpus": world, "ms": ms, "tflops_total": 2.0 * Q * N * D / ms / 1e9,
                      "tflops_per_gpu": 2.0 * Q * Nl * D / ms / 1e9, "planted_found": found}))
if world > 1:
    dist.destroy_process_group()
