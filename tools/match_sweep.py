"""BASELINE config 3: context-matching sweep, Q queries vs an N-row synthetic feature DB (D=23040),
bf16 storage, planted queries. Times the tcgen05 coarse pass + exact re-rank with CUDA events."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1_000_000)
ap.add_argument("--q", type=int, default=4096)
ap.add_argument("--d", type=int, default=23040)
ap.add_argument("--kc", type=int, default=8)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--dist", default="planted", choices=["planted", "iid"])
ap.add_argument("--storage", default="bf16", choices=["bf16", "fp32"])
a = ap.parse_args()
lib = _lib.load()
dev = "cuda"
N, Q, D = a.n, a.q, a.d
g = torch.Generator(device=dev).manual_seed(0)
F32 = a.storage == "fp32"
db16 = torch.empty((N, D), dtype=torch.float32 if F32 else torch.bfloat16, device=dev)
norm = torch.empty((N,), dtype=torch.float32, device=dev)
CH = 16384
t0 = time.time()
for s in range(0, N, CH):
    m = min(CH, N - s)
    if F32:
        db16[s:s + m].normal_(generator=g)
        _lib.check(lib.mocha_db_norms_f32(_lib.ptr(db16[s:s + m]), m, D, _lib.ptr(norm[s:s + m]), _lib.stream_ptr()))
    else:
        rows = torch.randn((m, D), generator=g, device=dev)
        _lib.check(lib.mocha_db_pack_bf16(_lib.ptr(rows), m, D, _lib.ptr(db16[s:s + m]), _lib.ptr(norm[s:s + m]), _lib.stream_ptr()))
torch.cuda.synchronize()
print(f"DB built: {N}x{D} {a.storage} = {db16.numel()*db16.element_size()/1e9:.1f} GB in {time.time()-t0:.1f}s", flush=True)
pick = torch.randint(0, N, (Q,), generator=g, device=dev)
if a.dist == "planted":
    q = db16[pick].float() + 0.05 * torch.randn((Q, D), generator=g, device=dev)
else:
    q = torch.randn((Q, D), generator=g, device=dev)
q = q.contiguous()
q16 = q.to(torch.bfloat16)
idx = torch.empty((Q, 1), dtype=torch.int64, device=dev)
dist = torch.empty((Q, 1), dtype=torch.float64, device=dev)
ws = torch.empty(lib.mocha_match_tc_workspace_bytes(Q, N, D, a.kc) + 1024, dtype=torch.uint8, device=dev)
times = []
for i in range(a.iters + 1):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if F32:
        _lib.check(lib.mocha_match_tc(_lib.ptr(q), None, Q, None, _lib.ptr(db16), _lib.ptr(norm), N, D, 1, a.kc, 0,
                                      _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    else:
        _lib.check(lib.mocha_match_tc(_lib.ptr(q), _lib.ptr(q16), Q, _lib.ptr(db16), None, _lib.ptr(norm), N, D, 1, a.kc, 0,
                                      _lib.ptr(idx), _lib.ptr(dist), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
    e1.record()
    torch.cuda.synchronize()
    if i > 0:
        times.append(e0.elapsed_time(e1))
ms = sum(times) / len(times)
flops = 2.0 * Q * N * D
res = {"workload": f"match sweep {Q}q x {N} x {D} {a.storage} {a.dist}", "ms": ms, "tflops": flops / ms / 1e9,
       "flops": flops}
if a.dist == "planted":
    res["planted_found"] = float((idx[:, 0] == pick).float().mean())
    # exact check of a subsample against fp64 brute force on the stored bf16 rows
    sub = torch.arange(0, Q, max(1, Q // 16), device=dev)[:16]
    best = torch.full((len(sub),), float("inf"), dtype=torch.float64, device=dev)
    besti = torch.zeros((len(sub),), dtype=torch.int64, device=dev)
    for s in range(0, N, 65536):
        blk = db16[s:s + 65536].float()
        d2 = (blk * blk).sum(1)[None, :].double() - 2.0 * (q[sub] @ blk.T).double()
        m, mi = d2.min(dim=1)
        upd = m < best
        best[upd] = m[upd]; besti[upd] = mi[upd] + s
    res["subsample_agree"] = float((besti == idx[sub, 0]).float().mean())
print(json.dumps(res))
