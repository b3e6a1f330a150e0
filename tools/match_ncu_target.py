"""One matcher query for an ncu capture: python tools/match_ncu_target.py [rows] [queries]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200.balltree import BallTree
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
D = 23040
g = torch.Generator(device="cuda").manual_seed(1)
db16 = torch.empty((N, D), dtype=torch.bfloat16, device="cuda")
for s in range(0, N, 65536):
    db16[s:s + 65536] = torch.randn((min(65536, N - s), D), generator=g, device="cuda").to(torch.bfloat16)
q = torch.randn((Q, D), generator=g, device="cuda")
# bf16-only storage: the tree API wants fp32 rows, so call the ABI directly
from mocha_sigasia2023_b200 import _lib
lib = _lib.load()
norm = (db16[:1].float() ** 2).sum(1).expand(N).contiguous() if False else torch.empty((N,), device="cuda")
for s in range(0, N, 65536):
    r = db16[s:s + 65536].float()
    norm[s:s + 65536] = (r * r).sum(1)
q16 = q.to(torch.bfloat16)
ws = torch.empty(lib.mocha_match_tc_workspace_bytes(Q, N, D, 8) + 1024, dtype=torch.uint8, device="cuda")
idx = torch.empty((Q, 2), dtype=torch.int64, device="cuda")
dd = torch.empty((Q, 2), dtype=torch.float64, device="cuda")
for _ in range(2):
    _lib.check(lib.mocha_match_tc(_lib.ptr(q), _lib.ptr(q16), Q, _lib.ptr(db16), None, _lib.ptr(norm), N, D, 2, 8, 0, _lib.ptr(idx),
                                  _lib.ptr(dd), _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
torch.cuda.synchronize()
print("done", int(idx[0, 0]))
