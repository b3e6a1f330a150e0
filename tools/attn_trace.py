"""In-kernel time line of the fused attention kernel (trace build):
    MOCHA_LIB=mocha_sigasia2023_b200/libmocha_b200_trace.so python tools/attn_trace.py [B H nq nkv dh]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mocha_sigasia2023_b200 import _lib

B, H, nq, nkv, dh = (int(x) for x in (sys.argv[1:6] + ["128", "4", "90", "90", "128"][len(sys.argv) - 1:]))
lib = _lib.load()
lib.mocha_debug_set_attn_trace.argtypes = [C.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
inner = H * dh
q = torch.randn((B * nq, inner), generator=g, device="cuda").bfloat16()
kv = torch.randn((B * nkv, 2 * inner), generator=g, device="cuda").bfloat16()
out = torch.empty((B, nq, inner), device="cuda", dtype=torch.bfloat16)
nct = min(148, B * H * ((nq + 127) // 128))
buf = torch.zeros(nct * 64, dtype=torch.int64, device="cuda")


def run():
    _lib.check(lib.mocha_attention_core(_lib.ptr(q), inner, C.c_void_p(kv.data_ptr()), 2 * inner, C.c_void_p(kv.data_ptr() + inner * 2),
                                        2 * inner, B, H, nq, nkv, dh, _lib.ptr(out), inner, _lib.stream_ptr()), "attn")


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.mocha_debug_set_attn_trace(_lib.ptr(buf))
run(); torch.cuda.synchronize()
lib.mocha_debug_set_attn_trace(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record(); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(nct, 64).astype(np.int64)
t0 = t[:, 14:15]
names = {0: "producer: Q/K buffers free", 1: "producer: V buffer free", 2: "MMA: Q/K landed", 3: "MMA: issuing S", 4: "MMA: P ready",
         5: "MMA: issuing P V", 8: "epi: scores ready", 9: "epi: row max known", 10: "epi: P written", 11: "epi: O accumulator ready",
         12: "epi: output handed to TMA"}
print(f"B={B} H={H} nq={nq} nkv={nkv} dh={dh}: {nct} CTAs, {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (untraced)")
print(f"  prologue done at {int(np.median(t[:, 15] - t[:, 14]))} cyc")
rows = []
for it in range(4):
    for s, nm in names.items():
        col = it * 16 + s
        ok = t[:, col] > 0
        if ok.any():
            rows.append((int(np.median((t[:, col] - t0[:, 0])[ok])), f"unit {it}: {nm}"))
for v, nm in sorted(rows):
    print(f"  {v:8d} cyc  {nm}")
