"""In-kernel time line of the fused block-tail kernel (trace build): python -m mocha_sigasia2023_b200.build --trace, then
    MOCHA_LIB=mocha_sigasia2023_b200/libmocha_b200_trace.so python tools/tail_trace.py [M K0 Hd act ln]
Prints SM-cycle stamps (relative to CTA start) of the pipeline roles, median over CTAs."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mocha_sigasia2023_b200 import _lib

M, K0, Hd, act, ln = (int(x) for x in (sys.argv[1:6] + ["11520", "512", "512", "2", "0"][len(sys.argv) - 1:]))
lib = _lib.load()
lib.mocha_debug_set_tail_trace.argtypes = [C.c_void_p]
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
A0, W0 = rn(M, K0).bfloat16(), (rn(256, K0) * K0 ** -0.5).bfloat16()
b0, R0 = rn(256), rn(M, 256)
g1, be1, g2, be2 = rn(256), rn(256), rn(256), rn(256)
W1, b1 = (rn(max(Hd, 1), 256) / 16).bfloat16(), rn(max(Hd, 1))
W2, b2 = (rn(256, max(Hd, 1)) * max(Hd, 1) ** -0.5).bfloat16(), rn(256)
O32, O16 = torch.empty(M, 256, device="cuda"), torch.empty(M, 256, device="cuda", dtype=torch.bfloat16)
nct = (M + 127) // 128
buf = torch.zeros(nct * 64, dtype=torch.int64, device="cuda")
P = _lib.ptr


def run():
    _lib.check(lib.mocha_block_tail(P(A0), K0, K0, P(W0), P(b0), P(R0), P(g1) if ln else None, P(be1) if ln else None, Hd, act,
                                    P(W1), P(b1), P(W2), P(b2), P(g2) if ln else None, P(be2) if ln else None, 1e-5, P(O32), P(O16),
                                    M, _lib.stream_ptr()), "tail")


for _ in range(3):
    run()
torch.cuda.synchronize()
lib.mocha_debug_set_tail_trace(P(buf))
run()
torch.cuda.synchronize()
lib.mocha_debug_set_tail_trace(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    run()
e1.record(); torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(nct, 64).astype(np.int64)
rel = t - t[:, :1]
names = {0: "CTA start", 1: "prologue + pdl_wait done", 2: "producer: prefix loads issued", 4: "MMA: first operands landed",
         5: "MMA: prefix issued", 6: "MMA: X ready", 24: "epi: out-proj accumulator ready", 25: "epi: LN1 stats done",
         26: "epi: stage P done", 40: "epi: FFN accumulator ready", 41: "epi: outputs handed to TMA", 42: "epi: stores drained",
         43: "CTA end"}
for j in range(4):
    names[8 + j] = f"MMA: G1({j}) issued"; names[16 + j] = f"MMA: G2({j}) issued"
    names[28 + 2 * j] = f"epi: hidden chunk {j} ready"; names[29 + 2 * j] = f"epi: hidden chunk {j} stored"
for c in range(2):
    names[56 + 2 * c] = f"   final chunk {c}: accumulator in registers"
    for k, nm in enumerate(("emit enter", "staging free (wait_group)", "boxes written", "proxy fence + syncwarp", "TMA issued")):
        names[44 + 6 * c + k] = f"   final chunk {c}: {nm}"
print(f"M={M} K0={K0} Hd={Hd} act={act} ln={ln}: {nct} CTAs, {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (untraced)")
order = sorted((k for k in names if (t[:, k] > 0).any()), key=lambda k: np.median(rel[:, k]))
for k in order:
    print(f"  {int(np.median(rel[:, k])):8d} cyc  (max {int(rel[:, k].max()):8d})  {names[k]}")
