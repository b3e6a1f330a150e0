"""In-kernel time line of the JointBlock temporal-conv GEMM (trace build, see tools/tc_trace.py)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib, packing, weights

lib = _lib.load()
lib.mocha_debug_set_trace.restype = C.c_int
lib.mocha_debug_set_trace.argtypes = [C.c_void_p]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pk = packing.PackedGenerator(weights.generator_state_dict(1777), weights.DEFAULT_MODEL_CFG, torch.device("cuda"))
rows = B * 60 * 24
x = torch.randn((rows, 256), device="cuda")
out = torch.empty((rows, 256), device="cuda")
ws = torch.empty(lib.mocha_embed_workspace_bytes(C.byref(pk.struct.dims), B), dtype=torch.uint8, device="cuda")
trace = torch.zeros((148, 32), dtype=torch.int64, device="cuda")
mode = int(os.environ.get("TC_DBG_MODE", "0"))
if mode:
    lib.mocha_debug_set_mode.restype = C.c_int
    lib.mocha_debug_set_mode.argtypes = [C.c_int]
    assert lib.mocha_debug_set_mode(mode) == 0
    print("debug mode", mode, "(1: no TMA loads, 2: no MMAs, 3: no epilogue work) - results are garbage by design")
for _ in range(3):
    trace.zero_()
    assert lib.mocha_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
    _lib.check(lib.mocha_bench_tconv(C.byref(pk.struct), _lib.ptr(x), B, _lib.ptr(out), 1, 1, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()))
    torch.cuda.synchronize()
t = trace.cpu()
t = t[t[:, 1] != 0]
names = {2: "setup done", 4: "loads issued t0", 5: "loads issued t1", 6: "loads issued t2", 7: "loads issued last",
         8: "operands landed t0", 9: "operands landed t1", 10: "operands landed t2", 11: "operands landed last",
         12: "mma committed t0", 13: "mma committed t1", 14: "mma committed t2", 15: "mma committed last",
         16: "acc ready t0", 17: "acc ready t1", 18: "acc ready t2", 19: "acc ready last",
         20: "drained t0", 21: "drained t1", 22: "drained t2", 23: "drained last", 24: "cta end"}
print(f"{t.shape[0]} CTAs, kernel span {int(t[:,25].max() - t[:,0].min())} ns")
for slot in (2, 8, 4, 12, 16, 20, 9, 5, 13, 17, 21, 10, 6, 14, 18, 22, 11, 7, 15, 19, 23, 24):
    v = t[:, slot]; ok = v != 0
    if ok.any():
        d = (v[ok] - t[ok, 1]).float()
        print(f"   {names[slot]:22s} median {d.median():9.0f}  max {d.max():9.0f} cyc")
