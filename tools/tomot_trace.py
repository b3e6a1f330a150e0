"""In-kernel time line of the LAST tcgen05 GEMM launch of an eager 128-clip bf16 frame: to_mot's 64 -> 64 temporal
convolution (tc_gemm_kernel<64, LinearEpiT<4>>, 1440 tiles of 128 x 64, K = 5 taps x 64). Trace build:
python -m mocha_sigasia2023_b200.build --trace; MOCHA_LIB=.../libmocha_b200_trace.so python tools/tomot_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib, workload

lib = _lib.load()
lib.mocha_debug_set_trace.restype = C.c_int
lib.mocha_debug_set_trace.argtypes = [C.c_void_p]
B = int(os.environ.get("CLIPS", "128"))
sess, *_ = workload.build_session(B, n_db=385, precision="bf16")
trace = torch.zeros((148, 32), dtype=torch.int64, device="cuda")
for f in range(3):
    inp = workload.step_inputs(B, seed=f)
    trace.zero_()
    assert lib.mocha_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
    sess.step_host(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
    torch.cuda.synchronize()
t = trace.cpu()
t = t[t[:, 1] != 0]
names = {2: "setup done", 4: "loads issued t0", 5: "loads issued t1", 6: "loads issued t2", 7: "loads issued last",
         8: "operands landed t0", 9: "operands landed t1", 10: "operands landed t2", 11: "operands landed last",
         12: "mma committed t0", 13: "mma committed t1", 14: "mma committed t2", 15: "mma committed last",
         16: "acc ready t0", 17: "acc ready t1", 18: "acc ready t2", 19: "acc ready last",
         20: "drained t0", 21: "drained t1", 22: "drained t2", 23: "drained last", 24: "cta end",
         26: "chunk0 loaded t0", 27: "chunk0 done t0", 29: "tile_begin start t1", 30: "tile_begin end t1"}
print(f"{t.shape[0]} CTAs, kernel span {int(t[:,25].max() - t[:,0].min())} ns")
for slot in (2, 8, 4, 12, 16, 26, 27, 20, 29, 30, 9, 5, 13, 17, 21, 10, 6, 14, 18, 22, 11, 7, 15, 19, 23, 24):
    v = t[:, slot]; ok = v != 0
    if ok.any():
        d = (v[ok] - t[ok, 1]).float()
        print(f"   {names[slot]:22s} median {d.median():9.0f}  max {d.max():9.0f} cyc")
