"""In-kernel time line of the tcgen05 GEMM pipeline roles (debug build: `python -m mocha_sigasia2023_b200.build --trace`).

    MOCHA_LIB=mocha_sigasia2023_b200/libmocha_b200_trace.so python tools/tc_trace.py 11520x768x256 ...

Prints, per shape, the median over CTAs of each event (SM cycles since CTA start) and the spread of
CTA start/end on the global timer.
"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib

lib = _lib.load()
lib.mocha_debug_set_trace.restype = C.c_int
lib.mocha_debug_set_trace.argtypes = [C.c_void_p]
shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]] or [(11520, 768, 256)]
trace = torch.zeros((148, 32), dtype=torch.int64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
NAMES = {2: "setup done", 4: "loads issued t0", 5: "loads issued t1", 6: "loads issued t2", 7: "loads issued t3+",
         8: "operands landed t0", 9: "operands landed t1", 10: "operands landed t2", 11: "operands landed t3+",
         12: "mma committed t0", 13: "mma committed t1", 14: "mma committed t2", 15: "mma committed t3+",
         16: "acc ready t0", 17: "acc ready t1", 18: "acc ready t2", 19: "acc ready t3+",
         26: "  epi: 1st tmem ld done", 27: "  epi: 1st chunk done", 28: "  epi: 2nd chunk done", 20: "drained t0", 21: "drained t1", 22: "drained t2", 23: "drained t3+", 24: "cta end"}
for (M, N, K) in shapes:
    A = torch.randn((M, K), device="cuda"); W = torch.randn((N, K), device="cuda") / K ** 0.5
    W16 = W.to(torch.bfloat16); b = torch.randn((N,), device="cuda"); r = torch.randn((M, N), device="cuda")
    out = torch.empty((M, N), device="cuda")
    _lib.check(lib.mocha_register_bf16_blob(_lib.ptr(W), _lib.ptr(W16), W.numel()))
    ws = torch.empty(lib.mocha_linear_workspace_bytes(M, N, K, 1) + 1024, dtype=torch.uint8, device="cuda")
    for name, bias, res in (("plain", None, None), ("bias+res", b, r)):
        for cold in (True, False):
            for it in range(3):
                if cold:
                    flush.zero_()
                trace.zero_()
                assert lib.mocha_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.mocha_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), M, N, K, 0, 1,
                                            _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
                e1.record()
                torch.cuda.synchronize()
            dbg = (C.c_ulonglong * 16)()
            lib.mocha_debug_get_epi(dbg)
            print("   epilogue chunk (CTA 0, warp 2, last chunk) cycles: sts+sync %d | lds %d | bias/act/res %d | stg fp32 %d | bf16 %d | final sync %d"
                  % (dbg[1] - dbg[0], dbg[3] - dbg[2], dbg[4] - dbg[3], dbg[5] - dbg[4], dbg[6] - dbg[5], dbg[7] - dbg[6]))
            t = trace.cpu()
            live = t[:, 1] != 0
            t = t[live]
            g0 = t[:, 0].min()
            print(f"== M={M} N={N} K={K} {name} {'cold' if cold else 'warm'}: cast+gemm {e0.elapsed_time(e1)*1e3:.1f} us, "
                  f"{int(live.sum())} CTAs; CTA start spread {int(t[:,0].max()-g0)} ns, "
                  f"last CTA end {int(t[:,25].max()-g0)} ns after first start")
            for slot in (2, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 26, 27, 28, 20, 17, 21, 18, 22, 19, 23, 24):
                v = t[:, slot]
                ok = v != 0
                if ok.any():
                    d = (v[ok] - t[ok, 1]).float()
                    print(f"   {NAMES[slot]:22s} median {d.median():9.0f}  max {d.max():9.0f} cyc   ({int(ok.sum())} CTAs)")
