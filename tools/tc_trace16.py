"""Trace build: time line of bf16-in / bf16-out linear layers (the form the bf16 path launches).

    MOCHA_LIB=.../libmocha_b200_trace.so python tools/tc_trace16.py 11520x1024x256[:res] ...
"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib

lib = _lib.load()
lib.mocha_debug_set_trace.restype = C.c_int
lib.mocha_debug_set_trace.argtypes = [C.c_void_p]
lib.mocha_debug_linear_bf16.restype = C.c_int
lib.mocha_debug_linear_bf16.argtypes = [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p]
trace = torch.zeros((148, 32), dtype=torch.int64, device="cuda")
NAMES = {29: "  t1: unit decoded", 30: "  t1: tile_begin done", 2: "setup done", 8: "operands landed t0", 12: "mma committed t0", 16: "acc ready t0", 20: "drained t0",
         9: "operands landed t1", 13: "mma committed t1", 17: "acc ready t1", 21: "drained t1",
         10: "operands landed t2", 14: "mma committed t2", 18: "acc ready t2", 22: "drained t2", 24: "cta end"}
for spec in sys.argv[1:]:
    shape, _, flags = spec.partition(":")
    M, N, K = (int(x) for x in shape.split("x"))
    A = torch.randn((M, K), device="cuda").to(torch.bfloat16)
    W = (torch.randn((N, K), device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda") if "res" in flags else None
    o32 = torch.empty((M, N), device="cuda") if ("res" in flags or "f32" in flags) else None
    o16 = torch.empty((M, N), device="cuda", dtype=torch.bfloat16) if "f32" not in flags else None
    for it in range(4):
        trace.zero_()
        assert lib.mocha_debug_set_trace(C.c_void_p(trace.data_ptr())) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.mocha_debug_linear_bf16(_lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(r), _lib.ptr(o32), _lib.ptr(o16), M, N, K, 0,
                                         _lib.stream_ptr())
        _lib.check(rc)
        e1.record()
        torch.cuda.synchronize()
    t = trace.cpu(); t = t[t[:, 1] != 0]
    g0 = t[:, 0].min()
    print(f"== {spec}: {e0.elapsed_time(e1)*1e3:.1f} us (events), {t.shape[0]} CTAs, last CTA end {int(t[:,25].max()-g0)} ns after first start")
    for slot in (2, 8, 12, 16, 20, 29, 30, 9, 13, 17, 21, 10, 14, 18, 22, 24):
        v = t[:, slot]; ok = v != 0
        if ok.any():
            d = (v[ok] - t[ok, 1]).float()
            print(f"   {NAMES[slot]:22s} median {d.median():9.0f}  max {d.max():9.0f} cyc   ({int(ok.sum())} CTAs)")
    dbg = (C.c_ulonglong * 16)()
    lib.mocha_debug_get_epi(dbg)
    print("   TMA-store chunk (CTA 0, warp 2, last chunk) cycles: bias/act in registers %d | wait for previous store %d | staging stores %d | fence %d | issue %d"
          % (dbg[8] - dbg[14], dbg[9] - dbg[8], dbg[11] - dbg[9], dbg[12] - dbg[11], dbg[13] - dbg[12]))
    ref = A.float() @ W.float().T + b + (r if r is not None else 0)
    out = o16.float() if o16 is not None else o32
    print("   max rel err", float((out - ref).abs().max() / ref.abs().max()))
