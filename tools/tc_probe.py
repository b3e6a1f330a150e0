"""Time tcgen05 linear layers of given shapes with CUDA events (micro-benchmark for kernel work)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib

lib = _lib.load()
shapes = [(184320, 256, 192), (184320, 256, 1280), (11520, 1536, 256), (11520, 256, 256), (11520, 256, 512),
          (11520, 512, 256), (23296, 768, 256), (90, 256, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, N, K) in shapes:
    A = torch.randn((M, K), device="cuda")
    W = torch.randn((N, K), device="cuda") / K ** 0.5
    W16 = W.to(torch.bfloat16)
    b = torch.randn((N,), device="cuda")
    r = torch.randn((M, N), device="cuda")
    out = torch.empty((M, N), device="cuda")
    _lib.check(lib.mocha_register_bf16_blob(_lib.ptr(W), _lib.ptr(W16), W.numel()))
    ws = torch.empty(lib.mocha_linear_workspace_bytes(M, N, K, 1) + 1024, dtype=torch.uint8, device="cuda")
    for name, bias, res in (("plain", None, None), ("bias+res", b, r)):
        ts = []
        for i in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.mocha_linear(_lib.ptr(A), _lib.ptr(W), _lib.ptr(bias), _lib.ptr(res), _lib.ptr(out), M, N, K, 0, 1,
                                        _lib.ptr(ws), ws.numel(), _lib.stream_ptr()))
            e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1))
        ms = sum(ts) / len(ts)
        print(f"M={M} N={N} K={K} {name:9s}: {ms*1e3:8.1f} us (cast+gemm)  {2.0*M*N*K/ms/1e9:8.1f} TFLOP/s", flush=True)
    ref = (A.to(torch.bfloat16).float() @ W16.float().T) + b + r
    print("   max rel err", float((out - ref).abs().max() / ref.abs().max()))
