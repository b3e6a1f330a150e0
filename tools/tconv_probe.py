"""Launch the dominant kernel of the batched path (JointBlock temporal conv on tcgen05) once, for ncu."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib, packing, weights

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pk = packing.PackedGenerator(weights.generator_state_dict(1777), weights.DEFAULT_MODEL_CFG, torch.device("cuda"))
rows = B * 60 * 24
x = torch.randn((rows, 256), device="cuda")
out = torch.empty((rows, 256), device="cuda")
ws = torch.empty(lib.mocha_embed_workspace_bytes(C.byref(pk.struct.dims), B), dtype=torch.uint8, device="cuda")
for _ in range(2):
    _lib.check(lib.mocha_bench_tconv(C.byref(pk.struct), _lib.ptr(x), B, _lib.ptr(out), 1, 1, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok")
