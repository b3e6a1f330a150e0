"""Run the bf16 tensor-core encoder (mocha_encoder_fwd, 128 clips) three times - a target for
`ncu -k regex:tc_gemm_kernel --set full --import-source on` source-level views of the GEMM epilogues."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import _lib, packing, weights

lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
pk = packing.PackedGenerator(weights.generator_state_dict(1777), weights.DEFAULT_MODEL_CFG, torch.device("cuda"))
d = pk.struct.dims
n = (d.T // d.tp) * d.P
tokens = torch.randn((B, n, d.D), device="cuda")
enc = torch.empty_like(tokens)
ws = torch.empty(lib.mocha_encoder_workspace_bytes(C.byref(d), B), dtype=torch.uint8, device="cuda")
for _ in range(3):
    _lib.check(lib.mocha_encoder_fwd(C.byref(pk.struct), _lib.ptr(tokens), B, _lib.ptr(enc), _lib.MOCHA_BF16, _lib.ptr(ws), ws.numel(),
                                     _lib.stream_ptr()))
torch.cuda.synchronize()
print("ok")
