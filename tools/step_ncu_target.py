"""A few eager (un-captured) frames of the 128-clip bf16 step for ncu captures of individual kernels:
    ncu --set full --clock-control none -k regex:'style_mlp|embed_graph_agg' -c 4 python tools/step_ncu_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mocha_sigasia2023_b200 import workload
B = int(os.environ.get("CLIPS", "128"))
sess, *_ = workload.build_session(B, n_db=385, precision="bf16")
for f in range(4):
    inp = workload.step_inputs(B, seed=f)
    sess.step_host(inp["X"], inp["src_hips_vel"], inp["src_rvel"], inp["src_rang"], inp["contacts"], inp["eps"])
torch.cuda.synchronize()
print("done")
