"""attention kernels only, BASELINE shape — for `ncu --set full --import-source on`."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops
dev = "cuda"; torch.manual_seed(0)
H, S, D = 32, 2688, 2048
rnd = lambda *s: torch.randn(*s, device=dev).bfloat16()
q, k, v = rnd(1, H, S, 64), rnd(1, H, S, 64), rnd(1, H, S, 64)
ao = torch.empty(1, S, D, device=dev, dtype=torch.bfloat16); lse = torch.empty(1, H, S, device=dev)
dout = rnd(1, S, D); dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
delta = torch.empty(ops.attn_bwd_ws_floats(1, H, S, S), device=dev)
for _ in range(2):
    ops.attn_fwd(q, k, v, None, ao, lse, 1, H, S, S, 0.125)
    ops.attn_bwd(q, k, v, None, ao, dout, lse, delta, dq, dk, dv, 1, H, S, S, 0.125)
torch.cuda.synchronize(); print("done")
