"""Launches the hot kernels in isolation at the BASELINE shapes (for `ncu --set full`): FFN-up GEMM (+GELU epilogue),
FFN-down GEMM (gated residual), dX GEMM (MN-major B), self / cross attention fwd + bwd, norm+modulate, fused q|k|v norm+rope."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
R, D, H, S = 2688, 2048, 32, 2688
rnd = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).bfloat16()  # noqa: E731
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 3
x, W1, b1 = rnd(R, D), rnd(4 * D, D, sc=0.02), rnd(4 * D, sc=0.02)
f, pre = torch.empty(R, 4 * D, device=dev, dtype=torch.bfloat16), torch.empty(R, 4 * D, device=dev, dtype=torch.bfloat16)
W2, b2, res = rnd(D, 4 * D, sc=0.02), rnd(D, sc=0.02), rnd(R, D)
tab, temb = rnd(6, D, sc=0.3), rnd(1, 6 * D, sc=0.3)
h = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
q, k, v = rnd(1, H, S, 64), rnd(1, H, S, 64), rnd(1, H, S, 64)
ao = torch.empty(1, S, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(1, H, S, device=dev)
dout = rnd(1, S, D)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
delta = torch.empty(ops.attn_bwd_ws_floats(1, H, S, S), device=dev)
qkv = rnd(R, 3 * D)
cos = torch.randn(S, D // 2, device=dev)
sin = torch.randn(S, D // 2, device=dev)
L = 128
kc, vc = rnd(1, H, L, 64), rnd(1, H, L, 64)
dkc, dvc = torch.empty_like(kc), torch.empty_like(vc)
biasc = torch.zeros(1, L, device=dev); biasc[:, 77:] = -10000.0
deltac = torch.zeros(ops.attn_bwd_ws_floats(1, H, S, L), device=dev)
q2, k2, v2 = torch.empty_like(q), torch.empty_like(q), torch.empty_like(q)
dqkv = torch.empty_like(qkv)
for _ in range(n_rep):
    ops.gemm(x, W1, f, M=R, N=4 * D, K=D, bias=b1, epi=ops.EPI_GELU, out2=pre)                       # FFN up
    ops.gemm(f, W2, h, M=R, N=D, K=4 * D, bias=b2, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[5],
             gate_temb=temb[:, 5 * D:], temb_stride=6 * D, rows_per_sample=S)                         # FFN down
    ops.gemm(res, W2, f, M=R, N=4 * D, K=D, b_mn=True, epi=ops.EPI_MUL_DGELU, aux=pre)               # dX (MN-major B)
    ops.attn_fwd(q, k, v, None, ao, lse, 1, H, S, S, 0.125)
    ops.attn_bwd(q, k, v, None, ao, dout, lse, delta, dq, dk, dv, 1, H, S, S, 0.125)
    ops.attn_fwd(q, kc, vc, biasc, ao, lse, 1, H, S, L, 0.125)                                       # cross attention
    ops.attn_bwd(q, kc, vc, biasc, ao, dout, lse, deltac, dq, dkc, dvc, 1, H, S, L, 0.125)
    ops.norm_modulate_fwd(x, h, tab[0], temb, tab[1], temb[:, D:], 6 * D, R, D, S, 1e-6)
    ops.qkv_norm_rope_fwd(qkv, 3 * D, 0, (tab[0], tab[1], None), 0b011, cos, sin, (q2, k2, v2), 1, S, H, 1e-5)
    ops.qkv_norm_rope_bwd((q2, k2, v2), qkv, 3 * D, 0, (tab[0], tab[1], None), 0b011, cos, sin, dqkv, 3 * D, 0, 1, S, H, 1e-5)
torch.cuda.synchronize()
print("done")
