"""Per-tile phase times of one consumer warp of the attention kernels (needs tools/micro/libb2d_trace.so)."""
import ctypes
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from finetrainers_b200 import lib  # noqa: E402

lib.LIB_PATH = os.path.join(HERE, "libb2d_trace.so")
from finetrainers_b200 import ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)
B, H, S, D = 1, 32, 2688, 2048
rnd = lambda *s: torch.randn(*s, device=dev).bfloat16()  # noqa: E731
q, k, v = rnd(B, H, S, 64), rnd(B, H, S, 64), rnd(B, H, S, 64)
ao = torch.empty(B, S, D, device=dev, dtype=torch.bfloat16)
lse = torch.empty(B, H, S, device=dev)
dout = rnd(B, S, D)
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
delta = torch.empty(ops.attn_bwd_ws_floats(B, H, S, S), device=dev)
for _ in range(3):
    ops.attn_fwd(q, k, v, None, ao, lse, B, H, S, S, 0.125)
    ops.attn_bwd(q, k, v, None, ao, dout, lse, delta, dq, dk, dv, B, H, S, S, 0.125)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 16384)()
assert lib.load().b2d_trace_read(buf) == 0
t = list(buf)


def show(name, base, n, labels):
    print(name)
    rows = []
    for j in range(n):
        r = t[base + j * 8: base + j * 8 + 5]
        if r[0] == 0:
            break
        rows.append(r)
    for j, r in enumerate(rows[:6] + rows[-3:]):
        print("  tile", j if j < 6 else len(rows) - 3 + j - 6, " ".join(f"{labels[i]}={r[i + 1] - r[i]:5d}" for i in range(4)),
              f"total={r[4] - r[0]:5d}", f"gap_to_next={(rows[rows.index(r) + 1][0] - r[4]) if rows.index(r) + 1 < len(rows) else 0:5d}")
    if len(rows) > 4:
        mid = rows[2:-1]
        avg = [sum(r[i + 1] - r[i] for r in mid) / len(mid) for i in range(4)]
        per = (mid[-1][0] - mid[0][0]) / (len(mid) - 1)
        print("  steady-state avg:", " ".join(f"{labels[i]}={avg[i]:7.1f}" for i in range(4)), f" period={per:7.1f} clk/tile")


show("forward softmax warp (CTA (3,5)): per 64-key tile", 0, 42, ["wait_S", "ld", "compute", "st+arrive"])
for nm, base in (("backward dQ pass, warpgroup 0", 4096), ("backward dQ pass, warpgroup 1", 4096 + 1024),
                 ("backward dK/dV pass, warpgroup 0", 4096 + 2048), ("backward dK/dV pass, warpgroup 1", 4096 + 2048 + 1024)):
    show(nm + ": per 64-row tile of that warpgroup", base, 16, ["pre", "wait_S", "compute", "st+arrive"])


def show_mma(name, base, n=56):
    rows = [t[base + i * 4: base + i * 4 + 4] for i in range(n) if t[base + i * 4] != 0]
    if len(rows) < 12:
        return
    mid = rows[4:-4]
    w_ds = sum(r[1] - r[0] for r in mid) / len(mid)
    issue_out = sum(r[2] - r[1] for r in mid) / len(mid)
    w_y = sum(r[3] - r[2] for r in mid) / len(mid)
    per = (mid[-1][0] - mid[0][0]) / (len(mid) - 1)
    print(f"{name}: MMA-issuing thread per tile: wait_dS={w_ds:6.1f} accumulate+commit={issue_out:6.1f} wait_Y={w_y:6.1f} "
          f"issue_SdP+loop={per - w_ds - issue_out - w_y:6.1f}  period={per:6.1f} clk/tile")


show_mma("backward dQ pass", 8192)
show_mma("backward dK/dV pass", 8192 + 2048)
for nm, o in (("dQ", 16000), ("dK/dV", 16008)):
    c0, g0, c1, g1 = t[o:o + 4]
    if g1 > g0:
        print(f"backward {nm} pass, CTA (3,5): {c1 - c0} clk in {(g1 - g0) / 1e3:.1f} us -> SM clock {(c1 - c0) / (g1 - g0) * 1e3:.0f} MHz")
