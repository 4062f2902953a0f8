"""Correctness + timing sweep of the GEMM tilings at the step's shapes (M = 2688 tokens): one CTA per 128 x bn tile
versus CTA pairs sharing a 256 x bn tile (tcgen05 cta_group::2), bn in {256, 192, 160, 128}, plus the library's own
automatic choice.  Each line: max error relative to the fp32 torch result's max magnitude and the mean device time of 20
back-to-back launches with the activations rotated over 3 buffer sets.     python tools/gemm_variants.py [--quick]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops  # noqa: E402


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device="cuda") * scale).bfloat16()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


def timeit(fn, n=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    quick = "--quick" in sys.argv
    torch.manual_seed(0)
    M = 2688
    # name, N, K, b_mn, epilogue kind, K2 (LoRA extension), a2 groups
    shapes = [("ffn_up   N=8192 K=2048 gelu+pre", 8192, 2048, False, "gelu", 0, 0),
              ("qkv      N=6144 K=2048 +lora3", 6144, 2048, False, "plain", 64, 3),
              ("out_proj N=2048 K=2048 gate+lora", 2048, 2048, False, "gate", 64, 1),
              ("ffn_down N=2048 K=8192 gate", 2048, 8192, False, "gate", 0, 0),
              ("dx_w2    N=8192 K=2048 b_mn dgelu", 8192, 2048, True, "dgelu", 0, 0),
              ("dx_w1    N=2048 K=8192 b_mn", 2048, 8192, True, "plain", 0, 0),
              ("dx_qkv   N=2048 K=6144 b_mn +lora", 2048, 6144, True, "plain", 192, 1),
              ("dx_o     N=2048 K=2048 b_mn +lora", 2048, 2048, True, "plain", 64, 1)]
    if quick:
        shapes = shapes[:1] + shapes[3:4]
    bad = 0
    for name, N, K, b_mn, kind, K2, groups in shapes:
        A = [rnd(M, K) for _ in range(3)]
        Bm = rnd(K, N, scale=0.05) if b_mn else rnd(N, K, scale=0.05)
        bias = rnd(N)
        ref = A[0].float() @ (Bm.float() if b_mn else Bm.float().t())
        kw = dict(M=M, N=N, K=K, b_mn=b_mn)
        if K2:
            u = rnd(M, max(1, groups) * K2, scale=0.3)
            Bl = rnd(K2, N, scale=0.05) if b_mn else rnd(N, K2, scale=0.05)
            kw.update(A2=u, B2=Bl, K2=K2, a2_group_n=(N // groups if groups > 1 else 0))
            gN = N // max(1, groups)
            for j in range(max(1, groups)):
                uj = u[:, j * K2:(j + 1) * K2].float()
                ref[:, j * gN:(j + 1) * gN] += uj @ (Bl[:, j * gN:(j + 1) * gN].float() if b_mn else Bl[j * gN:(j + 1) * gN].float().t())
        if kind == "gelu":
            kw.update(bias=bias, epi=ops.EPI_GELU)
            want = F.gelu(ref + bias.float(), approximate="tanh")
        elif kind == "gate":
            res, tab, temb = rnd(M, N), rnd(6, N, scale=0.3), rnd(1, 6 * N, scale=0.3)
            kw.update(bias=bias, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[5], gate_temb=temb[:, 5 * N:], temb_stride=6 * N,
                      rows_per_sample=M)
            want = res.float() + (tab[5].float()[None] + temb[:, 5 * N:].float()) * (ref + bias.float())
        elif kind == "dgelu":
            aux = rnd(M, N)
            kw.update(epi=ops.EPI_MUL_DGELU, aux=aux)
            x = aux.float().requires_grad_(True)
            F.gelu(x, approximate="tanh").sum().backward()
            want = ref * x.grad
        else:
            want = ref
        outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
        outs2 = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(3)] if kind == "gelu" else None
        print(name, flush=True)
        configs = [("auto", 0, 0)] + [(f"{'pair' if cp == 2 else '1cta'}-{bn}", bn, cp) for cp in (1, 2) for bn in (256, 192, 160, 128)]
        best = None
        for label, bn, cp in configs:
            if groups > 1 and bn and (N // groups) % bn:
                continue
            def run(i, bn=bn, cp=cp):
                extra = dict(out2=outs2[i % 3]) if outs2 else {}
                ops.gemm(A[i % 3], Bm, outs[i % 3], block_n=bn, cta_pair=cp, **extra, **kw)
            run(0)
            torch.cuda.synchronize()
            e = rel(outs[0], want)
            ok = e < 1e-2
            bad += (not ok)
            us = timeit(run)
            tf = 2.0 * M * N * (K + K2) / us / 1e6
            if ok and label != "auto" and (best is None or us < best[1]):
                best = (label, us)
            print(f"   {label:10s} err={e:.2e} {'OK  ' if ok else 'FAIL'} {us:7.1f} us {tf:7.1f} TFLOP/s", flush=True)
        print(f"   best: {best}", flush=True)
    print("GEMM_VARIANTS_DONE bad=%d" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
