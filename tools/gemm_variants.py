"""Correctness + timing of the GEMM kernel variants at the step's shapes (run once per variant: the CTA-pair switch is
read once per process).   B2D_GEMM_2CTA=1 python tools/gemm_variants.py [--time]
Prints one line per case: max error relative to the fp32 torch result's max magnitude, and (with --time) the mean
device time of 20 back-to-back launches with operands rotated over 3 buffer sets."""
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
    do_time = "--time" in sys.argv
    torch.manual_seed(0)
    tag = "2cta" if os.environ.get("B2D_GEMM_2CTA") == "1" else "1cta"
    bad = 0
    M = 2688
    # (name, N, K, b_mn, epilogue)
    cases = [("ffn_up  N=8192 K=2048 gelu+pre", 8192, 2048, False, "gelu"),
             ("qkv     N=6144 K=2048 bias+lora", 6144, 2048, False, "lora"),
             ("dx_w2   N=8192 K=2048 b_mn dgelu", 8192, 2048, True, "dgelu"),
             ("plain   N=2048 K=2048 bn256", 2048, 2048, False, "plain"),
             ("ragged  M=300 N=520 K=200 bn256", 520, 200, False, "plain_ragged"),
             ("gate_res N=2048 K=8192 bn256", 2048, 8192, False, "gate"),
             ("batch3  N=512 K=256 b_mn bn256", 512, 256, True, "batch")]
    for name, N, K, b_mn, kind in cases:
        Mi = 300 if kind == "plain_ragged" else M
        A = rnd(Mi, K)
        Bm = rnd(K, N, scale=0.05) if b_mn else rnd(N, K, scale=0.05)
        ref = A.float() @ (Bm.float() if b_mn else Bm.float().t())
        out = torch.zeros(Mi, N, device="cuda", dtype=torch.bfloat16)
        kw = dict(M=Mi, N=N, K=K, b_mn=b_mn, block_n=256)
        errs = []
        if kind == "gelu":
            bias = rnd(N)
            out2 = torch.zeros_like(out)
            ops.gemm(A, Bm, out, bias=bias, epi=ops.EPI_GELU, out2=out2, **kw)
            pre = ref + bias.float()
            errs = [rel(out, F.gelu(pre, approximate="tanh")), rel(out2, pre)]
        elif kind == "lora":
            bias = rnd(N)
            rp = 64
            u = rnd(Mi, 3 * rp, scale=0.3)
            Bl = rnd(N, rp, scale=0.05)
            ops.gemm(A, Bm, out, bias=bias, A2=u, B2=Bl, K2=rp, a2_group_n=N // 3, **kw)
            r2 = ref + bias.float()
            for j in range(3):
                r2[:, j * (N // 3):(j + 1) * (N // 3)] += u[:, j * rp:(j + 1) * rp].float() @ Bl[j * (N // 3):(j + 1) * (N // 3)].float().t()
            errs = [rel(out, r2)]
        elif kind == "dgelu":
            aux = rnd(Mi, N)
            ops.gemm(A, Bm, out, epi=ops.EPI_MUL_DGELU, aux=aux, **kw)
            x = aux.float().requires_grad_(True)
            F.gelu(x, approximate="tanh").sum().backward()
            errs = [rel(out, ref * x.grad)]
        elif kind == "gate":
            res, tab, temb = rnd(Mi, N), rnd(6, N, scale=0.3), rnd(1, 6 * N, scale=0.3)
            bias = rnd(N)
            ops.gemm(A, Bm, out, bias=bias, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[5], gate_temb=temb[:, 5 * N:],
                     temb_stride=6 * N, rows_per_sample=Mi, **kw)
            gate = tab[5].float()[None] + temb[:, 5 * N:].float()
            errs = [rel(out, res.float() + gate * (ref + bias.float()))]
        elif kind == "batch":
            nb = 3
            A3 = rnd(nb * Mi, K)
            B3 = rnd(nb * K, N, scale=0.05)
            o3 = torch.zeros(nb * Mi, N, device="cuda", dtype=torch.bfloat16)
            ops.gemm(A3, B3, o3, M=Mi, N=N, K=K, b_mn=True, block_n=256, batch=nb, a_boff=(Mi, 0), b_boff=(K, 0), c_boff=Mi * N)
            r3 = torch.cat([A3[i * Mi:(i + 1) * Mi].float() @ B3[i * K:(i + 1) * K].float() for i in range(nb)])
            errs = [rel(o3, r3)]
        else:
            ops.gemm(A, Bm, out, **kw)
            errs = [rel(out, ref)]
        torch.cuda.synchronize()
        ok = all(e < 1e-2 for e in errs)
        bad += (not ok)
        line = f"[{tag}] {name:36s} err={['%.2e' % e for e in errs]} {'OK' if ok else 'FAIL'}"
        if do_time and kind in ("gelu", "lora", "dgelu", "plain", "gate"):
            sets = []
            for _ in range(3):
                sets.append((rnd(Mi, K), torch.empty(Mi, N, device="cuda", dtype=torch.bfloat16), torch.empty(Mi, N, device="cuda", dtype=torch.bfloat16)))
            if kind == "gelu":
                f = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], bias=bias, epi=ops.EPI_GELU, out2=sets[i % 3][2], **kw)
            elif kind == "lora":
                f = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], bias=bias, A2=u, B2=Bl, K2=rp, a2_group_n=N // 3, **kw)
            elif kind == "dgelu":
                f = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], epi=ops.EPI_MUL_DGELU, aux=aux, **kw)
            elif kind == "gate":
                f = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], bias=bias, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[5],
                                       gate_temb=temb[:, 5 * N:], temb_stride=6 * N, rows_per_sample=Mi, **kw)
            else:
                f = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], **kw)
            us = timeit(f)
            line += f"  {us:7.1f} us  {2.0 * Mi * N * K / us / 1e6:7.1f} TFLOP/s"
            if kind in ("plain", "gate"):  # also the auto-picked tile for comparison
                kw2 = dict(kw); kw2["block_n"] = 0
                if kind == "plain":
                    f2 = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], **kw2)
                else:
                    f2 = lambda i: ops.gemm(sets[i % 3][0], Bm, sets[i % 3][1], bias=bias, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[5],
                                            gate_temb=temb[:, 5 * N:], temb_stride=6 * N, rows_per_sample=Mi, **kw2)
                us2 = timeit(f2)
                line += f"   (auto tile: {us2:7.1f} us {2.0 * Mi * N * K / us2 / 1e6:7.1f} TFLOP/s)"
        print(line, flush=True)
    print("GEMM_VARIANTS_DONE bad=%d" % bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
