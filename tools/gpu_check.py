"""Diagnostic GPU sweep (not a pytest): runs every libb2d kernel against torch fp32 references and prints one line
per case without stopping at the first failure.  Usage: python tools/gpu_check.py [gemm] [elem] [attn] ..."""
import math
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from finetrainers_b200 import ops  # noqa: E402

dev = "cuda"
RESULTS = []


def report(name, got, ref, tol, extra=""):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    bad = not (err <= tol * scale) or math.isnan(err)
    RESULTS.append((name, not bad))
    print(f"[{'FAIL' if bad else ' ok '}] {name:58s} max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e} {extra}",
          flush=True)
    return not bad


def rnd(*shape, scale=1.0, dtype=torch.bfloat16):
    return (torch.randn(*shape, device=dev) * scale).to(dtype)


def check_gemm():
    torch.manual_seed(0)
    # --- basic K-major x K-major, all tile widths
    for (M, N, K) in [(256, 256, 128), (2688, 2048, 2048), (100, 72, 200), (4, 2048, 256), (128, 12288, 2048)]:
        A = rnd(M, K)
        B = rnd(N, K, scale=0.05)
        ref = A.float() @ B.float().t()
        for bn in (64, 128, 192, 256):
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            try:
                ops.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)
                torch.cuda.synchronize()
                report(f"gemm KK M{M} N{N} K{K} bn{bn}", out, ref, 1e-2)
            except Exception as e:  # noqa
                print("EXC", M, N, K, bn, e)
                RESULTS.append((f"gemm KK {M} {N} {K} {bn}", False))
    # auto tile
    M, N, K = 2688, 6144, 2048
    A = rnd(M, K); B = rnd(N, K, scale=0.05)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, B, out, M=M, N=N, K=K)
    report("gemm KK auto 2688x6144x2048", out, A.float() @ B.float().t(), 1e-2)

    # --- B MN-major (dX = dY W): B given as [K, N]
    for (M, N, K) in [(256, 128, 128), (2688, 2048, 2048), (2688, 2048, 8192), (200, 192, 136)]:
        A = rnd(M, K)
        Bt = rnd(K, N, scale=0.05)
        ref = A.float() @ Bt.float()
        for bn in (64, 128, 160, 256):
            out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
            ops.gemm(A, Bt, out, M=M, N=N, K=K, b_mn=True, block_n=bn)
            report(f"gemm K/MN M{M} N{N} K{K} bn{bn}", out, ref, 1e-2)

    # --- both MN-major (dW = dY^T X), fp32 atomic, split-K, transposed
    for (M, N, K) in [(128, 64, 128), (2048, 64, 2688), (2048, 192, 2688), (304, 64, 1000)]:
        At = rnd(K, M)
        Bt = rnd(K, N, scale=0.05)
        ref = At.float().t() @ Bt.float()
        for splits in (1, 3):
            if splits > (K + 63) // 64 - 1:
                continue
            out = torch.zeros(M, N, device=dev, dtype=torch.float32)
            ops.gemm(At, Bt, out, M=M, N=N, K=K, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC, splits=splits, block_n=64)
            report(f"gemm MN/MN f32atomic M{M} N{N} K{K} splits{splits}", out, ref, 2e-3)
        outT = torch.zeros(N, M, device=dev, dtype=torch.float32)
        ops.gemm(At, Bt, outT, M=M, N=N, K=K, a_mn=True, b_mn=True, epi=ops.EPI_F32_ATOMIC_T, splits=2, block_n=64,
                 alpha=0.5)
        report(f"gemm MN/MN f32atomicT M{M} N{N} K{K}", outT, 0.5 * ref.t(), 2e-3)
    # A MN-major, B K-major
    M, N, K = 256, 128, 192
    At = rnd(K, M); B = rnd(N, K, scale=0.05)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(At, B, out, M=M, N=N, K=K, a_mn=True)
    report("gemm MN/K 256x128x192", out, At.float().t() @ B.float().t(), 1e-2)

    # --- epilogues
    M, N, K = 2688, 2048, 512
    A = rnd(M, K); B = rnd(N, K, scale=0.05); bias = rnd(N)
    acc = A.float() @ B.float().t()
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias)
    report("epi STORE+bias", out, acc + bias.float(), 1e-2)
    out2 = torch.zeros_like(out)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, epi=ops.EPI_GELU, out2=out2)
    pre = acc + bias.float()
    report("epi GELU out", out, F.gelu(pre, approximate="tanh"), 1e-2)
    report("epi GELU out2(pre)", out2, pre, 1e-2)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, epi=ops.EPI_SILU)
    report("epi SILU", out, F.silu(pre), 1e-2)
    # gate-res with per-sample gate (2 samples of 1344 rows)
    res = rnd(M, N); tab = rnd(6, N, scale=0.3); temb = rnd(2, 6 * N, scale=0.3)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, epi=ops.EPI_GATE_RES, res=res, gate_table=tab[2], gate_temb=temb[:, 2 * N:],
             gate2_table=tab[5], gate2_temb=temb[:, 5 * N:], out2=out2, temb_stride=6 * N, rows_per_sample=1344)
    gate = (tab[2].float()[None] + temb[:, 2 * N:3 * N].float()).repeat_interleave(1344, 0)
    gate2 = (tab[5].float()[None] + temb[:, 5 * N:6 * N].float()).repeat_interleave(1344, 0)
    ref = res.float() + gate * pre
    report("epi GATE_RES out", out, ref, 1e-2)
    report("epi GATE_RES out2", out2, ref.bfloat16().float() * gate2, 1e-2)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, epi=ops.EPI_GATE_RES, res=res)
    report("epi RES (no gate)", out, res.float() + pre, 1e-2)
    aux = rnd(M, N)
    ops.gemm(A, B, out, M=M, N=N, K=K, epi=ops.EPI_MUL_DGELU, aux=aux)
    x = aux.float().requires_grad_(True)
    F.gelu(x, approximate="tanh").sum().backward()
    report("epi MUL_DGELU", out, acc * x.grad, 1e-2)
    o32 = torch.zeros(M, N, device=dev, dtype=torch.float32)
    ops.gemm(A, B, o32, M=M, N=N, K=K, bias=bias, epi=ops.EPI_F32_STORE)
    report("epi F32_STORE", o32, pre, 2e-3)

    # --- LoRA extension operands
    M, N, K, r = 2688, 6144, 2048, 64
    A = rnd(M, K); B = rnd(N, K, scale=0.05); U = rnd(M, 3 * r); BL = rnd(N, r, scale=0.1); bias = rnd(N)
    out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(A, B, out, M=M, N=N, K=K, bias=bias, A2=U, B2=BL, K2=r, a2_group_n=2048, block_n=256)
    ref = A.float() @ B.float().t() + bias.float()
    for j in range(3):
        ref[:, j * 2048:(j + 1) * 2048] += U[:, j * r:(j + 1) * r].float() @ BL[j * 2048:(j + 1) * 2048].float().t()
    report("gemm ext fwd-LoRA (a2_group_n)", out, ref, 1e-2)
    # bwd: [dy | du] [W ; A]  with W [N=6144(contract), Kin=2048] MN-major and A_lora [192, 2048]
    dY = rnd(M, 6144); W = rnd(6144, 2048, scale=0.05); dU = rnd(M, 192); AL = rnd(192, 2048, scale=0.1)
    out = torch.zeros(M, 2048, device=dev, dtype=torch.bfloat16)
    ops.gemm(dY, W, out, M=M, N=2048, K=6144, b_mn=True, A2=dU, B2=AL, K2=192)
    report("gemm ext bwd-LoRA (MN-major B2)", out, dY.float() @ W.float() + dU.float() @ AL.float(), 1e-2)

    # --- batch offsets: du_j = dy[:, j*2048:(j+1)*2048] @ BL[j*2048:(j+1)*2048, :]  (B MN-major [K=2048 rows, r cols])
    du = torch.zeros(M, 3 * r, device=dev, dtype=torch.bfloat16)
    ops.gemm(dY, BL, du, M=M, N=r, K=2048, b_mn=True, batch=3, a_boff=(0, 2048), b_boff=(2048, 0), c_boff=r, ldc=3 * r,
             block_n=64)
    ref = torch.cat([dY[:, j * 2048:(j + 1) * 2048].float() @ BL[j * 2048:(j + 1) * 2048].float() for j in range(3)], 1)
    report("gemm batched offsets (du_cat)", du, ref, 1e-2)

    # --- timing
    for (M, N, K, bn) in [(2688, 2048, 2048, 0), (2688, 6144, 2048, 0), (2688, 8192, 2048, 0), (2688, 2048, 8192, 0),
                          (2688, 2048, 2048, 128), (2688, 2048, 2048, 256), (2688, 8192, 2048, 256), (2688, 8192, 2048, 128),
                          (2688, 8192, 2048, 192)]:
        A = rnd(M, K); B = rnd(N, K, scale=0.05)
        out = torch.zeros(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gemm(A, B, out, M=M, N=N, K=K, block_n=bn)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            torch.matmul(A, B.t(), out=out)
        t1.record(); torch.cuda.synchronize()
        ms_t = t0.elapsed_time(t1) / 20
        print(f"[time] gemm {M}x{N}x{K} bn={bn}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TFLOP/s | cuBLAS {ms_t*1e3:.1f} us {2*M*N*K/ms_t/1e9:.0f} TFLOP/s", flush=True)


def check_elem():
    torch.manual_seed(1)
    B_, S, D = 2, 1344, 2048
    R = B_ * S
    x = rnd(R, D); tab = rnd(6, D, scale=0.3); temb = rnd(B_, 6 * D, scale=0.3)
    y = torch.empty_like(x)
    for ln in (False, True):
        ops.norm_modulate_fwd(x, y, tab[0], temb[:, 0:], tab[1], temb[:, D:], 6 * D, R, D, S, 1e-6, ln)
        xf = x.float().requires_grad_(True)
        shift = (tab[0].float()[None] + temb[:, :D].float()).repeat_interleave(S, 0)
        scale = (tab[1].float()[None] + temb[:, D:2 * D].float()).repeat_interleave(S, 0)
        n = F.layer_norm(xf, (D,), eps=1e-6) if ln else F.rms_norm(xf, (D,), eps=1e-6)
        ref = n * (1 + scale) + shift
        report(f"norm_modulate_fwd ln={ln}", y, ref, 1e-2)
        dy = rnd(R, D); dxin = rnd(R, D)
        ref.backward(dy.float())
        dx = torch.empty_like(x); o2 = torch.empty_like(x)
        ops.norm_modulate_bwd(dy, x, dxin, dx, tab[1], temb[:, D:], 6 * D, R, D, S, 1e-6, ln, gate2_tab=tab[5],
                              gate2_emb=temb[:, 5 * D:], out2=o2)
        refdx = dxin.float() + xf.grad
        report(f"norm_modulate_bwd ln={ln}", dx, refdx, 1e-2)
        g2 = (tab[5].float()[None] + temb[:, 5 * D:].float()).repeat_interleave(S, 0)
        report(f"norm_modulate_bwd out2 ln={ln}", o2, refdx.bfloat16().float() * g2, 1e-2)
        ops.norm_modulate_bwd(dy, x, None, dx, tab[1], temb[:, D:], 6 * D, R, D, S, 1e-6, ln)
        report(f"norm_modulate_bwd no-accum ln={ln}", dx, xf.grad, 1e-2)
    cs = torch.empty_like(x)
    ops.colscale(x, cs, tab[2], temb[:, 2 * D:], 6 * D, R, D, S)
    g = (tab[2].float()[None] + temb[:, 2 * D:3 * D].float()).repeat_interleave(S, 0)
    report("colscale", cs, x.float() * g, 1e-2)

    # rope table vs oracle-style torch computation
    Fr, Hh, Ww = 7, 16, 24
    S = Fr * Hh * Ww
    cos = torch.empty(S, D // 2, device=dev); sin = torch.empty(S, D // 2, device=dev)
    sf, sh, sw = (8 / 25) / 20, 32 / 2048, 32 / 2048
    ops.rope_table(cos, sin, Fr, Hh, Ww, D, sf, sh, sw)
    from oracle.ltx_oracle import ltx_rope_table, apply_rotary_emb  # test-only
    rc, rs = ltx_rope_table(Fr, Hh, Ww, D, [8 / 25, 32, 32], 1, "cpu")
    report("rope_table cos", cos.cpu(), rc[0][:, 0::2], 5e-3)
    report("rope_table sin", sin.cpu(), rs[0][:, 0::2], 5e-3)
    cos = rc[0].to(dev).contiguous(); sin = rs[0].to(dev).contiguous()
    cos_p = rc[0][:, 0::2].to(dev).contiguous(); sin_p = rs[0][:, 0::2].to(dev).contiguous()

    # qknorm + rope
    Bq, H = 2, 32
    qkv = rnd(Bq * S, 3 * D); w = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16()
    dst = torch.empty(Bq, H, S, 64, device=dev, dtype=torch.bfloat16)
    for which, norm, rope in ((0, True, True), (1, True, False), (2, False, False)):
        ops.qknorm_rope_fwd(qkv, 3 * D, which * D, w, cos_p if rope else None, sin_p if rope else None, dst, Bq, S, H, norm, 1e-5)
        xf = qkv[:, which * D:(which + 1) * D].float().reshape(Bq, S, D).requires_grad_(True)
        n = F.rms_norm(xf, (D,), weight=w.float(), eps=1e-5) if norm else xf
        if rope:
            xr, xi = n.unflatten(2, (-1, 2)).unbind(-1)
            rot = torch.stack([-xi, xr], dim=-1).flatten(2)
            n = n * cos[None] + rot * sin[None]
        ref = n.unflatten(2, (H, 64)).transpose(1, 2)
        report(f"qknorm_rope_fwd which={which} norm={norm} rope={rope}", dst, ref, 1e-2)
        dyh = rnd(Bq, H, S, 64)
        ref.backward(dyh.float())
        dx = torch.zeros(Bq * S, 3 * D, device=dev, dtype=torch.bfloat16)
        ops.qknorm_rope_bwd(dyh, qkv, 3 * D, which * D, w, cos_p if rope else None, sin_p if rope else None, dx, 3 * D,
                            which * D, Bq, S, H, norm, 1e-5)
        report(f"qknorm_rope_bwd which={which}", dx[:, which * D:(which + 1) * D], xf.grad.reshape(Bq * S, D), 1e-2)

    # fused 3-segment launch (q: norm+rope, k: norm+rope, v: copy) against the same references
    wq = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16(); wk = (1 + 0.1 * torch.randn(D, device=dev)).bfloat16()
    dsts = [torch.empty(Bq, H, S, 64, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    ops.qkv_norm_rope_fwd(qkv, 3 * D, 0, (wq, wk, None), 0b011, cos_p, sin_p, dsts, Bq, S, H, 1e-5)
    dys = [rnd(Bq, H, S, 64) for _ in range(3)]
    dx = torch.zeros(Bq * S, 3 * D, device=dev, dtype=torch.bfloat16)
    ops.qkv_norm_rope_bwd(dys, qkv, 3 * D, 0, (wq, wk, None), 0b011, cos_p, sin_p, dx, 3 * D, 0, Bq, S, H, 1e-5)
    for which, wt in ((0, wq), (1, wk), (2, None)):
        xf = qkv[:, which * D:(which + 1) * D].float().reshape(Bq, S, D).requires_grad_(True)
        n = F.rms_norm(xf, (D,), weight=wt.float(), eps=1e-5) if wt is not None else xf
        if which < 2:
            xr, xi = n.unflatten(2, (-1, 2)).unbind(-1)
            rot = torch.stack([-xi, xr], dim=-1).flatten(2)
            n = n * cos[None] + rot * sin[None]
        ref = n.unflatten(2, (H, 64)).transpose(1, 2)
        report(f"qkv_norm_rope_fwd seg={which}", dsts[which], ref, 1e-2)
        ref.backward(dys[which].float())
        report(f"qkv_norm_rope_bwd seg={which}", dx[:, which * D:(which + 1) * D], xf.grad.reshape(Bq * S, D), 1e-2)
    for fn, name in ((lambda: ops.qkv_norm_rope_fwd(qkv, 3 * D, 0, (wq, wk, None), 0b011, cos_p, sin_p, dsts, Bq, S, H, 1e-5), "qkv_norm_rope_fwd"),
                     (lambda: ops.qkv_norm_rope_bwd(dys, qkv, 3 * D, 0, (wq, wk, None), 0b011, cos_p, sin_p, dx, 3 * D, 0, Bq, S, H, 1e-5), "qkv_norm_rope_bwd"),
                     (lambda: ops.norm_modulate_fwd(x, y, tab[0], temb[:, 0:], tab[1], temb[:, D:], 6 * D, R, D, S // 2 * 2 // 2 if False else 1344, 1e-6), "norm_modulate_fwd")):
        for _ in range(3):
            fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        print(f"[time] {name} rows={Bq * S}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us (eager, may be host-bound)", flush=True)

    # prep + loss
    Bp, Cc = 2, 128
    lat = rnd(Bp, Cc, Fr, Hh, Ww); noise = rnd(Bp, Cc, Fr, Hh, Ww)
    mean = torch.randn(Bp, Cc, device=dev) * 0.1; std = 1 + 0.1 * torch.rand(Bp, Cc, device=dev)
    sigma = torch.tensor([0.3, 0.811], device=dev); sff = torch.tensor([0.1, 0.25], device=dev)
    xt = torch.empty(Bp, S, Cc, device=dev, dtype=torch.bfloat16); tg = torch.empty_like(xt)
    from oracle.ltx_oracle import normalize_latents, pack_latents, flow_match_xt
    for ff in (None, sff):
        ops.prep_noise_pack(lat, noise, mean, std, sigma, ff, xt, tg, Bp, Cc, Fr, Hh * Ww)
        x0 = normalize_latents(lat, mean, std)
        sg = sigma.view(Bp, 1, 1, 1, 1)
        if ff is None:
            noisy = flow_match_xt(x0, noise, sg)
        else:
            noisy = torch.cat([flow_match_xt(x0[:, :, :1], noise[:, :, :1], ff.view(Bp, 1, 1, 1, 1)),
                               flow_match_xt(x0[:, :, 1:], noise[:, :, 1:], sg)], 2)
        ref_xt = pack_latents(noisy).to(torch.bfloat16)
        ref_tg = pack_latents(noise) - pack_latents(x0)
        report(f"prep x_t (ff={'y' if ff is not None else 'n'})", xt, ref_xt, 0.0, "bit-exact")
        report(f"prep target (ff={'y' if ff is not None else 'n'})", tg, ref_tg, 0.0, "bit-exact")
    pred = rnd(Bp, S, Cc)
    wgt = torch.tensor([1.0, 2.5], device=dev)
    loss = torch.zeros(1, device=dev); dpred = torch.empty_like(pred); ws = torch.empty(1024, device=dev)
    ops.loss_mse(pred, tg, wgt, 1.0, loss, dpred, ws, Bp, S * Cc)
    pf = pred.float().requires_grad_(True)
    l = (wgt.view(Bp, 1, 1) * (pf - tg.float()).pow(2)).mean((1, 2)).mean()
    l.backward()
    report("loss value", loss, l.detach().view(1), 1e-5)
    report("loss dpred", dpred, pf.grad, 1e-2)
    # sinusoid
    t = torch.tensor([0.0, 1.0, 500.0, 999.0], device=dev)
    so = torch.empty(4, 256, device=dev, dtype=torch.bfloat16)
    ops.timestep_sinusoid(t, so, 4)
    from oracle.ltx_oracle import sinusoid_256
    report("timestep_sinusoid", so, sinusoid_256(t.cpu()).to(dev), 1e-2)
    # cast / sumsq / adamw
    n = 1000003
    src = torch.randn(n, device=dev); dstb = torch.empty(n, device=dev, dtype=torch.bfloat16)
    ops.cast_f32_bf16(src, dstb, n, 0.5)
    report("cast_f32_bf16", dstb, (src * 0.5).bfloat16(), 0.0, "bit-exact")
    ss = torch.zeros(1, device=dev)
    ops.sumsq(src, n, ss, ws)
    report("sumsq", ss, (src.double() ** 2).sum().float().view(1), 1e-5)
    p = torch.randn(n, device=dev); g = torch.randn(n, device=dev); m = torch.zeros(n, device=dev); v = torch.zeros(n, device=dev)
    pr = torch.nn.Parameter(p.clone()); pr.grad = g.clone()
    opt = torch.optim.AdamW([pr], lr=1e-2, betas=(0.9, 0.99), weight_decay=1e-2, eps=1e-8)
    ss.zero_(); ops.sumsq(g, n, ss, ws)
    for step in (1, 2):
        torch.nn.utils.clip_grad_norm_([pr], 1.0)
        opt.step()
        ops.adamw_clip(p, g, m, v, n, ss, 1.0, 1e-2, 0.9, 0.99, 1e-8, 1e-2, step)
        report(f"adamw_clip step{step}", p, pr.detach(), 1e-5)
        g2 = torch.randn(n, device=dev)
        g.copy_(g2); pr.grad = g2.clone()
        ss.zero_(); ops.sumsq(g, n, ss, ws)


def check_attn():
    torch.manual_seed(0)
    cases = [(2, 8, 256, 256, False), (1, 32, 2688, 2688, False), (2, 4, 2688, 128, True), (1, 2, 200, 72, True),
             (1, 2, 128, 128, False), (1, 32, 2688, 128, True), (5, 32, 300, 128, True), (3, 2, 1000, 100, False)]
    for (B_, H, Sq, Sk, use_bias) in cases:
        q = rnd(B_, H, Sq, 64); k = rnd(B_, H, Sk, 64); v = rnd(B_, H, Sk, 64)
        bias = None
        if use_bias:
            lens = torch.randint(1, Sk + 1, (B_,), device=dev)
            mask = torch.arange(Sk, device=dev)[None] < lens[:, None]
            bias = ((1 - mask.float()) * -10000.0).contiguous()
        out = torch.zeros(B_, Sq, H * 64, device=dev, dtype=torch.bfloat16)
        lse = torch.zeros(B_, H, Sq, device=dev)
        scale = 1.0 / 8.0
        name = f"B{B_} H{H} Sq{Sq} Sk{Sk} bias={use_bias}"
        try:
            ops.attn_fwd(q, k, v, bias, out, lse, B_, H, Sq, Sk, scale)
            torch.cuda.synchronize()
        except Exception as e:
            print("EXC attn_fwd", name, e); RESULTS.append((name, False)); continue
        qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
        am = bias[:, None, None, :] if bias is not None else None
        with torch.nn.attention.sdpa_kernel(torch.nn.attention.SDPBackend.MATH):
            ref = F.scaled_dot_product_attention(qf, kf, vf, attn_mask=am)
        report(f"attn_fwd {name}", out, ref.transpose(1, 2).flatten(2), 5e-3 / max(ref.abs().max().item(), 1e-6) if False else 1e-2)
        s = (qf @ kf.transpose(-1, -2)) * scale + (am if am is not None else 0)
        report(f"attn_fwd lse {name}", lse, torch.logsumexp(s, -1), 1e-3)
        dout = rnd(B_, Sq, H * 64)
        ref.backward(dout.float().unflatten(2, (H, 64)).transpose(1, 2))
        dq = torch.zeros_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v)
        ws = torch.zeros(ops.attn_bwd_ws_floats(B_, H, Sq, Sk), device=dev)
        try:
            ops.attn_bwd(q, k, v, bias, out, dout, lse, ws, dq, dk, dv, B_, H, Sq, Sk, scale)
            torch.cuda.synchronize()
        except Exception as e:
            print("EXC attn_bwd", name, e); RESULTS.append((name + " bwd", False)); continue
        report(f"attn_bwd dq {name}", dq, qf.grad, 2e-2)
        report(f"attn_bwd dk {name}", dk, kf.grad, 2e-2)
        report(f"attn_bwd dv {name}", dv, vf.grad, 2e-2)
    # timing at the LTX shape
    B_, H, S = 1, 32, 2688
    q = rnd(B_, H, S, 64); k = rnd(B_, H, S, 64); v = rnd(B_, H, S, 64)
    out = torch.zeros(B_, S, H * 64, device=dev, dtype=torch.bfloat16); lse = torch.zeros(B_, H, S, device=dev)
    dout = rnd(B_, S, H * 64); dq = torch.zeros_like(q); dk = torch.zeros_like(k); dv = torch.zeros_like(v)
    ws = torch.zeros(ops.attn_bwd_ws_floats(B_, H, S, S), device=dev)
    for fn, name, flops in ((lambda: ops.attn_fwd(q, k, v, None, out, lse, B_, H, S, S, 0.125), "attn_fwd", 4 * S * S * 64 * H),
                            (lambda: ops.attn_bwd(q, k, v, None, out, dout, lse, ws, dq, dk, dv, B_, H, S, S, 0.125), "attn_bwd", 10 * S * S * 64 * H)):
        try:
            for _ in range(3):
                fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"[time] {name} S=2688 H=32: {ms*1e3:.1f} us  {flops/ms/1e9:.0f} TFLOP/s", flush=True)
        except Exception as e:
            print("EXC timing", name, e)
    # cross attention at the LTX shape: 2688 queries x 128 text keys, key bias
    L = 128
    kc = rnd(B_, H, L, 64); vc = rnd(B_, H, L, 64); dkc = torch.zeros_like(kc); dvc = torch.zeros_like(vc)
    biasc = torch.zeros(B_, L, device=dev); biasc[:, 77:] = -10000.0
    wsc = torch.zeros(ops.attn_bwd_ws_floats(B_, H, S, L), device=dev)
    for fn, name in ((lambda: ops.attn_fwd(q, kc, vc, biasc, out, lse, B_, H, S, L, 0.125), "cross attn_fwd"),
                     (lambda: ops.attn_bwd(q, kc, vc, biasc, out, dout, lse, wsc, dq, dkc, dvc, B_, H, S, L, 0.125), "cross attn_bwd")):
        try:
            for _ in range(3):
                fn()
            # these launches are short enough to be host-bound from Python: time a CUDA-graph replay of 20 calls
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                fn()
                side.synchronize()
                with torch.cuda.graph(gr, stream=side):
                    for _ in range(20):
                        fn()
            gr.replay(); torch.cuda.synchronize()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                gr.replay()
            e1.record(); torch.cuda.synchronize()
            print(f"[time] {name} Sq=2688 Sk=128 H=32: {e0.elapsed_time(e1) / 100 * 1e3:.1f} us (graph replay)", flush=True)
        except Exception as e:
            print("EXC timing", name, e)
    qt = q.clone().requires_grad_(True)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        F.scaled_dot_product_attention(qt, k, v)
    e0.record()
    for _ in range(10):
        F.scaled_dot_product_attention(qt, k, v)
    e1.record(); torch.cuda.synchronize()
    print(f"[time] torch sdpa fwd S=2688: {e0.elapsed_time(e1)/10*1e3:.1f} us", flush=True)


def _build_pair(cfg_kwargs, rank, seed=0, lora_b_std=0.02):
    """oracle (CPU fp32 math, bf16-valued base weights) and the B200 model with identical parameters."""
    from oracle import ltx_oracle as O
    from finetrainers_b200.model import B200LTXTransformer, LTXConfig
    ocfg = O.LTXConfig(**cfg_kwargs)
    om = O.LTXTransformerOracle(ocfg)
    O.add_lora(om, rank, rank)
    O.synthetic_init_(om, seed=seed, lora_b_std=lora_b_std)
    with torch.no_grad():
        for n, p in om.named_parameters():
            if "lora_" not in n:
                p.copy_(p.to(torch.bfloat16).float())
    bm = B200LTXTransformer(LTXConfig(**cfg_kwargs), torch.bfloat16, "cuda")
    bm.add_adapter(rank, rank)
    missing = bm.load_state_dict(om.state_dict(), strict=True)
    bm.prepare()
    return O, om, bm


def _run_b200(bm, batch, steps=1):
    from finetrainers_b200.trainer import SFTTrainStep
    st = SFTTrainStep(bm, flow_weighting_scheme="none")
    import random
    random.seed(12345)  # keep first-frame conditioning off in parity runs: random.random() >= 0.1 for this seed? force below
    st.spec.first_frame_conditioning_p = 0.0
    cond = {"encoder_hidden_states": batch["encoder_hidden_states"].cuda(), "encoder_attention_mask": batch["encoder_attention_mask"].cuda()}
    lat = {"latents": batch["latents"].cuda(), "latents_mean": batch["latents_mean"].cuda(), "latents_std": batch["latents_std"].cuda()}
    st.micro_step(cond, lat, sigmas=batch["sigmas"].view(-1).cuda(), noise=batch["noise"].cuda())
    torch.cuda.synchronize()
    B, S = batch["latents"].shape[0], batch["latents"].shape[2] * batch["latents"].shape[3] * batch["latents"].shape[4]
    ws = bm._workspace(B, S, batch["encoder_hidden_states"].shape[1])
    return st, st.loss_buf.item(), ws["pred"].view(B, S, -1).float().cpu()


def check_model():
    torch.manual_seed(0)
    for (rank, lb) in ((64, 0.02), (16, 0.02)):
        cfgk = dict(in_channels=32, out_channels=32, num_attention_heads=4, attention_head_dim=64, cross_attention_dim=256,
                    num_layers=2, caption_channels=128)
        O, om, bm = _build_pair(cfgk, rank, lora_b_std=lb)
        batch = O.make_synthetic_batch(om.cfg, 2, 2, 4, 9, text_len=24, seed=7)
        loss_o, pred_o = O.oracle_step(om, {k: (v.float() if v.is_floating_point() else v) for k, v in batch.items()})
        st, loss_b, pred_b = _run_b200(bm, batch)
        report(f"model small r={rank}: pred", pred_b, pred_o, 3e-2)
        rel = abs(loss_b - loss_o.item()) / abs(loss_o.item())
        ok = rel < 1e-3
        RESULTS.append((f"model small r={rank}: loss", ok))
        print(f"[{' ok ' if ok else 'FAIL'}] model small r={rank}: loss b200={loss_b:.6f} oracle={loss_o.item():.6f} rel={rel:.2e}")
        og = dict(om.named_parameters())
        worst = 0.0
        gmax = max(p.grad.abs().max().item() for n, p in om.named_parameters() if "lora_" in n)
        for n, p in bm.named_parameters():
            if "lora_" in n:
                g_o = og[n].grad
                g_b = p.grad.float().cpu()
                denom = max(g_o.abs().max().item(), 5e-2 * gmax)
                e = (g_b - g_o).abs().max().item() / denom
                worst = max(worst, e)
                if e > 5e-2:
                    print(f"   grad mismatch {n}: rel_max_err={e:.3e} ref_max={denom:.3e}")
        ok = worst < 5e-2
        RESULTS.append((f"model small r={rank}: lora grads", ok))
        print(f"[{' ok ' if ok else 'FAIL'}] model small r={rank}: LoRA grads worst rel-to-max err {worst:.3e}")
        # bf16 oracle (reference-typed) for context
        om16 = om.to(torch.bfloat16)
        for n, p in om16.named_parameters():
            if "lora_" in n:
                p.data = p.data.float()
        om16.zero_grad()
        l16, p16 = O.oracle_step(om16, batch, backward=False)
        print(f"   context: bf16-typed oracle loss={l16.item():.6f} rel-to-fp32-oracle={abs(l16.item()-loss_o.item())/loss_o.item():.2e}; "
              f"pred err bf16-oracle={(p16.float()-pred_o).abs().max().item():.3e} b200={(pred_b-pred_o).abs().max().item():.3e}")
        # optimizer step runs
        st.optimizer_step()
        torch.cuda.synchronize()
        del bm, st


def check_model_full():
    """LTX-2B config, 49x512x768 (S=2688), B=1, r=64: timing + parity of loss vs the CPU oracle."""
    cfgk = dict()
    from finetrainers_b200.model import B200LTXTransformer, LTXConfig
    from finetrainers_b200.trainer import SFTTrainStep
    from oracle import ltx_oracle as O
    t0 = time.time()
    ocfg = O.LTXConfig()
    om = O.LTXTransformerOracle(ocfg)
    O.add_lora(om, 64, 64)
    O.synthetic_init_(om, seed=0, lora_b_std=0.02)
    om = om.to(torch.bfloat16)
    for n, p in om.named_parameters():
        if "lora_" in n:
            p.data = p.data.float()
    print(f"oracle built {time.time()-t0:.1f}s", flush=True)
    bm = B200LTXTransformer(LTXConfig(), torch.bfloat16, "cuda")
    bm.add_adapter(64, 64)
    bm.load_state_dict(om.state_dict(), strict=True)
    bm.prepare()
    batch = O.make_synthetic_batch(ocfg, 1, 7, 16, 24, seed=1234)
    st, loss_b, pred_b = _run_b200(bm, batch)
    print(f"b200 full-size loss {loss_b:.6f}  ({time.time()-t0:.1f}s)", flush=True)
    # timing
    cond = {"encoder_hidden_states": batch["encoder_hidden_states"].cuda(), "encoder_attention_mask": batch["encoder_attention_mask"].cuda()}
    def one():
        lat = {"latents": batch["latents"].cuda(), "latents_mean": batch["latents_mean"].cuda(), "latents_std": batch["latents_std"].cuda()}
        st.train_step(dict(cond), lat, sigmas=batch["sigmas"].view(-1).cuda(), noise=batch["noise"].cuda())
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        one()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"[time] full train step: {ms:.2f} ms -> {2688/ms*1e3:.0f} tokens/s  (algorithmic {2688*8.88e9/ms/1e9:.0f} TFLOP/s)", flush=True)
    # fp32-math oracle on CPU (weights are bf16-valued); forward only to bound the time
    om32 = om.float()
    tc = time.time()
    with torch.no_grad():
        loss_o, pred_o = O.oracle_step(om32, {k: (v.float() if v.is_floating_point() else v) for k, v in batch.items()}, backward=False)
    print(f"oracle fwd {time.time()-tc:.1f}s", flush=True)
    rel = abs(loss_b - loss_o.item()) / abs(loss_o.item())
    ok = rel < 1e-3
    RESULTS.append(("model full: loss", ok))
    print(f"[{' ok ' if ok else 'FAIL'}] model full: loss b200={loss_b:.6f} oracle={loss_o.item():.6f} rel={rel:.2e}")
    report("model full: pred", pred_b, pred_o, 5e-2)


if __name__ == "__main__":
    which = sys.argv[1:] or ["gemm", "elem", "attn"]
    print(torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    for w in which:
        try:
            {"gemm": check_gemm, "elem": check_elem, "attn": check_attn, "model": check_model, "full": check_model_full}[w]()
        except Exception:
            traceback.print_exc()
            RESULTS.append((w + " (exception)", False))
    nfail = sum(1 for _, ok in RESULTS if not ok)
    print(f"SUMMARY: {len(RESULTS) - nfail}/{len(RESULTS)} ok, {nfail} failed, {time.time()-t0:.1f}s")
    for n, ok in RESULTS:
        if not ok:
            print("  FAILED:", n)
