"""Thin tensor->pointer wrappers over the libb2d C ABI (include/b2d.h).  PyTorch only supplies device memory and the
current stream; all arithmetic happens in the sm_100a kernels.  No fallbacks: a missing library or a non-CUDA tensor
raises."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import lib as _l
from .lib import GemmDesc, check

EPI_STORE, EPI_GELU, EPI_SILU, EPI_GATE_RES, EPI_MUL_DGELU, EPI_F32_ATOMIC, EPI_F32_ATOMIC_T, EPI_F32_STORE = range(8)

LAUNCH_COUNT = 0  # number of libb2d kernels launched (bench.py reports it as gpu_launches)
TIMING = False    # when True every wrapper brackets its launch with CUDA events on the current stream
KERNEL_TIMES = {}  # tag -> [(start_event, end_event), ...]
CONTEXT = ""      # optional call-site label set by the model (profiling only): tags become "<CONTEXT>/<tag>"


class _Timed:
    __slots__ = ("tag", "e0")

    def __init__(self, tag):
        self.tag = tag

    def __enter__(self):
        if TIMING:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *a):
        if TIMING:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            KERNEL_TIMES.setdefault(f"{CONTEXT}/{self.tag}" if CONTEXT else self.tag, []).append((self.e0, e1))


def collect_kernel_times():
    """{tag: (total_ms, launches)} — call after torch.cuda.synchronize()."""
    return {k: (sum(a.elapsed_time(b) for a, b in v), len(v)) for k, v in KERNEL_TIMES.items()}


def _ptr(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise _l.B2DError("libb2d ops need CUDA tensors (there is no CPU fallback)")
    return C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _count(n=1):
    global LAUNCH_COUNT
    LAUNCH_COUNT += n


def gemm(A: torch.Tensor, B: torch.Tensor, out: torch.Tensor, *, M: int, N: int, K: int, lda=None, ldb=None, ldc=None,
         a_mn=False, b_mn=False, A2=None, B2=None, K2=0, lda2=None, ldb2=None, a2_group_n=0, splits=1, batch=1,
         a_boff=(0, 0), b_boff=(0, 0), c_boff=0, epi=EPI_STORE, alpha=1.0, out2=None, ldc2=None, bias=None, res=None,
         ldres=None, aux=None, ldaux=None, gate_table=None, gate_temb=None, gate2_table=None, gate2_temb=None,
         temb_stride=0, rows_per_sample=0, block_n=0, max_ctas=0, a2_boff_row=0, b2_boff_row=0, bias_boff=0,
         cta_pair=0, tag="gemm"):
    """C = epilogue(alpha * (opA(A) opB(B)^T + A2 B2^T)).  See include/b2d.h b2d_gemm_desc."""
    d = GemmDesc()
    d.A = A.data_ptr(); d.lda = lda if lda is not None else A.stride(0)
    d.B = B.data_ptr(); d.ldb = ldb if ldb is not None else B.stride(0)
    if K2:
        d.A2 = A2.data_ptr(); d.lda2 = lda2 if lda2 is not None else A2.stride(0)
        d.B2 = B2.data_ptr(); d.ldb2 = ldb2 if ldb2 is not None else B2.stride(0)
    d.M, d.N, d.K, d.K2 = M, N, K, K2
    d.a_mn_major, d.b_mn_major = int(a_mn), int(b_mn)
    d.a2_group_n = a2_group_n
    d.splits, d.batch = splits, batch
    d.a_boff_row, d.a_boff_col = a_boff
    d.b_boff_row, d.b_boff_col = b_boff
    d.c_boff = c_boff
    d.epi = epi
    d.alpha = alpha
    d.out = out.data_ptr(); d.ldc = ldc if ldc is not None else out.stride(0)
    if out2 is not None:
        d.out2 = out2.data_ptr(); d.ldc2 = ldc2 if ldc2 is not None else out2.stride(0)
    if bias is not None:
        d.bias = bias.data_ptr()
    if res is not None:
        d.res = res.data_ptr(); d.ldres = ldres if ldres is not None else res.stride(0)
    if aux is not None:
        d.aux = aux.data_ptr(); d.ldaux = ldaux if ldaux is not None else aux.stride(0)
    if gate_table is not None:
        d.gate_table = gate_table.data_ptr(); d.gate_temb = gate_temb.data_ptr()
    if gate2_table is not None:
        d.gate2_table = gate2_table.data_ptr(); d.gate2_temb = gate2_temb.data_ptr()
    d.temb_stride = temb_stride
    d.rows_per_sample = rows_per_sample
    d.block_n = block_n
    d.max_ctas = max_ctas
    d.a2_boff_row, d.b2_boff_row, d.bias_boff = a2_boff_row, b2_boff_row, bias_boff
    d.cta_pair = cta_pair
    with _Timed(tag):
        check(_l.load().b2d_gemm(C.byref(d), _stream()), "gemm")
    _count()
    return out


def norm_modulate_fwd(x, y, shift_tab, shift_emb, scale_tab, scale_emb, emb_stride, rows, D, rows_per_sample, eps,
                      layer_norm=False):
    with _Timed("norm_modulate_fwd"):
        check(_l.load().b2d_norm_modulate_fwd(_ptr(x), _ptr(y), _ptr(shift_tab), _ptr(shift_emb), _ptr(scale_tab),
                                              _ptr(scale_emb), C.c_int64(emb_stride), rows, D, rows_per_sample,
                                              C.c_float(eps), int(layer_norm), _stream()), "norm_modulate_fwd")
    _count()
    return y


def norm_modulate_bwd(dy, x, dx_in, dx_out, scale_tab, scale_emb, emb_stride, rows, D, rows_per_sample, eps,
                      layer_norm=False, gate2_tab=None, gate2_emb=None, out2=None):
    with _Timed("norm_modulate_bwd"):
        check(_l.load().b2d_norm_modulate_bwd(_ptr(dy), _ptr(x), _ptr(dx_in), _ptr(dx_out), _ptr(scale_tab),
                                              _ptr(scale_emb), _ptr(gate2_tab), _ptr(gate2_emb), _ptr(out2),
                                              C.c_int64(emb_stride), rows, D, rows_per_sample, C.c_float(eps),
                                              int(layer_norm), _stream()), "norm_modulate_bwd")
    _count()
    return dx_out


def colscale(x, out, tab, emb, emb_stride, rows, D, rows_per_sample):
    with _Timed("colscale"):
        check(_l.load().b2d_colscale(_ptr(x), _ptr(out), _ptr(tab), _ptr(emb), C.c_int64(emb_stride), rows, D,
                                     rows_per_sample, _stream()), "colscale")
    _count()
    return out


def qknorm_rope_fwd(src, ld, col_off, weight, cos, sin, dst, B, S, H, norm, eps):
    with _Timed("qknorm_rope_fwd"):
        check(_l.load().b2d_qknorm_rope_fwd(_ptr(src), C.c_int64(ld), C.c_int64(col_off), _ptr(weight), _ptr(cos),
                                            _ptr(sin), _ptr(dst), B, S, H, int(norm), C.c_float(eps), _stream()),
              "qknorm_rope_fwd")
    _count()
    return dst


def qknorm_rope_bwd(dyh, x, ld, col_off, weight, cos, sin, dx, ld_dx, dx_col_off, B, S, H, norm, eps):
    with _Timed("qknorm_rope_bwd"):
        check(_l.load().b2d_qknorm_rope_bwd(_ptr(dyh), _ptr(x), C.c_int64(ld), C.c_int64(col_off), _ptr(weight),
                                            _ptr(cos), _ptr(sin), _ptr(dx), C.c_int64(ld_dx), C.c_int64(dx_col_off), B,
                                            S, H, int(norm), C.c_float(eps), _stream()), "qknorm_rope_bwd")
    _count()
    return dx


def qkv_norm_rope_fwd(src, ld, col_off, weights, rope_mask, cos, sin, dsts, B, S, H, eps, rows_per_w=0, w_stride=0):
    """nseg = len(dsts) consecutive D-wide segments of src rows -> head-split dsts in ONE launch (b2d.h)."""
    n = len(dsts)
    w = list(weights) + [None] * (3 - n)
    d = list(dsts) + [None] * (3 - n)
    with _Timed("qknorm_rope_fwd"):
        check(_l.load().b2d_qkv_norm_rope_fwd(_ptr(src), C.c_int64(ld), C.c_int64(col_off), n, _ptr(w[0]), _ptr(w[1]),
                                              _ptr(w[2]), int(rope_mask), _ptr(cos), _ptr(sin), _ptr(d[0]), _ptr(d[1]),
                                              _ptr(d[2]), B, S, H, C.c_float(eps), int(rows_per_w), C.c_int64(w_stride),
                                              _stream()), "qkv_norm_rope_fwd")
    _count()


def qkv_norm_rope_bwd(dys, x, ld, col_off, weights, rope_mask, cos, sin, dx, ld_dx, dx_col_off, B, S, H, eps,
                      rows_per_w=0, w_stride=0):
    n = len(dys)
    w = list(weights) + [None] * (3 - n)
    d = list(dys) + [None] * (3 - n)
    with _Timed("qknorm_rope_bwd"):
        check(_l.load().b2d_qkv_norm_rope_bwd(_ptr(d[0]), _ptr(d[1]), _ptr(d[2]), _ptr(x), C.c_int64(ld),
                                              C.c_int64(col_off), n, _ptr(w[0]), _ptr(w[1]), _ptr(w[2]), int(rope_mask),
                                              _ptr(cos), _ptr(sin), _ptr(dx), C.c_int64(ld_dx), C.c_int64(dx_col_off), B,
                                              S, H, C.c_float(eps), int(rows_per_w), C.c_int64(w_stride), _stream()),
              "qkv_norm_rope_bwd")
    _count()
    return dx


def rope_table(cos, sin, F, H, W, D, sf, sh, sw):
    check(_l.load().b2d_rope_table(_ptr(cos), _ptr(sin), F, H, W, D, C.c_float(sf), C.c_float(sh), C.c_float(sw),
                                   _stream()), "rope_table")
    _count()


def attn_fwd(q, k, v, key_bias, out, lse, B, H, Sq, Sk, scale):
    with _Timed("attn_fwd"):
        check(_l.load().b2d_attn_fwd(_ptr(q), _ptr(k), _ptr(v), _ptr(key_bias), _ptr(out), _ptr(lse), B, H, Sq, Sk,
                                     C.c_float(scale), _stream()), "attn_fwd")
    _count()
    return out


def attn_bwd_ws_floats(B, H, Sq, Sk):
    """fp32 elements b2d_attn_bwd needs in delta_ws (include/b2d.h)."""
    return 2 * B * H * Sq + (2 * B * H * Sk * 64 + B * H if Sk <= 512 else 0)


def attn_bwd(q, k, v, key_bias, out, dout, lse, delta_ws, dq, dk, dv, B, H, Sq, Sk, scale):
    if delta_ws.numel() < attn_bwd_ws_floats(B, H, Sq, Sk):
        raise _l.B2DError(f"attn_bwd workspace too small: {delta_ws.numel()} < {attn_bwd_ws_floats(B, H, Sq, Sk)} floats")
    with _Timed("attn_bwd"):
        check(_l.load().b2d_attn_bwd(_ptr(q), _ptr(k), _ptr(v), _ptr(key_bias), _ptr(out), _ptr(dout), _ptr(lse),
                                     _ptr(delta_ws), _ptr(dq), _ptr(dk), _ptr(dv), B, H, Sq, Sk, C.c_float(scale),
                                     _stream()), "attn_bwd")
    if Sk <= 128:
        _count(1)  # single key tile: ONE fused delta/dQ/dK/dV kernel
    else:
        # delta + dK/dV + dQ, plus two fp32->bf16 converts when the few-key-tiles split path runs
        split = (B * H * ((Sk + 127) // 128) < 96) and Sk <= 512 and ((Sq + 63) // 64) >= 8
        _count(5 if split else 3)


def prep_noise_pack(latents, noise, mean, std, sigma, sigma_ff, x_t, target, B, Cc, F, HW):
    check(_l.load().b2d_prep_noise_pack(_ptr(latents), _ptr(noise), _ptr(mean), _ptr(std), _ptr(sigma),
                                        _ptr(sigma_ff), _ptr(x_t), _ptr(target), B, Cc, F, HW, _stream()), "prep")
    _count()


def loss_mse(pred, target, weight, loss_scale, loss_out, dpred, partial_ws, B, per_sample):
    with _Timed("loss"):
        check(_l.load().b2d_loss_mse(_ptr(pred), _ptr(target), _ptr(weight), C.c_float(loss_scale), _ptr(loss_out),
                                     _ptr(dpred), _ptr(partial_ws), B, C.c_int64(per_sample), _stream()), "loss_mse")
    _count(2)


def timestep_sinusoid(t, out, n):
    check(_l.load().b2d_timestep_sinusoid(_ptr(t), _ptr(out), n, _stream()), "timestep_sinusoid")
    _count()


def cast_f32_bf16(src, dst, n, scale=1.0):
    with _Timed("cast"):
        check(_l.load().b2d_cast_f32_bf16(_ptr(src), _ptr(dst), C.c_int64(n), C.c_float(scale), _stream()), "cast")
    _count()


def sumsq(x, n, out, partial_ws):
    with _Timed("sumsq"):
        check(_l.load().b2d_sumsq(_ptr(x), C.c_int64(n), _ptr(out), _ptr(partial_ws), _stream()), "sumsq")
    _count(2)


def adamw_clip(p, g, m, v, n, sumsq_t, max_norm, lr, beta1, beta2, eps, wd, step, grad_div=1.0):
    with _Timed("adamw"):
        check(_l.load().b2d_adamw_clip(_ptr(p), _ptr(g), _ptr(m), _ptr(v), C.c_int64(n), _ptr(sumsq_t),
                                       C.c_float(max_norm), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
                                       C.c_float(eps), C.c_float(wd), int(step), C.c_float(grad_div), _stream()), "adamw")
    _count()
