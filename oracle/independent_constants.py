"""Second, independent derivation of the UPSTREAM-ONLY constants the oracle restates — TEST INFRASTRUCTURE.

The block arithmetic of the reference lives in diffusers / peft, which are neither vendored nor installable here, so the
oracle's block dataflow cannot be pinned against reference-held vectors ("parity unpinned", oracle/ltx_oracle.py header).
What CAN be done offline is to derive every upstream constant a second time, from the published definition, in a
different style (scalar float64 loops, no torch, no code shared with ``ltx_oracle.py``), and to check the two against
each other (tests/test_oracle_golden.py::test_independent_derivation_matches_oracle).  A transcription slip in one of the
two (cos/sin order, padding side, frequency spacing, an off-by-one in the sigma table) then shows up as a mismatch.

Provenance of each constant (diffusers 0.32-0.33 sources; symbol = where the default is set):

| constant | value | upstream symbol |
|---|---|---|
| timestep sinusoid width | 256 channels, ``[cos | sin]`` halves | ``PixArtAlphaCombinedTimestepSizeEmbeddings``: ``Timesteps(num_channels=256, flip_sin_to_cos=True, downscale_freq_shift=0)`` |
| sinusoid frequencies | ``exp(-ln(10000) * i / 128)``, i = 0..127 | ``get_timestep_embedding(max_period=10000, scale=1)``: exponent ``/(half_dim - downscale_freq_shift)`` |
| time-embed MLP | Linear(256,D) - SiLU - Linear(D,D); then SiLU - Linear(D,6D) | ``TimestepEmbedding(act_fn="silu")``, ``AdaLayerNormSingle.linear`` |
| caption projection | Linear(4096,D) - GELU(tanh) - Linear(D,D) | ``PixArtAlphaTextProjection(act_fn="gelu_tanh")`` |
| q/k norm | RMSNorm over all heads (D), affine, eps 1e-5 | ``Attention(qk_norm="rms_norm_across_heads", eps=1e-5)`` default ``eps`` |
| block norms | RMSNorm(D, eps 1e-6, no affine) | ``LTXVideoTransformerBlock(norm_eps=1e-6, elementwise_affine=False)``; dims pinned in-repo by ``tests/models/ltx_video/_test_tp.py:29-59,186-245`` |
| final norm | LayerNorm(D, eps 1e-6, no affine) | ``LTXVideoTransformer3DModel.norm_out`` |
| scale_shift_table | block ``[6, D]``, final ``[2, D]``, init ``randn / sqrt(D)`` | ``LTXVideoTransformerBlock.scale_shift_table`` / model ``scale_shift_table`` |
| RoPE theta, bases | theta 10000; base_num_frames 20, base_height 2048, base_width 2048 | ``LTXVideoRotaryPosEmbed.__init__`` defaults |
| RoPE frequencies | ``theta ** linspace(0, 1, D // 6) * pi / 2 * (2 g - 1)``; layout frequency-major then (f,h,w); each value repeated for the pair; ``D % 6`` leading dims padded with cos = 1, sin = 0 | ``LTXVideoRotaryPosEmbed.forward`` |
| RoPE grid scaling | ``g_f = f * (8/25) * patch_t / 20``, ``g_h = h * 32 * patch / 2048``, ``g_w`` likewise | ``rope_interpolation_scale`` built at ``finetrainers/models/ltx_video/base_specification.py:325-334`` (in-repo) |
| flow-match sigmas | ``sigmas[i] = (1000 - i) / 1000``, then a trailing 0 | ``FlowMatchEulerDiscreteScheduler(num_train_timesteps=1000, shift=1.0)`` |
| SDPA scale | ``1 / sqrt(64)`` | ``F.scaled_dot_product_attention`` default |
| key-mask bias | ``(1 - mask) * -10000`` | in-repo: ``finetrainers/patches/models/ltx_video/patch.py:55-57`` |
| LoRA | ``y = W x + b + (alpha / r) B A x``, A kaiming-uniform(a = sqrt 5), B = 0 | peft ``LoraLayer.update_layer`` / ``Linear.forward``; policy in-repo ``trainer.py:120-136`` |
"""
import math


def sinusoid_256(t: float):
    """256 numbers: cos(t f_i) for i = 0..127 followed by sin(t f_i)."""
    out = [0.0] * 256
    for i in range(128):
        f = math.exp(-math.log(10000.0) * i / 128.0)
        out[i] = math.cos(t * f)
        out[128 + i] = math.sin(t * f)
    return out


def rope_entry(f: int, h: int, w: int, col: int, dim: int, scale_f: float, scale_h: float, scale_w: float,
               theta: float = 10000.0, base_f: float = 20.0, base_h: float = 2048.0, base_w: float = 2048.0):
    """(cos, sin) applied to channel ``col`` of the token at latent position (f, h, w)."""
    nf = dim // 6
    pad = dim % 6
    if col < pad:
        return 1.0, 0.0
    pair = (col - pad) // 2                 # every rotary value serves two consecutive channels
    freq_index, axis = divmod(pair, 3)      # frequency-major, then the (f, h, w) axis
    g = (f * scale_f / base_f, h * scale_h / base_h, w * scale_w / base_w)[axis]
    x = freq_index / (nf - 1)               # linspace(0, 1, nf)
    ang = (theta ** x) * (math.pi / 2.0) * (2.0 * g - 1.0)
    return math.cos(ang), math.sin(ang)


def flow_match_sigmas(n: int = 1000):
    return [(n - i) / n for i in range(n)] + [0.0]


def gelu_tanh(x: float) -> float:
    return 0.5 * x * (1.0 + math.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def rms_norm(row, weight=None, eps=1e-6):
    ms = sum(v * v for v in row) / len(row)
    r = 1.0 / math.sqrt(ms + eps)
    return [v * r * (weight[i] if weight is not None else 1.0) for i, v in enumerate(row)]
