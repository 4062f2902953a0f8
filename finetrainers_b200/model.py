"""B200-native LTX-Video DiT: the ``torch.nn.Module`` finetrainers' ``LTXVideoModelSpecification.forward`` calls
(``/root/reference/finetrainers/models/ltx_video/base_specification.py:336-342``), re-implemented as one autograd node
whose forward AND backward are sequences of libb2d kernels (tcgen05 GEMMs with fused epilogues, fused norm/modulate,
q/k-norm + RoPE, tcgen05 attention) instead of the diffusers module graph
(``/root/reference/finetrainers/patches/models/ltx_video/patch.py:38-127`` + diffusers ``LTXVideoTransformerBlock``).

* Parameter FQNs are diffusers/peft compatible (``transformer_blocks.0.attn1.to_q.lora_A.default.weight`` ...), so
  ``state_dict`` / LoRA export (``base_specification.py:379-397``) keep working.  The parameters are views into packed
  device buffers (``[Wq;Wk;Wv]``, flat fp32 LoRA master / grad buffers) so the kernels see fused operands with no copies.
* All block activations needed by backward are kept resident (≈5.5 GB at 49x512x768, B=1: trivial on 180 GB HBM3e), so
  there is NO recompute pass, unlike the reference's ``checkpoint_wrapper`` (``utils/activation_checkpoint.py:40-49``).
* The timestep embedding is evaluated on the B distinct timesteps, not on B*S rows (``patch.py:67-79`` flattens B*S).
* LoRA (peft semantics: ``y = Wx + b + (alpha/r) B A x``) runs in the same tcgen05 accumulator as the base GEMM
  (K-extension operands); master weights and gradients are fp32 (``trainer.py:130-136``), GEMM operands bf16.
* There is no fallback: without libb2d.so / an sm_100 device every call raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, asdict
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

LORA_TARGETS = ("to_q", "to_k", "to_v", "to_out.0")  # examples/training/sft/ltx_video/crush_smol_lora/train.sh:77


@dataclass
class LTXConfig:
    """``tests/models/ltx_video/_test_tp.py:29-59`` (real size)."""
    in_channels: int = 128
    out_channels: int = 128
    patch_size: int = 1
    patch_size_t: int = 1
    num_attention_heads: int = 32
    attention_head_dim: int = 64
    cross_attention_dim: int = 2048
    num_layers: int = 28
    caption_channels: int = 4096
    norm_eps: float = 1e-6
    qk_norm_eps: float = 1e-5
    ffn_mult: int = 4

    @property
    def inner_dim(self) -> int:
        return self.num_attention_heads * self.attention_head_dim

    def to_dict(self):
        return asdict(self)


class ParamLinear(nn.Module):
    """Parameter container with nn.Linear's attribute names; the math happens in libb2d."""

    def __init__(self, in_features, out_features, bias=True, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(torch.empty(out_features, in_features, dtype=dtype, device=device))
        self.bias = nn.Parameter(torch.empty(out_features, dtype=dtype, device=device)) if bias else None

    def forward(self, *a, **k):
        raise RuntimeError("ParamLinear holds parameters only; the b200 engine executes the fused step")


class LoraLinear(nn.Module):
    """peft-compatible naming: base_layer / lora_A.default / lora_B.default (fp32 adapters)."""

    def __init__(self, base: ParamLinear, r: int, alpha: float):
        super().__init__()
        dev = base.weight.device
        self.base_layer = base
        self.lora_A = nn.ModuleDict({"default": ParamLinear(base.in_features, r, False, torch.float32, dev)})
        self.lora_B = nn.ModuleDict({"default": ParamLinear(r, base.out_features, False, torch.float32, dev)})
        self.r, self.lora_alpha, self.scaling = r, alpha, alpha / r
        bound = 1.0 / math.sqrt(base.in_features)  # kaiming_uniform_(a=sqrt(5)) == U(-1/sqrt(fan_in), +)
        with torch.no_grad():
            self.lora_A["default"].weight.uniform_(-bound, bound)
            self.lora_B["default"].weight.zero_()


class _NormW(nn.Module):
    def __init__(self, dim, dtype, device):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(dim, dtype=dtype, device=device))


class _Attn(nn.Module):
    def __init__(self, cfg: LTXConfig, dtype, device):
        super().__init__()
        d = cfg.inner_dim
        self.norm_q = _NormW(d, dtype, device)
        self.norm_k = _NormW(d, dtype, device)
        self.to_q = ParamLinear(d, d, True, dtype, device)
        self.to_k = ParamLinear(d, d, True, dtype, device)
        self.to_v = ParamLinear(d, d, True, dtype, device)
        self.to_out = nn.ModuleList([ParamLinear(d, d, True, dtype, device), nn.Dropout(0.0)])


class _GELUProj(nn.Module):
    def __init__(self, d_in, d_out, dtype, device):
        super().__init__()
        self.proj = ParamLinear(d_in, d_out, True, dtype, device)


class _FF(nn.Module):
    def __init__(self, d, mult, dtype, device):
        super().__init__()
        self.net = nn.ModuleList([_GELUProj(d, d * mult, dtype, device), nn.Dropout(0.0),
                                  ParamLinear(d * mult, d, True, dtype, device)])


class _Block(nn.Module):
    def __init__(self, cfg: LTXConfig, dtype, device):
        super().__init__()
        d = cfg.inner_dim
        self.norm1 = nn.Identity()
        self.attn1 = _Attn(cfg, dtype, device)
        self.norm2 = nn.Identity()
        self.attn2 = _Attn(cfg, dtype, device)
        self.ff = _FF(d, cfg.ffn_mult, dtype, device)
        self.scale_shift_table = nn.Parameter(torch.empty(6, d, dtype=dtype, device=device))


class _TimestepEmbedder(nn.Module):
    def __init__(self, d, dtype, device):
        super().__init__()
        self.linear_1 = ParamLinear(256, d, True, dtype, device)
        self.linear_2 = ParamLinear(d, d, True, dtype, device)


class _Emb(nn.Module):
    def __init__(self, d, dtype, device):
        super().__init__()
        self.timestep_embedder = _TimestepEmbedder(d, dtype, device)


class _AdaSingle(nn.Module):
    def __init__(self, d, dtype, device):
        super().__init__()
        self.emb = _Emb(d, dtype, device)
        self.linear = ParamLinear(d, 6 * d, True, dtype, device)


class _TextProj(nn.Module):
    def __init__(self, c, d, dtype, device):
        super().__init__()
        self.linear_1 = ParamLinear(c, d, True, dtype, device)
        self.linear_2 = ParamLinear(d, d, True, dtype, device)


def _base(m):
    return m.base_layer if isinstance(m, LoraLinear) else m


class _StepFn(torch.autograd.Function):
    """One autograd node for the whole 28-block stack (forward and hand-written backward)."""

    @staticmethod
    def forward(ctx, anchor, model, hidden_states, encoder_hidden_states, tvals, key_bias, num_frames, height, width,
                rope_scale):
        ctx.model = model
        out = model._forward_impl(hidden_states, encoder_hidden_states, tvals, key_bias, num_frames, height, width,
                                  rope_scale)
        ctx.gen = model._fwd_gen
        return out

    @staticmethod
    def backward(ctx, dpred):
        if ctx.gen != ctx.model._fwd_gen:
            raise RuntimeError("B200LTXTransformer keeps ONE set of saved activations: another forward ran between this "
                               "forward and its backward (validation pass, second micro-batch, ...); call backward first")
        ctx.model._backward_impl(dpred)
        return (None,) * 10


class B200LTXTransformer(nn.Module):
    def __init__(self, cfg: Optional[LTXConfig] = None, dtype=torch.bfloat16, device="cuda"):
        super().__init__()
        cfg = cfg or LTXConfig()
        if cfg.attention_head_dim != 64:
            raise ValueError("b200 attention kernels are specialised for attention_head_dim == 64")
        if cfg.patch_size != 1 or cfg.patch_size_t != 1:
            raise ValueError("LTX-Video uses patch_size = patch_size_t = 1")
        if cfg.cross_attention_dim != cfg.inner_dim:
            raise ValueError("cross_attention_dim must equal inner_dim (LTX-Video)")
        self.cfg = cfg
        self.config = cfg  # diffusers-style attribute
        d = cfg.inner_dim
        self.proj_in = ParamLinear(cfg.in_channels, d, True, dtype, device)
        self.scale_shift_table = nn.Parameter(torch.empty(2, d, dtype=dtype, device=device))
        self.time_embed = _AdaSingle(d, dtype, device)
        self.caption_projection = _TextProj(cfg.caption_channels, d, dtype, device)
        self.transformer_blocks = nn.ModuleList([_Block(cfg, dtype, device) for _ in range(cfg.num_layers)])
        self.norm_out = nn.Identity()
        self.proj_out = ParamLinear(d, cfg.out_channels, True, dtype, device)
        self.gradient_checkpointing = False  # accepted for API compatibility; nothing is recomputed
        self.lora_rank = 0
        self.lora_scaling = 1.0
        self._prepared = False
        self._ws: Dict[Tuple, Dict[str, torch.Tensor]] = {}
        self._rope: Dict[Tuple, Tuple[torch.Tensor, torch.Tensor]] = {}
        self._anchor = torch.zeros((), dtype=torch.float32, device=device, requires_grad=True)
        self._saved_key = None
        self._fsdp = None  # fsdp.FSDPState once apply_fsdp2 ran
        self._fwd_gen = 0
        self.skip_block0_dx = True

    # ------------------------------------------------------------------------------------------------
    # adapters / packing
    # ------------------------------------------------------------------------------------------------
    def _resolve_lora_targets(self, target_modules) -> List[str]:
        """peft's matching rule (``peft.tuners.tuners_utils.check_target_module_exists``): a ``str`` is a regex that must
        FULL-match the module name, a list matches by exact name or ``.``-suffix.  Returns the matched linear FQNs."""
        import re
        names = [n for n, m in self.named_modules() if isinstance(m, ParamLinear) and ".lora_" not in n]
        names = [n[:-len(".base_layer")] if n.endswith(".base_layer") else n for n in names]
        if isinstance(target_modules, str):
            return [n for n in names if re.fullmatch(target_modules, n)]
        tm = list(target_modules)
        return [n for n in names if any(n == t or n.endswith("." + t) for t in tm)]

    def add_adapter(self, adapter_config=None, lora_alpha: Optional[float] = None, target_modules=None,
                    adapter_name: str = "default"):
        """``transformer.add_adapter(LoraConfig(r=, lora_alpha=, init_lora_weights=True, target_modules=))`` exactly as
        ``SFTTrainer._prepare_trainable_parameters`` calls it (trainer.py:120-128): the first positional argument is a
        peft-style config object (anything with ``.r``, ``.lora_alpha``, ``.target_modules`` and optionally
        ``.init_lora_weights``).  ``add_adapter(64, 64)`` (rank, alpha) is kept as a shorthand.  ``target_modules`` follows
        peft's matching rule and must select exactly the attention projections the engine fuses
        (to_q|to_k|to_v|to_out.0 of attn1 + attn2 in every block: the reference default regex, config.py:26)."""
        init = True
        if adapter_config is None:
            rank = 64
        elif hasattr(adapter_config, "r"):
            rank = int(adapter_config.r)
            if lora_alpha is None:
                lora_alpha = getattr(adapter_config, "lora_alpha", None)
            if target_modules is None:
                target_modules = getattr(adapter_config, "target_modules", None)
            init = getattr(adapter_config, "init_lora_weights", True)
            if getattr(adapter_config, "lora_dropout", 0.0):
                raise NotImplementedError("lora_dropout > 0 is not supported (the reference never sets it)")
        else:
            rank = int(adapter_config)
        if adapter_name != "default":
            raise NotImplementedError("only the 'default' adapter name is supported")
        if self.lora_rank:
            raise ValueError("an adapter is already attached")
        if rank <= 0:
            raise ValueError("LoRA rank must be positive")
        if init not in (True, "gaussian"):
            raise NotImplementedError(f"init_lora_weights={init!r}: only True (kaiming-uniform A, zero B) and 'gaussian'")
        if target_modules is None:
            target_modules = LORA_TARGETS
        if isinstance(target_modules, (set, frozenset)):
            target_modules = sorted(target_modules)
        want = sorted(f"transformer_blocks.{i}.{a}.{t}" for i in range(len(self.transformer_blocks))
                      for a in ("attn1", "attn2") for t in LORA_TARGETS)
        got = sorted(self._resolve_lora_targets(target_modules))
        if got != want:
            extra = [n for n in got if n not in want][:4]
            missing = [n for n in want if n not in got][:4]
            raise NotImplementedError("b200 engine fuses LoRA on to_q|to_k|to_v|to_out.0 of attn1+attn2 of every block; "
                                      f"target_modules selects a different set (extra: {extra}, missing: {missing})")
        alpha = float(lora_alpha if lora_alpha is not None else rank)
        for p in self.parameters():
            p.requires_grad_(False)
        for blk in self.transformer_blocks:
            for attn in (blk.attn1, blk.attn2):
                attn.to_q = LoraLinear(attn.to_q, rank, alpha)
                attn.to_k = LoraLinear(attn.to_k, rank, alpha)
                attn.to_v = LoraLinear(attn.to_v, rank, alpha)
                attn.to_out[0] = LoraLinear(attn.to_out[0], rank, alpha)
        if init == "gaussian":  # peft: A ~ N(0, 1/r), B = 0
            with torch.no_grad():
                for n, p in self.named_parameters():
                    if "lora_A" in n:
                        p.normal_(0.0, 1.0 / rank)
        self.lora_rank = rank
        self.lora_scaling = alpha / rank
        self.peft_config = {adapter_name: adapter_config if hasattr(adapter_config, "r") else None}
        self._lora_init = init
        self._lora_targets = target_modules
        self._prepared = False

    def _apply(self, fn, recurse=True):
        """``.to()`` / ``.cuda()`` / ``.float()`` replace parameter storage when the dtype or device changes, which would
        silently detach the parameters from the packed buffers the kernels read.  Re-pack in that case (values are taken
        from the moved parameters; the fp32 LoRA masters stay fp32, the reference's own policy under DDP,
        trainer.py:130-136)."""
        was = getattr(self, "_prepared", False)
        probe = self.proj_in.weight
        probes = [self.proj_in.weight]
        if len(self.transformer_blocks):
            blk = self.transformer_blocks[0]
            probes.append(blk.ff.net[2].weight)
            if self.lora_rank:
                probes.append(blk.attn1.to_q.lora_A["default"].weight)
        before = [(q.data_ptr(), q.dtype) for q in probes]
        stash = self.lora_flat.clone() if (was and self.lora_rank) else None  # a dtype cast must not round the fp32 masters
        out = super()._apply(fn, recurse)
        if self._anchor.device != probe.device:
            self._anchor = torch.zeros((), dtype=torch.float32, device=probe.device, requires_grad=True)
        if was and before != [(q.data_ptr(), q.dtype) for q in probes]:
            self._prepared = False
            self._ws.clear()
            self._rope.clear()
            self.prepare()
            if stash is not None:
                self.lora_flat.copy_(stash.to(self.lora_flat.device))
        return out

    def lora_parameters(self) -> List[nn.Parameter]:
        return [p for n, p in self.named_parameters() if "lora_" in n]

    def lora_state_dict(self) -> Dict[str, torch.Tensor]:
        """What ``peft.get_peft_model_state_dict`` returns for the adapters (adapter name stripped), contiguous CPU
        tensors — the ``transformer_state_dict`` finetrainers hands to ``LTXPipeline.save_lora_weights``
        (``base_specification.py:379-397``, ``trainer.py:279-306``)."""
        return {n.replace(".default.weight", ".weight"): p.detach().to("cpu").contiguous().clone()
                for n, p in self.named_parameters() if "lora_" in n}

    def save_lora_weights(self, directory: str, metadata: Optional[Dict[str, str]] = None) -> str:
        """Writes ``pytorch_lora_weights.safetensors`` with diffusers' ``transformer.`` key prefix (loadable with
        ``pipe.load_lora_weights``), plus the LoRA config as metadata like the reference's save hook."""
        import json
        import os
        from safetensors.torch import save_file
        os.makedirs(directory, exist_ok=True)
        sd = {"transformer." + k: v for k, v in self.lora_state_dict().items()}
        tm = getattr(self, "_lora_targets", LORA_TARGETS)
        # same keys, order and formatting as the reference's save hook (trainer.py:284-290)
        meta = {"format": "pt", "lora_config": json.dumps({"r": self.lora_rank, "lora_alpha": self.lora_rank * self.lora_scaling,
                                                           "init_lora_weights": getattr(self, "_lora_init", True),
                                                           "target_modules": tm if isinstance(tm, str) else list(tm)},
                                                          indent=4)}
        meta.update(metadata or {})
        path = os.path.join(directory, "pytorch_lora_weights.safetensors")
        save_file(sd, path, metadata=meta)
        return path

    # ---- flat-buffer layouts -----------------------------------------------------------------------------------------
    def _block_specs(self):
        """(key, shape) of the tensors of ONE block's flat unit, in storage order (every size is a multiple of 8 elements,
        so every view starts 16-byte aligned as TMA requires)."""
        d, f = self.cfg.inner_dim, self.cfg.ffn_mult * self.cfg.inner_dim
        return [("Wqkv", (3 * d, d)), ("bqkv", (3 * d,)), ("Wo", (d, d)), ("bo", (d,)), ("Wq2", (d, d)), ("bq2", (d,)),
                ("Wo2", (d, d)), ("bo2", (d,)), ("W1", (f, d)), ("b1", (f,)), ("W2", (d, f)), ("b2", (d,)),
                ("nq1", (d,)), ("nk1", (d,)), ("nq2", (d,)), ("sst", (6, d))]

    def _block_params(self, blk):
        """(key, [parameters packed into that view, in order]) for one block."""
        a1, a2 = blk.attn1, blk.attn2
        return [("Wqkv", [_base(a1.to_q).weight, _base(a1.to_k).weight, _base(a1.to_v).weight]),
                ("bqkv", [_base(a1.to_q).bias, _base(a1.to_k).bias, _base(a1.to_v).bias]),
                ("Wo", [_base(a1.to_out[0]).weight]), ("bo", [_base(a1.to_out[0]).bias]),
                ("Wq2", [_base(a2.to_q).weight]), ("bq2", [_base(a2.to_q).bias]),
                ("Wo2", [_base(a2.to_out[0]).weight]), ("bo2", [_base(a2.to_out[0]).bias]),
                ("W1", [blk.ff.net[0].proj.weight]), ("b1", [blk.ff.net[0].proj.bias]),
                ("W2", [blk.ff.net[2].weight]), ("b2", [blk.ff.net[2].bias]),
                ("nq1", [a1.norm_q.weight]), ("nk1", [a1.norm_k.weight]), ("nq2", [a2.norm_q.weight]),
                ("sst", [blk.scale_shift_table])]

    def _root_specs(self):
        cfg = self.cfg
        d, nl = cfg.inner_dim, cfg.num_layers
        return [("proj_in.w", (d, cfg.in_channels)), ("proj_in.b", (d,)), ("t1.w", (d, 256)), ("t1.b", (d,)),
                ("t2.w", (d, d)), ("t2.b", (d,)), ("ada.w", (6 * d, d)), ("ada.b", (6 * d,)),
                ("c1.w", (d, cfg.caption_channels)), ("c1.b", (d,)), ("c2.w", (d, d)), ("c2.b", (d,)),
                ("sst", (2, d)), ("proj_out.w", (cfg.out_channels, d)), ("proj_out.b", (cfg.out_channels,)),
                ("Wkv2_all", (nl, 2 * d, d)), ("bkv2_all", (nl, 2 * d)), ("nk2_all", (nl, d))]

    def _root_params(self):
        te, cp = self.time_embed, self.caption_projection
        kw, kb, kn = [], [], []
        for blk in self.transformer_blocks:
            kw += [_base(blk.attn2.to_k).weight, _base(blk.attn2.to_v).weight]
            kb += [_base(blk.attn2.to_k).bias, _base(blk.attn2.to_v).bias]
            kn.append(blk.attn2.norm_k.weight)
        return [("proj_in.w", [self.proj_in.weight]), ("proj_in.b", [self.proj_in.bias]),
                ("t1.w", [te.emb.timestep_embedder.linear_1.weight]), ("t1.b", [te.emb.timestep_embedder.linear_1.bias]),
                ("t2.w", [te.emb.timestep_embedder.linear_2.weight]), ("t2.b", [te.emb.timestep_embedder.linear_2.bias]),
                ("ada.w", [te.linear.weight]), ("ada.b", [te.linear.bias]),
                ("c1.w", [cp.linear_1.weight]), ("c1.b", [cp.linear_1.bias]), ("c2.w", [cp.linear_2.weight]),
                ("c2.b", [cp.linear_2.bias]), ("sst", [self.scale_shift_table]),
                ("proj_out.w", [self.proj_out.weight]), ("proj_out.b", [self.proj_out.bias]),
                ("Wkv2_all", kw), ("bkv2_all", kb), ("nk2_all", kn)]

    FLAT_ALIGN = 2048  # elements: every flat unit is padded so that it splits evenly over up to 8 ranks in 16-byte pieces

    @classmethod
    def _flat_numel(cls, specs):
        n = sum((math.prod(shape) + 7) // 8 * 8 for _, shape in specs)
        return (n + cls.FLAT_ALIGN - 1) // cls.FLAT_ALIGN * cls.FLAT_ALIGN

    @staticmethod
    def _carve(flat, specs):
        out, o = {}, 0
        for key, shape in specs:
            n = math.prod(shape)
            out[key] = flat[o:o + n].view(shape)
            o += (n + 7) // 8 * 8
        return out

    @torch.no_grad()
    def prepare(self):
        """Pack weights into the fused layouts the kernels consume and re-point the parameters into them."""
        cfg = self.cfg
        d = cfg.inner_dim
        r = self.lora_rank
        rp = ((r + 63) // 64) * 64 if r else 0
        self.rpad = rp
        dev = self.proj_in.weight.device
        self._blk = []
        # flat fp32 LoRA master + grad (padded rank) and bf16 operand copy
        per_blk = (8 * rp * d) * 2 if r else 0  # A:[3rp+rp+rp+2rp+rp, d] ; B:[(3+1+1+2+1) d, rp]
        self._per_blk = per_blk
        nl = cfg.num_layers
        if r:
            self.lora_flat = torch.zeros(nl * per_blk, dtype=torch.float32, device=dev)
            self.lora_grad_flat = torch.zeros_like(self.lora_flat)
            self.lora_bf16 = torch.zeros(nl * per_blk, dtype=torch.bfloat16, device=dev)
        # the text-side K/V projection of cross attention reads only the caption embedding, so all blocks' [Wk2;Wv2], biases
        # and norm_k weights are stacked: one batched launch per step instead of one per block
        wdt = self.proj_in.weight.dtype
        # ---- base weights: ONE flat buffer per DiT block (the FSDP-2 sharding unit, ptd.py:482-499) plus one "root" flat
        # buffer for everything outside the blocks; the module parameters become views of that storage
        self._blk_flat = []
        root_specs = self._root_specs()
        self._root_flat = torch.empty(self._flat_numel(root_specs), dtype=wdt, device=dev)
        root_views = self._carve(self._root_flat, root_specs)
        for (key, _), (_, params) in zip(root_specs, self._root_params()):
            v, o = root_views[key], 0
            for prm in params:
                n = prm.numel()
                seg = v.reshape(-1)[o:o + n].view(prm.shape)
                seg.copy_(prm.data)
                prm.data = seg
                o += n
        self._Wkv2_all, self._bkv2_all, self._nk2_all = root_views["Wkv2_all"], root_views["bkv2_all"], root_views["nk2_all"]
        self._root_views = root_views
        specs = self._block_specs()
        for li, blk in enumerate(self.transformer_blocks):
            a1, a2 = blk.attn1, blk.attn2
            flat = torch.empty(self._flat_numel(specs), dtype=wdt, device=dev)
            e = self._carve(flat, specs)
            for key, params in self._block_params(blk):
                v, o = e[key], 0
                for prm in params:
                    n = prm.numel()
                    seg = v.reshape(-1)[o:o + n].view(prm.shape)
                    seg.copy_(prm.data)
                    prm.data = seg
                    o += n
            self._blk_flat.append(flat)
            # the text-side K/V projection weights of every block live (stacked) in the root unit
            e["Wkv2"], e["bkv2"], e["nk2"] = self._Wkv2_all[li], self._bkv2_all[li], self._nk2_all[li]
            if r:
                base = li * per_blk
                off = [base]

                def carve(rows, cols):
                    n = rows * cols
                    s = off[0]
                    off[0] += n
                    return (self.lora_flat[s:s + n].view(rows, cols), self.lora_grad_flat[s:s + n].view(rows, cols),
                            self.lora_bf16[s:s + n].view(rows, cols))

                groups = {"qkv": [a1.to_q, a1.to_k, a1.to_v], "o": [a1.to_out[0]], "q2": [a2.to_q],
                          "kv2": [a2.to_k, a2.to_v], "o2": [a2.to_out[0]]}
                for gname, mods in groups.items():
                    n_ad = len(mods)
                    A, gA, Ab = carve(n_ad * rp, d)
                    Bm, gB, Bb = carve(n_ad * d, rp)
                    for j, m in enumerate(mods):
                        A[j * rp:j * rp + r].copy_(m.lora_A["default"].weight.data)
                        Bm[j * d:(j + 1) * d, :r].copy_(m.lora_B["default"].weight.data)
                        m.lora_A["default"].weight.data = A[j * rp:j * rp + r]
                        m.lora_B["default"].weight.data = Bm[j * d:(j + 1) * d, :r]
                        m.lora_A["default"].weight.grad = gA[j * rp:j * rp + r]
                        m.lora_B["default"].weight.grad = gB[j * d:(j + 1) * d, :r]
                    e["A_" + gname], e["gA_" + gname], e["Ab_" + gname] = A, gA, Ab
                    e["B_" + gname], e["gB_" + gname], e["Bb_" + gname] = Bm, gB, Bb
            self._blk.append(e)
        self._prepared = True
        self._ws.clear()
        return self

    @torch.no_grad()
    def _rebind_flat_storage(self, block_flats, root_flat):
        """FSDP-2: move the base-weight views of every block onto the given full-size buffers (gather slots shared by
        several blocks) and of the root unit onto ``root_flat``, then drop the private per-block storage.  The buffers'
        contents are only valid while the owning unit is resident (fsdp.FSDPState schedules that)."""
        specs = self._block_specs()
        for li, (blk, flat) in enumerate(zip(self.transformer_blocks, block_flats)):
            views = self._carve(flat, specs)
            for key, params in self._block_params(blk):
                v, o = views[key], 0
                for prm in params:
                    n = prm.numel()
                    prm.data = v.reshape(-1)[o:o + n].view(prm.shape)
                    o += n
            self._blk[li].update(views)
        rv = self._carve(root_flat, self._root_specs())
        for (key, _), (_, params) in zip(self._root_specs(), self._root_params()):
            v, o = rv[key], 0
            for prm in params:
                n = prm.numel()
                prm.data = v.reshape(-1)[o:o + n].view(prm.shape)
                o += n
        self._Wkv2_all, self._bkv2_all, self._nk2_all = rv["Wkv2_all"], rv["bkv2_all"], rv["nk2_all"]
        for li in range(len(self._blk)):
            self._blk[li]["Wkv2"], self._blk[li]["bkv2"], self._blk[li]["nk2"] = self._Wkv2_all[li], self._bkv2_all[li], self._nk2_all[li]
        self._root_views = rv
        self._blk_flat = None
        self._root_flat = None

    def _attach_lora_grads(self):
        """(Re-)attach .grad views after an external ``zero_grad(set_to_none=True)``; returns True if any was missing."""
        missing = False
        for e, blk in zip(self._blk, self.transformer_blocks):
            a1, a2 = blk.attn1, blk.attn2
            groups = {"qkv": [a1.to_q, a1.to_k, a1.to_v], "o": [a1.to_out[0]], "q2": [a2.to_q],
                      "kv2": [a2.to_k, a2.to_v], "o2": [a2.to_out[0]]}
            d, r, rp = self.cfg.inner_dim, self.lora_rank, self.rpad
            for gname, mods in groups.items():
                for j, m in enumerate(mods):
                    pa, pb = m.lora_A["default"].weight, m.lora_B["default"].weight
                    if pa.grad is None or pb.grad is None:
                        missing = True
                        pa.grad = e["gA_" + gname][j * rp:j * rp + r]
                        pb.grad = e["gB_" + gname][j * d:(j + 1) * d, :r]
        return missing

    # ------------------------------------------------------------------------------------------------
    # workspace
    # ------------------------------------------------------------------------------------------------
    def _workspace(self, B, S, L):
        key = (B, S, L)
        ws = self._ws.get(key)
        if ws is not None:
            return ws
        cfg = self.cfg
        d, H, nl, rp = cfg.inner_dim, cfg.num_attention_heads, cfg.num_layers, self.rpad
        R, RL = B * S, B * L
        dev = self.proj_in.weight.device
        bf = dict(dtype=torch.bfloat16, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        ws = {}

        def z(name, *shape, kw=bf):
            ws[name] = torch.zeros(*shape, **kw)

        # embeds
        z("tsin", B, 256); z("t1", B, d); z("t2s", B, d); z("embedded", B, d); z("temb", B, 6 * d)
        z("c1", RL, d); z("enc", RL, d)
        # per-block saved activations
        z("h", nl + 1, R, d)            # h[l] = input of block l; h[nl] = final hidden
        z("n1", nl, R, d); z("qkv", nl, R, 3 * d)
        z("qh", nl, B, H, S, 64); z("kh", nl, B, H, S, 64); z("vh", nl, B, H, S, 64)
        z("ao", nl, R, d); z("lse", nl, B, H, S, kw=f32)
        z("h1", nl, R, d); z("q2", nl, R, d); z("q2h", nl, B, H, S, 64)
        z("kv2", nl, RL, 2 * d); z("k2h", nl, B, H, L, 64); z("v2h", nl, B, H, L, 64)
        z("ao2", nl, R, d); z("lse2", nl, B, H, S, kw=f32)
        z("h2", nl, R, d); z("ffpre", nl, R, cfg.ffn_mult * d)
        if rp:
            z("u_qkv", nl, R, 3 * rp); z("u_o", nl, R, rp); z("u_q2", nl, R, rp); z("u_kv2", nl, RL, 2 * rp)
            z("u_o2", nl, R, rp)
            # per-block copies of every adapter's output gradient dy and of du = s*dy*B: the weight gradients dA/dB of all
            # 28 blocks are computed at the END of backward as a handful of block-batched GEMMs (1.9 GB at B=1)
            z("dy_o2", nl, R, d); z("dy_q2", nl, R, d); z("dy_kv2", nl, RL, 2 * d); z("dy_o", nl, R, d)
            z("dy_qkv", nl, R, 3 * d)
            z("du_o2", nl, R, rp); z("du_q2", nl, R, rp); z("du_kv2", nl, RL, 2 * rp); z("du_o", nl, R, rp)
            z("du_qkv", nl, R, 3 * rp)
        # scratch shared by all blocks
        z("n2", R, d); z("f", R, cfg.ffn_mult * d); z("y", R, d); z("pred", R, cfg.out_channels)
        z("dh", R, d); z("g", R, d); z("dwide", R, cfg.ffn_mult * d); z("dn", R, d); z("da", R, d)
        z("dqh", B, H, S, 64); z("dkh", B, H, S, 64); z("dvh", B, H, S, 64)
        z("dk2h", nl, B, H, L, 64); z("dv2h", nl, B, H, L, 64)   # kept per block: one batched norm-bwd at the end
        z("delta", max(ops.attn_bwd_ws_floats(B, H, S, S), ops.attn_bwd_ws_floats(B, H, S, L)), kw=f32)
        self._ws[key] = ws
        return ws

    def _rope_tables(self, Fr, Hh, Ww, rope_scale):
        key = (Fr, Hh, Ww, tuple(float(x) for x in rope_scale))
        t = self._rope.get(key)
        if t is None:
            d = self.cfg.inner_dim
            dev = self.proj_in.weight.device
            cos = torch.empty(Fr * Hh * Ww, d // 2, dtype=torch.float32, device=dev)
            sin = torch.empty_like(cos)
            # diffusers LTXVideoRotaryPosEmbed: grid * scale * patch / base (base_num_frames 20, base_h = base_w = 2048)
            ops.rope_table(cos, sin, Fr, Hh, Ww, d, rope_scale[0] * self.cfg.patch_size_t / 20.0,
                           rope_scale[1] * self.cfg.patch_size / 2048.0, rope_scale[2] * self.cfg.patch_size / 2048.0)
            t = (cos, sin)
            self._rope[key] = t
        return t

    # ------------------------------------------------------------------------------------------------
    # public forward (diffusers signature; patch.py:38-51)
    # ------------------------------------------------------------------------------------------------
    def forward(self, hidden_states, encoder_hidden_states, timestep, encoder_attention_mask=None, num_frames=None,
                height=None, width=None, rope_interpolation_scale=None, return_dict=False, **kwargs):
        if not self._prepared:
            self.prepare()
        B = hidden_states.shape[0]
        # finetrainers passes per-token timesteps that are constant per sample (base_specification.py:319-320)
        tvals = timestep.reshape(B, -1)[:, 0].to(torch.float32).contiguous()
        key_bias = None
        if encoder_attention_mask is not None:
            m = encoder_attention_mask
            if m.ndim == 3:
                m = m[:, 0]
            key_bias = ((1.0 - m.to(torch.float32)) * -10000.0).contiguous()  # patch.py:55-57
        if rope_interpolation_scale is None:
            rope_interpolation_scale = (1.0, 1.0, 1.0)
        out = _StepFn.apply(self._anchor, self, hidden_states, encoder_hidden_states, tvals, key_bias, int(num_frames),
                            int(height), int(width), tuple(float(x) for x in rope_interpolation_scale))
        return (out,)

    # ------------------------------------------------------------------------------------------------
    # forward implementation
    # ------------------------------------------------------------------------------------------------
    def refresh_lora_operands(self):
        """fp32 master -> bf16 GEMM operands (one flat cast per step)."""
        if self.lora_rank:
            ops.cast_f32_bf16(self.lora_flat, self.lora_bf16, self.lora_flat.numel(), 1.0)

    def _lin(self, x, W, bias, out, M, N, K, lora=None, **kw):
        """out = epi(x W^T + bias [+ u B^T]);  lora = (Ab [n*rp,K], Bb [N,rp], u [M,n*rp], n_ad)."""
        if lora is not None:
            Ab, Bb, u, n_ad = lora
            rp = self.rpad
            ops.gemm(x, Ab, u, M=M, N=n_ad * rp, K=K, alpha=self.lora_scaling, tag="lora_u")
            ops.gemm(x, W, out, M=M, N=N, K=K, bias=bias, A2=u, B2=Bb, K2=rp, a2_group_n=(N // n_ad if n_ad > 1 else 0),
                     **kw)
        else:
            ops.gemm(x, W, out, M=M, N=N, K=K, bias=bias, **kw)

    def _forward_impl(self, hidden_states, ehs, tvals, key_bias, Fr, Hh, Ww, rope_scale):
        cfg = self.cfg
        d, H, nl, rp = cfg.inner_dim, cfg.num_attention_heads, cfg.num_layers, self.rpad
        B, S, Cin = hidden_states.shape
        L = ehs.shape[1]
        assert S == Fr * Hh * Ww, "sequence length must equal num_frames*height*width (patch size 1)"
        R, RL = B * S, B * L
        ws = self._workspace(B, S, L)
        self._saved_key = (B, S, L, Fr, Hh, Ww, rope_scale)
        self._fwd_gen += 1
        self._key_bias = key_bias
        cos, sin = self._rope_tables(Fr, Hh, Ww, rope_scale)
        x_in = hidden_states.reshape(R, Cin).to(torch.bfloat16).contiguous()
        ehs2 = ehs.reshape(RL, cfg.caption_channels).to(torch.bfloat16).contiguous()
        self.refresh_lora_operands()
        fs = self._fsdp
        if fs is not None:
            fs.begin_forward()  # all-gather the root unit and the first two blocks (communication stream)
        te = self.time_embed
        # ---- timestep embedding on the B distinct timesteps (K2)
        ops.timestep_sinusoid(tvals, ws["tsin"], B)
        ops.gemm(ws["tsin"], te.emb.timestep_embedder.linear_1.weight, ws["t1"], M=B, N=d, K=256,
                 bias=te.emb.timestep_embedder.linear_1.bias, epi=ops.EPI_SILU)
        ops.gemm(ws["t1"], te.emb.timestep_embedder.linear_2.weight, ws["t2s"], M=B, N=d, K=d,
                 bias=te.emb.timestep_embedder.linear_2.bias, epi=ops.EPI_SILU, out2=ws["embedded"])
        ops.gemm(ws["t2s"], te.linear.weight, ws["temb"], M=B, N=6 * d, K=d, bias=te.linear.bias)
        temb = ws["temb"]
        # ---- caption projection (K3), patch embed (K1)
        cp = self.caption_projection
        ops.gemm(ehs2, cp.linear_1.weight, ws["c1"], M=RL, N=d, K=cfg.caption_channels, bias=cp.linear_1.bias,
                 epi=ops.EPI_GELU)
        ops.gemm(ws["c1"], cp.linear_2.weight, ws["enc"], M=RL, N=d, K=d, bias=cp.linear_2.bias)
        enc = ws["enc"]
        ops.gemm(x_in, self.proj_in.weight, ws["h"][0], M=R, N=d, K=Cin, bias=self.proj_in.bias)
        scale = 1.0 / math.sqrt(cfg.attention_head_dim)
        # ---- cross-attention K/V of ALL blocks (functions of `enc` only): LoRA-down, fused [Wk2;Wv2] projection with the
        # LoRA-up K-extension, and k-norm + head split, each as ONE block-batched launch
        ops.CONTEXT = "f.kv2"
        e0, pb = self._blk[0], self._per_blk
        kv2_all = ws["kv2"].view(nl * RL, 2 * d)
        if rp:
            u_all = ws["u_kv2"].view(nl * RL, 2 * rp)
            ops.gemm(enc, e0["Ab_kv2"], u_all, M=RL, N=2 * rp, K=d, batch=nl, b_boff=(pb // d, 0), c_boff=RL * 2 * rp,
                     alpha=self.lora_scaling, tag="lora_u")
            ops.gemm(enc, self._Wkv2_all.view(nl * 2 * d, d), kv2_all, M=RL, N=2 * d, K=d, bias=self._bkv2_all, batch=nl,
                     b_boff=(2 * d, 0), c_boff=RL * 2 * d, bias_boff=2 * d, A2=u_all, B2=e0["Bb_kv2"], K2=rp, a2_group_n=d,
                     a2_boff_row=RL, b2_boff_row=pb // rp)
        else:
            ops.gemm(enc, self._Wkv2_all.view(nl * 2 * d, d), kv2_all, M=RL, N=2 * d, K=d, bias=self._bkv2_all, batch=nl,
                     b_boff=(2 * d, 0), c_boff=RL * 2 * d, bias_boff=2 * d)
        ops.qkv_norm_rope_fwd(kv2_all, 2 * d, 0, (self._nk2_all, None), 0, None, None, (ws["k2h"], ws["v2h"]), nl * B, L, H,
                              cfg.qk_norm_eps, rows_per_w=RL, w_stride=d)
        for l in range(nl):
            if fs is not None:
                fs.pre_block_forward(l)
            e = self._blk[l]
            sst = e["sst"]
            h_in, n1 = ws["h"][l], ws["n1"][l]
            # K5: RMSNorm + modulate (shift_msa = row 0, scale_msa = row 1)
            ops.CONTEXT = "f.self"
            ops.norm_modulate_fwd(h_in, n1, sst[0], temb[:, 0:], sst[1], temb[:, d:], 6 * d, R, d, S, cfg.norm_eps)
            # K6: fused QKV (+LoRA)
            self._lin(n1, e["Wqkv"], e["bqkv"], ws["qkv"][l], R, 3 * d, d,
                      lora=(e["Ab_qkv"], e["Bb_qkv"], ws["u_qkv"][l], 3) if rp else None)
            # K7: q/k RMSNorm + RoPE + head split
            ops.qkv_norm_rope_fwd(ws["qkv"][l], 3 * d, 0, (e["nq1"], e["nk1"], None), 0b011, cos, sin,
                                  (ws["qh"][l], ws["kh"][l], ws["vh"][l]), B, S, H, cfg.qk_norm_eps)
            # K8: self attention
            ops.attn_fwd(ws["qh"][l], ws["kh"][l], ws["vh"][l], None, ws["ao"][l], ws["lse"][l], B, H, S, S, scale)
            # K9: out proj + gated residual (gate_msa = row 2)
            self._lin(ws["ao"][l], e["Wo"], e["bo"], ws["h1"][l], R, d, d,
                      lora=(e["Ab_o"], e["Bb_o"], ws["u_o"][l], 1) if rp else None,
                      epi=ops.EPI_GATE_RES, res=h_in, gate_table=sst[2], gate_temb=temb[:, 2 * d:], temb_stride=6 * d,
                      rows_per_sample=S)
            # K10: cross attention (no pre-norm, no gate)
            ops.CONTEXT = "f.cross"
            h1 = ws["h1"][l]
            self._lin(h1, e["Wq2"], e["bq2"], ws["q2"][l], R, d, d,
                      lora=(e["Ab_q2"], e["Bb_q2"], ws["u_q2"][l], 1) if rp else None)
            ops.qknorm_rope_fwd(ws["q2"][l], d, 0, e["nq2"], None, None, ws["q2h"][l], B, S, H, True, cfg.qk_norm_eps)
            ops.attn_fwd(ws["q2h"][l], ws["k2h"][l], ws["v2h"][l], key_bias, ws["ao2"][l], ws["lse2"][l], B, H, S, L, scale)
            self._lin(ws["ao2"][l], e["Wo2"], e["bo2"], ws["h2"][l], R, d, d,
                      lora=(e["Ab_o2"], e["Bb_o2"], ws["u_o2"][l], 1) if rp else None,
                      epi=ops.EPI_GATE_RES, res=h1)
            # K11/K12: norm2 + modulate (rows 3,4), FFN with GELU epilogue, gated residual (row 5)
            ops.CONTEXT = "f.ffn"
            h2 = ws["h2"][l]
            ops.norm_modulate_fwd(h2, ws["n2"], sst[3], temb[:, 3 * d:], sst[4], temb[:, 4 * d:], 6 * d, R, d, S,
                                  cfg.norm_eps)
            ops.gemm(ws["n2"], e["W1"], ws["f"], M=R, N=cfg.ffn_mult * d, K=d, bias=e["b1"], epi=ops.EPI_GELU,
                     out2=ws["ffpre"][l], tag="ffn_up")
            ops.gemm(ws["f"], e["W2"], ws["h"][l + 1], M=R, N=d, K=cfg.ffn_mult * d, bias=e["b2"], epi=ops.EPI_GATE_RES,
                     res=h2, gate_table=sst[5], gate_temb=temb[:, 5 * d:], temb_stride=6 * d, rows_per_sample=S)
            if fs is not None:
                fs.post_block_forward(l)  # block l's weights are no longer read: its slot takes block l + 2
        ops.CONTEXT = "f.head"
        # K13: final LayerNorm + modulate (table rows 0 = shift, 1 = scale; embedded_timestep), proj_out
        t2 = self.scale_shift_table.data
        ops.norm_modulate_fwd(ws["h"][nl], ws["y"], t2[0], ws["embedded"], t2[1], ws["embedded"], d, R, d, S, 1e-6, True)
        ops.gemm(ws["y"], self.proj_out.weight, ws["pred"], M=R, N=cfg.out_channels, K=d, bias=self.proj_out.bias)
        ops.CONTEXT = ""
        return ws["pred"].view(B, S, cfg.out_channels)

    # ------------------------------------------------------------------------------------------------
    # backward implementation (LoRA: dX through every op, dW only for adapters)
    # ------------------------------------------------------------------------------------------------
    def _splits(self, tiles, kb):
        sm = 148
        s = max(1, min(kb, sm // max(1, tiles)))
        per = -(-kb // s)
        return -(-kb // per)

    def _lora_du(self, dy, du, e, g, M, N, n_ad):
        """du_j = s * dy_j B_j for the n_ad adapters packed in dy [M, N] -> du [M, n_ad*rp]."""
        rp = self.rpad
        Nj = N // n_ad
        ops.gemm(dy, e["Bb_" + g], du, M=M, N=rp, K=Nj, b_mn=True, batch=n_ad, a_boff=(0, Nj), b_boff=(Nj, 0),
                 c_boff=rp, ldc=n_ad * rp, alpha=self.lora_scaling, tag="lora_du")
        return du

    def _lora_wgrads(self, ws, R, RL, lo=0, hi=None):
        """dA / dB of every adapter of blocks [lo, hi): 13 block-batched split-free GEMMs (dB_j += dy_j^T u_j ;
        dA += du^T x computed as (x^T du)^T), accumulating into the flat fp32 gradient buffer.  The whole model in one go
        at the end of backward, or one block range at a time so that the range's slice of the flat gradient is final -
        and its all-reduce can start - while earlier blocks are still in backward (trainer: DDP overlap)."""
        cfg = self.cfg
        d, nl, rp, pb = cfg.inner_dim, cfg.num_layers, self.rpad, self._per_blk
        hi = nl if hi is None else hi
        nr = hi - lo
        e0 = self._blk[lo]
        # kv2 has no dX consumer, so its du is also produced here, block-batched per adapter
        for j in range(2):
            ops.gemm(ws["dy_kv2"].view(nl * RL, 2 * d)[lo * RL:, j * d:], e0["Bb_kv2"][j * d:],
                     ws["du_kv2"].view(nl * RL, 2 * rp)[lo * RL:, j * rp:],
                     M=RL, N=rp, K=d, lda=2 * d, ldb=rp, ldc=2 * rp, b_mn=True, batch=nr, a_boff=(RL, 0), b_boff=(pb // rp, 0),
                     c_boff=RL * 2 * rp, alpha=self.lora_scaling, tag="lora_du")
        groups = (("qkv", ws["dy_qkv"], ws["n1"], ws["u_qkv"], ws["du_qkv"], 3, R, R),
                  ("o", ws["dy_o"], ws["ao"], ws["u_o"], ws["du_o"], 1, R, R),
                  ("q2", ws["dy_q2"], ws["h1"], ws["u_q2"], ws["du_q2"], 1, R, R),
                  ("kv2", ws["dy_kv2"], ws["enc"], ws["u_kv2"], ws["du_kv2"], 2, RL, 0),
                  ("o2", ws["dy_o2"], ws["ao2"], ws["u_o2"], ws["du_o2"], 1, R, R))
        for g, dy, x, u, du, n_ad, M, x_stride in groups:
            N = n_ad * d
            dy2, u2, du2 = dy.view(nl * M, N), u.view(nl * M, n_ad * rp), du.view(nl * M, n_ad * rp)
            x2 = x.view(-1, d)
            # the contraction runs over the M token rows of ONE block: stacking blocks along that axis is only legal when
            # M is a whole number of 64-row k-blocks (otherwise the k-tail would read the next block's rows, not zeros)
            if M % 64 == 0:
                spans = [(lo, nr)]
            else:
                spans = [(l, 1) for l in range(lo, hi)]
            for (l0, nb) in spans:
                for j in range(n_ad):
                    ops.gemm(dy2[l0 * M:, j * d:], u2[l0 * M:, j * rp:], self._blk[l0]["gB_" + g][j * d:], M=d, N=rp, K=M,
                             lda=N, ldb=n_ad * rp, ldc=rp, a_mn=True, b_mn=True, batch=nb, a_boff=(M, 0), b_boff=(M, 0),
                             c_boff=pb, epi=ops.EPI_F32_ATOMIC, block_n=64 if rp == 64 else 128, tag="lora_dB")
                ops.gemm(x2[l0 * x_stride:], du2[l0 * M:], self._blk[l0]["gA_" + g], M=d, N=n_ad * rp, K=M, lda=d,
                         ldb=n_ad * rp, ldc=d, a_mn=True, b_mn=True, batch=nb, a_boff=(x_stride, 0), b_boff=(M, 0),
                         c_boff=pb, epi=ops.EPI_F32_ATOMIC_T, block_n=64, tag="lora_dA")

    def _backward_impl(self, dpred):
        """The whole backward in one call (autograd path / single graph)."""
        self._backward_head(dpred)
        self._backward_blocks(self.cfg.num_layers - 1, 0)
        self._backward_tail(0, self.cfg.num_layers)
        if self._fsdp is not None:
            self._fsdp.end_backward()

    def _bwd_ctx(self):
        cfg = self.cfg
        B, S, L, Fr, Hh, Ww, rope_scale = self._saved_key
        ws = self._workspace(B, S, L)
        cos, sin = self._rope_tables(Fr, Hh, Ww, rope_scale)
        return cfg, B, S, L, ws, cos, sin

    def _backward_head(self, dpred):
        """proj_out / final LayerNorm+modulate backward: leaves dh (residual-stream gradient) and g (dh x gate_mlp of the
        last block) in the workspace."""
        cfg, B, S, L, ws, cos, sin = self._bwd_ctx()
        d, nl, rp = cfg.inner_dim, cfg.num_layers, self.rpad
        R = B * S
        temb = ws["temb"]
        if not rp:
            raise NotImplementedError("full-rank fine-tuning backward (base dW) is not built yet; use add_adapter()")
        if self._attach_lora_grads():
            self.lora_grad_flat.zero_()
        dp = dpred.reshape(R, cfg.out_channels).to(torch.bfloat16).contiguous()
        # head: dy = dpred Wout ; dh = LN-modulate bwd ; g = dh * gate_mlp(last block)
        ops.gemm(dp, self.proj_out.weight, ws["dn"], M=R, N=d, K=cfg.out_channels, b_mn=True)
        t2 = self.scale_shift_table.data
        last = self._blk[nl - 1]["sst"]
        ops.norm_modulate_bwd(ws["dn"], ws["h"][nl], None, ws["dh"], t2[1], ws["embedded"], d, R, d, S, 1e-6, True)
        # (embedded has stride d, temb stride 6d: the gate of the last block is applied by a separate colscale)
        ops.colscale(ws["dh"], ws["g"], last[5], temb[:, 5 * d:], 6 * d, R, d, S)

    def _backward_blocks(self, l_hi, l_lo):
        """Backward through blocks l_hi, l_hi - 1, ..., l_lo (dX through every op; per-block dy / du of the adapters are
        stored for the batched weight-gradient GEMMs)."""
        cfg, B, S, L, ws, cos, sin = self._bwd_ctx()
        d, H, nl, rp = cfg.inner_dim, cfg.num_attention_heads, cfg.num_layers, self.rpad
        R, RL = B * S, B * L
        key_bias = self._key_bias
        temb = ws["temb"]
        scale = 1.0 / math.sqrt(cfg.attention_head_dim)
        dh, g = ws["dh"], ws["g"]
        fs = self._fsdp
        for l in range(l_hi, l_lo - 1, -1):
            if fs is not None:
                fs.pre_block_backward(l)
            e = self._blk[l]
            sst = e["sst"]
            dh2, dq2, dkv2, dyo, dqkv = ws["dy_o2"][l], ws["dy_q2"][l], ws["dy_kv2"][l], ws["dy_o"][l], ws["dy_qkv"][l]
            # ---- FFN: dfp = (g W2) * gelu'(pre) ; dn2 = dfp W1 ; dh2 = dh + norm_bwd(dn2; h2, scale_mlp=row 4)
            ops.CONTEXT = "b.ffn"
            ops.gemm(g, e["W2"], ws["dwide"], M=R, N=cfg.ffn_mult * d, K=d, b_mn=True, epi=ops.EPI_MUL_DGELU,
                     aux=ws["ffpre"][l])
            ops.gemm(ws["dwide"], e["W1"], ws["dn"], M=R, N=d, K=cfg.ffn_mult * d, b_mn=True)
            ops.norm_modulate_bwd(ws["dn"], ws["h2"][l], dh, dh2, sst[4], temb[:, 4 * d:], 6 * d, R, d, S, cfg.norm_eps)
            # ---- cross attention out-proj (no gate): da2 = dh2 W_o2 + du A
            ops.CONTEXT = "b.cross"
            du = self._lora_du(dh2, ws["du_o2"][l], e, "o2", R, d, 1)
            ops.gemm(dh2, e["Wo2"], ws["da"], M=R, N=d, K=d, b_mn=True, A2=du, B2=e["Ab_o2"], K2=rp)
            ops.attn_bwd(ws["q2h"][l], ws["k2h"][l], ws["v2h"][l], key_bias, ws["ao2"][l], ws["da"], ws["lse2"][l],
                         ws["delta"], ws["dqh"], ws["dk2h"][l], ws["dv2h"][l], B, H, S, L, scale)
            ops.qknorm_rope_bwd(ws["dqh"], ws["q2"][l], d, 0, e["nq2"], None, None, dq2, d, 0, B, S, H, True,
                                cfg.qk_norm_eps)
            du = self._lora_du(dq2, ws["du_q2"][l], e, "q2", R, d, 1)
            # dh1 = dh2 + dq2 W_q2 + du A ; gated copy (gate_msa, row 2) = dy of the self-attention out-proj
            ops.gemm(dq2, e["Wq2"], dh, M=R, N=d, K=d, b_mn=True, A2=du, B2=e["Ab_q2"], K2=rp,
                     epi=ops.EPI_GATE_RES, res=dh2, gate2_table=sst[2], gate2_temb=temb[:, 2 * d:], out2=dyo,
                     temb_stride=6 * d, rows_per_sample=S)
            # ---- self attention out-proj (gated): dattn = g W_o + du A
            ops.CONTEXT = "b.self"
            du = self._lora_du(dyo, ws["du_o"][l], e, "o", R, d, 1)
            ops.gemm(dyo, e["Wo"], ws["da"], M=R, N=d, K=d, b_mn=True, A2=du, B2=e["Ab_o"], K2=rp)
            ops.attn_bwd(ws["qh"][l], ws["kh"][l], ws["vh"][l], None, ws["ao"][l], ws["da"], ws["lse"][l], ws["delta"],
                         ws["dqh"], ws["dkh"], ws["dvh"], B, H, S, S, scale)
            ops.qkv_norm_rope_bwd((ws["dqh"], ws["dkh"], ws["dvh"]), ws["qkv"][l], 3 * d, 0, (e["nq1"], e["nk1"], None),
                                  0b011, cos, sin, dqkv, 3 * d, 0, B, S, H, cfg.qk_norm_eps)
            du = self._lora_du(dqkv, ws["du_qkv"][l], e, "qkv", R, 3 * d, 3)
            if l == 0 and self.skip_block0_dx:
                if fs is not None:
                    fs.post_block_backward(l)
                break  # nothing trainable upstream of block 0's adapters (proj_in / embeds are frozen)
            ops.gemm(dqkv, e["Wqkv"], ws["dn"], M=R, N=d, K=3 * d, b_mn=True, A2=du, B2=e["Ab_qkv"], K2=3 * rp)
            # dh0 = dh1 + norm_bwd(dn1; h_in, scale_msa=row 1) ; g = dh0 * gate_mlp of block l-1
            if fs is not None and l > 0:
                fs.pre_block_backward(l - 1)  # the op below reads block l-1's gate row: its all-gather must have landed
            prev = self._blk[l - 1]["sst"] if l > 0 else None
            ops.norm_modulate_bwd(ws["dn"], ws["h"][l], dh, dh, sst[1], temb[:, d:], 6 * d, R, d, S, cfg.norm_eps,
                                  gate2_tab=prev[5] if l > 0 else None, gate2_emb=temb[:, 5 * d:] if l > 0 else None,
                                  out2=g if l > 0 else None)
            if fs is not None:
                fs.post_block_backward(l)
        ops.CONTEXT = ""

    def _backward_tail(self, lo, hi):
        """Adapter gradients of blocks [lo, hi): the text-side k-norm backward of those blocks in one launch (its output
        only feeds the kv2 adapter gradients), then the block-batched dA / dB GEMMs."""
        cfg, B, S, L, ws, cos, sin = self._bwd_ctx()
        d, H, nl = cfg.inner_dim, cfg.num_attention_heads, cfg.num_layers
        R, RL = B * S, B * L
        ops.CONTEXT = "b.kv2"
        ops.qkv_norm_rope_bwd((ws["dk2h"][lo:hi], ws["dv2h"][lo:hi]), ws["kv2"].view(nl * RL, 2 * d)[lo * RL:hi * RL], 2 * d, 0,
                              (self._nk2_all[lo:hi], None), 0, None, None, ws["dy_kv2"].view(nl * RL, 2 * d)[lo * RL:hi * RL],
                              2 * d, 0, (hi - lo) * B, L, H, cfg.qk_norm_eps, rows_per_w=RL, w_stride=d)
        ops.CONTEXT = "b.wgrad"
        self._lora_wgrads(ws, R, RL, lo, hi)
        ops.CONTEXT = ""
