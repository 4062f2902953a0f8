"""``LTXVideoModelSpecification`` hot-path mirror: the ``forward`` contract finetrainers' ``SFTTrainer`` calls
(``/root/reference/finetrainers/models/modeling_utils.py:183-186``;
LTX: ``/root/reference/finetrainers/models/ltx_video/base_specification.py:271-345``), same argument names, same dict
mutation (``pop`` of latents / latents_mean / latents_std, insertion of ``hidden_states``), same return triple
``(pred, target, sigmas)``.  Normalise + noising + packing + target run as ONE libb2d kernel (K15) instead of ~12 ATen
launches; everything else is delegated to the B200 transformer module.
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Tuple

import torch

from . import ops
from .model import B200LTXTransformer, LTXConfig


class LTXVideoModelSpecification:
    # TODO(aryan)-marked constants of the reference forward (base_specification.py:281-282, 325-334)
    first_frame_conditioning_p = 0.1
    min_first_frame_sigma = 0.25
    frame_rate = 25
    temporal_compression_ratio = 8
    vae_spatial_compression_ratio = 32

    def __init__(self, transformer_config: Optional[LTXConfig] = None, transformer_dtype=torch.bfloat16):
        self.transformer_config = transformer_config or LTXConfig()
        self.transformer_dtype = transformer_dtype

    # -- load_diffusion_models (base_specification.py:173-190) with random-init weights (no hub access here)
    def load_diffusion_models(self, device="cuda") -> Dict[str, object]:
        transformer = B200LTXTransformer(self.transformer_config, self.transformer_dtype, device)
        return {"transformer": transformer, "scheduler": FlowMatchSchedulerTable()}

    # -- modeling_utils.py:156-181
    def collate_conditions(self, data):
        from .data import collate
        return collate(data)

    def collate_latents(self, data):
        from .data import collate
        return collate(data)

    def forward(self, transformer: B200LTXTransformer, condition_model_conditions: Dict[str, torch.Tensor],
                latent_model_conditions: Dict[str, torch.Tensor], sigmas: torch.Tensor,
                generator: Optional[torch.Generator] = None, compute_posterior: bool = True,
                noise: Optional[torch.Tensor] = None, **kwargs) -> Tuple[torch.Tensor, ...]:
        if not compute_posterior:
            raise NotImplementedError("posterior sampling (precomputed DiagonalGaussian latents) is outside the hot path")
        latents = latent_model_conditions.pop("latents")
        latents_mean = latent_model_conditions.pop("latents_mean")
        latents_std = latent_model_conditions.pop("latents_std")
        B, C, Fr, Hh, Ww = latents.shape
        dev = latents.device
        latents = latents.to(torch.bfloat16).contiguous()
        if noise is None:
            # same draw as the reference: torch.zeros_like(latents).normal_(generator=generator) (:296)
            noise = torch.zeros_like(latents).normal_(generator=generator)
        else:
            noise = noise.to(torch.bfloat16).contiguous()
        sig = sigmas.reshape(B).to(torch.float32).contiguous()
        sig_ff = None
        if random.random() < self.first_frame_conditioning_p:
            # base_specification.py:298-310: first latent frame gets sigma_1 = min(U[0,1) * sigma, 0.25)
            ff = torch.rand_like(sig) * sig
            sig_ff = torch.min(ff, torch.full_like(sig, self.min_first_frame_sigma)).contiguous()
        S = Fr * Hh * Ww
        x_t = torch.empty(B, S, C, dtype=torch.bfloat16, device=dev)
        target = torch.empty_like(x_t)
        ops.prep_noise_pack(latents, noise, latents_mean.reshape(B, C).to(torch.float32).contiguous(),
                            latents_std.reshape(B, C).to(torch.float32).contiguous(), sig, sig_ff, x_t, target, B, C, Fr,
                            Hh * Ww)
        sig_tok = sig.view(B, 1, 1).expand(B, S, 1)
        timesteps = (sig_tok * 1000.0).long()  # fp32 multiply then truncate, as the reference (:320)
        latent_model_conditions["hidden_states"] = x_t
        latent_frame_rate = self.frame_rate / self.temporal_compression_ratio
        rope_interpolation_scale = [1 / latent_frame_rate, self.vae_spatial_compression_ratio,
                                    self.vae_spatial_compression_ratio]
        latent_model_conditions.setdefault("num_frames", Fr)
        latent_model_conditions.setdefault("height", Hh)
        latent_model_conditions.setdefault("width", Ww)
        pred = transformer(**latent_model_conditions, **condition_model_conditions, timestep=timesteps,
                           rope_interpolation_scale=rope_interpolation_scale, return_dict=False)[0]
        return pred, target, sig_tok


class FlowMatchSchedulerTable:
    """diffusers ``FlowMatchEulerDiscreteScheduler()`` defaults as far as the trainer reads them
    (``scheduler.sigmas``, ``config.num_train_timesteps``; utils/diffusion.py:66-74): sigmas[i] = (1000 - i)/1000, then 0."""

    class _Cfg:
        num_train_timesteps = 1000

    def __init__(self):
        self.config = self._Cfg()
        ts = torch.linspace(1, 1000, 1000, dtype=torch.float32).flip(0)
        self.sigmas = torch.cat([ts / 1000.0, torch.zeros(1)])
