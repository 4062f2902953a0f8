"""The SFT train-step body of finetrainers' ``SFTTrainer._train``
(``/root/reference/finetrainers/trainer/sft_trainer/trainer.py:397-529``) rebuilt around the B200 engine.

Kept from the reference: sigma sampling (``utils/diffusion.py:38-63,84-114``), loss weighting (``:117-130``), the
loss definition (``trainer.py:474-481``), clip-then-AdamW ordering (``:488-503``), gradient accumulation, and the
per-step metrics (``global_avg_loss``, ``global_max_loss``, ``grad_norm``; ``:507-520``).

Changed for B200: loss + dloss/dpred is one kernel; LoRA gradients land in one flat fp32 buffer that is all-reduced in
place (DDP) and consumed by one fused clip+AdamW kernel; the three scalar reductions are one 3-float all-reduce; the
host never synchronises inside a step unless the caller asks for the metrics (``sync_metrics``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch

from . import ops
from .model import B200LTXTransformer
from .specification import LTXVideoModelSpecification, FlowMatchSchedulerTable


def compute_density_for_timestep_sampling(weighting_scheme: str, batch_size: int, logit_mean: float = 0.0,
                                          logit_std: float = 1.0, mode_scale: float = 1.29, device="cpu",
                                          generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """utils/diffusion.py:38-63."""
    if weighting_scheme == "logit_normal":
        u = torch.normal(mean=logit_mean, std=logit_std, size=(batch_size,), device=device, generator=generator)
        u = torch.nn.functional.sigmoid(u)
    elif weighting_scheme == "mode":
        u = torch.rand(size=(batch_size,), device=device, generator=generator)
        u = 1 - u - mode_scale * (torch.cos(math.pi * u / 2) ** 2 - 1 + u)
    else:
        u = torch.rand(size=(batch_size,), device=device, generator=generator)
    return u


def prepare_sigmas(scheduler, sigmas: torch.Tensor, batch_size: int, num_train_timesteps: int,
                   flow_weighting_scheme: str = "none", flow_logit_mean: float = 0.0, flow_logit_std: float = 1.0,
                   flow_mode_scale: float = 1.29, device="cpu", generator=None) -> torch.Tensor:
    """utils/diffusion.py:84-114 (flow-match branch)."""
    w = compute_density_for_timestep_sampling(flow_weighting_scheme, batch_size, flow_logit_mean, flow_logit_std,
                                              flow_mode_scale, device, generator)
    indices = (w * num_train_timesteps).long()
    return sigmas[indices]


def prepare_loss_weights(sigmas: torch.Tensor, flow_weighting_scheme: str = "none") -> torch.Tensor:
    """utils/diffusion.py:117-130 -> diffusers compute_loss_weighting_for_sd3."""
    if flow_weighting_scheme == "sigma_sqrt":
        return (sigmas ** -2.0).float()
    if flow_weighting_scheme == "cosmap":
        bot = 1 - 2 * sigmas + 2 * sigmas ** 2
        return 2 / (math.pi * bot)
    return torch.ones_like(sigmas)


def expand_tensor_dims(t: torch.Tensor, ndim: int) -> torch.Tensor:
    """utils/torch.py:219-221."""
    while t.ndim < ndim:
        t = t.unsqueeze(-1)
    return t


class SFTTrainStep:
    """One optimizer step = ``gradient_accumulation_steps`` micro-steps of forward/loss/backward, then
    all-reduce (DDP) + clip + AdamW on the flat LoRA buffers."""

    def __init__(self, transformer: B200LTXTransformer, spec: Optional[LTXVideoModelSpecification] = None, *,
                 lr: float = 5e-5, beta1: float = 0.9, beta2: float = 0.99, weight_decay: float = 1e-4,
                 eps: float = 1e-8, max_grad_norm: float = 1.0, gradient_accumulation_steps: int = 1,
                 flow_weighting_scheme: str = "logit_normal", flow_logit_mean: float = 0.0,
                 flow_logit_std: float = 1.0, flow_mode_scale: float = 1.29, seed: int = 42,
                 process_group=None):
        self.transformer = transformer
        self.spec = spec or LTXVideoModelSpecification(transformer.cfg)
        self.scheduler = FlowMatchSchedulerTable()
        self.lr, self.beta1, self.beta2, self.wd, self.eps = lr, beta1, beta2, weight_decay, eps
        self.max_grad_norm = max_grad_norm
        self.grad_accum = gradient_accumulation_steps
        self.scheme = flow_weighting_scheme
        self.flow_logit_mean, self.flow_logit_std, self.flow_mode_scale = flow_logit_mean, flow_logit_std, flow_mode_scale
        if not transformer._prepared:
            transformer.prepare()
        dev = transformer.proj_in.weight.device
        self.device = dev
        self.generator = torch.Generator(device=dev).manual_seed(seed)
        self.scheduler_sigmas = self.scheduler.sigmas.to(dev)
        n = transformer.lora_flat.numel()
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_buf = torch.zeros(1, dtype=torch.float32, device=dev)
        self.loss_acc = torch.zeros(1, dtype=torch.float32, device=dev)
        self.partial = torch.zeros(1024, dtype=torch.float32, device=dev)
        self.metrics = torch.zeros(3, dtype=torch.float32, device=dev)
        self.opt_step = 0
        self.micro = 0
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (
            torch.distributed.is_available() and torch.distributed.is_initialized()) else 1
        self._dpred = {}

    # -- forward + loss + backward of one micro-batch (trainer.py:436-483)
    def micro_step(self, condition_model_conditions: Dict[str, torch.Tensor],
                   latent_model_conditions: Dict[str, torch.Tensor], sigmas: Optional[torch.Tensor] = None,
                   noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        B = latent_model_conditions["latents"].shape[0]
        if sigmas is None:
            sigmas = prepare_sigmas(self.scheduler, self.scheduler_sigmas, B, self.scheduler.config.num_train_timesteps,
                                    self.scheme, self.flow_logit_mean, self.flow_logit_std, self.flow_mode_scale,
                                    self.device, self.generator)
        sigmas = expand_tensor_dims(sigmas, latent_model_conditions["latents"].ndim)
        pred, target, sig_tok = self.spec.forward(self.transformer, condition_model_conditions,
                                                  latent_model_conditions, sigmas, generator=self.generator,
                                                  noise=noise)
        weights = prepare_loss_weights(sig_tok[:, 0, 0].float(), self.scheme).contiguous()
        key = tuple(pred.shape)
        dpred = self._dpred.get(key)
        if dpred is None:
            dpred = torch.empty(pred.shape, dtype=torch.bfloat16, device=pred.device)
            self._dpred[key] = dpred
        per_sample = pred.shape[1] * pred.shape[2]
        ops.loss_mse(pred, target, weights, 1.0 / self.grad_accum, self.loss_buf, dpred, self.partial, B, per_sample)
        pred.backward(dpred)
        self.loss_acc += self.loss_buf
        self.micro += 1
        return self.loss_buf

    # -- clip + AdamW (+ DDP all-reduce) (trainer.py:486-520)
    def optimizer_step(self, sync_metrics: bool = False):
        tr = self.transformer
        g = tr.lora_grad_flat
        if self.world > 1:
            # DDP: average the flat fp32 gradient buffer in place over NVLink (ptd.py:462-463 replicate(bucket_cap_mb=100))
            torch.distributed.all_reduce(g, op=torch.distributed.ReduceOp.AVG, group=self.pg)
        self.sumsq.zero_()
        ops.sumsq(g, g.numel(), self.sumsq, self.partial)
        self.opt_step += 1
        self.metrics[0:1] = self.sumsq.sqrt()
        self.metrics[1:2] = self.loss_acc
        self.metrics[2:3] = self.loss_acc
        ops.adamw_clip(tr.lora_flat, g, self.exp_avg, self.exp_avg_sq, g.numel(), self.sumsq, self.max_grad_norm, self.lr,
                       self.beta1, self.beta2, self.eps, self.wd, self.opt_step, 1.0)
        self.loss_acc.zero_()
        self.micro = 0
        if not sync_metrics:
            return None
        m = self.metrics.clone()
        if self.world > 1:
            avg = m[:2].clone()
            torch.distributed.all_reduce(avg, op=torch.distributed.ReduceOp.AVG, group=self.pg)
            mx = m[2:].clone()
            torch.distributed.all_reduce(mx, op=torch.distributed.ReduceOp.MAX, group=self.pg)
            m = torch.cat([avg, mx])
        grad_norm, avg_loss, max_loss = m.tolist()
        return {"train/grad_norm": grad_norm, "train/global_avg_loss": avg_loss, "train/global_max_loss": max_loss}

    def train_step(self, condition_model_conditions, latent_model_conditions, sigmas=None, noise=None,
                   sync_metrics=False):
        self.micro_step(condition_model_conditions, latent_model_conditions, sigmas, noise)
        if self.micro % self.grad_accum == 0:
            return self.optimizer_step(sync_metrics)
        return None
