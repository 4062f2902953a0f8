"""Parallel backend for the data-parallel hot path: the subset of finetrainers' ``BaseParallelBackend``
(``/root/reference/finetrainers/parallel/base.py:9-115``) that ``SFTTrainer`` touches for DDP training, one process per
GPU over NCCL/NVLink (``parallel/ptd.py:41-279``; DDP = ``replicate(bucket_cap_mb=100)`` ``ptd.py:462-463``).

B200 design: the LoRA gradients already live in ONE flat fp32 buffer written by the backward kernels, so "DDP" is a
single in-place all-reduce (AVG) of that buffer on the NVSwitch fabric (235 MB at r=64: ~0.5 ms, NVLS-capable) issued
right after backward — no reducer hooks, no per-parameter buckets — and the three scalar metrics reductions
(``parallel/utils.py:6-19``, ``trainer.py:512-518``) fold into one 3-float all-reduce.  The path shards over independent
samples; there is no activation traffic.  FSDP-2 / HSDP / CP / TP / PP are outside round 1.
"""
from __future__ import annotations

import datetime
import os
from contextlib import contextmanager
from typing import Any, Dict, Optional

import torch
import torch.distributed as dist


def _avg_all_reduce(t: torch.Tensor, group=None) -> torch.Tensor:
    """AVG all-reduce that also works on the gloo backend (CPU tests): SUM then divide."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        t.div_(dist.get_world_size(group))
    return t


def dist_mean(x: torch.Tensor, group=None) -> float:
    """parallel/utils.py:17-19."""
    assert x.numel() == 1
    return _avg_all_reduce(x.clone(), group).item()


def dist_max(x: torch.Tensor, group=None) -> float:
    """parallel/utils.py:13-15."""
    assert x.numel() == 1
    y = x.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(y, op=dist.ReduceOp.MAX, group=group)
    return y.item()


def fused_step_metrics(grad_norm: torch.Tensor, loss: torch.Tensor, group=None) -> Dict[str, float]:
    """grad_norm AVG, loss AVG, loss MAX (trainer.py:507-520) with two tiny collectives and ONE host sync."""
    avg = torch.stack([grad_norm.reshape(()), loss.reshape(())]).float()
    mx = loss.reshape(1).float().clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        _avg_all_reduce(avg, group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    g, l = avg.tolist()
    return {"train/grad_norm": g, "train/global_avg_loss": l, "train/global_max_loss": mx.item()}


def allreduce_flat_grads(flat_grad: torch.Tensor, group=None, chunk_bytes: int = 0) -> torch.Tensor:
    """Average the flat gradient buffer in place.  ``chunk_bytes`` > 0 issues several collectives (bucket_cap_mb-style)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat_grad
    if chunk_bytes <= 0:
        return _avg_all_reduce(flat_grad, group)
    n = max(1, chunk_bytes // flat_grad.element_size())
    for s in range(0, flat_grad.numel(), n):
        _avg_all_reduce(flat_grad[s:s + n], group)
    return flat_grad


class B200ParallelBackend:
    """Same surface as ``PytorchDTensorParallelBackend`` for the calls the SFT loop makes (ptd.py:41-279)."""

    def __init__(self, world_size: Optional[int] = None, dp_degree: Optional[int] = None, backend: str = "nccl",
                 timeout: int = 180, device_type: str = "cuda", **unsupported_degrees):
        for k, v in unsupported_degrees.items():
            if k in ("pp_degree", "dp_shards", "cp_degree", "tp_degree") and v not in (None, 1):
                raise NotImplementedError(f"{k}={v}: only data-parallel replication (DDP) is built (SURVEY §2.2)")
        self._device_type = device_type
        if not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=timeout))
        self._world = dist.get_world_size() if dist.is_initialized() else 1
        self._rank = dist.get_rank() if dist.is_initialized() else 0
        self._local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if world_size is not None and world_size != self._world:
            raise ValueError(f"world_size {world_size} != launched world {self._world}")
        self._dp_degree = dp_degree or self._world
        if self._dp_degree != self._world:
            raise ValueError("the b200 backend is pure data parallel: dp_degree must equal world_size")
        if device_type == "cuda":
            torch.cuda.set_device(self._local_rank)
        self.tracker = None

    # --- model / optimizer preparation -------------------------------------------------------------------------
    def apply_ddp(self, model: torch.nn.Module, device_mesh=None) -> torch.nn.Module:
        """Replicas start identical (broadcast from rank 0); gradient averaging is done on the flat buffer by
        ``SFTTrainStep.optimizer_step`` / ``allreduce_flat_grads``."""
        if self._world > 1:
            with torch.no_grad():
                if hasattr(model, "lora_flat") and getattr(model, "_prepared", False):
                    dist.broadcast(model.lora_flat, src=0)
                else:
                    for p in model.parameters():
                        dist.broadcast(p.data, src=0)
        model._b200_ddp = True
        return model

    def apply_fsdp2(self, *args, **kwargs):
        raise NotImplementedError("FSDP-2 (per-block bf16 all-gather + fp32 reduce-scatter, ptd.py:466-499) is the next "
                                  "multi-GPU row; round 1 ships DDP for the LoRA config")

    def apply_context_parallel(self, *args, **kwargs):
        raise NotImplementedError("context parallelism is out of scope: LTX has no CP plan in the reference either")

    def prepare_model(self, model):
        return model

    def prepare_optimizer(self, optimizer):
        return optimizer

    def prepare_dataset(self, dataset):
        return dataset

    def prepare_dataloader(self, dataset, batch_size=1, num_workers=0, pin_memory=True):
        return torch.utils.data.DataLoader(dataset, batch_size=batch_size, num_workers=num_workers, pin_memory=pin_memory)

    def get_mesh(self, name: Optional[str] = None):
        return {"dp": None, "dp_cp": None, "dp_replicate": None}.get(name) if name else {"dp": None, "dp_cp": None}

    def get_checkpointer(self, *args, **kwargs):
        raise NotImplementedError("checkpointing stays with the caller (parameter FQNs are diffusers/peft compatible)")

    def wait_for_everyone(self):
        if self._world > 1:
            dist.barrier()

    @contextmanager
    def main_process_first(self):
        if not self.is_main_process:
            self.wait_for_everyone()
        yield
        if self.is_main_process:
            self.wait_for_everyone()

    def destroy(self):
        if dist.is_initialized():
            dist.destroy_process_group()

    def log(self, metrics: Dict[str, Any], step: int) -> None:
        if self.is_main_process and self.tracker is not None:
            self.tracker.log(metrics, step)

    # --- properties ---------------------------------------------------------------------------------------------
    @property
    def world_size(self):
        return self._world

    @property
    def rank(self):
        return self._rank

    @property
    def local_rank(self):
        return self._local_rank

    @property
    def is_main_process(self):
        return self._rank == 0

    @property
    def is_local_main_process(self):
        return self._local_rank == 0

    @property
    def device(self):
        return torch.device(self._device_type, self._local_rank) if self._device_type == "cuda" else torch.device("cpu")

    @property
    def pipeline_parallel_enabled(self):
        return False

    @property
    def data_parallel_enabled(self):
        return self._world > 1

    @property
    def data_replication_enabled(self):
        return self._world > 1

    @property
    def data_sharding_enabled(self):
        return False

    @property
    def context_parallel_enabled(self):
        return False

    @property
    def tensor_parallel_enabled(self):
        return False
