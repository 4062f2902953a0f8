"""Learning-rate schedules of the reference step (``--lr_scheduler``: finetrainers/optimizer.py:191-229 builds a
``LambdaLR`` from one of the factor functions at :250-432; the example run uses ``constant_with_warmup`` with 1000 warm-up
steps, examples/training/sft/ltx_video/crush_smol_lora/train.sh:91-92).

The b200 step has no optimizer object: the fused clip+AdamW kernel takes the step's learning rate as a scalar argument,
so the schedule is a pure host function ``factor(step)`` with LambdaLR's convention — optimizer step number k (1-based)
runs at ``lr * factor(k - 1)``.  Pinned against the reference's own lambdas by tests/golden/lr_golden.json.
"""
from __future__ import annotations

import math
from typing import Callable, Optional

SCHEDULES = ("constant", "constant_with_warmup", "piecewise_constant", "linear", "cosine", "cosine_with_restarts",
             "polynomial")


def _warmup(step: int, n_warm: int) -> Optional[float]:
    return step / max(1, n_warm) if step < n_warm else None


def lr_factor_fn(name: str = "constant", *, num_warmup_steps: int = 0, num_training_steps: Optional[int] = None,
                 num_cycles: float = 1, power: float = 1.0, lr_init: float = 1e-3, lr_end: float = 1e-7,
                 step_rules: Optional[str] = None) -> Callable[[int], float]:
    name = name.lower()
    w = int(num_warmup_steps or 0)
    if name == "constant":
        return lambda step: 1.0
    if name == "constant_with_warmup":
        return lambda step: step / max(1.0, w) if step < w else 1.0
    if name == "piecewise_constant":
        if not step_rules:
            raise ValueError("piecewise_constant needs step_rules, e.g. '1:10,0.1:20,0.005'")
        *head, last = step_rules.split(",")
        table = sorted((int(b), float(m)) for m, b in (r.split(":") for r in head))
        tail = float(last)

        def piecewise(step: int) -> float:
            for bound, mult in table:
                if step < bound:
                    return mult
            return tail

        return piecewise
    if num_training_steps is None:
        raise ValueError(f"lr scheduler '{name}' needs num_training_steps")
    total = int(num_training_steps)
    span = max(1, total - w)
    if name == "linear":
        def linear(step: int) -> float:
            f = _warmup(step, w)
            return f if f is not None else max(0.0, (total - step) / span)
        return linear
    if name == "cosine":
        def cosine(step: int) -> float:
            f = _warmup(step, w)
            if f is not None:
                return f
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * float(num_cycles) * 2.0 * ((step - w) / span))))
        return cosine
    if name == "cosine_with_restarts":
        def restarts(step: int) -> float:
            f = _warmup(step, w)
            if f is not None:
                return f
            prog = (step - w) / span
            if prog >= 1.0:
                return 0.0
            return max(0.0, 0.5 * (1.0 + math.cos(math.pi * ((float(num_cycles) * prog) % 1.0))))
        return restarts
    if name == "polynomial":
        if not lr_init > lr_end:
            raise ValueError(f"lr_end ({lr_end}) must be smaller than initial lr ({lr_init})")

        def poly(step: int) -> float:
            f = _warmup(step, w)
            if f is not None:
                return f
            if step > total:
                return lr_end / lr_init
            remaining = 1 - (step - w) / (total - w)
            return ((lr_init - lr_end) * remaining ** power + lr_end) / lr_init
        return poly
    raise ValueError(f"Unsupported scheduler: {name}")
