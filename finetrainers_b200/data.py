"""Precomputed latent / condition reader — the step immediately BEFORE the hot path.

Reads the reference's on-disk format (``/root/reference/finetrainers/data/precomputation.py:319-345,413-420``):
``{data_type}-{index}.pt`` written with ``torch.save(dict)`` and read with ``torch.load(weights_only=True)``; rank ``r``
of ``world_size`` owns indices ``r*num_items + i`` exactly as ``PrecomputedDataIterable.__iter__`` does.

The reference loads each item synchronously onto the GPU from the training thread (``map_location=torch.device(rank)``).
At >60 k tokens/s that is the next bottleneck, so this reader stages items in pinned host memory and issues the H2D copy
of item i+1 on a side stream while step i runs; the consumer gets device tensors plus an event to wait on.
"""
from __future__ import annotations

import pathlib
from typing import Any, Dict, Iterator, Optional

import torch


def save_item(item: Dict[str, Any], index: int, directory, data_type: str) -> None:
    """precomputation.py:413-415."""
    directory = pathlib.Path(directory)
    directory.mkdir(parents=True, exist_ok=True)
    torch.save(item, (directory / f"{data_type}-{index}.pt").as_posix())


def load_item(index: int, directory, data_type: str, map_location=None) -> Dict[str, Any]:
    """precomputation.py:418-420."""
    return torch.load((pathlib.Path(directory) / f"{data_type}-{index}.pt").as_posix(), map_location=map_location,
                      weights_only=True)


class PrecomputedReader:
    def __init__(self, save_dir, data_type: str, rank: int = 0, world_size: int = 1,
                 device: Optional[torch.device] = None, prefetch: bool = True):
        self.dir = pathlib.Path(save_dir)
        self.data_type = data_type
        self.rank, self.world_size = rank, world_size
        self.device = device
        self.num_items = len(list(self.dir.glob(f"{data_type}-*.pt")))
        self.prefetch = prefetch and device is not None and torch.device(device).type == "cuda"
        self._stream = torch.cuda.Stream(device) if self.prefetch else None
        self.requires_data = False

    def __len__(self) -> int:
        return self.num_items

    def _stage(self, index: int):
        item = load_item(index, self.dir, self.data_type, map_location="cpu")
        if self.device is None:
            return item, None
        if not self.prefetch:
            return {k: (v.to(self.device) if torch.is_tensor(v) else v) for k, v in item.items()}, None
        out = {}
        with torch.cuda.stream(self._stream):
            for k, v in item.items():
                out[k] = v.pin_memory().to(self.device, non_blocking=True) if torch.is_tensor(v) else v
            ev = torch.cuda.Event()
            ev.record(self._stream)
        return out, ev

    def __iter__(self) -> Iterator[Dict[str, Any]]:
        nxt = self._stage(self.rank * self.num_items + 0) if self.num_items else None
        for i in range(self.num_items):
            cur, ev = nxt
            if i + 1 < self.num_items:
                nxt = self._stage(self.rank * self.num_items + i + 1)  # overlaps with the caller's step i
            else:
                self.requires_data = True
            if ev is not None:
                torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur
