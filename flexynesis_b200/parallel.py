"""Data-parallel plumbing: one process per GPU, samples sharded by rank, parameters replicated,
one sum all-reduce of the flat gradient arena per step, then the same clip + Adam on every rank with
grad_scale = 1 / world (clipping acts on the averaged gradient, as it would on one GPU holding the global batch).

The reference has no multi-GPU path (`pl.Trainer(devices=1)`, flexynesis/main.py:223); the semantics here are
PyTorch-DDP's: per-rank BatchNorm statistics and per-rank whole-batch functionals (Cox risk sets, MMD), mean of the
per-rank losses. torch.distributed is used for the plumbing only (NCCL on GPUs; the CPU tests run the same code over
gloo).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard [start, stop) of n samples; every rank gets the same count (the remainder n % world is
    dropped so that all ranks run the same number of equally sized steps)."""
    per = n // world
    return rank * per, (rank + 1) * per


class DatasetShard:
    """View of a MultiOmicDataset duck type restricted to this rank's samples (tensors are sliced, not copied)."""

    def __init__(self, dataset, rank: int, world: int):
        lo, hi = shard_range(len(dataset), rank, world)
        self.rank, self.world, self.range = rank, world, (lo, hi)
        self.dat = {k: v[lo:hi] for k, v in dataset.dat.items()}
        self.ann = {k: v[lo:hi] for k, v in dataset.ann.items()}
        self.samples = list(dataset.samples[lo:hi])
        self.features = dataset.features
        self.variable_types = dataset.variable_types

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return ({k: v[i] for k, v in self.dat.items()}, {k: v[i] for k, v in self.ann.items()}, self.samples[i])


class GradAllReduce:
    """Sum all-reduce of the flat gradient arena over the default process group. `chunks` > 1 splits the arena into
    equal slices issued back to back (each slice can start as soon as the backward kernels that fill it have been
    queued on the stream; NCCL pipelines them)."""

    def __init__(self, world: int, chunks: int = 1, group=None):
        self.world, self.chunks, self.group = world, max(int(chunks), 1), group
        self.calls = 0

    def __call__(self, flat: torch.Tensor) -> torch.Tensor:
        self.calls += 1
        if self.world <= 1:
            return flat
        if self.chunks == 1:
            dist.all_reduce(flat, group=self.group)
            return flat
        n = flat.numel()
        per = (n + self.chunks - 1) // self.chunks
        for c in range(self.chunks):
            lo, hi = c * per, min((c + 1) * per, n)
            if lo < hi:
                dist.all_reduce(flat[lo:hi], group=self.group)
        return flat


def broadcast_parameters(arena_flat: torch.Tensor, buffers: Dict[str, torch.Tensor], src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and BatchNorm buffers."""
    dist.broadcast(arena_flat, src, group=group)
    for b in buffers.values():
        if b.dtype.is_floating_point or b.dtype in (torch.int64, torch.int32):
            dist.broadcast(b, src, group=group)
