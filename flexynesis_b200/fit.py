"""Fit loop of the engine: reproduces the step policy the reference gets from Lightning
(flexynesis/main.py:212-225, :289-318): shuffled drop_last batches -> training_step -> backward ->
clip_grad_norm_(1.0) -> Adam.step, validation loss per epoch -- with the whole step captured in a CUDA graph.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional

import torch

from . import _lib as L


class GraphedStep:
    """Captures model.fit_step(batch) (forward + backward [+ all-reduce] + clip + Adam + plane refresh) for a batch held
    in static device buffers and replays it. With `resplit_inputs` the fp32 -> operand-plane split of the batch is
    part of the graph (the caller overwrites the static buffers between replays); without it the planes of an
    unchanged resident batch are reused (full-batch training).

    allreduce: optional callable(flat_grad_tensor) run eagerly between the backward graph and the update graph
    (data-parallel training: NCCL all-reduce of the flat gradient arena)."""

    def __init__(self, model, batch, lr: Optional[float] = None, resplit_inputs: bool = False,
                 allreduce: Optional[Callable] = None, grad_scale: float = 1.0, warmup: int = 2):
        self.model, self.batch = model, batch
        self.allreduce = allreduce
        groups, _ = model._split_batch(batch)
        eng = model.engine(groups[0][0].device)
        self.eng = eng
        lr = float(model.config["lr"] if lr is None else lr)
        eng.inputs.enabled = not resplit_inputs
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                if hasattr(allreduce, "step"):
                    g, y = model._split_batch(batch)
                    self.ws = eng.forward_backward(g, y, None)
                    allreduce.step(lr)
                else:
                    self.ws = model.fit_step(batch, lr=lr, allreduce=allreduce, grad_scale=grad_scale)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        if allreduce is None:
            self.g1, self.g2 = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(self.g1):
                self.ws = model.fit_step(batch, lr=lr, grad_scale=grad_scale)
        else:
            self.g1, self.g2 = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(self.g1):
                g, y = model._split_batch(batch)
                self.ws = eng.forward_backward(g, y, None)
            if not hasattr(allreduce, "step"):             # NCCL all-reduce + the single-GPU optimizer kernel
                self.g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g2):
                    eng.optimizer_step(lr, 1.0, grad_scale)
        self.lr = lr
        self.launches_per_step = L.launch_count() - n0 + (3 if self.g2 is None and allreduce is not None else 0)

    def __call__(self):
        self.g1.replay()
        if self.allreduce is not None:
            if self.g2 is not None:
                self.allreduce(self.eng.arena.grad)
                self.g2.replay()
            else:
                self.allreduce.step(self.lr)               # NvlsDataParallel: fused reduce-scatter / Adam / all-gather
        return self.ws

    def losses(self) -> Dict[str, torch.Tensor]:
        return self.eng.losses(self.ws)


def fit(model, dataset, batch_size: int, epochs: int, device="cuda", val_dataset=None, seed: int = 0,
        log_every: int = 0):
    """Train `model` on `dataset` (MultiOmicDataset duck type) with the engine's fused steps. Returns the per-epoch
    history [{'train_loss': ..., 'val_loss': ...}]."""
    from .data import DeviceBatcher
    model.to(device)
    model.train()
    loader = DeviceBatcher(dataset, batch_size, device, shuffle=True, drop_last=True, seed=seed)
    val = DeviceBatcher(val_dataset, len(val_dataset), device, shuffle=False, drop_last=False) if val_dataset else None
    history = []
    graphed = None
    for epoch in range(epochs):
        tot, nb = None, 0
        for batch in loader:
            if loader.full_batch:
                if graphed is None:
                    graphed = GraphedStep(model, batch)
                ws = graphed()
            else:
                ws = model.fit_step(batch)
            t = model.engine().losses(ws)["__total__"].detach().clone()
            tot = t if tot is None else tot + t
            nb += 1
        rec = {"train_loss": float(tot / max(nb, 1))}
        if val is not None:
            model.eval()
            with torch.no_grad():
                for vb in val:
                    rec["val_loss"] = float(model.validation_step(vb, 0, log=False))
            model.train()
        history.append(rec)
        if log_every and (epoch + 1) % log_every == 0:
            print(f"[fxn] epoch {epoch + 1}: {rec}")
    return history
