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
        if eng.sync is not None:
            raise NotImplementedError("global-batch sync (parallel.GlobalBatchSync) issues host-driven collectives between "
                                      "kernels and runs eagerly: use model.fit_step, not a captured graph")
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


class HostStreamTrainer:
    """Trains from batches that live in (pinned) HOST memory: every step's inputs are copied host -> device and the
    step's loss is read back, with the copy of step i+1 overlapped with the compute of step i (two static device buffer
    sets, one captured graph per set, one copy stream). This is the end-to-end path `bench.py` reports as `e2e`.

    make_batch(buffers) -> batch tuple maps a list of device tensors (same order as `host_tensors`) to the model's
    batch format."""

    def __init__(self, model, host_tensors, make_batch, lr: Optional[float] = None, allreduce=None, grad_scale: float = 1.0):
        self.host = [t if t.is_pinned() else t.pin_memory() for t in host_tensors]
        dev = next(model.parameters()).device
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.sets, self.steps, self.ready, self.consumed = [], [], [], []
        for _ in range(2):
            bufs = [torch.empty_like(t, device=dev) for t in self.host]
            for d, s in zip(bufs, self.host):
                d.copy_(s)
            self.sets.append(bufs)
            self.steps.append(GraphedStep(model, make_batch(bufs), lr=lr, resplit_inputs=True, allreduce=allreduce,
                                          grad_scale=grad_scale))
            self.ready.append(torch.cuda.Event())
            self.consumed.append(torch.cuda.Event())
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host)
        self.launches_per_step = self.steps[0].launches_per_step
        self._next = 0
        self._primed = False

    def _start_copy(self, k: int, host_tensors=None):
        src = self.host if host_tensors is None else host_tensors
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[k])          # the step that last read this buffer set has finished
            for d, s in zip(self.sets[k], src):
                d.copy_(s, non_blocking=True)
            self.ready[k].record(self.copy_stream)

    def step(self, next_host_tensors=None) -> float:
        """One training step on the batch whose copy was started by the previous call; starts the copy of the next batch
        (`next_host_tensors`, default: the same host batch again) and returns this step's loss (device -> host read)."""
        k = self._next
        cur = torch.cuda.current_stream()
        if not self._primed:
            self.consumed[0].record(cur); self.consumed[1].record(cur)
            self._start_copy(k)
            self._primed = True
        cur.wait_event(self.ready[k])
        self.steps[k]()
        self.consumed[k].record(cur)
        self._start_copy(1 - k, next_host_tensors)                 # overlaps with the step just queued
        self._next = 1 - k
        return float(self.steps[k].losses()["__total__"])          # synchronises with the step, not with the copy


def fit(model, dataset, batch_size: int, epochs: int, device="cuda", val_dataset=None, seed: int = 0,
        log_every: int = 0):
    """Train `model` on `dataset` (MultiOmicDataset duck type) with the engine's fused steps. Returns the per-epoch
    history [{'train_loss': ..., 'val_loss': ...}]."""
    from .data import DeviceBatcher, DeviceTripletBatcher
    model.to(device)
    model.train()
    if getattr(model, "main_var", None) is not None:       # MultiTripletNetwork: on-device triplet sampling
        loader = DeviceTripletBatcher(dataset, model.main_var, batch_size, device, shuffle=True, drop_last=True, seed=seed)
        loader.full_batch = False
    else:
        loader = DeviceBatcher(dataset, batch_size, device, shuffle=True, drop_last=True, seed=seed)
    if val_dataset is None:
        val = None
    elif getattr(model, "main_var", None) is not None:     # triplet batches for the validation objective as well
        vbase = getattr(val_dataset, "dataset", val_dataset)
        nval = int((~torch.isnan(torch.as_tensor(vbase.ann[model.main_var]).float())).sum())
        val = DeviceTripletBatcher(val_dataset, model.main_var, max(nval, 1), device, shuffle=False, drop_last=False,
                                   seed=seed + 1)
    else:
        val = DeviceBatcher(val_dataset, len(val_dataset), device, shuffle=False, drop_last=False)
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
