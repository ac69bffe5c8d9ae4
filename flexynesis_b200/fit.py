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
            import torch.distributed as dist
            if not (dist.is_initialized() and dist.get_backend(eng.sync.group) == "nccl"):
                raise NotImplementedError("global-batch sync (parallel.GlobalBatchSync) inside a captured graph needs the NCCL "
                                          "backend (its collectives are captured with the kernels); over gloo use "
                                          "model.fit_step")
        lr = float(model.config["lr"] if lr is None else lr)
        eng.inputs.enabled = not resplit_inputs
        eng.ws_tag = f"graph{id(self)}"            # private workspace: eager calls of the same batch size cannot alias it
        try:
            self._capture(model, batch, eng, lr, allreduce, grad_scale, warmup)
        finally:
            eng.ws_tag = ""

    def _capture(self, model, batch, eng, lr, allreduce, grad_scale, warmup):
        # The warm-up steps (they allocate the workspace and load every kernel before capture) must not count as training:
        # parameters, Adam moments, counters and BatchNorm buffers are put back afterwards, so the first replay is step 1.
        a = eng.arena
        saved = [t.clone() for t in (a.flat, a.exp_avg, a.exp_avg_sq, a.step, eng.noise_step)]
        bufs = [(b, b.clone()) for b in model.buffers()]
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
            with torch.no_grad():
                for dst, src in zip((a.flat, a.exp_avg, a.exp_avg_sq, a.step, eng.noise_step), saved):
                    dst.copy_(src)
                for b, src in bufs:
                    b.copy_(src)
            eng.wplanes.refresh()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = L.launch_count()
        if allreduce is None:
            self.g1, self.g2 = torch.cuda.CUDAGraph(), None
            with torch.cuda.graph(self.g1):
                self.ws = model.fit_step(batch, lr=lr, grad_scale=grad_scale)
        else:
            self.g1, self.g2 = torch.cuda.CUDAGraph(), None
            self.fused_dp = hasattr(allreduce, "step") and getattr(allreduce, "capturable", False)
            with torch.cuda.graph(self.g1):
                g, y = model._split_batch(batch)
                self.ws = eng.forward_backward(g, y, None)
                if self.fused_dp:                          # reduce-scatter / Adam / all-gather + in-stream barriers: same graph
                    allreduce.step(lr)
            if not hasattr(allreduce, "step"):             # NCCL all-reduce + the single-GPU optimizer kernel
                self.g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g2):
                    eng.optimizer_step(lr, 1.0, grad_scale)
        self.lr = lr
        self.launches_per_step = L.launch_count() - n0 + (
            3 if self.g2 is None and allreduce is not None and not getattr(self, "fused_dp", False) else 0)

    def __call__(self):
        self.g1.replay()
        if self.allreduce is not None:
            if self.g2 is not None:
                self.allreduce(self.eng.arena.grad)
                self.g2.replay()
            elif not self.fused_dp:
                self.allreduce.step(self.lr)               # NvlsDataParallel with host barriers: outside the graph
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


def _validation_loss(model, val) -> float:
    """Mean of validation_step over the batches of `val`, weighted by batch size (what Lightning's on_epoch mean logs)."""
    tot, rows = 0.0, 0
    for vb in val:
        first = vb[0]
        while isinstance(first, dict):
            first = next(iter(first.values()))
        n = int(first.shape[0])
        tot += float(model.validation_step(vb, 0, log=False)) * n
        rows += n
    return tot / max(rows, 1)


def fit(model, dataset, batch_size: int, epochs: int, device="cuda", val_dataset=None, seed: int = 0,
        log_every: int = 0, graph: bool = True):
    """Train `model` on `dataset` with the engine's fused steps: the reference's Lightning policy (flexynesis/main.py:212-225,
    :289-318: shuffled drop_last batches -> training_step -> backward -> clip_grad_norm_(1.0) -> Adam.step, a validation
    pass per epoch) with the batch gathered ON THE DEVICE into static buffers and the whole step replayed from ONE CUDA
    graph -- for mini-batches too (the reference's default regime, B in {32, 64, 128}, main.py:183-190): the batcher
    refills the buffers, the captured step re-splits them into operand planes. `dataset` is a MultiOmicDataset duck type,
    a TripletMultiOmicDataset (MultiTripletNetwork) or a MultiOmicDatasetNW (GNN: `node_features_tensor`). Validation runs
    in batches of `batch_size` rows. Returns the per-epoch history [{'train_loss': ..., 'val_loss': ...}]."""
    from .data import DeviceBatcher, DeviceNodeBatcher, DeviceTripletBatcher
    model.to(device)
    model.train()
    triplet = getattr(model, "main_var", None) is not None
    graph_ds = hasattr(dataset, "node_features_tensor") and not hasattr(dataset, "dat")
    if triplet:                                              # MultiTripletNetwork: on-device triplet sampling
        loader = DeviceTripletBatcher(dataset, model.main_var, batch_size, device, shuffle=True, drop_last=True, seed=seed)
        loader.full_batch = False
    elif graph_ds:
        loader = DeviceNodeBatcher(dataset, batch_size, device, shuffle=True, drop_last=True, seed=seed)
    else:
        loader = DeviceBatcher(dataset, batch_size, device, shuffle=True, drop_last=True, seed=seed)
    if val_dataset is None:
        val = None
    elif triplet:                                            # triplet batches for the validation objective as well
        vbase = getattr(val_dataset, "dataset", val_dataset)
        nval = int((~torch.isnan(torch.as_tensor(vbase.ann[model.main_var]).float())).sum())
        val = DeviceTripletBatcher(val_dataset, model.main_var, max(min(nval, batch_size), 1), device, shuffle=False,
                                   drop_last=False, seed=seed + 1)
    elif graph_ds:
        val = DeviceNodeBatcher(val_dataset, min(len(val_dataset), batch_size), device, shuffle=False, drop_last=False)
    else:
        val = DeviceBatcher(val_dataset, min(len(val_dataset), batch_size), device, shuffle=False, drop_last=False)
    history = []
    graphed = None
    for epoch in range(epochs):
        tot, nb = None, 0
        for batch in loader:
            if graph and graphed is None:
                # full batch: the resident planes are split once; mini-batches: the split of the static buffers is captured
                graphed = GraphedStep(model, batch, resplit_inputs=not loader.full_batch)
            ws = graphed() if graph else model.fit_step(batch)
            t = model.engine().losses(ws)["__total__"].detach().clone()
            tot = t if tot is None else tot + t
            nb += 1
        rec = {"train_loss": float(tot / max(nb, 1)) if tot is not None else float("nan")}
        if val is not None:
            model.eval()
            with torch.no_grad():
                rec["val_loss"] = _validation_loss(model, val)
            model.train()
        history.append(rec)
        if log_every and (epoch + 1) % log_every == 0:
            print(f"[fxn] epoch {epoch + 1}: {rec}")
    return history
