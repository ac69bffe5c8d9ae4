"""Independent hyper-parameter trials, one per GPU.

`HyperparameterTuning.perform_tuning` (flexynesis/main.py:352-368) runs its `n_iter` trials one after the other on one
device; the trials of one ask/tell round are independent, so an 8-GPU box can run eight at once with no communication
at all ("replicas only" in the scope contract's terms). `run_trials` is that scheduler: one worker process per device
pulls (index, config) pairs from a queue, calls `objective(config, device)` and reports (index, result). The objective
is the caller's -- typically: build a flexynesis_b200 model on `device`, `fit`, return the validation loss -- and must be
picklable (a module-level function)."""
from __future__ import annotations

import traceback
from typing import Any, Callable, List, Optional, Sequence

import torch
import torch.multiprocessing as mp


def _worker(device: str, objective: Callable, tasks, results):
    if device.startswith("cuda"):
        torch.cuda.set_device(torch.device(device))
    while True:
        item = tasks.get()
        if item is None:
            return
        idx, cfg = item
        try:
            results.put((idx, objective(cfg, device), None))
        except Exception:                                    # report, keep serving the queue
            results.put((idx, None, traceback.format_exc()))


def run_trials(objective: Callable[[dict, str], Any], configs: Sequence[dict], devices: Optional[Sequence[str]] = None,
               timeout: Optional[float] = None) -> List[Any]:
    """Run objective(config, device) for every config, at most one trial per device at a time; returns the results in the
    order of `configs`. A trial that raises makes run_trials raise RuntimeError with its traceback after the other
    trials have finished. devices defaults to every visible CUDA device."""
    if devices is None:
        devices = [f"cuda:{i}" for i in range(torch.cuda.device_count())]
    if not devices:
        raise RuntimeError("run_trials: no devices")
    ctx = mp.get_context("spawn")
    tasks, results = ctx.Queue(), ctx.Queue()
    for item in enumerate(configs):
        tasks.put(item)
    for _ in devices:
        tasks.put(None)
    procs = [ctx.Process(target=_worker, args=(d, objective, tasks, results), daemon=True) for d in devices]
    for p in procs:
        p.start()
    out: List[Any] = [None] * len(configs)
    errors = []
    for _ in range(len(configs)):
        idx, res, err = results.get(timeout=timeout)
        if err is not None:
            errors.append((idx, err))
        out[idx] = res
    for p in procs:
        p.join(30)
    if errors:
        raise RuntimeError("trial(s) failed:\n" + "\n".join(f"[trial {i}]\n{e}" for i, e in errors))
    return out
