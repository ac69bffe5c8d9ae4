"""Independent hyper-parameter trials, one per GPU.

`HyperparameterTuning.perform_tuning` (flexynesis/main.py:352-368) runs its `n_iter` trials one after the other on one
device; the trials of one ask/tell round are independent, so an 8-GPU box can run eight at once with no communication
at all ("replicas only" in the scope contract's terms). `run_trials` is that scheduler: one worker process per device
pulls (index, config) pairs from a queue, calls `objective(config, device)` and reports (index, result). The objective
is the caller's -- typically: build a flexynesis_b200 model on `device`, `fit`, return the validation loss -- and must be
picklable (a module-level function)."""
from __future__ import annotations

import queue
import time
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
        results.put((idx, "__started__", device))            # lets the parent name the trial a dead worker was holding
        try:
            results.put((idx, objective(cfg, device), None))
        except Exception:                                    # report, keep serving the queue
            results.put((idx, None, traceback.format_exc()))


def run_trials(objective: Callable[[dict, str], Any], configs: Sequence[dict], devices: Optional[Sequence[str]] = None,
               timeout: Optional[float] = None) -> List[Any]:
    """Run objective(config, device) for every config, at most one trial per device at a time; returns the results in the
    order of `configs`. A trial that raises makes run_trials raise RuntimeError with its traceback after the other
    trials have finished; a worker PROCESS that dies (out-of-memory kill, CUDA fault, segfault in the native library)
    fails the trial it was holding instead of hanging the caller. `timeout` bounds the wait for any single result.
    devices defaults to every visible CUDA device."""
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
    running = {}                                             # device -> trial index it last started
    done, last = 0, time.monotonic()
    while done < len(configs):
        try:
            idx, res, err = results.get(timeout=1.0)
        except queue.Empty:
            dead = [(d, p) for d, p in zip(devices, procs) if not p.is_alive() and p.exitcode not in (0, None)]
            for d, p in dead:
                if d in running:                             # the trial this worker held is lost with it
                    i = running.pop(d)
                    errors.append((i, f"worker on {d} died with exit code {p.exitcode}"))
                    done += 1
            if not any(p.is_alive() for p in procs) and done < len(configs):
                errors.append((-1, "every worker exited before all trials were served"))
                break
            if timeout is not None and time.monotonic() - last > timeout:
                errors.append((-1, f"no result within {timeout} s"))
                break
            continue
        last = time.monotonic()
        if isinstance(res, str) and res == "__started__":
            running[err] = idx
            continue
        for d, i in list(running.items()):
            if i == idx:
                running.pop(d)
        if err is not None:
            errors.append((idx, err))
        out[idx] = res
        done += 1
    for p in procs:
        if p.is_alive() and errors and errors[-1][0] == -1:
            p.terminate()
        p.join(30)
    if errors:
        raise RuntimeError("trial(s) failed:\n" + "\n".join(f"[trial {i}]\n{e}" for i, e in errors))
    return out
