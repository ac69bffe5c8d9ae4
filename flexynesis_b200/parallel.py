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


class NvlsDataParallel:
    """Data-parallel optimizer step over NVSwitch multicast (csrc/dp.cu): the gradient arena is reduce-scattered by
    `multimem.ld_reduce` (in-switch sum), each rank clips with the global norm and runs Adam on its 1/W slice only, and
    the updated parameters are all-gathered by `multimem.st`. Needs torch symmetric memory with multicast support
    (NVLS: every GPU behind one NVSwitch domain); `available()` says whether this process group has it -- callers fall
    back to GradAllReduce (NCCL) + the single-GPU optimizer kernel and SAY SO in their report.

    Usage:  with NvlsDataParallel.arena_allocation(): eng = model.engine(dev)     # arenas land in symmetric memory
            dp = NvlsDataParallel(eng)                                               # rendezvous (collective)
            ...backward...; dp.step(lr)                                              # collective
    """

    _pending = []          # tensors handed out by the allocator, rendezvoused by the next constructor

    @staticmethod
    def available() -> bool:
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
        except Exception:
            return False
        return dist.is_initialized() and dist.get_world_size() > 1 and torch.cuda.is_available()

    @staticmethod
    def _alloc(numel, device):
        import torch.distributed._symmetric_memory as symm_mem
        t = symm_mem.empty(numel, dtype=torch.float32, device=device)
        t.zero_()
        NvlsDataParallel._pending.append(t)
        return t

    class arena_allocation:
        def __enter__(self):
            from .engine import ParamArena
            NvlsDataParallel._pending.clear()
            ParamArena.allocator = NvlsDataParallel._alloc

        def __exit__(self, *exc):
            from .engine import ParamArena
            ParamArena.allocator = None
            return False

    def __init__(self, eng, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib as L
        self.L, self.eng = L, eng
        a = eng.arena
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        group = group or dist.group.WORLD
        if not any(t.data_ptr() == a.flat.data_ptr() for t in NvlsDataParallel._pending):
            raise RuntimeError("the engine's arenas are not in symmetric memory: build it under NvlsDataParallel.arena_allocation()")
        self.h_flat = symm_mem.rendezvous(a.flat, group)
        self.h_grad = symm_mem.rendezvous(a.grad, group)
        self.partials = symm_mem.empty(64, dtype=torch.float32, device=a.device)
        self.partials.zero_()
        self.h_part = symm_mem.rendezvous(self.partials, group)
        for h in (self.h_flat, self.h_grad, self.h_part):
            if not h.multicast_ptr:
                raise RuntimeError("symmetric memory without multicast (NVLS) support on this system")
        per = a.numel // self.world
        per -= per % 4
        self.begin = self.rank * per
        self.end = a.numel if self.rank == self.world - 1 else (self.rank + 1) * per
        self.scratch = torch.zeros(4, dtype=torch.float32, device=a.device)        # 16 bytes: double sum + counter
        # in-stream barrier state (csrc/dp.cu dp_barrier_kernel): symmetric arrival counters + a local epoch counter
        self.flags = symm_mem.empty(16, dtype=torch.int32, device=a.device)
        self.flags.zero_()
        self.h_flags = symm_mem.rendezvous(self.flags, group)
        self.epoch = torch.zeros(48, dtype=torch.int32, device=a.device)      # [0..15] epochs, [16..47] kernel scratch
        env = __import__("os").environ
        self.host_barriers = bool(int(env.get("FXN_DP_HOST_BARRIERS", "0"))) or not self.h_flags.multicast_ptr
        self.fused_barriers = bool(int(env.get("FXN_DP_FUSED_BARRIERS", "0")))       # barriers 1 and 2 inside their kernels
        # (measured at 2 ranks: no gain over the stand-alone barrier kernels -- the wait is rank skew, not launch overhead)
        # gradient reduce-scatter: multimem.ld_reduce in the switch (default) or peer loads added in rank order (<= 8 ranks)
        mode = env.get("FXN_DP_REDUCE", "multimem")     # at 2 ranks: multimem 30 us, peer loads 39 us for the 4.3 MB slice
        self.peer_grads = [int(p) for p in self.h_grad.buffer_ptrs] if mode == "p2p" else None
        # Reduce-scatter fused into the weight-gradient GEMMs (opt-in: FXN_DP_FUSED_RS=1, <= 8 ranks): every rank owns an inbox
        # arena; a GEMM's stream-K reductions for elements of another rank's slice travel over NVLink into that rank's inbox
        # while the GEMM runs, and the reduce kernel adds own + inbox instead of pulling through the switch. Correct (bench
        # dp_check.fused_reduce_scatter) but NOT faster on this box: the remote reductions slow the GEMM by ~9 us and the
        # reduce kernel does not get shorter (its time is not set by the bytes it pulls) -- 0.461 vs 0.456 ms per step at
        # 2 ranks, profiles/r02_dp_timeline_cfg2_n2_fused_rs.log. Kept as an experiment switch.
        self.rs = None
        if self.world <= 8 and bool(int(env.get("FXN_DP_FUSED_RS", "0"))) and a.numel < (1 << 31):
            self.inbox = symm_mem.empty(a.numel, dtype=torch.float32, device=a.device)
            self.inbox.zero_()
            self.h_inbox = symm_mem.rendezvous(self.inbox, group)
            self.rs = L.ReduceScatterContext(self.world, self.rank, per, a.grad.data_ptr(), a.numel,
                                             [int(p) for p in self.h_inbox.buffer_ptrs])
            L.RS = self.rs
        # make every rank start from rank 0's parameters
        dist.broadcast(a.flat, 0, group=group)
        eng.wplanes.refresh()
        torch.cuda.synchronize()
        self.h_flat.barrier(channel=0)

    def _barrier(self, handle, slot=None):
        slot = 2 if slot is None else slot
        if self.host_barriers:                           # torch symmetric-memory barrier, enqueued by the host (not capturable)
            handle.barrier(channel=0)
        else:                                            # one device thread per rank, in stream order (capturable)
            self.L.dp_barrier(self.h_flags.multicast_ptr, self.flags.data_ptr(), self.epoch.data_ptr(), slot, self.world)

    def _sync(self, slot):
        return (self.h_flags.multicast_ptr, self.flags.data_ptr(), self.epoch.data_ptr(), slot, self.world)

    @property
    def capturable(self) -> bool:
        return not self.host_barriers

    def step(self, lr: float, max_norm: float = 1.0, pull_all: bool = False):
        """Collective: call on every rank after its backward pass has been queued on the current stream. With the
        in-stream barriers the whole sequence is plain kernel launches and can be captured in the same CUDA graph as the
        backward pass. pull_all: the gradient arena was filled by something other than this engine's backward pass (tests,
        bench.dp_consistency): pull every element through the switch, ignore the ranges the GEMMs reduce-scatter."""
        L, a = self.L, self.eng.arena
        dev_sync = not self.host_barriers and self.fused_barriers
        if not dev_sync:
            self._barrier(self.h_grad)                   # every rank's gradients are complete and visible
        L.dp_reduce_sumsq(self.h_grad.multicast_ptr, a.grad.data_ptr(), self.begin, self.end, 1.0 / self.world,
                          self.h_part.multicast_ptr, self.rank, self.scratch.data_ptr(), a.step.data_ptr(),
                          sync=self._sync(0) if dev_sync else None, peers=self.peer_grads,
                          inbox=self.inbox.data_ptr() if self.rs is not None else None,
                          ranges=self.rs.merged_ranges() if (self.rs is not None and not pull_all) else None)
        if not dev_sync:
            self._barrier(self.h_part)                   # all partial norms have landed everywhere
        L.dp_adam_bcast(self.h_flat.multicast_ptr, a.flat.data_ptr(), a.grad.data_ptr(), a.exp_avg.data_ptr(),
                        a.exp_avg_sq.data_ptr(), self.begin, self.end, self.partials.data_ptr(), self.world, lr, max_norm,
                        a.step.data_ptr(), a.grad_norm.data_ptr(), sync=self._sync(1) if dev_sync else None)
        self._barrier(self.h_flat)                       # all slices of the new parameters have landed
        self.eng.wplanes.refresh()


class GlobalBatchSync:
    """Collectives that make an N-rank step equal to ONE step on the concatenated batch (SURVEY.md section 8e, items 1, 2
    and 6) instead of PyTorch-DDP's per-rank semantics:

      * BatchNorm: every rank Chan-merges its row-tile partials into one (sum, M2) record, the records are all-gathered
        and the norm kernel merges them (`stat_rows` = global rows); backward sum-all-reduces (sum g, sum g*xhat)
        between the reduction and the apply pass (fxn_bn_act_bwd phases 1 / 2).
      * MSE / cross-entropy means run over the valid labels of the GLOBAL batch: the valid counts are all-reduced and
        each rank normalises by count_global / world, so the mean over ranks of the rank losses (and of their
        gradients, which is what the gradient all-reduce forms) is the global mean.
      * Cox: risk scores, durations and events are all-gathered and the partial likelihood is evaluated over the global
        risk sets; each rank back-propagates the coefficients of its own rows.

    Every rank must hold the same number of rows per step. The supervised_vae MMD term couples all pairs of the batch
    and stays per-rank (documented in DESIGN.md). Works over any torch.distributed backend (NCCL on GPUs; gloo in the
    tests, also for CUDA tensors)."""

    def __init__(self, group=None):
        if not dist.is_initialized():
            raise RuntimeError("GlobalBatchSync needs an initialised torch.distributed process group")
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.calls = 0

    def all_gather(self, t: torch.Tensor) -> torch.Tensor:
        """[...] on every rank -> [world, ...] (rank-major), same on every rank."""
        self.calls += 1
        t = t.contiguous()
        out = [torch.empty_like(t) for _ in range(self.world)]
        dist.all_gather(out, t, group=self.group)
        return torch.stack(out, 0)

    def all_reduce(self, t: torch.Tensor) -> torch.Tensor:
        """In-place sum over ranks (t must be contiguous)."""
        self.calls += 1
        dist.all_reduce(t, group=self.group)
        return t


def merge_stat_records(records: torch.Tensor, rows_per_rank: int):
    """Host restatement of what fxn_bn_act_fwd does with gathered records (used by the CPU protocol test):
    records [world, 2, cols] = per-rank (sum, M2 about the rank mean) over rows_per_rank rows -> (mean, biased var) of the
    world * rows_per_rank rows, by Chan's parallel formula."""
    world = records.shape[0]
    n = float(rows_per_rank)
    total = records[:, 0].sum(0)
    mean = total / (world * n)
    d = records[:, 0] / n - mean
    m2 = (records[:, 1] + n * d * d).sum(0)
    return mean, m2 / (world * n)


# ----------------------------------------------------------------------------------------------------
# host placement: one process per GPU, on the CPU cores next to it
# ----------------------------------------------------------------------------------------------------
def parse_cpulist(text: str):
    """'0-3,8,10-11' (the format of /sys/devices/system/node/node*/cpulist) -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs (None when the platform does not say, e.g. single-socket boxes report -1)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index)
        bdf = f"{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        with open(os.path.join("/sys/bus/pci/devices", bdf, "numa_node")) as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def pin_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPU cores of the NUMA node its GPU hangs off, so that the pinned staging buffers of the
    host -> device batch path (fit.HostStreamTrainer, data.load_matrix_npy) are allocated and filled on the socket whose
    PCIe root the DMA goes through. Returns the core list, or None when the topology is unknown (nothing is changed)."""
    import os
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:
        return None
