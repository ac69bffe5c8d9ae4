"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sample sharding, the flat gradient arena layout and
the all-reduce + 1/W scaling must reproduce the mean of the per-rank gradients that the oracle computes on each shard."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.restatement import Noise, Spec, Trainer, init_params, synthetic_batch

VT = {"y": "numerical", "c": "categorical"}
SPEC = dict(model="DirectPred", input_dims=[40, 24], latent_dim=16, hidden_dim_factor=0.4, supervisor_hidden_dim=8,
            variables=["c", "y"], variable_types=VT, num_classes={"c": 3})
N = 96


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _DS:
    def __init__(self, dat, ann):
        self.dat, self.ann, self.variable_types = dat, ann, VT
        self.features = {k: list(range(v.shape[1])) for k, v in dat.items()}
        self.samples = [f"s{i}" for i in range(N)]

    def __len__(self):
        return N


def _shard_grads(P0, spec, dat, ann, lo, hi, seed):
    P = {k: v.clone() for k, v in P0.items()}
    tr = Trainer(P, spec, 1e-3)
    torch.manual_seed(seed)                        # dropout masks of the shard
    res = tr.step(({k: v[lo:hi] for k, v in dat.items()}, {k: v[lo:hi] for k, v in ann.items()}, None), Noise())
    return res["grads"]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        import flexynesis_b200 as fx
        from flexynesis_b200.engine import ParamArena
        from flexynesis_b200.parallel import DatasetShard, GradAllReduce, broadcast_parameters, shard_range
        spec = Spec(**SPEC)
        torch.manual_seed(0)
        P0 = init_params(spec)
        dat, ann = synthetic_batch(spec, N, 0)
        full = _DS(dat, ann)
        shard = DatasetShard(full, rank, world)
        lo, hi = shard_range(N, rank, world)
        assert len(shard) == N // world and shard.samples[0] == f"s{lo}"
        ctor = _DS(dat, {k: torch.nan_to_num(v, nan=0.0) for k, v in ann.items()})
        torch.manual_seed(100 + rank)              # ranks start from different inits ...
        model = fx.DirectPred({"latent_dim": 16, "hidden_dim_factor": 0.4, "supervisor_hidden_dim": 8, "lr": 1e-3},
                              ctor, ["c", "y"], device_type="cpu")
        arena = ParamArena(model, torch.device("cpu"))
        broadcast_parameters(arena.flat, dict(model.named_buffers()))   # ... and must agree after the broadcast
        gathered = [torch.empty_like(arena.flat) for _ in range(world)]
        dist.all_gather(gathered, arena.flat)
        assert all(torch.equal(g, gathered[0]) for g in gathered)
        # per-rank oracle gradients of this rank's shard -> arena layout -> all-reduce -> 1 / W
        mine = _shard_grads(P0, spec, shard.dat, shard.ann, 0, len(shard), 7 + rank)
        for k in arena.names:
            if mine[k] is not None:
                arena.view(k, arena.grad).copy_(mine[k])
        GradAllReduce(world, chunks=3)(arena.grad)
        arena.grad.mul_(1.0 / world)
        # expectation computed locally from both shards
        want = {k: None for k in arena.names}
        for r in range(world):
            a, b = shard_range(N, r, world)
            g = _shard_grads(P0, spec, dat, ann, a, b, 7 + r)
            for k in arena.names:
                if g[k] is not None:
                    want[k] = g[k] / world if want[k] is None else want[k] + g[k] / world
        worst = 0.0
        for k in arena.names:
            got = arena.view(k, arena.grad)
            if want[k] is None:
                assert float(got.abs().max()) == 0.0, k
                continue
            worst = max(worst, float((got - want[k]).abs().max()) / max(float(want[k].abs().max()), 1e-12))
        out[rank] = worst
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_gradients_allreduce_to_global_mean():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world
        for r in range(world):
            assert out[r] < 1e-6, out[r]


def test_shard_ranges_partition_the_dataset():
    from flexynesis_b200.parallel import shard_range
    for n, w in [(96, 2), (4096, 8), (1001, 4)]:
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert all(b - a == n // w for a, b in spans)
