"""Global-batch data parallelism: with parallel.GlobalBatchSync, W ranks that each hold 1/W of a
batch must reproduce ONE step on the concatenated batch -- BatchNorm statistics, MSE / cross-entropy means over the
globally valid labels, Cox risk sets -- so "N GPUs == 1 GPU on the concatenated batch" is a parity statement:

  * CPU (not gpu): the protocol itself (record merge, count normalisation, two-phase BatchNorm backward) over gloo,
    world size 2, against autograd on the global batch;
  * GPU: the engine on two ranks (sharing cuda:0, gloo) against the CPU oracle's step on the global batch.
"""
import os
import socket
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


# ------------------------------------------------------------------------------------------------
# CPU: the protocol over gloo
# ------------------------------------------------------------------------------------------------
def _protocol_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from flexynesis_b200.parallel import GlobalBatchSync, merge_stat_records
    sync = GlobalBatchSync()
    g = torch.Generator().manual_seed(0)
    B, d, h = 24, 7, 5
    X = torch.randn(world * B, d, generator=g, dtype=torch.float64)
    W1 = torch.randn(h, d, generator=g, dtype=torch.float64)
    gamma, beta = torch.rand(h, generator=g, dtype=torch.float64) + 0.5, torch.randn(h, generator=g, dtype=torch.float64)
    w2 = torch.randn(h, generator=g, dtype=torch.float64)
    y = torch.randn(world * B, generator=g, dtype=torch.float64)
    y[torch.rand(world * B, generator=g) < 0.3] = float("nan")          # ragged: valid counts differ between ranks
    x, yl = X[rank * B:(rank + 1) * B], y[rank * B:(rank + 1) * B]
    # ---- forward: local (sum, M2) record -> all-gather -> Chan merge ----
    z = x @ W1.T
    rec = torch.stack([z.sum(0), ((z - z.mean(0)) ** 2).sum(0)])
    mean, var = merge_stat_records(sync.all_gather(rec), B)
    rstd = torch.rsqrt(var + 1e-5)
    xhat = (z - mean) * rstd
    a = torch.relu(xhat * gamma + beta)
    o = a @ w2
    valid = ~torch.isnan(yl)
    cnt = torch.tensor([float(valid.sum())], dtype=torch.float64)
    sync.all_reduce(cnt)
    denom = cnt / world                                                 # count_global / world
    diff = torch.where(valid, o - torch.nan_to_num(yl), torch.zeros_like(o))
    loss_r = (diff ** 2).sum() / denom
    # ---- backward of the rank loss, BatchNorm in two phases ----
    do = 2 * diff / denom
    dw2 = a.T @ do
    da = do[:, None] * w2[None, :]
    gpre = da * (xhat * gamma + beta > 0)
    sums = torch.stack([gpre.sum(0), (gpre * xhat).sum(0)])             # phase 1: this rank's share of dbeta / dgamma
    dbeta, dgamma = sums[0].clone(), sums[1].clone()
    sync.all_reduce(sums)                                               # between the phases
    n = world * B
    dz = gamma * rstd * (gpre - sums[0] / n - xhat * sums[1] / n)       # phase 2
    dW1 = dz.T @ x
    grads = torch.cat([dW1.flatten(), dgamma, dbeta, dw2])
    sync.all_reduce(grads)
    grads /= world
    lr_ = loss_r.clone().reshape(1)
    sync.all_reduce(lr_)
    if rank == 0:
        q.put((grads, lr_ / world, (X, W1, gamma, beta, w2, y)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_global_batch_protocol_matches_single_process_autograd():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_protocol_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, loss, (X, W1, gamma, beta, w2, y) = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    W1, gamma, beta, w2 = [t.clone().requires_grad_(True) for t in (W1, gamma, beta, w2)]
    z = X @ W1.T
    xhat = (z - z.mean(0)) * torch.rsqrt(z.var(0, unbiased=False) + 1e-5)
    o = torch.relu(xhat * gamma + beta) @ w2
    valid = ~torch.isnan(y)
    ref = ((o[valid] - y[valid]) ** 2).mean()
    ref.backward()
    want = torch.cat([W1.grad.flatten(), gamma.grad, beta.grad, w2.grad])
    assert abs(float(loss) - float(ref)) < 1e-12
    assert float((grads - want).abs().max()) < 1e-12 * max(1.0, float(want.abs().max()))


# ------------------------------------------------------------------------------------------------
# GPU: engine on two ranks vs the oracle on the global batch
# ------------------------------------------------------------------------------------------------
VT = {"y": "numerical", "c": "categorical", "e": "numerical", "t": "numerical"}


@pytest.mark.gpu
@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [2, 3])
def test_sharded_engine_step_equals_global_batch_step(tmp_path, world):
    from oracle.restatement import Spec
    from test_gpu_parity import Report, grad_keep_mask, oracle_reference
    spec = Spec(model="DirectPred", input_dims=[260, 120], latent_dim=40, hidden_dim_factor=0.2, supervisor_hidden_dim=16,
                variables=["y", "c", "e"], variable_types=VT, num_classes={"c": 4}, surv_event_var="e", surv_time_var="t")
    B = 96 * world                       # 96 rows per rank: not a multiple of the 128-row statistics tile
    P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=2)
    path = str(tmp_path / "rec.pt")
    torch.save(dict(spec=spec.__dict__, batch=batch, lr=1e-3, P0=P0,
                    steps=[dict(P_before=s["P_before"], noise=s["noise"]) for s in steps]), path)
    port = str(_free_port())
    worker = os.path.join(ROOT, "tests", "helpers", "global_batch_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port, path]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=500) == 0
    out = torch.load(path + ".out", weights_only=False)
    assert out["sync_calls"] > 0
    rep = Report()
    for s, (st, got) in enumerate(zip(steps, out["steps"])):
        for k, v in st["outputs"].items():
            rep.close(f"step{s} outputs[{k}]", got["logits"][k], v)
        for k, v in st["losses"].items():
            rep.close(f"step{s} loss[{k}]", torch.tensor(got["losses"]["__total__" if k == "train_loss" else k]), v, atol=1e-5)
        gmax = max(float(g.abs().max()) for g in st["grads"].values() if g is not None)
        for k, g in st["grads"].items():
            if g is None:
                continue
            atol = 1e-4 * gmax if float(g.abs().max()) < 1e-3 * gmax else 1e-6
            rep.close(f"step{s} grad[{k}]", got["grads"][k], g, atol=atol, keep=grad_keep_mask(k, g, st["flagged"]))
        for k, b in got["buffers"].items():
            if "running_" in k:
                rep.close(f"step{s} buffer[{k}]", b, st["P_after"][k], rtol=1e-4)
    rep.finish()
