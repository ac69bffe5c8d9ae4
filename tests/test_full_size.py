"""Parity at BASELINE.json's FULL sizes (configs 2-5 as mapped in SURVEY.md section 8): one training step of the engine
against the CPU oracle on the same seeded inputs and replayed noise (the oracle needs seconds per step at these
sizes), plus a size-independent property (row-permutation invariance of loss and gradients) on config 2.
Tolerance: 1e-3 relative (BASELINE.json north_star), as in test_gpu_parity.py."""
import pytest
import torch

from oracle.restatement import Spec, synthetic_graph
from test_gpu_parity import (Report, build_model, compare_step, masks_from_noise, oracle_reference, sync_state, to_cuda)

pytestmark = pytest.mark.gpu
VT = {"y": "numerical", "c": "categorical", "e": "numerical", "t": "numerical"}

FULL = {
    # config 2: DirectPred, 2 omics 4096x5000 + 4096x3000, intermediate fusion, 5-class head
    "cfg2": (Spec(model="DirectPred", input_dims=[5000, 3000], latent_dim=256, hidden_dim_factor=0.1024,
                  supervisor_hidden_dim=32, variables=["c"], variable_types=VT, num_classes={"c": 5}), 4096),
    # config 3: supervised_vae, same omics, latent 128, MMD + reconstruction + Cox head
    "cfg3": (Spec(model="supervised_vae", input_dims=[5000, 3000], latent_dim=128, hidden_dim_factor=0.1024,
                  supervisor_hidden_dim=32, variables=["e"], variable_types=VT, surv_event_var="e", surv_time_var="t"), 4096),
    # config 5, one rank's share: DirectPred early fusion 4096x24000 -> 1024 -> 512, regression + 5-class heads
    "cfg5_rank": (Spec(model="DirectPred", input_dims=[24000], latent_dim=512, hidden_dim_factor=0.04267,
                       supervisor_hidden_dim=256, variables=["y", "c"], variable_types=VT, num_classes={"c": 5}), 4096),
}


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", list(FULL))
def test_full_size_step_matches_oracle(name):
    spec, B = FULL[name]
    P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=1)
    model = build_model(spec, batch, 1e-3, P0)
    model.train()
    cb = to_cuda(batch)
    rep = Report()
    sync_state(model, steps[0]["P_before"])
    compare_step(rep, model, spec, batch, cb, steps[0], 0, steps[0]["P_before"], 1e-3)
    rep.finish()


@pytest.mark.timeout(900)
def test_full_size_gnn_matches_oracle():
    """config 4 graph at full size (2000 genes, 20 000 directed edges, GCN 1 -> 32 -> 32, fc 64000 -> 128): a training step
    at B = 512 (the oracle's autograd graph at B = 4096 needs tens of GB), and the eval-mode forward at B = 4096."""
    spec = Spec(model="GNN", input_dims=[1], latent_dim=128, supervisor_hidden_dim=32, variables=["y"], variable_types=VT,
                node_count=2000, node_embedding_dim=32, num_convs=2, activation="relu")
    edge_index = synthetic_graph(2000, 20000, 0)
    P0, batch, steps, _ = oracle_reference(spec, 512, 1e-3, steps=1, edge_index=edge_index)
    model = build_model(spec, batch, 1e-3, P0, edge_index)
    model.train()
    rep = Report()
    sync_state(model, steps[0]["P_before"])
    steps[0]["flagged"] = {}
    compare_step(rep, model, spec, batch, to_cuda(batch), steps[0], 0, steps[0]["P_before"], 1e-3)
    # eval forward at the full batch
    from oracle.restatement import Noise, forward
    g = torch.Generator().manual_seed(7)
    x = torch.randn(4096, 2000, 1, generator=g)
    sync_state(model, steps[0]["P_before"])
    model.eval()
    with torch.no_grad():
        # eval mode has no batch coupling: the oracle runs in 512-row chunks to bound host memory
        Pe = {k: v.clone() for k, v in steps[0]["P_before"].items()}
        want = torch.cat([forward(Pe, spec, (x[i:i + 512], {"y": torch.zeros(512)}, None), False, Noise(),
                                  edge_index)["outputs"]["y"] for i in range(0, 4096, 512)], 0)
        got = model.forward(x.cuda())
    rep.close("eval outputs[y] at B=4096", got["y"], want)
    rep.finish()


@pytest.mark.timeout(600)
def test_full_size_row_permutation_invariance():
    """Size-independent property at config 2's full size: permuting the samples of the batch (inputs, labels and dropout
    masks together) changes neither the loss nor any parameter gradient beyond fp32 summation-order noise. Exercises
    every row tile, the BatchNorm partial merge and the wgrad reductions over the whole batch without an oracle."""
    spec, B = FULL["cfg2"]
    from oracle.restatement import init_params, synthetic_batch
    torch.manual_seed(0)
    P0 = init_params(spec)
    dat, y = synthetic_batch(spec, B, 3)
    model = build_model(spec, (dat, y, None), 1e-3, P0)
    model.train()
    eng = model.engine()
    g = torch.Generator().manual_seed(11)
    h = [e.layer_1.out_features for e in model.encoders]
    masks = {f"encoders.{i}.dropout": (torch.rand(B, h[i], generator=g) > 0.1).to(torch.uint8).cuda() for i in range(2)}
    masks["MLPs.c.dropout"] = (torch.rand(B, 32, generator=g) > 0.1).to(torch.uint8).cuda()

    def run(perm):
        d = {k: v[perm].cuda() for k, v in dat.items()}
        yy = {k: v[perm].cuda() for k, v in y.items()}
        mk = {k: v[perm.cuda()].contiguous() for k, v in masks.items()}
        groups, lab = model._split_batch((d, yy, None))
        ws = eng.forward_backward(groups, lab, mk)
        return float(eng.losses(ws)["__total__"]), eng.arena.grad.clone()

    l0, g0 = run(torch.arange(B))
    l1, g1 = run(torch.randperm(B, generator=g))
    assert abs(l0 - l1) <= 1e-5 * abs(l0), (l0, l1)
    scale = float(g0.abs().max())
    assert float((g0 - g1).abs().max()) <= 2e-5 * scale
