"""Parity at BASELINE.json's FULL sizes (configs 2-5 as mapped in SURVEY.md section 8): one training step of the engine
against the CPU oracle on the same seeded inputs and replayed noise (the oracle needs seconds per step at these
sizes), plus a size-independent property (row-permutation invariance of loss and gradients) on config 2.
Tolerance: 1e-3 relative (BASELINE.json north_star), as in test_gpu_parity.py."""
import pytest
import torch

from oracle.restatement import Spec, synthetic_graph
from test_gpu_parity import (Report, build_model, compare_step, gnn_oracle_steps, masks_from_noise, oracle_reference,
                             sync_state, to_cuda)

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


def _center_gap(v: torch.Tensor) -> torch.Tensor:
    """v [B, units]: values whose SIGN is an activation gate. Returns a per-unit shift that puts zero in the middle of the
    widest gap between consecutive sorted values near the zero crossing."""
    s, _ = torch.sort(v.detach(), 0)
    B = s.shape[0]
    idx = (s < 0).sum(0)                                        # first non-negative position per unit
    pos = (idx[None, :] + torch.arange(-8, 9)[:, None]).clamp(0, B - 1)
    win = torch.gather(s, 0, pos)                               # 17 consecutive sorted values around the crossing
    gaps = win[1:] - win[:-1]
    j = gaps.argmax(0)
    lo = torch.gather(win, 0, j[None, :])[0]
    hi = torch.gather(win, 0, (j + 1)[None, :])[0]
    delta = -(lo + hi) / 2
    assert float((v + delta).abs().min()) > 1e-5, "could not establish a gate margin"
    return delta


def _bn_train(P, prefix, z):
    return (z - z.mean(0)) * torch.rsqrt(z.var(0, unbiased=False) + 1e-5) * P[prefix + ".weight"] + P[prefix + ".bias"]


def gate_safe_state(spec, P, batch):
    """ReLU / LeakyReLU gates are discontinuities of the gradient. At these sizes (millions of gated values per layer)
    a seeded init always has pre-activations within fp32 rounding (1e-6) of zero, where two correct implementations
    may gate differently and one sample's term of every upstream gradient flips with it (1/sqrt(B) of an entry, far
    above 1e-3). This nudges the bias that sits right in front of each gate (< 1e-2, per unit, in forward order) so that
    every gated value keeps a margin of >= 1e-5 from zero; the comparison that follows is then strict everywhere.
    Returns the noise tensors (dropout masks, epsilon, MMD priors) drawn once and replayed on both sides."""
    import copy
    import torch.nn.functional as Fn
    from oracle.restatement import Noise, forward, vae_encoder, _fused_embedding
    torch.manual_seed(1000)
    n0 = Noise()
    with torch.no_grad():
        forward(copy.deepcopy(P), spec, batch, True, n0)
        given = dict(n0.record)
        xs = list(batch[0].values())
        if spec.model == "DirectPred":
            for i, x in enumerate(xs):
                z = Fn.linear(x, P[f"encoders.{i}.layer_1.weight"], P[f"encoders.{i}.layer_1.bias"])
                P[f"encoders.{i}.batchnorm.bias"] += _center_gap(_bn_train(P, f"encoders.{i}.batchnorm", z))
            emb = _fused_embedding(copy.deepcopy(P), spec, xs, True, Noise(given))
        else:
            for i, x in enumerate(xs):
                z = Fn.linear(x, P[f"encoders.{i}.hidden_layers.0.weight"], P[f"encoders.{i}.hidden_layers.0.bias"])
                P[f"encoders.{i}.hidden_layers.0.bias"] += _center_gap(z)
            Pc = copy.deepcopy(P)
            means, logvars = zip(*[vae_encoder(Pc, f"encoders.{i}", x, True) for i, x in enumerate(xs)])
            mean = Fn.linear(torch.cat(means, 1), P["FC_mean.weight"], P["FC_mean.bias"])
            log_var = Fn.linear(torch.cat(logvars, 1), P["FC_log_var.weight"], P["FC_log_var.bias"])
            emb = mean + log_var * given["epsilon"]
            for i in range(len(xs)):
                z = Fn.linear(emb, P[f"decoders.{i}.hidden_layers.0.weight"], P[f"decoders.{i}.hidden_layers.0.bias"])
                P[f"decoders.{i}.hidden_layers.0.bias"] += _center_gap(z)
        for v in spec.variables:
            z = Fn.linear(emb, P[f"MLPs.{v}.layer_1.weight"], P[f"MLPs.{v}.layer_1.bias"])
            P[f"MLPs.{v}.batchnorm.bias"] += _center_gap(_bn_train(P, f"MLPs.{v}.batchnorm", z))
    return given


@pytest.mark.timeout(900)
@pytest.mark.parametrize("name", list(FULL))
def test_full_size_step_matches_oracle(name):
    from oracle.restatement import init_params, synthetic_batch
    spec, B = FULL[name]
    torch.manual_seed(0)
    P = init_params(spec)
    dat, y = synthetic_batch(spec, B, 0)
    batch = (dat, y, None)
    given = gate_safe_state(spec, P, batch)
    P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=1, batch=batch, P=P, given=given)
    st = steps[0]
    st["flagged"] = {}                    # every gate has a verified margin: strict comparison of all gradients
    model = build_model(spec, batch, 1e-3, P0)
    model.train()
    cb = to_cuda(batch)
    rep = Report()
    sync_state(model, st["P_before"])
    compare_step(rep, model, spec, batch, cb, st, 0, st["P_before"], 1e-3)
    rep.finish()


@pytest.mark.timeout(900)
def test_full_size_gnn_matches_oracle():
    """config 4 graph at full size (2000 genes, 20 000 directed edges, GCN 1 -> 32 -> 32, fc 64000 -> 128): a training step
    at B = 512 (the oracle's autograd graph at B = 4096 needs tens of GB), and the eval-mode forward at B = 4096."""
    spec = Spec(model="GNN", input_dims=[1], latent_dim=128, supervisor_hidden_dim=32, variables=["y"], variable_types=VT,
                node_count=2000, node_embedding_dim=32, num_convs=2, activation="relu")
    edge_index = synthetic_graph(2000, 20000, 0)
    P0, batch, steps = gnn_oracle_steps(spec, 512, 1e-3, 1, edge_index)
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
