"""GPU parity: the B200 engine (through the drop-in classes and the C ABI) against
 (a) the golden vectors recorded from the reference's own source (tests/golden/*.pt), and
 (b) the CPU oracle (oracle/restatement.py) on seeded synthetic inputs at larger sizes.
Tolerance: 1e-3 relative (BASELINE.json north_star), with an absolute floor for quantities that are analytically
zero; noise (dropout masks, epsilon, MMD prior) is replayed from the recorded / oracle-drawn tensors.
"""
import glob
import os

import pytest
import torch

from oracle.restatement import Noise, Spec, Trainer, forward, synthetic_batch

pytestmark = pytest.mark.gpu
GOLDEN_DIR = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-3


class _DS:
    """dataset duck type for the constructors"""

    def __init__(self, dat, ann, variable_types):
        self.dat, self.variable_types = dat, variable_types
        self.ann = {k: torch.nan_to_num(v, nan=0.0) for k, v in ann.items()}
        self.features = {k: list(range(v.shape[1])) for k, v in dat.items()}
        self.samples = [f"s{i}" for i in range(next(iter(dat.values())).shape[0])]

    def __len__(self):
        return len(self.samples)


def _close(tag, got, ref, rtol=RTOL, atol=1e-6):
    got, ref = got.detach().double().cpu().flatten(), ref.detach().double().cpu().flatten()
    scale = max(float(ref.abs().max()), 1e-30)
    err = float((got - ref).abs().max())
    assert err <= atol + rtol * scale, f"{tag}: abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})"


class _GDS:
    """MultiOmicDatasetNW duck type: node features [B, N, F], labels, one shared edge_index"""

    def __init__(self, x, ann, variable_types, edge_index):
        self.node_features_tensor, self.edge_index, self.variable_types = x, edge_index, variable_types
        self.ann = {k: torch.nan_to_num(v, nan=0.0) for k, v in ann.items()}
        self.samples = [f"s{i}" for i in range(x.shape[0])]

    def __getitem__(self, i):
        return self.node_features_tensor[i], {k: v[i] for k, v in self.ann.items()}, self.samples[i]

    def __len__(self):
        return len(self.samples)


def build_model(spec: Spec, batch, lr, P0=None, edge_index=None):
    import flexynesis_b200 as fx
    cfg = {"latent_dim": spec.latent_dim, "hidden_dim_factor": spec.hidden_dim_factor,
           "supervisor_hidden_dim": spec.supervisor_hidden_dim, "lr": lr,
           "node_embedding_dim": spec.node_embedding_dim, "num_convs": spec.num_convs, "activation": spec.activation}
    targets = [v for v in spec.variables if v != spec.surv_event_var]
    if spec.model == "GNN":
        ds = _GDS(batch[0], batch[1], spec.variable_types, edge_index)
        model = fx.GNN(cfg, ds, targets, surv_event_var=spec.surv_event_var, surv_time_var=spec.surv_time_var,
                       use_loss_weighting=spec.use_loss_weighting, device_type="gpu", gnn_conv_type=spec.conv)
        if P0 is not None:
            model.load_state_dict(P0, strict=True)
        return model.cuda()
    if spec.model == "MultiTripletNetwork":
        ds = _DS(batch[0], batch[3], spec.variable_types)
        cls = fx.MultiTripletNetwork
    else:
        ds = _DS(batch[0], batch[1], spec.variable_types)
        cls = getattr(fx, spec.model)
    extra = {}
    if spec.model == "CrossModalPred":
        keys = list(batch[0].keys())
        extra = dict(input_layers=[keys[i] for i in spec.in_idx] if spec.in_idx is not None else None,
                     output_layers=[keys[i] for i in spec.out_idx] if spec.out_idx is not None else None)
    model = cls(cfg, ds, targets, surv_event_var=spec.surv_event_var, surv_time_var=spec.surv_time_var,
                use_loss_weighting=spec.use_loss_weighting, device_type="gpu", **extra)
    if P0 is not None:
        model.load_state_dict(P0, strict=True)
    return model.cuda()


def to_cuda(batch):
    def mv(o):
        if torch.is_tensor(o):
            return o.cuda()
        if isinstance(o, dict):
            return {k: mv(v) for k, v in o.items()}
        return o
    return tuple(mv(o) for o in batch)


def masks_from_noise(noise):
    """Replayed noise for the engine: dropout keep-masks as uint8, Gaussian draws (epsilon, MMD priors) as fp32."""
    out = {}
    for k, v in noise.items():
        if "dropout" in k:
            out[k] = (v != 0).to(torch.uint8).cuda().contiguous()
        else:
            out[k] = v.float().cuda().contiguous()
    return out


class Report:
    """Collects every deviation of a scenario so that one GPU run shows all of them."""

    def __init__(self):
        self.errors = []

    def close(self, tag, got, ref, rtol=RTOL, atol=1e-6, keep=None):
        got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
        if keep is not None:
            got, ref = got[keep], ref[keep]
        if ref.numel() == 0:
            return
        scale = max(float(ref.abs().max()), 1e-30)
        err = float((got - ref).abs().max())
        if os.environ.get("FXN_PARITY_VERBOSE"):
            print(f"[parity] {tag}: rel {err / scale:.3e} (scale {scale:.3e})")
        if not err <= atol + rtol * scale:
            self.errors.append(f"{tag}: abs err {err:.3e} vs scale {scale:.3e} (rel {err / scale:.3e})")
        if keep is not None:
            self.skipped = getattr(self, "skipped", 0) + int((~keep).sum())

    def finish(self):
        if getattr(self, "skipped", 0):
            print(f"[parity] {self.skipped} gradient elements excluded (hidden units with a ReLU gate within 2e-4 of zero)")
        assert not self.errors, "\n".join(self.errors)


def gate_margin_units(spec, P, batch, res, masks, margin=2e-4):
    """ReLU gates are discontinuities of the gradient: a pre-activation within rounding distance of 0 may fall on either
    side in two correct implementations, which changes that hidden unit's gradient by O(1/B). Returns, per MLP block
    prefix, a bool vector of hidden units whose smallest |BN output| over the batch is below `margin` (computed with
    the oracle's parameters); the comparison skips exactly those units' layer_1.weight rows and batchnorm grads."""
    import torch.nn.functional as Fn
    out = {}

    def block(prefix, x):
        if prefix + ".layer_1.weight" not in P:
            # Encoder / Decoder: Linear -> LeakyReLU(0.2) -> BN; the gate sits directly on the Linear output
            z = Fn.linear(x, P[prefix + ".hidden_layers.0.weight"], P[prefix + ".hidden_layers.0.bias"])
            out[prefix] = z.abs().min(0).values < margin * float(z.abs().max())
            return
        z = Fn.linear(x, P[prefix + ".layer_1.weight"], P[prefix + ".layer_1.bias"])
        yv = (z - z.mean(0)) * torch.rsqrt(z.var(0, unbiased=False) + 1e-5) * P[prefix + ".batchnorm.weight"] \
            + P[prefix + ".batchnorm.bias"]
        out[prefix] = out.get(prefix, torch.zeros(z.shape[1], dtype=torch.bool)) | (yv.abs().min(0).values < margin)

    with torch.no_grad():
        if spec.model == "MultiTripletNetwork":
            for d in batch[:3]:
                for i, x in enumerate(d.values()):
                    block(f"encoders.{i}", x)
        elif spec.model == "CrossModalPred":
            layers = list(batch[0].values())
            for i, li in enumerate(spec.in_idx if spec.in_idx is not None else range(len(layers))):
                block(f"encoders.{i}", layers[li])
        elif spec.model != "GNN":
            for i, x in enumerate(batch[0].values()):
                block(f"encoders.{i}", x)
        for v in spec.variables:
            block(f"MLPs.{v}", res["embedding"].detach())
        if spec.model in ("supervised_vae", "CrossModalPred"):
            nd = len(spec.out_idx) if (spec.model == "CrossModalPred" and spec.out_idx is not None) else len(spec.input_dims)
            for i in range(nd):
                block(f"decoders.{i}", res["embedding"].detach())
    return out


def grad_keep_mask(name, g, flagged):
    """elements of parameter `name` that are compared strictly (everything except flagged hidden units)"""
    for prefix, units in flagged.items():
        if name.startswith(prefix + ".") and bool(units.any()):
            if name.endswith("layer_1.weight") or name.endswith("hidden_layers.0.weight"):
                return (~units)[:, None].expand_as(g)
            if name.endswith(("batchnorm.weight", "batchnorm.bias", "hidden_layers.0.bias")):
                return ~units
    return None


def sync_state(model, P):
    """Load the oracle's parameters/buffers into the engine-backed model (copies in place into the arena)."""
    model.load_state_dict({k: v.detach() for k, v in P.items()}, strict=True)


def compare_step(rep, model, spec, batch, cb, st, s, P_before_cpu, lr):
    eng = model.engine()
    masks = masks_from_noise(st["noise"])
    groups, y = model._split_batch(cb)
    ws = eng.forward_backward(groups, y, masks)
    for k, v in st["outputs"].items():
        rep.close(f"step{s} outputs[{k}]", ws["heads"]["logits"][k], v)
    if st.get("embedding") is not None and spec.model in ("supervised_vae", "CrossModalPred"):
        rep.close(f"step{s} z", eng.embedding(ws), st["embedding"])
    vals = eng.losses(ws)
    for k, v in st["losses"].items():
        if k == "train_loss":
            rep.close(f"step{s} total", vals["__total__"], v)
        else:
            rep.close(f"step{s} loss[{k}]", vals[k], v, atol=1e-5)
    flagged = st.get("flagged", {})
    gmax = max(float(g.abs().max()) for g in st["grads"].values() if g is not None)
    for k, g in st["grads"].items():
        got = eng.arena.view(k, eng.arena.grad)
        if g is None:
            if float(got.abs().max()) != 0.0:
                rep.errors.append(f"step{s} {k} must have no gradient")
            continue
        # element-wise, relative to the largest element of this parameter's gradient; parameters whose whole gradient
        # is rounding noise (biases in front of a BatchNorm) are compared against the model-wide scale instead
        atol = 1e-4 * gmax if float(g.abs().max()) < 1e-3 * gmax else 1e-6
        rep.close(f"step{s} grad[{k}]", got, g, atol=atol, keep=grad_keep_mask(k, g, flagged))
    return eng


GOLDEN = [p for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.pt")))
          if os.path.basename(p).startswith(("directpred", "triplet", "supervised_vae", "gnn", "crossmodal"))]


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_engine_matches_reference_golden(path):
    """Every step of every golden scenario, starting from the reference's own pre-step state."""
    g = torch.load(path, weights_only=False)
    spec = Spec(**g["spec"])
    rep = Report()
    model = build_model(spec, g["batch"], g["lr"], g["P0"], g.get("edge_index"))
    if g["eval_outputs0"] is not None and spec.model == "GNN":
        model.eval()
        with torch.no_grad():
            out = model.forward(g["batch"][0].cuda())
        for k, v in g["eval_outputs0"].items():
            rep.close(f"initial eval outputs[{k}]", out[k], v)
    if g["eval_outputs0"] is not None and spec.model == "DirectPred":
        model.eval()
        with torch.no_grad():
            out = model.forward([x.cuda() for x in g["batch"][0].values()])
        for k, v in g["eval_outputs0"].items():
            rep.close(f"initial eval outputs[{k}]", out[k], v)
    model.train()
    cb = to_cuda(g["batch"])
    for s, st in enumerate(g["steps"]):
        sync_state(model, st["P_before"])
        # flag near-zero ReLU gates with the oracle forward at this state
        P = {k: v.clone() for k, v in st["P_before"].items()}
        res = forward(P, spec, g["batch"], True, Noise(st["noise"]), g.get("edge_index"))
        st["flagged"] = gate_margin_units(spec, st["P_before"], g["batch"], res, None)
        compare_step(rep, model, spec, g["batch"], cb, st, s, st["P_before"], g["lr"])
    rep.finish()


def gate_safe_conv_masks(spec, P, x, edge_index, drawn, margin=1e-4):
    """Dropout masks for the flexGCN layers that also drop every element whose ReLU gate is marginal.

    A conv layer has B * N * emb gates; a forward value that agrees with the oracle to 1e-5 (the bf16x3 tensor-core
    transform of GNNEngine.gemm_layers) still lands a few of the millions of near-zero pre-activations on the other side
    of the gate, and every flip moves the layer's gradient sums by one element's worth (measured: 1e-3 .. 7e-3 of the
    gradient scale, i.e. about sqrt(flip probability), against 6e-6 for the fp32 transform). MLP blocks deal with this by
    skipping hidden units with a marginal gate (gate_margin_units); a conv layer has no per-unit gradient to skip, so the
    marginal ELEMENTS (|BatchNorm output| < margin, computed with the oracle) are dropped by the explicit dropout mask
    that both implementations replay: their gradient is then zero whichever way the gate falls, and everything else is
    held to RTOL. Returns (noise dict for Noise(given=...), number of elements dropped this way)."""
    from oracle.restatement import ACTS, CONVS, batchnorm
    Pc = {k: v.detach().clone() for k, v in P.items()}
    given, dropped = dict(drawn), 0
    with torch.no_grad():
        for k in range(spec.num_convs):
            h = CONVS[spec.conv](Pc, f"encoders.0.convs.{k}", x, edge_index)
            yb = batchnorm(Pc, f"encoders.0.bns.{k}", h.reshape(-1, h.shape[2]), True).view_as(h)
            m = drawn[f"encoders.0.dropout.{k}"].clone()
            if spec.activation == "relu":
                marg = yb.abs() < margin
                dropped += int((marg & (m != 0)).sum())
                m[marg] = 0
            given[f"encoders.0.dropout.{k}"] = m
            x = ACTS[spec.activation](yb) * m / 0.8
    return given, dropped


def gnn_oracle_steps(spec, B, lr, steps, edge_index):
    """oracle_reference for a GNN scenario with gate-safe conv dropout masks: every step is run twice on the CPU (once to
    draw the noise, once with the marginal gates dropped) and starts from the previous step's parameters."""
    P0 = batch = None
    P, out = None, []
    for s in range(steps):
        Pd, batch, drawn, _ = oracle_reference(spec, B, lr, steps=1, edge_index=edge_index, P=P, batch=batch)
        if P0 is None:
            P0 = Pd
        start = drawn[0]["P_before"]
        given, dropped = gate_safe_conv_masks(spec, start, batch[0], edge_index, drawn[0]["noise"])
        _, _, st, P = oracle_reference(spec, B, lr, steps=1, edge_index=edge_index, batch=batch,
                                       P={k: v.clone() for k, v in start.items()}, given=given)
        print(f"[parity] step {s}: {dropped} marginal conv-layer gates dropped through the replayed dropout mask")
        out.append(st[0])
    return P0, batch, out


def oracle_reference(spec, B, lr, steps, seed=0, batch=None, edge_index=None, P=None, given=None):
    """Run the CPU oracle for `steps` steps; records pre-step state, noise, results and Adam moments. `P` (optional):
    start from these parameters instead of a seeded init; `given` (optional): noise tensors replayed in every step."""
    torch.manual_seed(seed)
    from oracle.restatement import init_params
    if P is None:
        P = init_params(spec)
    P0 = {k: v.clone() for k, v in P.items()}
    if batch is None:
        dat, y = synthetic_batch(spec, B, seed)
        batch = (dat, y, None)
        if spec.model == "GNN":
            g = torch.Generator().manual_seed(seed + 1)
            batch = (torch.randn(B, spec.node_count, spec.input_dims[0], generator=g), y, None)
    tr = Trainer(P, spec, lr, edge_index=edge_index)
    out = []
    for s in range(steps):
        torch.manual_seed(1000 + s)
        noise = Noise(given)
        P_before = {k: v.detach().clone() for k, v in P.items()}
        adam_before = {k: {kk: (vv.clone() if torch.is_tensor(vv) else vv) for kk, vv in tr.opt.state[P[k]].items()}
                       for k in tr.names if P[k] in tr.opt.state}
        res = tr.step(batch, noise)
        out.append(dict(P_before=P_before, adam_before=adam_before, noise=noise.record,
                        embedding=res["embedding"].detach(),
                        outputs={k: v.detach() for k, v in res["outputs"].items()},
                        losses={**{k: v.detach() for k, v in res["losses"].items()}, "train_loss": res["total"].detach()},
                        grads=res["grads"], grad_norm=res["grad_norm"].detach(),
                        flagged=gate_margin_units(spec, P_before, batch, res, None),
                        P_after={k: v.detach().clone() for k, v in P.items()}))
    return P0, batch, out, {k: v.detach().clone() for k, v in P.items()}


VT = {"y": "numerical", "c": "categorical", "e": "numerical", "t": "numerical"}
CASES = {
    # BASELINE.json config 1: DirectPred 512x1000 -> 128 -> 64, one regression target
    "cfg1": (Spec(model="DirectPred", input_dims=[1000], latent_dim=64, hidden_dim_factor=0.128,
                  supervisor_hidden_dim=32, variables=["y"], variable_types=VT), 512),
    # config-2 architecture at reduced batch/features with ragged sizes (h = 307-like odd widths)
    "cfg2_small": (Spec(model="DirectPred", input_dims=[1000, 603], latent_dim=72, hidden_dim_factor=0.1024,
                        supervisor_hidden_dim=20, variables=["c"], variable_types=VT, num_classes={"c": 5}), 500),
    # multitask + survival + weighting, batch not a multiple of the 128-row tile
    "multitask": (Spec(model="DirectPred", input_dims=[700, 300, 120], latent_dim=48, hidden_dim_factor=0.2,
                       supervisor_hidden_dim=16, variables=["y", "c", "e"], variable_types=VT, num_classes={"c": 7},
                       surv_event_var="e", surv_time_var="t"), 333),
    # BASELINE.json config 3 architecture (supervised_vae, 2 omics, Cox head, MMD + reconstruction) at reduced size,
    # ragged widths (h = 61 / 36, batch not a multiple of 128)
    "svae_cox": (Spec(model="supervised_vae", input_dims=[600, 360], latent_dim=24, hidden_dim_factor=0.1024,
                      supervisor_hidden_dim=16, variables=["e"], variable_types=VT, surv_event_var="e",
                      surv_time_var="t"), 300),
    # latent 128 as in config 3, classification + regression heads, one modality, batch = 512
    "svae_heads": (Spec(model="supervised_vae", input_dims=[400], latent_dim=128, hidden_dim_factor=0.2,
                        supervisor_hidden_dim=32, variables=["y", "c"], variable_types=VT, num_classes={"c": 4}), 512),
    # CrossModalPred: encode layers 0 and 2, reconstruct layers 1 and 2 (one layer only decoded, one in both sets)
    "crossmodal": (Spec(model="CrossModalPred", input_dims=[300, 220, 150], latent_dim=32, hidden_dim_factor=0.15,
                        supervisor_hidden_dim=16, variables=["y", "e"], variable_types=VT, surv_event_var="e",
                        surv_time_var="t", in_idx=[0, 2], out_idx=[1, 2]), 260),
    # defaults: every layer encoded and reconstructed (then it coincides with supervised_vae up to the hidden clamp)
    "crossmodal_all": (Spec(model="CrossModalPred", input_dims=[200, 90], latent_dim=24, hidden_dim_factor=0.2,
                            supervisor_hidden_dim=8, variables=["c"], variable_types=VT, num_classes={"c": 3}), 200),
}


@pytest.mark.parametrize("name", list(CASES))
def test_engine_matches_oracle(name):
    """Forward/backward parity at three successive oracle states, plus an exact test of the fused clip+Adam kernel:
    with the oracle's gradients and Adam moments loaded, one optimizer_step must reproduce the oracle's next state."""
    spec, B = CASES[name]
    lr = 1e-3
    P0, batch, steps, _ = oracle_reference(spec, B, lr, steps=3)
    model = build_model(spec, batch, lr, P0)
    model.train()
    cb = to_cuda(batch)
    rep = Report()
    for s, st in enumerate(steps):
        sync_state(model, st["P_before"])
        eng = compare_step(rep, model, spec, batch, cb, st, s, st["P_before"], lr)
        a = eng.arena
        # ---- clip + Adam with identical inputs ----
        a.grad.zero_(); a.exp_avg.zero_(); a.exp_avg_sq.zero_()
        for k, g in st["grads"].items():
            if g is not None:
                a.view(k, a.grad).copy_(g)
            ad = st["adam_before"].get(k)
            if ad:
                a.view(k, a.exp_avg).copy_(ad["exp_avg"])
                a.view(k, a.exp_avg_sq).copy_(ad["exp_avg_sq"])
        a.step.fill_(s)
        eng.optimizer_step(lr, 1.0)
        rep.close(f"step{s} grad_norm", a.grad_norm, st["grad_norm"], rtol=1e-5)
        sd = model.state_dict()
        for k, g in st["grads"].items():
            if g is None:
                rep.close(f"step{s} untouched {k}", sd[k], st["P_before"][k], rtol=0, atol=0)
            else:
                rep.close(f"step{s} adam {k}", sd[k], st["P_after"][k], rtol=2e-6, atol=2e-7)
    rep.finish()


GNN_CASES = {
    # BASELINE.json config 4 architecture (1 feature per node, GCN 1 -> 32 -> 32, fc N*32 -> 128) at reduced size
    "cfg4_small": (Spec(model="GNN", input_dims=[1], latent_dim=128, supervisor_hidden_dim=32, variables=["y"],
                        variable_types=VT, node_count=300, node_embedding_dim=32, num_convs=2, activation="relu"), 96, 3000),
    # ragged: 3 features per node, embedding 12 (not a multiple of 8), 3 convolutions, two heads, nodes without in-edges
    "ragged": (Spec(model="GNN", input_dims=[3], latent_dim=20, supervisor_hidden_dim=8, variables=["y", "c"],
                    variable_types=VT, num_classes={"c": 3}, node_count=77, node_embedding_dim=12, num_convs=3,
                    activation="relu"), 50, 120),
    # GraphConv (the reference CLI's default convolution) and SAGEConv: neighbour sum / mean + root weight; graphs with
    # isolated nodes, duplicate edges and self loops (synthetic_graph draws with replacement)
    "graphconv": (Spec(model="GNN", input_dims=[1], latent_dim=32, supervisor_hidden_dim=16, variables=["y"],
                       variable_types=VT, node_count=300, node_embedding_dim=32, num_convs=2, activation="relu",
                       conv="GC"), 64, 2500),
    "graphconv_ragged": (Spec(model="GNN", input_dims=[3], latent_dim=20, supervisor_hidden_dim=8, variables=["y", "c"],
                              variable_types=VT, num_classes={"c": 3}, node_count=77, node_embedding_dim=12, num_convs=3,
                              activation="tanh", conv="GC"), 50, 120),
    "sage": (Spec(model="GNN", input_dims=[2], latent_dim=24, supervisor_hidden_dim=8, variables=["y"],
                  variable_types=VT, node_count=150, node_embedding_dim=16, num_convs=2, activation="relu",
                  conv="SAGE"), 40, 400),
}


@pytest.mark.parametrize("name", list(GNN_CASES))
def test_gnn_matches_oracle(name):
    from oracle.restatement import synthetic_graph
    spec, B, n_edges = GNN_CASES[name]
    lr = 1e-3
    edge_index = synthetic_graph(spec.node_count, n_edges, 0)
    P0, batch, steps = gnn_oracle_steps(spec, B, lr, 2, edge_index)
    model = build_model(spec, batch, lr, P0, edge_index)
    model.train()
    cb = to_cuda(batch)
    rep = Report()
    for s, st in enumerate(steps):
        sync_state(model, st["P_before"])
        st["flagged"] = {}
        compare_step(rep, model, spec, batch, cb, st, s, st["P_before"], lr)
    rep.finish()


def test_gnn_fp32_transform_is_strict(monkeypatch):
    """The config-4 scenario with the per-node transform on the fp32 CUDA-core kernels (FXN_GCN_FUSED=1, the path layers
    with emb % 16 != 0 take) and the noise exactly as drawn -- no gate-safe masks: every gradient within RTOL."""
    from oracle.restatement import synthetic_graph
    monkeypatch.setenv("FXN_GCN_FUSED", "1")
    spec, B, n_edges = GNN_CASES["cfg4_small"]
    edge_index = synthetic_graph(spec.node_count, n_edges, 0)
    P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=2, edge_index=edge_index)
    model = build_model(spec, batch, 1e-3, P0, edge_index)
    model.train()
    assert not model.engine().gemm_layers
    cb = to_cuda(batch)
    rep = Report()
    for s, st in enumerate(steps):
        sync_state(model, st["P_before"])
        st["flagged"] = {}
        compare_step(rep, model, spec, batch, cb, st, s, st["P_before"], 1e-3)
    rep.finish()


def test_free_running_trajectory_tracks_oracle():
    """Engine and oracle run 6 steps independently (own optimizer state, same masks). Chaotic sensitivity to ReLU gate
    flips and to Adam's sign-like first updates rules out element-wise equality, so the check is on the loss curve and
    on the norm of the parameter displacement."""
    spec, B = CASES["multitask"]
    lr = 1e-3
    P0, batch, steps, P_final = oracle_reference(spec, B, lr, steps=6)
    model = build_model(spec, batch, lr, P0)
    model.train()
    cb = to_cuda(batch)
    eng = model.engine()
    for s, st in enumerate(steps):
        ws = model.fit_step(cb, masks=masks_from_noise(st["noise"]))
        tot = float(eng.losses(ws)["__total__"])
        ref = float(st["losses"]["train_loss"])
        assert abs(tot - ref) <= 5e-3 * abs(ref), f"step {s}: loss {tot} vs oracle {ref}"
    sd = model.state_dict()
    num = den = 0.0
    for k, v in P_final.items():
        if not v.dtype.is_floating_point or k.endswith(("layer_1.bias", "running_mean", "running_var")):
            continue
        if k.endswith("layer_out.bias") or k == "fusion_block.bias":
            continue        # analytically-zero-gradient biases random-walk in both implementations
        num += float((sd[k].cpu().double() - v.double()).pow(2).sum())
        den += float((v.double() - P0[k].double()).pow(2).sum())
    assert num <= (0.1 ** 2) * den, f"parameter displacement differs: {num ** 0.5:.3e} vs {den ** 0.5:.3e}"


def test_all_missing_labels_give_zero_loss_and_no_nan():
    spec, B = CASES["multitask"]
    dat, y = synthetic_batch(spec, 200, 3)
    y = {k: torch.full_like(v, float("nan")) for k, v in y.items()}
    model = build_model(spec, (dat, y, None), 1e-3)
    model.train()
    loss = model.training_step(to_cuda((dat, y, None)), 0, log=False)
    eng = model.engine()
    vals = eng.losses(eng.ws[200])
    for k in spec.variables:
        assert float(vals[k]) == 0.0
    assert torch.isfinite(loss)
    assert torch.isfinite(eng.arena.grad).all()


def test_training_step_autograd_contract():
    """loss.backward() must populate .grad of every parameter with the engine's gradients (Lightning contract)."""
    spec, B = CASES["cfg2_small"]
    P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=1)
    model = build_model(spec, batch, 1e-3, P0)
    model.train()
    masks = masks_from_noise(steps[0]["noise"])
    loss = model.training_step(to_cuda(batch), 0, log=False, masks=masks)
    loss.backward()
    gmax = max(float(g.abs().max()) for g in steps[0]["grads"].values() if g is not None)
    for k, p in model.named_parameters():
        g = steps[0]["grads"][k]
        if g is None:
            continue
        assert p.grad is not None, k
        _close(f"autograd grad[{k}]", p.grad, g, atol=1e-4 * gmax)
    opt = model.configure_optimizers()
    torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()   # torch optimizer on arena-backed parameters must work too


def test_feature_importance_matches_oracle_attribution():
    """compute_feature_importance: the engine's eval-mode input-gradient pass through head,
    fusion and encoders, integrated along captum's Gauss-Legendre path, against torch autograd on the oracle."""
    from oracle.restatement import attribution_path, feature_importance_sums, init_params
    spec = Spec(model="DirectPred", input_dims=[300, 170], latent_dim=40, hidden_dim_factor=0.2, supervisor_hidden_dim=16,
                variables=["y", "c"], variable_types=VT, num_classes={"c": 3})
    torch.manual_seed(0)
    P = init_params(spec)
    for k in P:                                   # non-trivial running statistics and affine parameters
        if k.endswith("running_mean"):
            P[k] = torch.randn_like(P[k]) * 0.3
        elif k.endswith("running_var"):
            P[k] = torch.rand_like(P[k]) + 0.5
        elif k.endswith("batchnorm.weight"):
            P[k] = torch.rand_like(P[k]) + 0.5
        elif k.endswith("batchnorm.bias"):
            P[k] = torch.randn_like(P[k]) * 0.2
    N = 700
    dat, y = synthetic_batch(spec, N, 0)
    model = build_model(spec, (dat, y, None), 1e-3, P)
    ds = _DS(dat, y, spec.variable_types)
    rep = Report()
    for var, C in (("c", 3), ("y", 1)):
        df = model.compute_feature_importance(ds, var, steps_or_samples=5, batch_size=256)
        alphas, weights = attribution_path("IntegratedGradients", 5)
        want = [[torch.zeros(d) for d in spec.input_dims] for _ in range(C)]
        for s in range(0, N, 256):
            part = feature_importance_sums(P, spec, {k: v[s:s + 256] for k, v in dat.items()}, var, alphas, weights)
            for c in range(C):
                for j in range(2):
                    want[c][j] += part[c][j]
        assert list(df.columns) == ["target_variable", "target_class", "target_class_label", "layer", "name", "importance"]
        assert len(df) == C * sum(spec.input_dims)
        for c in range(C):
            for j, layer in enumerate(dat.keys()):
                got = torch.tensor(df[(df.target_class == c) & (df.layer == layer)]["importance"].to_numpy())
                rep.close(f"importance[{var}][class {c}][{layer}]", got, want[c][j] / N)
        assert var in model.feature_importances
    rep.finish()


def test_device_triplet_batcher_follows_the_reference_sampling_law():
    """DeviceTripletBatcher against TripletMultiOmicDataset's rules (data.py:1088-1151):
    anchors = samples with a label, each once per epoch; positive = another sample of the anchor's class; negative from
    a different class group (NaN labels form the group "NA"); gathered rows are the dataset's rows, bit for bit."""
    from flexynesis_b200 import DeviceTripletBatcher
    g = torch.Generator().manual_seed(0)
    N = 500
    lab = torch.randint(0, 4, (N,), generator=g).float()
    lab[torch.rand(N, generator=g) < 0.1] = float("nan")
    lab[7] = 9.0                                               # a class with a single member
    dat = {"a": torch.randn(N, 37, generator=g), "b": torch.randn(N, 12, generator=g)}
    ds = _DS(dat, {"c": lab}, {"c": "categorical"})
    ds.ann = {"c": lab}
    bt = DeviceTripletBatcher(ds, "c", 64, "cuda", seed=1)
    seen = []
    labc = lab.cuda()
    for anchor, pos, neg, y in bt:
        a_idx = (anchor["a"][:, None, :] == dat["a"].cuda()[None, :, :]).all(-1).float().argmax(1)
        p_idx = (pos["a"][:, None, :] == dat["a"].cuda()[None, :, :]).all(-1).float().argmax(1)
        n_idx = (neg["a"][:, None, :] == dat["a"].cuda()[None, :, :]).all(-1).float().argmax(1)
        assert torch.equal(anchor["b"], dat["b"].cuda()[a_idx]) and torch.equal(pos["b"], dat["b"].cuda()[p_idx])
        assert torch.equal(y["c"], labc[a_idx])
        assert not torch.isnan(labc[a_idx]).any()
        assert torch.equal(labc[p_idx], labc[a_idx])
        single = labc[a_idx] == 9.0
        assert bool(((p_idx != a_idx) | single).all())
        ln = labc[n_idx]
        assert bool((torch.isnan(ln) | (ln != labc[a_idx])).all())
        seen.append(a_idx.cpu())
    seen = torch.cat(seen)
    assert seen.unique().numel() == seen.numel() == len(bt) * 64
    # negatives reach every other group, including NA
    _, q = bt.sample_indices(torch.full((4000,), int(torch.nonzero(lab == 0)[0]), device="cuda"))
    lq = labc[q]
    assert bool(torch.isnan(lq).any()) and set(lq[~torch.isnan(lq)].unique().tolist()) == {1.0, 2.0, 3.0, 9.0}


def test_fit_loop_trains_directpred_and_triplet_with_validation():
    """flexynesis_b200.fit.fit (the Lightning-policy loop, main.py:212-225 / :289-318): mini-batches from the device
    batchers (plain and triplet), CUDA-graphed or eager steps, per-epoch validation; the training loss must go down."""
    import numpy as np
    import flexynesis_b200 as fx
    vt = {"y": "numerical", "c": "categorical"}
    tr = fx.SyntheticMultiOmicDataset([120, 60], 512, vt, {"c": 3}, seed=0)
    va = fx.SyntheticMultiOmicDataset([120, 60], 128, vt, {"c": 3}, seed=1)
    # make the labels learnable: class = argmax of three input features, y = a linear read-out
    for ds in (tr, va):
        x = ds.dat["layer0"]
        ds.ann["c"] = x[:, :3].argmax(1).float()
        ds.ann["y"] = x[:, 3] - 0.5 * x[:, 4]
    cfg = {"latent_dim": 16, "hidden_dim_factor": 0.25, "supervisor_hidden_dim": 8, "lr": 5e-3}

    class V:
        pass

    def view(ds):
        v = V()
        v.dat, v.features, v.variable_types, v.ann, v.samples = ds.dat, ds.features, ds.variable_types, ds.clean_ann(), ds.samples
        return v
    torch.manual_seed(0)
    m = fx.DirectPred(cfg, view(tr), ["c", "y"], device_type="gpu")
    hist = fx.fit.fit(m, tr, batch_size=128, epochs=6, val_dataset=va)
    assert len(hist) == 6 and all(np.isfinite(h["train_loss"]) and np.isfinite(h["val_loss"]) for h in hist)
    assert min(h["train_loss"] for h in hist[-2:]) < hist[0]["train_loss"]
    hist_full = fx.fit.fit(fx.DirectPred(cfg, view(tr), ["c", "y"], device_type="gpu"), tr, batch_size=512, epochs=10)
    assert min(h["train_loss"] for h in hist_full[-3:]) < hist_full[0]["train_loss"]   # full batch: the CUDA-graphed path
    torch.manual_seed(0)
    t = fx.MultiTripletNetwork(cfg, view(tr), ["c", "y"], device_type="gpu")
    hist_t = fx.fit.fit(t, tr, batch_size=128, epochs=5, val_dataset=va)
    assert all(np.isfinite(h["train_loss"]) and np.isfinite(h["val_loss"]) for h in hist_t)
    assert min(h["train_loss"] for h in hist_t[-2:]) < hist_t[0]["train_loss"]
