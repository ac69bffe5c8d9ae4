"""Host-side logic that needs no GPU: the safetensors round trip of the drop-in classes' state_dict (the reference
saves models that way, flexynesis/__main__.py:1563-1569, and loads them strictly, inference.py:378-379), the
trials-per-device scheduler, and the tile-width planner of concurrent encoder GEMMs."""
import os

import pytest
import torch

import flexynesis_b200 as fx
from flexynesis_b200.engine import concurrent_plan, split_groups
from flexynesis_b200.trials import run_trials


class _DS:
    def __init__(self, dims, ann, vt):
        self.dat = {f"l{i}": torch.randn(12, d) for i, d in enumerate(dims)}
        self.ann, self.variable_types = ann, vt
        self.features = {k: [f"f{j}" for j in range(v.shape[1])] for k, v in self.dat.items()}
        self.samples = [f"s{i}" for i in range(12)]

    def __len__(self):
        return len(self.samples)


@pytest.mark.parametrize("cls", ["DirectPred", "supervised_vae", "CrossModalPred"])
def test_state_dict_round_trips_through_safetensors(tmp_path, cls):
    from safetensors.torch import load_file, save_file
    torch.manual_seed(0)
    ann = {"y": torch.randn(12), "c": torch.tensor([0., 1, 2] * 4)}
    ds = _DS([40, 30], ann, {"y": "numerical", "c": "categorical"})
    cfg = {"latent_dim": 8, "hidden_dim_factor": 0.25, "supervisor_hidden_dim": 4, "lr": 1e-3}
    make = lambda: getattr(fx, cls)(cfg, ds, ["y", "c"], device_type="cpu")
    a, b = make(), make()
    path = str(tmp_path / "m.safetensors")
    save_file(a.state_dict(), path)
    b.load_state_dict(load_file(path), strict=True)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)


def _objective(cfg, device):
    import time
    time.sleep(0.5)                                           # long enough for every worker to be up and pulling
    if cfg.get("boom"):
        raise ValueError("boom")
    if cfg.get("die"):
        os._exit(3)                                           # a worker that dies without raising (OOM kill, CUDA fault)
    return (cfg["x"] ** 2, device, os.getpid())


@pytest.mark.timeout(300)
def test_run_trials_spreads_trials_over_devices_and_keeps_order():
    res = run_trials(_objective, [{"x": i} for i in range(7)], devices=["cpu:0", "cpu:1", "cpu:2"], timeout=120)
    assert [r[0] for r in res] == [i * i for i in range(7)]
    assert {r[1] for r in res} <= {"cpu:0", "cpu:1", "cpu:2"}
    assert len({r[2] for r in res}) >= 2                      # really ran in several worker processes
    with pytest.raises(RuntimeError, match="boom"):
        run_trials(_objective, [{"x": 1}, {"x": 2, "boom": True}], devices=["cpu:0"], timeout=120)


@pytest.mark.timeout(300)
def test_run_trials_fails_the_trial_of_a_dead_worker_instead_of_hanging():
    with pytest.raises(RuntimeError, match="died with exit code 3"):
        run_trials(_objective, [{"x": 1}, {"x": 2, "die": True}, {"x": 3}], devices=["cpu:0", "cpu:1"], timeout=120)


def test_concurrent_plan():
    assert concurrent_plan(4096, [512], [5000]) == [(0, 0)]                     # a single branch plans for the whole chip
    plan = concurrent_plan(4096, [512, 307], [5000, 3000])                      # config 2: the branches share the 74 SM pairs
    assert sum(g for _, g in plan) == 74 and all(g >= 2 for _, g in plan)
    for (bn, g), h in zip(plan, [512, 307]):
        assert bn % 32 == 0 and 64 <= bn <= 256
        tn = -(-h // bn)
        assert (tn - 1) * bn < h                                                # no tile column lies entirely outside N
    assert plan[0][1] > plan[1][1]                                              # the bigger contraction gets more of the chip
    assert split_groups([3.0, 1.0]) == [56, 18] and split_groups([1.0]) == [0]
    assert sum(split_groups([5.0, 3.0, 1.0])) == 74


def test_vae_feature_importance_satisfies_completeness():
    """supervised_vae / CrossModalPred attribution (torch path): with a deterministic latent code (FC_log_var zeroed, so
    z = mean) Integrated Gradients must satisfy completeness -- sum over features of the signed attributions equals
    f(x) - f(0) per sample -- which the quadrature reaches at 64 Gauss-Legendre points; and the returned frame has the
    reference's layout."""
    from flexynesis_b200.models import attribution_path
    torch.manual_seed(0)
    ann = {"y": torch.randn(12), "c": torch.tensor([0., 1, 2] * 4)}
    ds = _DS([40, 30], ann, {"y": "numerical", "c": "categorical"})
    cfg = {"latent_dim": 8, "hidden_dim_factor": 0.25, "supervisor_hidden_dim": 4, "lr": 1e-3}
    m = fx.CrossModalPred(cfg, ds, ["y", "c"], input_layers=["l0"], output_layers=["l1"], device_type="cpu")
    with torch.no_grad():
        m.FC_log_var.weight.zero_(); m.FC_log_var.bias.zero_()
    m.eval()
    df = m.compute_feature_importance(ds, "c", steps_or_samples=8, batch_size=5)
    assert list(df.columns) == ["target_variable", "target_class", "target_class_label", "layer", "name", "importance"]
    assert len(df) == 3 * 40 and set(df.layer) == {"l0"} and (df.importance >= 0).all()
    # completeness, straight from the definition
    x = ds.dat["l0"]
    alphas, weights = attribution_path("IntegratedGradients", 64)
    G = torch.zeros_like(x)
    for al, w in zip(alphas, weights):
        xk = (x * float(al)).requires_grad_(True)
        out = m.forward([xk])[4]["y"]
        G += float(w) * torch.autograd.grad(out[:, 0].sum(), [xk])[0]
    with torch.no_grad():
        f1 = m.forward([x])[4]["y"][:, 0]
        f0 = m.forward([torch.zeros_like(x)])[4]["y"][:, 0]
    assert torch.allclose((x * G).sum(1), f1 - f0, atol=2e-3), ((x * G).sum(1) - (f1 - f0)).abs().max()


def test_gnn_feature_importance_layout_and_completeness():
    """GNN attribution (torch path): frame layout of the reference (one row per class, layer and gene) and the
    completeness axiom of Integrated Gradients on the graph encoder."""
    from flexynesis_b200.models import attribution_path
    from oracle.restatement import synthetic_graph
    torch.manual_seed(0)
    B, N, F = 10, 30, 2

    class G:
        pass
    ds = G()
    ds.node_features_tensor = torch.randn(B, N, F)
    ds.edge_index = synthetic_graph(N, 70, 0)
    ds.variable_types = {"y": "numerical"}
    ds.ann = {"y": torch.randn(B)}
    ds.samples = [f"s{i}" for i in range(B)]
    ds.common_features = [f"g{i}" for i in range(N)]
    ds.multiomic_dataset = G()
    ds.multiomic_dataset.dat = {"rna": None, "cnv": None}
    ds.multiomic_dataset.variable_types, ds.multiomic_dataset.ann = ds.variable_types, ds.ann
    G.__len__ = lambda self: B
    G.__getitem__ = lambda self, i: (self.node_features_tensor[i], {"y": self.ann["y"][i]}, self.samples[i])
    cfg = {"latent_dim": 6, "supervisor_hidden_dim": 4, "lr": 1e-3, "node_embedding_dim": 8, "num_convs": 2,
           "activation": "tanh"}
    m = fx.GNN(cfg, ds, ["y"], device_type="cpu", gnn_conv_type="GC")
    m.eval()
    df = m.compute_feature_importance(ds, "y", steps_or_samples=6, batch_size=4)
    assert list(df.columns) == ["target_variable", "target_class", "target_class_label", "layer", "name", "importance"]
    assert len(df) == 2 * N and list(df.layer.unique()) == ["rna", "cnv"] and list(df.name[:3]) == ["g0", "g1", "g2"]
    x = ds.node_features_tensor
    alphas, weights = attribution_path("IntegratedGradients", 48)
    Gr = torch.zeros_like(x)
    for al, w in zip(alphas, weights):
        xk = (x * float(al)).requires_grad_(True)
        Gr += float(w) * torch.autograd.grad(m.forward(xk)["y"][:, 0].sum(), [xk])[0]
    with torch.no_grad():
        f1, f0 = m.forward(x)["y"][:, 0], m.forward(torch.zeros_like(x))["y"][:, 0]
    assert torch.allclose((x * Gr).sum((1, 2)), f1 - f0, atol=2e-3)


def test_gemm_planner_invariants():
    """fxn_gemm_plan (the C library's host-side cost model, callable without a device): for the shapes of the five
    BASELINE configs and a sweep of ragged ones the plan must be launchable -- shared memory within the 227 KB opt-in
    limit, all CTA groups co-resident (74 pairs / 148 CTAs), tile width a legal UMMA N for the operand layout, the tiles
    covering N without an empty last tile, stream-K only for plain fp32 outputs, forced widths honoured."""
    from flexynesis_b200 import _lib as L
    shapes = [(4096, 512, 5000), (4096, 307, 3000), (4096, 256, 512), (4096, 32, 256), (512, 5000, 4096), (307, 3000, 4096),
              (4096, 1024, 24000), (1024, 24000, 4096), (4096, 5000, 512), (4096, 4096, 128), (200, 4096, 128),
              (512, 128, 1000), (128, 64, 128), (4096, 128, 64000), (128, 64000, 4096), (333, 307, 96), (1, 8, 8), (77, 5, 5000)]
    for (M, N, K) in shapes:
        for b_mn in (0, 1):
            for plain in (False, True):
                p = L.gemm_plan(M, N, K, 3, b_mn, plain, 0)
                cg, bn = p["cta_group"], p["block_n"]
                assert cg == (2 if M > 128 else 1)
                assert p["smem_bytes"] <= 227 * 1024, (M, N, K, p)
                assert 1 <= p["groups"] <= 148 // cg, (M, N, K, p)
                assert bn % ((64 if b_mn else 16) * cg) == 0 and 16 <= bn <= 256, (M, N, K, p)
                assert p["tiles_n"] * bn >= N and (p["tiles_n"] - 1) * bn < N, (M, N, K, p)
                assert p["tiles_m"] * 128 * cg >= M and (p["tiles_m"] - 1) * 128 * cg < M
                assert p["stages"] >= 2
                if not plain:
                    assert p["streamk"] == 0
    assert L.gemm_plan(4096, 512, 5000, 3, 0, False, 256)["block_n"] == 256
    assert L.gemm_plan(4096, 307, 3000, 3, 0, False, 160)["block_n"] == 160
    big = L.gemm_plan(512, 5000, 4096, 3, 1, True, 0)                 # config-2 weight gradient: stream-K over every SM
    assert big["streamk"] == 1 and big["groups"] == 74 and big["block_n"] in (128, 256)
    capped = L.gemm_plan(512, 5000, 4096, 3, 1, True, 0, max_groups=40)    # the engine splits the chip between two wgrads
    assert capped["groups"] == 40                                          # (80 tiles on 40 pairs: two whole rounds, no split tile)
    fix = L.gemm_plan(4096, 1024, 24000, 3, 0, False, 0, fix=True)         # config 5 forward: 64 tiles over all 74 pairs
    assert fix["streamk"] == 2 and fix["groups"] == 74
    assert L.gemm_plan(4096, 256, 512, 3, 0, False, 0, fix=True)["streamk"] == 0   # short K: whole tiles


def _golden(name):
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", name + ".pt"), weights_only=False)


@pytest.mark.parametrize("name", ["directpred_single", "directpred_fusion", "directpred_noweight", "triplet", "gnn"])
def test_cpu_resident_inference_matches_reference_eval_outputs(name):
    """Contract safety ("device moves"): with parameters and inputs on the CPU the drop-in classes run
    the torch formulation of their containers; in eval mode it must reproduce what the REFERENCE produced for the same
    state_dict (tests/golden/*.pt: `eval_outputs0`, recorded from the reference's own source)."""
    from oracle.restatement import Spec
    from test_gpu_parity import _DS, _GDS
    g = _golden(name)
    spec = Spec(**g["spec"])
    cfg = {"latent_dim": spec.latent_dim, "hidden_dim_factor": spec.hidden_dim_factor,
           "supervisor_hidden_dim": spec.supervisor_hidden_dim, "lr": g["lr"], "node_embedding_dim": spec.node_embedding_dim,
           "num_convs": spec.num_convs, "activation": spec.activation}
    targets = [v for v in spec.variables if v != spec.surv_event_var]
    kw = dict(surv_event_var=spec.surv_event_var, surv_time_var=spec.surv_time_var,
              use_loss_weighting=spec.use_loss_weighting, device_type="cpu")
    batch = g["batch"]
    if spec.model == "GNN":
        m = fx.GNN(cfg, _GDS(batch[0], batch[1], spec.variable_types, g["edge_index"]), targets, gnn_conv_type="GCN", **kw)
    elif spec.model == "MultiTripletNetwork":
        m = fx.MultiTripletNetwork(cfg, _DS(batch[0], batch[3], spec.variable_types), targets, **kw)
    else:
        m = fx.DirectPred(cfg, _DS(batch[0], batch[1], spec.variable_types), targets, **kw)
    m.load_state_dict(g["P0"], strict=True)
    m.eval()
    with torch.no_grad():
        if spec.model == "GNN":
            out = m.forward(batch[0])
        elif spec.model == "MultiTripletNetwork":
            out = m.forward(batch[0], batch[1], batch[2])[3]
        else:
            out = m.forward(list(batch[0].values()))
    assert g["eval_outputs0"] is not None
    for k, v in g["eval_outputs0"].items():
        assert torch.allclose(out[k], v, rtol=1e-5, atol=1e-6), (name, k, float((out[k] - v).abs().max()))


def test_predict_and_transform_formats_on_cpu():
    """predict -> {var: ndarray} with softmax rows for categorical variables; transform -> DataFrame [samples x E0..]
    indexed by sample name (direct_pred.py:296-415), here through the CPU-resident torch path of the drop-in classes."""
    import numpy as np
    torch.manual_seed(0)
    ann = {"y": torch.randn(12), "c": torch.tensor([0., 1, 2] * 4)}
    ds = _DS([40, 30], ann, {"y": "numerical", "c": "categorical"})
    cfg = {"latent_dim": 8, "hidden_dim_factor": 0.25, "supervisor_hidden_dim": 4, "lr": 1e-3}
    for cls in ("DirectPred", "supervised_vae", "CrossModalPred"):
        m = getattr(fx, cls)(cfg, ds, ["y", "c"], device_type="cpu")
        pred = m.predict(ds)
        assert set(pred) == {"y", "c"} and pred["y"].shape == (12, 1) and pred["c"].shape == (12, 3)
        assert np.allclose(pred["c"].sum(1), 1.0, atol=1e-5) and (pred["c"] >= 0).all()
        emb = m.transform(ds)
        assert list(emb.index) == ds.samples and list(emb.columns) == [f"E{i}" for i in range(8)]
        assert np.isfinite(emb.to_numpy()).all()
    dec = fx.CrossModalPred(cfg, ds, ["y"], input_layers=["l0"], output_layers=["l1"], device_type="cpu").decode(ds)
    assert set(dec) == {"l1"} and dec["l1"].shape == (30, 12) and list(dec["l1"].columns) == ds.samples


def test_parse_cpulist_and_numa_helpers():
    from flexynesis_b200.parallel import gpu_numa_node, parse_cpulist, pin_to_gpu_numa_node
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    assert parse_cpulist("5") == [5]
    # without a GPU the topology is unknown and nothing is changed
    import os
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert gpu_numa_node(0) is None and pin_to_gpu_numa_node(0) is None
    assert os.sched_getaffinity(0) == before or torch.cuda.is_available()


def test_load_matrix_npy_roundtrip(tmp_path):
    """on-disk matrix -> tensor (CPU route of the pinned-staging loader): values, dtype conversion, ragged last chunk,
    and the MultiOmicDataset duck type built from files"""
    import numpy as np
    from flexynesis_b200.data import dataset_from_npy, load_matrix_npy
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 11)).astype(np.float64)
    b = rng.integers(0, 50, (37, 5)).astype(np.int32)
    np.save(tmp_path / "a.npy", a)
    np.save(tmp_path / "b.npy", b)
    ta = load_matrix_npy(str(tmp_path / "a.npy"), "cpu", rows_per_chunk=8)
    assert ta.dtype == torch.float32 and ta.shape == (37, 11)
    assert torch.equal(ta, torch.from_numpy(a).float())
    ds = dataset_from_npy({"rna": str(tmp_path / "a.npy"), "cnv": str(tmp_path / "b.npy")},
                          {"y": torch.arange(37.0)}, {"y": "numerical"}, device="cpu")
    assert len(ds) == 37 and list(ds.dat) == ["rna", "cnv"] and len(ds.features["cnv"]) == 5
    x, y, name = ds[3]
    assert torch.equal(x["cnv"], torch.from_numpy(b[3]).float()) and float(y["y"]) == 3.0 and name == "s3"
    np.save(tmp_path / "c.npy", rng.standard_normal((5,)))
    with pytest.raises(ValueError):
        load_matrix_npy(str(tmp_path / "c.npy"), "cpu")
    with pytest.raises(ValueError):
        dataset_from_npy({"rna": str(tmp_path / "a.npy")}, {"y": torch.zeros(3)}, {"y": "numerical"}, device="cpu")


def test_reduce_scatter_context_claims_and_ranges():
    """host bookkeeping of the GEMM-fused reduce-scatter (flexynesis_b200._lib.ReduceScatterContext): which outputs qualify
    (contiguous, 16-byte aligned, 4-aligned length, inside the arena) and how the recorded ranges merge"""
    from flexynesis_b200._lib import ReduceScatterContext
    base, numel = 1 << 20, 10000
    rs = ReduceScatterContext(world=4, rank=1, per=2500 - 2500 % 4, base_ptr=base, numel=numel, inbox_ptrs=[11, 22, 33, 44])
    assert rs.claim(base + 4 * 128, 10, 20, 20) == 128                    # contiguous [10 x 20] at element 128
    assert rs.claim(base + 4 * 128, 10, 20, 24) is None                   # ldc != N: rows are not contiguous
    assert rs.claim(base + 4 * 129, 10, 20, 20) is None                   # not 16-byte aligned
    assert rs.claim(base + 4 * 128, 3, 5, 5) is None                      # 15 elements: not a multiple of 4
    assert rs.claim(base + 4 * 9990, 10, 20, 20) is None                  # runs past the end of the arena
    assert rs.claim(base - 16, 2, 4, 4) is None                           # in front of the arena
    rs.enabled = False
    assert rs.claim(base + 4 * 128, 10, 20, 20) is None
    rs.ranges.update({(0, 400), (400, 800), (1000, 1200), (1100, 1600)})
    assert rs.merged_ranges() == [[0, 800], [1000, 1600]]
    assert [int(p) for p in rs.inbox] == [11, 22, 33, 44]
