"""Generate tests/golden/*.pt by running the UNMODIFIED reference source (through oracle/ref_shim.py) on CPU.

Run in the build container only:  python -m oracle.make_golden
Each golden file holds, for one seeded scenario: the model spec, the initial state_dict, the batch, the
recorded noise of every step, and what the reference produced -- head outputs, fused embedding, every loss
term, the total loss, all parameter gradients, the pre-clip gradient norm, and the state_dict after the last
clip + Adam step. The generator also replays every scenario through oracle/restatement.py and refuses to
write a file the restatement does not reproduce (this is what pins the oracle).
"""
from __future__ import annotations

import copy
import os
import sys

import numpy as np
import torch

from . import ref_shim
from .restatement import Noise, Spec, Trainer, forward, synthetic_batch, synthetic_graph

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STEPS = 2
LR = 1e-3


def scenarios():
    vt = {"y": "numerical", "c": "categorical", "e": "numerical", "t": "numerical"}
    return {
        "directpred_single": dict(B=48, spec=Spec(
            model="DirectPred", input_dims=[40], latent_dim=16, hidden_dim_factor=0.3, supervisor_hidden_dim=8,
            variables=["y"], variable_types=vt)),
        "directpred_fusion": dict(B=64, spec=Spec(
            model="DirectPred", input_dims=[40, 24], latent_dim=16, hidden_dim_factor=0.4, supervisor_hidden_dim=8,
            variables=["c", "y", "e"], variable_types=vt, num_classes={"c": 5}, surv_event_var="e",
            surv_time_var="t")),
        "directpred_noweight": dict(B=64, spec=Spec(
            model="DirectPred", input_dims=[40, 24], latent_dim=16, hidden_dim_factor=0.4, supervisor_hidden_dim=8,
            variables=["c", "y"], variable_types=vt, num_classes={"c": 3}, use_loss_weighting=False)),
        "supervised_vae": dict(B=48, spec=Spec(
            model="supervised_vae", input_dims=[40, 24], latent_dim=8, hidden_dim_factor=0.4,
            supervisor_hidden_dim=8, variables=["c", "e"], variable_types=vt, num_classes={"c": 4},
            surv_event_var="e", surv_time_var="t")),
        "supervised_vae_nohead": dict(B=48, spec=Spec(
            model="supervised_vae", input_dims=[30], latent_dim=8, hidden_dim_factor=0.4,
            supervisor_hidden_dim=8, variables=[], variable_types=vt)),
        "triplet": dict(B=48, spec=Spec(
            model="MultiTripletNetwork", input_dims=[40, 24], latent_dim=16, hidden_dim_factor=0.4,
            supervisor_hidden_dim=8, variables=["c", "y"], variable_types=vt, num_classes={"c": 4})),
        "crossmodal": dict(B=48, spec=Spec(
            model="CrossModalPred", input_dims=[40, 24, 18], latent_dim=8, hidden_dim_factor=0.4,
            supervisor_hidden_dim=8, variables=["c", "y"], variable_types=vt, num_classes={"c": 3},
            in_idx=[0, 2], out_idx=[1, 2])),
        "gnn": dict(B=16, spec=Spec(
            model="GNN", input_dims=[2], latent_dim=12, supervisor_hidden_dim=8, variables=["y", "c"],
            variable_types=vt, num_classes={"c": 3}, node_count=30, node_embedding_dim=6, num_convs=2,
            activation="relu")),
    }


def build_reference(ref, spec: Spec, dat, y, edge_index=None):
    cfg = {"latent_dim": spec.latent_dim, "hidden_dim_factor": spec.hidden_dim_factor,
           "supervisor_hidden_dim": spec.supervisor_hidden_dim, "lr": LR, "epochs": 1, "batch_size": 32,
           "node_embedding_dim": spec.node_embedding_dim, "num_convs": spec.num_convs,
           "activation": spec.activation}
    targets = [v for v in spec.variables if v != spec.surv_event_var]
    kw = dict(config=cfg, target_variables=targets, surv_event_var=spec.surv_event_var,
              surv_time_var=spec.surv_time_var, use_loss_weighting=spec.use_loss_weighting, device_type="cpu")
    ann = {k: v.clone() for k, v in y.items()}
    # np.unique counts NaN as a class; the reference's DataImporter leaves NaN in ann too, so mimic by
    # giving the constructor an annotation without missing values
    ann_ctor = {k: torch.nan_to_num(v, nan=0.0) for k, v in ann.items()}
    if spec.model == "GNN":
        ds = ref_shim.RefGraphDataset(dat, ann_ctor, spec.variable_types, edge_index)
        return ref.gnn_early.GNN(dataset=ds, gnn_conv_type="GCN", **kw)
    ds = ref_shim.RefDataset(dat, ann_ctor, spec.variable_types)
    if spec.model == "CrossModalPred":
        keys = list(dat.keys())
        return ref.crossmodal_pred.CrossModalPred(
            dataset=ds, input_layers=[keys[i] for i in spec.in_idx] if spec.in_idx is not None else None,
            output_layers=[keys[i] for i in spec.out_idx] if spec.out_idx is not None else None, **kw)
    cls = {"DirectPred": lambda: ref.direct_pred.DirectPred,
           "supervised_vae": lambda: ref.supervised_vae.supervised_vae,
           "MultiTripletNetwork": lambda: ref.triplet_encoder.MultiTripletNetwork}[spec.model]()
    return cls(dataset=ds, **kw)


def make_batch(spec: Spec, B: int, seed: int):
    dat, y = synthetic_batch(spec, B, seed)
    if spec.model == "GNN":
        g = torch.Generator().manual_seed(seed + 1)
        x = torch.randn(B, spec.node_count, spec.input_dims[0], generator=g)
        return (x, y, None), synthetic_graph(spec.node_count, 3 * spec.node_count, seed)
    if spec.model == "MultiTripletNetwork":
        g = torch.Generator().manual_seed(seed + 1)
        perm_p, perm_n = torch.randperm(B, generator=g), torch.randperm(B, generator=g)
        pos = {k: v[perm_p] for k, v in dat.items()}
        neg = {k: v[perm_n] for k, v in dat.items()}
        return (dat, pos, neg, y), None
    return (dat, y, None), None


def eval_forward(model, spec, batch):
    """Eval-mode head outputs of the reference model (None for supervised_vae: its forward samples epsilon even
    in eval mode, supervised_vae.py:419-421)."""
    if spec.model in ("supervised_vae", "CrossModalPred"):
        return None
    model.eval()
    with torch.no_grad():
        if spec.model == "GNN":
            ev = model.forward(batch[0], model.edge_index)
        elif spec.model == "MultiTripletNetwork":
            ev = model.forward(batch[0], batch[1], batch[2])[3]
        else:
            ev = model.forward(list(batch[0].values()))
    return {k: v.clone() for k, v in ev.items()}


def validation_record(model, spec, batch):
    """The reference's validation_step on the batch in eval mode (direct_pred.py:262-294 and the other families' own
    versions): returned loss (= the UNWEIGHTED sum of the loss terms, :290), the logged per-variable losses, and the
    Gaussian draws it made (supervised_vae / CrossModalPred sample epsilon and the MMD prior in eval mode too)."""
    model.eval()
    logged = {}
    orig_log = getattr(model, "log_dict", None)
    model.log_dict = lambda d, **k: logged.update({kk: vv.detach().clone() for kk, vv in d.items()})
    with ref_shim.NoiseRecorder(model, triplet=(spec.model == "MultiTripletNetwork")) as rec:
        torch.manual_seed(77)
        with torch.no_grad():
            loss = model.validation_step(batch, 0)
    if orig_log is not None:
        model.log_dict = orig_log
    return dict(total=loss.detach().clone(), losses=logged, noise={k: v.detach().clone() for k, v in rec.record.items()})


def run_reference(name: str, sc) -> dict:
    ref = ref_shim.load()
    spec: Spec = sc["spec"]
    batch, edge_index = make_batch(spec, sc["B"], seed=0)
    torch.manual_seed(0)
    model = build_reference(ref, spec, batch[0], batch[3] if spec.model == "MultiTripletNetwork" else batch[1],
                            edge_index)
    P0 = copy.deepcopy(model.state_dict())
    ev0 = eval_forward(model, spec, batch)
    val0 = validation_record(model, spec, batch)
    model.train()
    opt = model.configure_optimizers()
    steps = []
    for s in range(STEPS):
        P_before = copy.deepcopy(model.state_dict())
        with ref_shim.NoiseRecorder(model, triplet=(spec.model == "MultiTripletNetwork")) as rec:
            torch.manual_seed(100 + s)
            opt.zero_grad(set_to_none=True)
            captured = {}
            if spec.model == "MultiTripletNetwork":
                orig = model.forward
                def fwd(*a, _o=orig, **k):
                    r = _o(*a, **k); captured["emb"], captured["outputs"] = r[0], r[3]; return r
                model.forward = fwd
            elif spec.model in ("supervised_vae", "CrossModalPred"):
                orig = model.forward
                def fwd(*a, _o=orig, **k):
                    r = _o(*a, **k); captured["emb"], captured["outputs"] = r[1], r[4]; return r
                model.forward = fwd
            else:
                orig = model.forward
                def fwd(*a, _o=orig, **k):
                    r = _o(*a, **k); captured["outputs"] = r; return r
                model.forward = fwd
            logged = {}
            model.log_dict = lambda d, **k: logged.update({kk: vv.detach().clone() for kk, vv in d.items()})
            loss = model.training_step(batch, 0)
            model.forward = orig
            loss.backward()
            grads = {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in model.named_parameters()}
            gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            opt.step()
        steps.append(dict(P_before=P_before, noise={k: v.detach().clone() for k, v in rec.record.items()},
                          outputs={k: v.detach().clone() for k, v in captured["outputs"].items()},
                          embedding=None if "emb" not in captured else captured["emb"].detach().clone(),
                          losses=logged, total=loss.detach().clone(), grads=grads, grad_norm=gnorm.detach().clone()))
    ev = eval_forward(model, spec, batch)
    return dict(name=name, spec=spec.__dict__, lr=LR, batch=batch, edge_index=edge_index, P0=P0, steps=steps,
                P_final=copy.deepcopy(model.state_dict()),
                eval_outputs0=ev0, eval_outputs=ev, val0=val0)


def significant_elements(step_grads, thresh=1e-3, global_thresh=1e-5):
    """{param: bool mask} of elements whose |grad|, in every step, exceeds both thresh * max|grad| of that
    parameter and global_thresh * the largest gradient element of the whole model (the second test removes
    parameters whose entire gradient is rounding noise)."""
    out = {}
    for grads in step_grads:
        gmax = max(float(gr.abs().max()) for gr in grads.values() if gr is not None)
        for k, gr in grads.items():
            if gr is None:
                continue
            m = (gr.abs() > thresh * gr.abs().max()) & (gr.abs() > global_thresh * gmax)
            out[k] = m if k not in out else (out[k] & m)
    return out


def check_oracle(g: dict, rtol=2e-5, atol=2e-6) -> float:
    """Replay a golden scenario through the restatement; return the worst relative deviation seen."""
    spec = Spec(**g["spec"])
    P = {k: v.clone() for k, v in g["P0"].items()}
    worst = 0.0

    def cmp(tag, a, b, extra_atol=0.0):
        nonlocal worst
        a, b = a.detach().double().flatten(), b.detach().double().flatten()
        scale = max(float(b.abs().max()), 1e-30)
        err = float((a - b).abs().max())
        rel = err / scale
        if extra_atol and err <= extra_atol:
            return
        if not (err <= atol + rtol * scale):
            raise AssertionError(f"{g['name']}: {tag} deviates: abs {err:.3e} rel {rel:.3e}")
        worst = max(worst, rel if err > atol else 0.0)

    if g["eval_outputs0"] is not None:
        res = forward(P, spec, g["batch"], False, Noise({}), g["edge_index"])
        for k, v in g["eval_outputs0"].items():
            cmp(f"initial eval outputs[{k}]", res["outputs"][k], v)
    if g.get("val0") is not None:
        res = forward(P, spec, g["batch"], False, Noise(g["val0"]["noise"]), g["edge_index"])
        cmp("validation_step total (unweighted sum)", res["val_total"], g["val0"]["total"])
        for k, v in g["val0"]["losses"].items():
            if k != "val_loss":
                cmp(f"validation_step loss[{k}]", res["losses"][k], v)
    tr = Trainer(P, spec, g["lr"], edge_index=g["edge_index"])
    for s, st in enumerate(g["steps"]):
        res = tr.step(g["batch"], Noise(st["noise"]))
        for k, v in st["outputs"].items():
            cmp(f"step{s} outputs[{k}]", res["outputs"][k], v)
        if st["embedding"] is not None:
            # biases whose gradient is analytically zero (they cancel in a - p, a - n and in the head BatchNorm)
            # random-walk by +-lr per step in BOTH implementations and shift the embedding by a constant
            # -> compare after removing the per-column constant offset, and bound the offset itself
            off = (res["embedding"].detach() - st["embedding"]).mean(0, keepdim=True) if s > 0 else 0.0
            cmp(f"step{s} embedding", res["embedding"].detach() - off, st["embedding"])
            if s > 0:
                assert float(off.abs().max()) < 10 * g["lr"] * s, "embedding offset beyond bias-noise bound"
        for k, v in st["losses"].items():
            if k == "train_loss":
                cmp(f"step{s} total", res["total"], v)
            else:
                cmp(f"step{s} loss[{k}]", res["losses"][k], v)
        cmp(f"step{s} grad_norm", res["grad_norm"], st["grad_norm"])
        for k, gr in st["grads"].items():
            if gr is None:
                assert res["grads"][k] is None, f"{k} should have no grad"
            elif k.endswith("layer_1.bias"):
                continue   # analytically zero through BatchNorm: rounding noise in both implementations
            else:
                cmp(f"step{s} grad[{k}]", res["grads"][k], gr)
    # Parameters after the Adam steps: Adam normalises each element's gradient by its own magnitude, so an element
    # whose gradient is analytically zero (biases feeding a BatchNorm) or tiny moves by +-lr per step driven by
    # rounding noise -- in the reference as well. Compare only elements with a significant gradient in every step.
    sig = significant_elements([st["grads"] for st in g["steps"]])
    for k, v in g["P_final"].items():
        a, b = P[k].detach().float(), v.float()
        if k in sig:
            if not bool(sig[k].any()):
                continue
            a, b = a[sig[k]], b[sig[k]]
        # running_mean of a BatchNorm absorbs the random walk of the bias in front of it (momentum 0.1 per step)
        noise_floor = g["lr"] * len(g["steps"]) if k.endswith("running_mean") else 0.0
        cmp(f"final {k}", a, b, noise_floor)
    if g["eval_outputs"] is not None:
        # after training the eval outputs inherit the +-lr bias noise through (b - running_mean): loose check only
        res = forward(P, spec, g["batch"], False, Noise({}), g["edge_index"])
        for k, v in g["eval_outputs"].items():
            cmp(f"final eval outputs[{k}]", res["outputs"][k], v, 20 * g["lr"] * len(g["steps"]))
    return worst


def main():
    if not ref_shim.available():
        print("reference tree not available; goldens can only be generated in the build container")
        return 1
    os.makedirs(OUT_DIR, exist_ok=True)
    torch.set_num_threads(1)           # bit-stable sums
    only = set(sys.argv[1:])            # optional: regenerate just the named scenarios
    for name, sc in scenarios().items():
        if only and name not in only:
            continue
        g = run_reference(name, sc)
        worst = check_oracle(g)
        path = os.path.join(OUT_DIR, name + ".pt")
        torch.save(g, path)
        print(f"{name:24s} oracle == reference (worst rel dev {worst:.2e}); wrote {os.path.getsize(path) / 1024:.0f} KiB")
    return 0


if __name__ == "__main__":
    sys.exit(main())
