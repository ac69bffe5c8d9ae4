"""Constructors of the drop-in classes fed the REFERENCE's own dataset objects: the unmodified
`MultiOmicDataset`, `TripletMultiOmicDataset` and `MultiOmicDatasetNW` classes of flexynesis/data.py (loaded through
oracle/ref_shim.py) and the `SimpleNamespace` that flexynesis/inference.py:116-122 builds when a saved model is re-created
for prediction. CPU only, and only where the reference tree is mounted (the build container): skipped elsewhere."""
from types import SimpleNamespace

import numpy as np
import pandas as pd
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted")

CFG = {"latent_dim": 16, "hidden_dim_factor": 0.25, "supervisor_hidden_dim": 8, "lr": 1e-3, "node_embedding_dim": 8,
       "num_convs": 2, "activation": "relu"}
VT = {"y": "numerical", "c": "categorical"}


def _reference_dataset(n=40, seed=0):
    ref = ref_shim.load()
    g = torch.Generator().manual_seed(seed)
    genes = [f"G{i}" for i in range(30)]
    dat = {"rna": torch.randn(n, 30, generator=g), "cnv": torch.randn(n, 20, generator=g)}
    features = {"rna": pd.Index(genes), "cnv": pd.Index(genes[:20])}
    c = torch.randint(0, 3, (n,), generator=g).float()
    c[:3] = torch.arange(3).float()
    ann = {"y": torch.randn(n, generator=g), "c": c}
    samples = [f"S{i}" for i in range(n)]
    ds = ref.data.MultiOmicDataset(dat, ann, dict(VT), features, samples, {"c": {0: "a", 1: "b", 2: "c"}})
    return ref, ds, genes


def _same_keys(model, ref_model):
    want = {k: tuple(v.shape) for k, v in ref_model.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == want
    model.load_state_dict(ref_model.state_dict(), strict=True)        # a reference checkpoint loads strictly


def test_directpred_and_svae_take_the_reference_multiomic_dataset():
    import flexynesis_b200 as fx
    ref, ds, _ = _reference_dataset()
    kw = dict(config=dict(CFG), dataset=ds, target_variables=["c", "y"], device_type="cpu")
    torch.manual_seed(0)
    _same_keys(fx.DirectPred(**kw), ref.direct_pred.DirectPred(**kw))
    _same_keys(fx.supervised_vae(**kw), ref.supervised_vae.supervised_vae(**kw))
    m = fx.DirectPred(**kw)
    out = m.predict(ds)                                # CPU-resident inference on the reference's dataset object
    assert out["c"].shape == (len(ds), 3) and out["y"].shape == (len(ds), 1)
    emb = m.transform(ds)
    assert list(emb.index) == list(ds.samples) and emb.shape == (len(ds), CFG["latent_dim"])


def test_triplet_network_takes_the_reference_triplet_dataset():
    import flexynesis_b200 as fx
    ref, ds, _ = _reference_dataset()
    tds = ref.data.TripletMultiOmicDataset(ds, "c")
    kw = dict(config=dict(CFG), target_variables=["c", "y"], device_type="cpu")
    # the reference trainer hands the constructor the TripletMultiOmicDataset's inner dataset (main.py:241-245)
    _same_keys(fx.MultiTripletNetwork(dataset=tds.dataset, **kw), ref.triplet_encoder.MultiTripletNetwork(dataset=tds.dataset, **kw))
    anchor, pos, neg, y = tds[0]
    assert set(anchor) == set(ds.dat) and set(y) == set(ds.ann)


def test_gnn_takes_the_reference_network_dataset():
    import flexynesis_b200 as fx
    ref, ds, genes = _reference_dataset()
    rng = np.random.default_rng(0)
    pairs = {tuple(sorted(rng.choice(len(genes), 2, replace=False))) for _ in range(80)}
    inter = pd.DataFrame([(genes[a], genes[b], 900) for a, b in pairs], columns=["protein1", "protein2", "combined_score"])
    nw = ref.data.MultiOmicDatasetNW(ds, inter)
    kw = dict(config=dict(CFG), dataset=nw, target_variables=["y"], device_type="cpu", gnn_conv_type="GCN")
    m = fx.GNN(**kw)
    _same_keys(m, ref.gnn_early.GNN(**kw))
    x, y, name = nw[0]
    assert x.shape == (m.encoders[0].fc.in_features // CFG["node_embedding_dim"], 2)
    # fit() dispatches to the node-feature batcher on exactly these fields
    assert hasattr(nw, "node_features_tensor") and hasattr(nw, "ann") and not hasattr(nw, "dat")


def test_constructor_accepts_the_inference_namespace():
    """flexynesis/inference.py:116-122 rebuilds a model from saved artifacts with a SimpleNamespace carrying only
    layers / features / dat (keys) / variable_types / ann (class lists)."""
    import flexynesis_b200 as fx
    ns = SimpleNamespace(layers=["rna", "cnv"], features={"rna": [f"G{i}" for i in range(30)], "cnv": [f"G{i}" for i in range(20)]},
                         dat={"rna": None, "cnv": None}, variable_types={"c": "categorical", "y": "numerical"},
                         ann={"c": ["a", "b", "c"], "y": np.array([0.0])})
    ref = ref_shim.load()
    kw = dict(config=dict(CFG), dataset=ns, target_variables=["c", "y"], device_type="cpu")
    _same_keys(fx.DirectPred(**kw), ref.direct_pred.DirectPred(**kw))
    _same_keys(fx.supervised_vae(**kw), ref.supervised_vae.supervised_vae(**kw))
