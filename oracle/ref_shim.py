"""Import shim that executes the reference's OWN hot-path source on CPU.

TEST INFRASTRUCTURE ONLY, and only usable in the build container: /root/reference does not exist on the
GPU box, so nothing in tests marked `gpu`, smoke() or bench.py may import this file. It is used by
oracle/make_golden.py (to generate tests/golden/*.pt) and by tests/test_oracle_vs_reference.py (skipped when
the reference tree is absent) to pin oracle/restatement.py.

`import flexynesis` fails here (skopt, lightning, torch_geometric, captum ... are not installed), so the shim
registers name-only stand-ins in sys.modules -- none of them contains hot-path arithmetic except GCNConv, whose
arithmetic lives in un-vendored torch_geometric and is therefore restated (oracle.restatement.gcn_conv) -- and
then loads the unmodified files flexynesis/modules.py and flexynesis/models/*.py from the reference tree.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch
from torch import nn

REFERENCE_ROOT = os.environ.get("FLEXYNESIS_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "flexynesis", "modules.py"))


class _LightningModule(nn.Module):
    """Stand-in for lightning.LightningModule: no-op logging, a `device` property."""

    def log(self, *a, **k):
        pass

    def log_dict(self, *a, **k):
        pass

    @property
    def device(self):
        try:
            return next(self.parameters()).device
        except StopIteration:
            return torch.device("cpu")


class _GCNConv(nn.Module):
    """torch_geometric.nn.GCNConv stand-in (parameters `lin.weight`, `bias`; glorot / zeros init)."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        nn.init.xavier_uniform_(self.lin.weight)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, edge_index):
        from .restatement import gcn_conv
        return gcn_conv({"c.lin.weight": self.lin.weight, "c.bias": self.bias}, "c", x, edge_index)


class _Unavailable(nn.Module):
    def __init__(self, *a, **k):
        raise NotImplementedError("only GCNConv is restated in the oracle shim")


_loaded = {}


def load():
    """Returns a namespace with the reference's modules: .modules, .direct_pred, .supervised_vae,
    .triplet_encoder, .gnn_early, .crossmodal_pred."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")

    def fake(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    tg = fake("torch_geometric")
    tg.__path__ = []
    fake("torch_geometric.data", Dataset=object, download_url=None, extract_gz=None)
    fake("torch_geometric.nn", GCNConv=_GCNConv, GATConv=_Unavailable, SAGEConv=_Unavailable, GraphConv=_Unavailable)
    fake("lightning", LightningModule=_LightningModule)
    fake("captum")
    fake("captum.attr", GradientShap=object, IntegratedGradients=object)

    def to_device_safe(t, device):
        return t.to(device)

    def create_device_from_string(s):
        return torch.device("cpu")

    pkg = fake("flexynesis")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "flexynesis")]
    fake("flexynesis.utils", to_device_safe=to_device_safe, create_device_from_string=create_device_from_string,
         create_covariate_matrix=None, get_variable_types=None)
    models = fake("flexynesis.models")
    models.__path__ = [os.path.join(REFERENCE_ROOT, "flexynesis", "models")]
    _loaded["modules"] = importlib.import_module("flexynesis.modules")
    _loaded["data"] = importlib.import_module("flexynesis.data")    # dataset containers only (data.py:945-1304)
    for name in ("direct_pred", "supervised_vae", "triplet_encoder", "gnn_early", "crossmodal_pred"):
        _loaded[name] = importlib.import_module("flexynesis.models." + name)
    return types.SimpleNamespace(**_loaded)


class RefDataset:
    """Duck-typed dataset accepted by the reference constructors (dat / features / variable_types / ann)."""

    def __init__(self, dat, ann, variable_types):
        self.dat = dat
        self.ann = ann
        self.variable_types = variable_types
        self.features = {k: [f"f{j}" for j in range(v.shape[1])] for k, v in dat.items()}
        self.samples = [f"s{i}" for i in range(next(iter(dat.values())).shape[0])]

    def __len__(self):
        return len(self.samples)


class RefGraphDataset:
    """Duck type of MultiOmicDatasetNW for the GNN constructor (gnn_early.py:82-86, :114-117)."""

    def __init__(self, x, ann, variable_types, edge_index):
        self.x, self.ann, self.variable_types, self.edge_index = x, ann, variable_types, edge_index

    def __getitem__(self, i):
        return self.x[i], {k: v[i] for k, v in self.ann.items()}, f"s{i}"

    def __len__(self):
        return self.x.shape[0]


class NoiseRecorder:
    """Context manager recording every random draw of a reference training_step under the oracle's site names:
    dropout masks via forward hooks on nn.Dropout, epsilon / MMD prior by wrapping torch.randn(_like)."""

    def __init__(self, model, triplet=False):
        self.model, self.triplet = model, triplet
        self.record = {}
        self._hooks = []
        self._calls = {}

    def __enter__(self):
        tags = ["anchor.", "positive.", "negative."]
        for name, mod in self.model.named_modules():
            if isinstance(mod, nn.Dropout):
                def hook(m, inp, out, name=name):
                    if not m.training:
                        return
                    x = inp[0]
                    mask = ((out != 0) | (x == 0)).to(x.dtype)
                    k = self._calls.get(name, 0)
                    self._calls[name] = k + 1
                    site = name
                    if self.triplet and name.startswith("encoders."):
                        site = tags[k] + name
                    if name.endswith("encoders.0.dropout") and hasattr(self.model.encoders[0], "convs"):
                        site = f"encoders.0.dropout.{k}"          # flexGCN shares one Dropout across convs
                    self.record[site] = mask
                self._hooks.append(mod.register_forward_hook(hook))
        self._randn, self._randn_like = torch.randn, torch.randn_like
        rec = self.record

        def randn(*a, **k):
            t = self._randn(*a, **k)
            i = sum(1 for s in rec if s.startswith("mmd_prior."))
            rec[f"mmd_prior.{i}"] = t
            return t

        def randn_like(x, **k):
            t = self._randn_like(x, **k)
            rec["epsilon"] = t
            return t

        torch.randn, torch.randn_like = randn, randn_like
        return self

    def __exit__(self, *exc):
        for h in self._hooks:
            h.remove()
        torch.randn, torch.randn_like = self._randn, self._randn_like
        return False
