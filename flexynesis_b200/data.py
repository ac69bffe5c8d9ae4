"""Dataset duck type + device-side batch feeder.

The engine consumes the reference's dataset objects unchanged (`MultiOmicDataset`: .dat {layer: [N, d] fp32},
.ann {var: [N]}, .variable_types, .features, .samples -- flexynesis/data.py:945-1003). `SyntheticMultiOmicDataset`
is the same duck type filled with the seeded synthetic matrices of SURVEY.md section 8d (no network, no files).
`DeviceBatcher` replaces DataLoader(shuffle=True, drop_last=True) + per-sample __getitem__ + default_collate
(flexynesis/main.py:289-298, data.py:980-995) by a permutation gather over the HBM-resident matrices.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import _lib as L


class SyntheticMultiOmicDataset:
    def __init__(self, input_dims: List[int], n: int, variable_types: Dict[str, str], num_classes: Dict[str, int] = None,
                 surv_event_var: Optional[str] = None, surv_time_var: Optional[str] = None, seed: int = 0,
                 missing: float = 0.05, layer_names: Optional[List[str]] = None):
        g = torch.Generator().manual_seed(seed)
        names = layer_names or [f"layer{i}" for i in range(len(input_dims))]
        self.dat = {k: torch.randn(n, d, generator=g) for k, d in zip(names, input_dims)}
        self.features = {k: [f"{k}_f{j}" for j in range(d)] for k, d in zip(names, input_dims)}
        self.samples = [f"s{i}" for i in range(n)]
        self.variable_types = dict(variable_types)
        self.label_mappings, self.feature_ann = {}, {}
        self.ann: Dict[str, torch.Tensor] = {}
        num_classes = num_classes or {}
        for var, kind in variable_types.items():
            if var == surv_time_var:
                self.ann[var] = 100.0 * torch.rand(n, generator=g)
                continue
            if var == surv_event_var:
                self.ann[var] = (torch.rand(n, generator=g) > 0.3).float()
                continue
            if kind == "numerical":
                v = torch.randn(n, generator=g)
                drop = torch.rand(n, generator=g) < missing
            else:
                c = num_classes[var]
                v = torch.randint(0, c, (n,), generator=g).float()
                v[:c] = torch.arange(c).float()
                drop = torch.rand(n, generator=g) < missing
                drop[:c] = False
            v[drop] = float("nan")
            self.ann[var] = v

    def __len__(self):
        return len(self.samples)

    def __getitem__(self, i):
        return ({k: v[i] for k, v in self.dat.items()}, {k: v[i] for k, v in self.ann.items()}, self.samples[i])

    def clean_ann(self):
        """Annotation with NaN replaced (what np.unique in the constructors should count classes on)."""
        return {k: torch.nan_to_num(v, nan=0.0) for k, v in self.ann.items()}


class DeviceBatcher:
    """Keeps every modality matrix and label vector resident in HBM and yields shuffled, drop_last batches built by
    one gather kernel per modality (fp32 rows; the engine splits them into operand planes)."""

    def __init__(self, dataset, batch_size: int, device, shuffle: bool = True, drop_last: bool = True, seed: int = 0):
        self.device = torch.device(device)
        self.dat = {k: v.to(self.device, torch.float32).contiguous() for k, v in dataset.dat.items()}
        self.ann = {k: torch.as_tensor(v).to(self.device, torch.float32).contiguous() for k, v in dataset.ann.items()}
        self.n = len(dataset)
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.full_batch = batch_size >= self.n
        self._tag = 0
        self._out = {k: torch.empty(batch_size, v.shape[1], device=self.device) for k, v in self.dat.items()} \
            if not self.full_batch else None

    def __len__(self):
        if self.full_batch:
            return 1
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def __iter__(self):
        if self.full_batch:
            # a full batch is permutation invariant: feed the resident matrices themselves (planes are split once)
            yield self.dat, self.ann, None
            return
        perm = torch.randperm(self.n, device=self.device, generator=self.gen) if self.shuffle \
            else torch.arange(self.n, device=self.device)
        nb = len(self)
        for b in range(nb):
            idx = perm[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
            rows = idx.numel()
            dat = {}
            for k, src in self.dat.items():
                out = self._out[k][:rows]
                L.gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), rows, src.shape[1], out.data_ptr(),
                              out.stride(0))
                self._tag += 1
                out._fxn_tag = self._tag   # written through a raw pointer: tell the engine's plane cache it changed
                dat[k] = out
            yield dat, {k: v[idx] for k, v in self.ann.items()}, None
