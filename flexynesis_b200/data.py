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
        # static batch buffers (features and labels): a captured training step reads them in place, batch after batch
        self._out = {k: torch.empty(batch_size, v.shape[1], device=self.device) for k, v in self.dat.items()} \
            if not self.full_batch else None
        self._yout = {k: torch.empty(batch_size, device=self.device) for k in self.ann} if not self.full_batch else None

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
            yy = {}
            for k, v in self.ann.items():
                yy[k] = torch.index_select(v, 0, idx, out=self._yout[k][:rows])
            yield dat, yy, None


class DeviceNodeBatcher:
    """Batches for the GNN from the reference's MultiOmicDatasetNW duck type (flexynesis/data.py:1154-1262):
    `node_features_tensor` [samples, nodes, features] and `ann` stay resident in HBM; a batch is one row gather over
    the [samples, nodes * features] view. Yields (x [B, nodes, features], y_dict, None) like default_collate over
    MultiOmicDatasetNW.__getitem__ (:1257-1262)."""

    def __init__(self, dataset, batch_size: int, device, shuffle: bool = True, drop_last: bool = True, seed: int = 0):
        self.device = torch.device(device)
        x = torch.as_tensor(dataset.node_features_tensor).to(self.device, torch.float32).contiguous()
        self.shape = tuple(x.shape[1:])
        self.x = x.view(x.shape[0], -1)
        self.ann = {k: torch.as_tensor(v).to(self.device, torch.float32).contiguous() for k, v in dataset.ann.items()}
        self.n = int(x.shape[0])
        self.batch_size, self.shuffle, self.drop_last = batch_size, shuffle, drop_last
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.full_batch = batch_size >= self.n
        self._tag = 0
        self._out = None if self.full_batch else torch.empty(batch_size, self.x.shape[1], device=self.device)
        self._yout = None if self.full_batch else {k: torch.empty(batch_size, device=self.device) for k in self.ann}

    def __len__(self):
        if self.full_batch:
            return 1
        return self.n // self.batch_size if self.drop_last else -(-self.n // self.batch_size)

    def __iter__(self):
        if self.full_batch:
            yield self.x.view(self.n, *self.shape), self.ann, None
            return
        perm = torch.randperm(self.n, device=self.device, generator=self.gen) if self.shuffle \
            else torch.arange(self.n, device=self.device)
        for b in range(len(self)):
            idx = perm[b * self.batch_size:(b + 1) * self.batch_size].contiguous()
            rows = idx.numel()
            out = self._out[:rows]
            L.gather_rows(self.x.data_ptr(), self.x.stride(0), idx.data_ptr(), rows, self.x.shape[1], out.data_ptr(),
                          out.stride(0))
            self._tag += 1
            xb = out.view(rows, *self.shape)
            xb._fxn_tag = self._tag
            yy = {k: torch.index_select(v, 0, idx, out=self._yout[k][:rows]) for k, v in self.ann.items()}
            yield xb, yy, None


class DeviceTripletBatcher:
    """(anchor, positive, negative, y_dict) batches for MultiTripletNetwork built ON THE DEVICE: replaces
    TripletMultiOmicDataset.__getitem__ + default_collate (flexynesis/data.py:1088-1151), which draw one positive and
    one negative per anchor with numpy / random on the host and stack 3 x B dict lookups per batch.

    Same sampling law as the reference: anchors are the samples whose `main_var` label is not NaN (each once per epoch,
    shuffled); the positive is uniform over the OTHER samples with the anchor's label; the negative label is uniform
    over the other label groups -- missing labels form one group "NA" that negatives may come from (:1122-1127) -- and
    the negative uniform inside it. (A label with a single sample makes the reference spin forever; here the anchor is
    its own positive.) Index arithmetic is a handful of torch ops on [B]-vectors; the rows are gathered by
    fxn_gather_rows from the HBM-resident matrices into three persistent buffers per layer."""

    def __init__(self, dataset, main_var: str, batch_size: int, device, shuffle: bool = True, drop_last: bool = True,
                 seed: int = 0):
        base = getattr(dataset, "dataset", dataset)
        self.device = torch.device(device)
        self.dat = {k: torch.as_tensor(v).to(self.device, torch.float32).contiguous() for k, v in base.dat.items()}
        self.ann = {k: torch.as_tensor(v).to(self.device, torch.float32).contiguous() for k, v in base.ann.items()}
        self.main_var, self.batch_size, self.shuffle, self.drop_last = main_var, batch_size, shuffle, drop_last
        lab = torch.as_tensor(base.ann[main_var]).to(torch.float64).cpu()
        nan = torch.isnan(lab)
        uniq = torch.unique(lab[~nan])
        gid = torch.searchsorted(uniq, torch.nan_to_num(lab, nan=float(uniq[0]) if uniq.numel() else 0.0))
        gid[nan] = uniq.numel()                                   # the "NA" group
        self.n_groups = int(uniq.numel()) + int(bool(nan.any()))
        if self.n_groups < 2:
            raise ValueError("triplet sampling needs at least two label groups")
        order = torch.sort(gid, stable=True).indices
        count = torch.bincount(gid, minlength=self.n_groups)
        start = torch.cumsum(count, 0) - count
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel()) - start[gid[order]]
        dev = self.device
        self.gid, self.order, self.count, self.start, self.rank = (t.to(dev) for t in (gid, order, count, start, rank))
        self.valid = torch.nonzero(~nan).flatten().to(dev)
        self.gen = torch.Generator(device=dev).manual_seed(seed)
        self._buf = [{k: torch.empty(batch_size, v.shape[1], device=dev) for k, v in self.dat.items()} for _ in range(3)]
        self._ybuf = {k: torch.empty(batch_size, device=dev) for k in self.ann}
        self._tag = 0

    def __len__(self):
        n = self.valid.numel()
        return n // self.batch_size if self.drop_last else -(-n // self.batch_size)

    def sample_indices(self, anchors: torch.Tensor):
        """positive / negative sample indices for a vector of anchor indices (device tensors)."""
        g = self.gid[anchors]
        cnt = self.count[g]
        u = torch.rand(anchors.numel(), 3, device=self.device, generator=self.gen, dtype=torch.float64)
        r = torch.minimum((u[:, 0] * (cnt - 1).clamp_min(1)).long(), (cnt - 2).clamp_min(0))
        r = r + (r >= self.rank[anchors]).long()                  # skip the anchor's own slot
        pos = torch.where(cnt > 1, self.order[self.start[g] + r.clamp_max(cnt - 1)], anchors)
        gn = torch.minimum((u[:, 1] * (self.n_groups - 1)).long(), torch.full_like(g, self.n_groups - 2))
        gn = gn + (gn >= g).long()                                # uniform over the OTHER groups
        cn = self.count[gn]
        neg = self.order[self.start[gn] + torch.minimum((u[:, 2] * cn).long(), cn - 1)]
        return pos, neg

    def _gather(self, slot: int, idx: torch.Tensor):
        out = {}
        idx = idx.contiguous()
        for k, src in self.dat.items():
            dst = self._buf[slot][k][:idx.numel()]
            L.gather_rows(src.data_ptr(), src.stride(0), idx.data_ptr(), idx.numel(), src.shape[1], dst.data_ptr(),
                          dst.stride(0))
            self._tag += 1
            dst._fxn_tag = self._tag
            out[k] = dst
        return out

    def __iter__(self):
        n = self.valid.numel()
        perm = self.valid[torch.randperm(n, device=self.device, generator=self.gen)] if self.shuffle else self.valid
        for b in range(len(self)):
            a = perm[b * self.batch_size:(b + 1) * self.batch_size]
            p, q = self.sample_indices(a)
            yy = {k: torch.index_select(v, 0, a, out=self._ybuf[k][:a.numel()]) for k, v in self.ann.items()}
            yield self._gather(0, a), self._gather(1, p), self._gather(2, q), yy


# ----------------------------------------------------------------------------------------------------
# on-disk matrices -> pinned host memory -> HBM
# ----------------------------------------------------------------------------------------------------
def load_matrix_npy(path: str, device="cuda", rows_per_chunk: int = 8192, dtype=torch.float32) -> torch.Tensor:
    """A [samples x features] `.npy` matrix (what `numpy.save` writes; any float / integer dtype, C order) as a tensor on
    `device`, without ever holding a second full copy on the host: the file is memory-mapped, and `rows_per_chunk` rows at
    a time are converted into one of two PINNED staging buffers and copied to the device asynchronously, so that the
    conversion of chunk i + 1 overlaps the DMA of chunk i.

    The reference reads modality matrices through pandas (`DataImporter.read_data`, data.py:155-190) or h5py
    (`H5DataImporter._read_h5_as_dataframe`, h5_dataloader.py:79-103: `/matrix` float32, samples as rows) into DataFrames
    and converts to tensors later; the matrix layout is the same [samples x features] float32 as here. (h5py is not
    installed in this image, so the HDF5 container itself is not read; `numpy.save(path, h5['matrix'][...])` converts.)"""
    import numpy as np
    arr = np.load(path, mmap_mode="r")
    if arr.ndim != 2:
        raise ValueError(f"{path}: expected a 2-D [samples x features] matrix, got shape {arr.shape}")
    n, d = arr.shape
    dev = torch.device(device)
    out = torch.empty(n, d, dtype=dtype, device=dev)
    if dev.type != "cuda":
        for r0 in range(0, n, rows_per_chunk):
            out[r0:r0 + rows_per_chunk] = torch.from_numpy(np.array(arr[r0:r0 + rows_per_chunk])).to(dtype)
        return out
    rows = min(rows_per_chunk, max(n, 1))
    stage = [torch.empty(rows, d, dtype=dtype).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event(), torch.cuda.Event()]
    copy_stream = torch.cuda.Stream(device=dev)
    for i, r0 in enumerate(range(0, n, rows)):
        k, m = i % 2, min(rows, n - r0)
        if i >= 2:
            done[k].synchronize()                        # the DMA that last used this staging buffer has finished
        stage[k][:m].numpy()[...] = arr[r0:r0 + m]      # numpy converts straight from the mapped file into pinned memory
        with torch.cuda.stream(copy_stream):
            out[r0:r0 + m].copy_(stage[k][:m], non_blocking=True)
            done[k].record(copy_stream)
    copy_stream.synchronize()
    return out


def dataset_from_npy(paths: Dict[str, str], ann: Dict[str, torch.Tensor], variable_types: Dict[str, str], device="cuda",
                     samples: Optional[List[str]] = None, features: Optional[Dict[str, List[str]]] = None):
    """MultiOmicDataset duck type (dat / ann / features / samples / variable_types, reference data.py:940-1000) whose
    modality matrices come from `.npy` files via load_matrix_npy and stay resident on `device` (DeviceBatcher's `.to(device)`
    is then a no-op: the matrices are never copied again)."""
    dat = {k: load_matrix_npy(p, device) for k, p in paths.items()}
    n = {v.shape[0] for v in dat.values()}
    if len(n) != 1:
        raise ValueError(f"modalities disagree on the number of samples: { {k: v.shape[0] for k, v in dat.items()} }")
    n = n.pop()
    ds = SyntheticMultiOmicDataset.__new__(SyntheticMultiOmicDataset)
    ds.dat = dat
    ds.ann = {k: torch.as_tensor(v).cpu() for k, v in ann.items()}     # labels stay on the host (the constructors count classes
                                                                         # with numpy); the batchers move them to the device
    for k, v in ds.ann.items():
        if v.shape[0] != n:
            raise ValueError(f"annotation {k!r} has {v.shape[0]} rows, the matrices have {n}")
    ds.variable_types = dict(variable_types)
    ds.samples = list(samples) if samples is not None else [f"s{i}" for i in range(n)]
    ds.features = features if features is not None else {k: [f"{k}_{j}" for j in range(v.shape[1])] for k, v in dat.items()}
    ds.label_mappings, ds.feature_ann = {}, {}
    return ds
