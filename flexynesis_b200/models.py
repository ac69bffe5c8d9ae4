"""Drop-in model classes: same names, constructor signatures, attributes, methods and state_dict keys as
flexynesis.models.{DirectPred, supervised_vae, MultiTripletNetwork, GNN}, with the training
arithmetic executed by the B200 engine instead of torch autograd.

What callers of the reference rely on and is kept:
  * `Model(config, dataset, target_variables, batch_variables=None, surv_event_var=None, surv_time_var=None,
    use_loss_weighting=True, device_type=None)`                                  direct_pred.py:30-40
  * `training_step(batch, batch_idx, log=True)` / `validation_step(...)` return a scalar loss tensor on which
    `.backward()` populates `.grad` of every parameter (Lightning then clips and steps)      :225-294
  * `forward`, `configure_optimizers`, `compute_loss`, `compute_total_loss`, `predict`, `transform`, `forward_target`
  * attributes `config, encoders, MLPs, log_vars, variables, target_variables, layers, input_dims, ...`
New, optional: `fit_step(batch)` -- forward + backward + clip(1.0) + Adam fused on the device and CUDA-graph
replayable; `fit(...)` in flexynesis_b200.fit drives it.
"""
from __future__ import annotations

import itertools
from typing import Dict, List

import numpy as np
import pandas as pd
import torch
import torch.nn.functional as F
from torch import nn

from .containers import MLP, Decoder, Encoder, flexGCN

try:  # Lightning is optional: the reference trains under pl.Trainer, this engine also ships its own fit loop
    import lightning as pl
    _Base = pl.LightningModule
except Exception:  # pragma: no cover - exercised on boxes without lightning
    class _Base(nn.Module):
        """Minimal stand-in for lightning.LightningModule (logging hooks + device property)."""

        def __init__(self):
            super().__init__()
            self.logged = {}

        def log(self, name, value, **kw):
            self.logged[name] = value

        def log_dict(self, d, **kw):
            self.logged.update(d)

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")


def _resolve_device(device_type) -> torch.device:
    s = (device_type or "auto")
    s = s.lower() if isinstance(s, str) else "auto"
    if s in ("gpu", "cuda", "auto") and torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    if s.startswith("cuda"):
        return torch.device(s)
    return torch.device("cpu")


class _EngineLoss(torch.autograd.Function):
    """Connects the engine's hand-written backward to autograd: the forward value is the engine's total loss, the
    backward hands out the gradients the engine already computed (scaled by grad_output)."""

    @staticmethod
    def forward(ctx, total, holder, *params):
        ctx.holder = holder
        return total.detach().clone()

    @staticmethod
    def backward(ctx, grad_out):
        arena = ctx.holder.arena
        grads = [grad_out * arena.view(name, arena.grad) for name in arena.names]
        return (None, None, *grads)


class _EngineModel(_Base):
    """Shared plumbing of the four models."""

    engine_groups = 1
    extra_losses = ()

    # ---- engine lifetime (private cache: never pickled, never in state_dict) ----
    def _make_engine(self, device):
        raise NotImplementedError

    def engine(self, device=None):
        eng = self.__dict__.get("_engine")
        if device is None:
            device = next(self.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("flexynesis_b200 trains on CUDA devices only (no CPU fallback); move the model with "
                               ".to('cuda') or pass device_type='gpu'")
        if eng is None or eng.device != device or not eng.arena.intact():
            if next(self.parameters()).device != device:
                self.to(device)
            eng = self._make_engine(device)
            self.__dict__["_engine"] = eng
        return eng

    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_engine", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k == "_engine":
                continue
            new.__dict__[k] = copy.deepcopy(v, memo)
        # parameters of the copy must own their storage (the source's are views into its arena)
        for p in new.parameters():
            p.data = p.data.clone()
        return new

    # ---- reference API shared by all models ----
    def configure_optimizers(self):
        return torch.optim.Adam(self.parameters(), lr=self.config["lr"])

    def compute_loss(self, var, y, y_hat):
        """Torch formulation of the per-variable loss (API compatibility; the engine fuses this into its head
        kernels). Missing labels (NaN, or -1 for categoricals) are excluded; no valid label -> 0."""
        numeric = self.variable_types[var] == "numerical"
        ok = ~torch.isnan(y) if numeric else ((y != -1) & ~torch.isnan(y))
        if int(ok.sum()) == 0:
            return torch.tensor(0.0, device=y_hat.device, requires_grad=True)
        if numeric:
            return F.mse_loss(torch.flatten(y_hat[ok]), y[ok].float())
        return F.cross_entropy(y_hat[ok], y[ok].long())

    def compute_total_loss(self, losses):
        if self.use_loss_weighting and len(losses) > 1:
            return sum(torch.exp(-self.log_vars[k]) * v + self.log_vars[k] for k, v in losses.items())
        return sum(losses.values())

    def _split_batch(self, batch):
        dat, y_dict = batch[0], batch[1]
        return [[dat[k] for k in dat.keys()]], y_dict

    def _on_cuda(self, groups) -> bool:
        return all(t.is_cuda for g in groups for t in g)

    def training_step(self, train_batch, batch_idx, log=True, masks=None):
        groups, y = self._split_batch(train_batch)
        dev = groups[0][0].device if groups[0][0].is_cuda else next(self.parameters()).device
        eng = self.engine(dev)
        ws = eng.forward_backward(groups, y, masks)
        vals = eng.losses(ws)
        total = _EngineLoss.apply(vals["__total__"], eng, *[eng.arena.params[n] for n in eng.arena.names])
        losses = {k: v.detach() for k, v in vals.items() if not k.startswith("__")}
        losses["train_loss"] = total
        if log:
            self.log_dict(losses, on_step=False, on_epoch=True, prog_bar=True)
        return total

    def validation_step(self, val_batch, batch_idx, log=True, masks=None):
        """The reference's validation objective: the UNWEIGHTED sum of the loss terms (direct_pred.py:262-294). `masks`
        (tests only): replayed Gaussian draws for the VAE families, whose forward samples in eval mode as well."""
        groups, y = self._split_batch(val_batch)
        dev = groups[0][0].device if groups[0][0].is_cuda else next(self.parameters()).device
        eng = self.engine(dev)
        ws = eng.evaluate(groups, y, train_mode=self.training, masks=masks)
        vals = eng.losses(ws)
        total = vals["__val_total__"].detach().clone()      # unweighted sum (direct_pred.py:290)
        losses = {k: v.detach().clone() for k, v in vals.items() if not k.startswith("__")}
        losses["val_loss"] = total
        if log:
            self.log_dict(losses, on_step=False, on_epoch=True, prog_bar=True)
        return total

    def fit_step(self, batch, lr=None, masks=None, grad_scale: float = 1.0, allreduce=None):
        """forward + backward + clip_grad_norm_(1.0) + Adam, all on the engine. `allreduce(flat_grad)` (optional) is
        called between backward and the update (data-parallel training)."""
        groups, y = self._split_batch(batch)
        eng = self.engine(groups[0][0].device if groups[0][0].is_cuda else None)
        ws = eng.forward_backward(groups, y, masks)
        if allreduce is not None:
            allreduce(eng.arena.grad)
        eng.optimizer_step(float(self.config["lr"] if lr is None else lr), 1.0, grad_scale)
        return ws

    # ---- inference helpers ----
    def _batches(self, dataset, batch_size):
        n = len(dataset)
        keys = list(dataset.dat.keys())
        for s in range(0, n, batch_size):
            yield [dataset.dat[k][s:s + batch_size] for k in keys], list(dataset.samples[s:s + batch_size])

    def _outputs_from_ws(self, eng, ws) -> Dict[str, torch.Tensor]:
        return {v: ws["heads"]["logits"][v].clone() for v in eng.heads.vars}


def attribution_path(method: str, steps_or_samples: int, generator=None):
    """(alphas, weights) of the straight-line path from the all-zero baseline to the input, as captum evaluates it for
    the reference (direct_pred.py:472-520): IntegratedGradients(method='gausslegendre', n_steps) -> Gauss-Legendre nodes
    and weights on [0, 1]; GradientShap(n_samples, zero baselines, stdevs = 0) -> uniform random alphas, weights 1/n."""
    n = int(steps_or_samples)
    if method == "IntegratedGradients":
        x, w = np.polynomial.legendre.leggauss(n)
        return list(0.5 * (1.0 + x)), list(0.5 * w)
    if method == "GradientShap":
        a = torch.rand(n, generator=generator).tolist()
        return a, [1.0 / n] * n
    raise ValueError(f"Unsupported method '{method}'. Choose 'IntegratedGradients' or 'GradientShap'.")


class DirectPred(_EngineModel):
    """Fully connected multi-omics network with supervisor heads (flexynesis/models/direct_pred.py)."""

    def __init__(self, config, dataset, target_variables, batch_variables=None, surv_event_var=None,
                 surv_time_var=None, use_loss_weighting=True, device_type=None):
        super().__init__()
        self.config = config
        self.target_variables = target_variables
        self.surv_event_var, self.surv_time_var = surv_event_var, surv_time_var
        if surv_event_var is not None and surv_time_var is not None:
            self.target_variables = self.target_variables + [surv_event_var]
        self.batch_variables = batch_variables
        self.variables = self.target_variables + batch_variables if batch_variables else self.target_variables
        self.feature_importances = {}
        self.use_loss_weighting = use_loss_weighting
        self.device_type = device_type
        if use_loss_weighting:
            self.log_vars = nn.ParameterDict({v: nn.Parameter(torch.zeros(1)) for v in self._loss_names()})
        self.variable_types = dataset.variable_types
        self.ann = dataset.ann
        self.layers = list(dataset.dat.keys())
        self.input_dims = [len(dataset.features[k]) for k in self.layers]
        latent = config["latent_dim"]
        self.encoders = nn.ModuleList(
            [MLP(d, int(d * config["hidden_dim_factor"]), latent) for d in self.input_dims])
        self.fusion_block = nn.Linear(latent * len(self.layers), latent) if len(self.layers) > 1 else None
        self.MLPs = nn.ModuleDict()
        for var in self.variables:
            classes = 1 if self.variable_types[var] == "numerical" else len(np.unique(self.ann[var]))
            self.MLPs[var] = MLP(latent, config["supervisor_hidden_dim"], classes)

    def _loss_names(self) -> List[str]:
        return list(self.variables)

    def _make_engine(self, device):
        from .engine import TrunkEngine
        return TrunkEngine(self, device, groups=1)

    def forward(self, x_list):
        """{var: head output}. CUDA inputs without input-gradients run on the engine; anything else (captum's
        requires_grad inputs, CPU-resident inference) uses the differentiable torch formulation of the containers."""
        x_list = list(x_list)
        use_engine = all(x.is_cuda and not x.requires_grad for x in x_list) and not (
            self.training and torch.is_grad_enabled())
        if use_engine:
            eng = self.engine(x_list[0].device)
            ws = eng.evaluate([x_list], None, train_mode=self.training)
            return self._outputs_from_ws(eng, ws)
        emb = self._embed_torch(x_list)
        return {var: mlp(emb) for var, mlp in self.MLPs.items()}

    def _embed_torch(self, x_list):
        cat = torch.cat([enc(x) for enc, x in zip(self.encoders, x_list)], dim=1)
        return self.fusion_block(cat) if self.fusion_block is not None else cat

    def predict(self, dataset):
        """{var: np.ndarray}: class probabilities for categorical variables, raw outputs otherwise (:296-351)."""
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        preds = {v: [] for v in self.variables}
        with torch.no_grad():
            for xs, _ in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                out = self.forward([x.to(device, torch.float32) for x in xs])
                for v in self.variables:
                    o = out[v].detach().float().cpu()
                    preds[v].append(torch.softmax(o, dim=1) if dataset.variable_types[v] == "categorical" else o)
        return {v: torch.cat(p).numpy() for v, p in preds.items()}

    def transform(self, dataset):
        """Fused embeddings as a DataFrame with columns E0.. indexed by sample name (:353-415)."""
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        embs, names = [], []
        with torch.no_grad():
            for xs, samples in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                xs = [x.to(device, torch.float32) for x in xs]
                if device.type == "cuda":
                    eng = self.engine(device)
                    ws = eng.evaluate([xs], None)
                    e = eng.embedding(ws).clone()
                else:
                    e = self._embed_torch(xs)
                embs.append(e.cpu())
                names.extend(samples)
        e = torch.cat(embs, 0)
        return pd.DataFrame(e.numpy(), index=names, columns=[f"E{i}" for i in range(e.shape[1])])

    def forward_target(self, *args):
        """captum adaptor (direct_pred.py:418-431): args = (*layer_tensors[steps, B, d], target_var, steps)."""
        inputs, target_var, steps = list(args[:-2]), args[-2], args[-1]
        outs = []
        for i in range(steps):
            outs.append(self.forward([x[i] for x in inputs])[target_var])
        return torch.cat(outs, dim=0)

    def _anchor_inputs(self, dataset):
        """{layer: [N x d]} matrices the attribution runs over (the triplet dataset wraps the plain one)."""
        base = getattr(dataset, "dataset", dataset)
        return base.dat

    def compute_feature_importance(self, dataset, target_var, method="IntegratedGradients", steps_or_samples=5,
                                   batch_size=512):
        """Mean |attribution| per feature, class and layer (direct_pred.py:432-590) with the attributions computed by the
        engine: for every path point alpha_k the eval-mode forward and the input-gradient backward run on the B200
        kernels (engine.input_gradients), attr = x * sum_k w_k * d out[:, class] / d x at alpha_k * x. captum is not
        involved (it is not importable here either); `attribution_path` restates its quadrature."""
        device = _resolve_device(self.device_type)
        if device.type != "cuda":
            raise RuntimeError("compute_feature_importance runs on the CUDA engine; there is no CPU path")
        was_training = self.training
        self.to(device)
        self.eval()
        eng = self.engine(device)
        dat = self._anchor_inputs(dataset)
        layers = list(dat.keys())
        n = next(iter(dat.values())).shape[0]
        if dataset.variable_types[target_var] == "numerical":
            num_class = 1
        else:
            ann = torch.as_tensor(np.asarray(getattr(dataset, "ann", self.ann)[target_var], dtype=np.float64))
            num_class = len(np.unique(ann.numpy()))
        alphas, weights = attribution_path(method, steps_or_samples)
        sums = [[torch.zeros(len(dataset.features[k]), dtype=torch.float64, device=device) for k in layers]
                for _ in range(num_class)]
        with torch.no_grad():
            for s in range(0, n, batch_size):
                xs = [torch.as_tensor(dat[k][s:s + batch_size]).to(device, torch.float32).contiguous() for k in layers]
                G = [torch.empty_like(x) for x in xs]
                for cls in range(num_class):
                    for g in G:
                        g.zero_()
                    for al, w in zip(alphas, weights):
                        eng.input_gradients([x * float(al) for x in xs], target_var, cls, w, G)
                    for j, (x, g) in enumerate(zip(xs, G)):
                        sums[cls][j] += (x * g).abs().sum(0).double()
        self.to("cpu")
        if was_training:
            self.train()
        mappings = getattr(dataset, "label_mappings", {}) or {}
        frames = []
        for i in range(num_class):
            for j, k in enumerate(layers):
                label = mappings[target_var].get(i) if target_var in mappings else ""
                frames.append(pd.DataFrame({"target_variable": target_var, "target_class": i, "target_class_label": label,
                                            "layer": k, "name": dataset.features[k],
                                            "importance": (sums[i][j] / n).float().cpu().numpy()}))
        df = pd.concat(frames, ignore_index=True)
        self.feature_importances[target_var] = df
        return df


class MultiTripletNetwork(DirectPred):
    """DirectPred trunk applied to (anchor, positive, negative) + triplet margin loss + heads on the anchor
    (flexynesis/models/triplet_encoder.py). The first target variable must be categorical (:69-75)."""

    def __init__(self, config, dataset, target_variables, batch_variables=None, surv_event_var=None,
                 surv_time_var=None, use_loss_weighting=True, device_type=None):
        main_var = target_variables[0]
        if dataset.variable_types[main_var] == "numerical":
            raise ValueError("The first target variable", main_var, " must be a categorical variable")
        super().__init__(config, dataset, target_variables, batch_variables, surv_event_var, surv_time_var,
                         use_loss_weighting, device_type)
        self.main_var = main_var

    def _loss_names(self):
        return list(self.variables) + ["triplet_loss"]      # ParameterDict order of the reference (:81-84)

    def _make_engine(self, device):
        from .engine import TrunkEngine
        return TrunkEngine(self, device, groups=3, extra_losses=(("triplet_loss", 1),))

    def _split_batch(self, batch):
        anchor, pos, neg, y = batch[0], batch[1], batch[2], batch[3]
        return [[d[k] for k in d.keys()] for d in (anchor, pos, neg)], y

    def concat_embeddings(self, dat):
        return self._embed_torch([dat[k] for k in dat.keys()])

    def forward(self, anchor, positive, negative):
        """(anchor_emb, positive_emb, negative_emb, {var: head output on the anchor})."""
        groups = [[d[k] for k in d.keys()] for d in (anchor, positive, negative)]
        use_engine = all(x.is_cuda and not x.requires_grad for g in groups for x in g) and not (
            self.training and torch.is_grad_enabled())
        if use_engine:
            eng = self.engine(groups[0][0].device)
            ws = eng.evaluate(groups, None, train_mode=self.training)
            return (eng.embedding(ws, 0).clone(), eng.embedding(ws, 1).clone(), eng.embedding(ws, 2).clone(),
                    self._outputs_from_ws(eng, ws))
        ea, ep, en = (self._embed_torch(g) for g in groups)
        return ea, ep, en, {var: mlp(ea) for var, mlp in self.MLPs.items()}

    def triplet_loss(self, anchor, positive, negative, margin=1.0):
        dp = (anchor - positive).pow(2).sum(1)
        dn = (anchor - negative).pow(2).sum(1)
        return torch.relu(dp - dn + margin).mean()

    def _anchor_only(self, xs):
        return [xs, xs, xs]

    def predict(self, dataset):
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        preds = {v: [] for v in self.variables}
        with torch.no_grad():
            for xs, _ in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                xs = [x.to(device, torch.float32) for x in xs]
                d = dict(zip(self.layers, xs))
                out = self.forward(d, d, d)[3]
                for v in self.variables:
                    o = out[v].detach().float().cpu()
                    preds[v].append(torch.softmax(o, dim=1) if dataset.variable_types[v] == "categorical" else o)
        return {v: torch.cat(p).numpy() for v, p in preds.items()}

    def transform(self, dataset):
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        embs, names = [], []
        with torch.no_grad():
            for xs, samples in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                xs = [x.to(device, torch.float32) for x in xs]
                d = dict(zip(self.layers, xs))
                embs.append(self.forward(d, d, d)[0].cpu())
                names.extend(samples)
        e = torch.cat(embs, 0)
        return pd.DataFrame(e.numpy(), index=names, columns=[f"E{i}" for i in range(e.shape[1])])


class supervised_vae(_EngineModel):
    """MMD-regularised variational autoencoder with supervisor heads on the latent code
    (flexynesis/models/supervised_vae.py:42-130). Constructor order of the sub-modules follows the reference so that
    the same torch seed gives the same initial parameters."""

    def __init__(self, config, dataset, target_variables, batch_variables=None, surv_event_var=None,
                 surv_time_var=None, use_loss_weighting=True, device_type=None):
        super().__init__()
        self.config = config
        self.dataset = dataset
        self.target_variables = target_variables
        self.surv_event_var, self.surv_time_var = surv_event_var, surv_time_var
        if surv_event_var is not None and surv_time_var is not None:
            self.target_variables = self.target_variables + [surv_event_var]
        self.batch_variables = batch_variables
        self.variables = self.target_variables + batch_variables if batch_variables else self.target_variables
        self.feature_importances = {}
        self.nan_detected = False
        self.device_type = device_type
        self.use_loss_weighting = use_loss_weighting
        if use_loss_weighting:
            self.log_vars = nn.ParameterDict(
                {v: nn.Parameter(torch.zeros(1)) for v in itertools.chain(self.variables, ["mmd_loss"])})
        self.variable_types = dataset.variable_types
        self.layers = list(dataset.dat.keys())
        self.input_dims = [len(dataset.features[k]) for k in self.layers]
        latent, n = config["latent_dim"], len(self.layers)
        hidden = [max(int(d * config["hidden_dim_factor"]), 2) for d in self.input_dims]
        self.encoders = nn.ModuleList([Encoder(d, [h], latent) for d, h in zip(self.input_dims, hidden)])
        self.FC_mean = nn.Linear(n * latent, latent)
        self.FC_log_var = nn.Linear(n * latent, latent)
        self.decoders = nn.ModuleList([Decoder(latent, [h], d) for d, h in zip(self.input_dims, hidden)])
        self.MLPs = nn.ModuleDict()
        for var in self.variables:
            classes = 1 if self.variable_types[var] == "numerical" else len(np.unique(dataset.ann[var]))
            self.MLPs[var] = MLP(latent, config["supervisor_hidden_dim"], classes)

    def _make_engine(self, device):
        from .engine import VAEEngine
        return VAEEngine(self, device)

    # ---- torch formulation (CPU-resident inference, captum) ----
    def multi_encoder(self, x_list):
        means, log_vars = zip(*[enc(x) for enc, x in zip(self.encoders, x_list)])
        return self.FC_mean(torch.cat(means, dim=1)), self.FC_log_var(torch.cat(log_vars, dim=1))

    def reparameterization(self, mean, var):
        return mean + var * torch.randn_like(var)

    def forward(self, x_list):
        """(x_hat_list, z, mean, log_var, {var: head output}); epsilon is drawn in eval mode too, as in the reference
        (supervised_vae.py:187-200)."""
        x_list = list(x_list)
        use_engine = all(x.is_cuda and not x.requires_grad for x in x_list) and not (
            self.training and torch.is_grad_enabled())
        if use_engine:
            eng = self.engine(x_list[0].device)
            ws = eng.evaluate([x_list], None, train_mode=self.training, want_xhat=True)
            Lt = eng.latent
            return ([t.clone() for t in ws["xhat"]], ws["z"][:, :Lt].clone(), ws["mean"][:, :Lt].clone(),
                    ws["s"][:, :Lt].clone(), self._outputs_from_ws(eng, ws))
        mean, log_var = self.multi_encoder(x_list)
        z = self.reparameterization(mean, log_var)
        return [dec(z) for dec in self.decoders], z, mean, log_var, {v: mlp(z) for v, mlp in self.MLPs.items()}

    def compute_kernel(self, x, y):
        dim = x.shape[1]
        d2 = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2.0 * x @ y.T
        return torch.exp(-d2.clamp_min(0.0) / float(dim * dim))

    def compute_mmd(self, x, y):
        return self.compute_kernel(x, x).mean() + self.compute_kernel(y, y).mean() - 2 * self.compute_kernel(x, y).mean()

    def MMD_loss(self, latent_dim, z, xhat, x):
        prior = torch.randn(200, latent_dim, device=z.device)
        return self.compute_mmd(prior, z) + torch.mean((xhat - x) ** 2)

    def _run_eval(self, dataset, want):
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        out = {v: [] for v in self.variables}
        embs, names = [], []
        with torch.no_grad():
            for xs, samples in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                _, z, _, _, outputs = self.forward([x.to(device, torch.float32) for x in xs])
                embs.append(z.detach().cpu())
                names.extend(samples)
                for v in self.variables:
                    o = outputs[v].detach().float().cpu()
                    out[v].append(torch.softmax(o, dim=1) if dataset.variable_types[v] == "categorical" else o)
        if want == "embedding":
            e = torch.cat(embs, 0)
            return pd.DataFrame(e.numpy(), index=names, columns=[f"E{i}" for i in range(e.shape[1])])
        return {v: torch.cat(p).numpy() for v, p in out.items()}

    def transform(self, dataset):
        """Latent codes z as a DataFrame (supervised_vae.py:383-436)."""
        return self._run_eval(dataset, "embedding")

    def predict(self, dataset):
        """{var: np.ndarray} (supervised_vae.py:438-492)."""
        return self._run_eval(dataset, "predict")

    def forward_target(self, *args):
        inputs, target_var, steps = list(args[:-2]), args[-2], args[-1]
        outs = []
        for i in range(steps):
            outs.append(self.forward([x[i] for x in inputs])[4][target_var])
        return torch.cat(outs, dim=0)

    def compute_feature_importance(self, dataset, target_var, method="IntegratedGradients", steps_or_samples=5,
                                   batch_size=512):
        """Mean |attribution| per feature, class and (input) layer with the reference's DataFrame layout
        (supervised_vae.py / crossmodal_pred.py `compute_feature_importance`). The integrand is evaluated with torch
        autograd on the container modules (any device) -- this model family's attribution is not an engine path; only
        captum's quadrature (`attribution_path`) is restated. As in the reference the latent code is re-drawn at every
        path point (`reparameterization` has no eval branch)."""
        device = _resolve_device(self.device_type)
        was_training = self.training
        self.to(device)
        self.eval()
        layers = list(self.layers)
        n = len(dataset.samples) if hasattr(dataset, "samples") else next(iter(dataset.dat.values())).shape[0]
        if dataset.variable_types[target_var] == "numerical":
            num_class = 1
        else:
            num_class = len(np.unique(np.asarray(dataset.ann[target_var], dtype=np.float64)))
        alphas, weights = attribution_path(method, steps_or_samples)
        sums = [[torch.zeros(len(dataset.features[k]), dtype=torch.float64) for k in layers] for _ in range(num_class)]
        for s in range(0, n, batch_size):
            xs = [torch.as_tensor(dataset.dat[k][s:s + batch_size]).to(device, torch.float32) for k in layers]
            for cls in range(num_class):
                G = [torch.zeros_like(x) for x in xs]
                for al, w in zip(alphas, weights):
                    xk = [(x * float(al)).detach().requires_grad_(True) for x in xs]
                    out = self.forward(xk)[4][target_var]
                    grads = torch.autograd.grad(out[:, cls].sum(), xk)
                    for g, d in zip(G, grads):
                        g += float(w) * d
                for j, (x, g) in enumerate(zip(xs, G)):
                    sums[cls][j] += (x * g).abs().sum(0).double().cpu()
        self.to("cpu")
        if was_training:
            self.train()
        mappings = getattr(dataset, "label_mappings", {}) or {}
        frames = []
        for i in range(num_class):
            for j, k in enumerate(layers):
                label = mappings[target_var].get(i) if target_var in mappings else ""
                frames.append(pd.DataFrame({"target_variable": target_var, "target_class": i, "target_class_label": label,
                                            "layer": k, "name": dataset.features[k],
                                            "importance": (sums[i][j] / n).float().numpy()}))
        df = pd.concat(frames, ignore_index=True)
        self.feature_importances[target_var] = df
        return df


class CrossModalPred(supervised_vae):
    """Cross-modality VAE (flexynesis/models/crossmodal_pred.py:31-187): Encoders over `input_layers`, Decoders into
    `output_layers` (each defaulting to every layer of the dataset), the same MMD + reconstruction + supervisor-head
    objective as supervised_vae (:293-351) with the reconstruction targets taken from the OUTPUT layers. Runs on the
    VAE engine, which keeps separate encoder / decoder layer sets."""

    def __init__(self, config, dataset, target_variables=None, batch_variables=None, surv_event_var=None,
                 surv_time_var=None, input_layers=None, output_layers=None, use_loss_weighting=True, device_type=None):
        _EngineModel.__init__(self)
        self.config = config
        self.target_variables = list(target_variables or [])
        self.surv_event_var, self.surv_time_var = surv_event_var, surv_time_var
        if surv_event_var is not None and surv_time_var is not None:
            self.target_variables = self.target_variables + [surv_event_var]
        self.batch_variables = batch_variables
        self.variables = self.target_variables + batch_variables if batch_variables else self.target_variables
        self.variable_types = dataset.variable_types
        self.ann = dataset.ann
        self.input_layers = list(input_layers) if input_layers else list(dataset.dat.keys())
        self.output_layers = list(output_layers) if output_layers else list(dataset.dat.keys())
        self.layers = self.input_layers
        self.feature_importances = {}
        self.nan_detected = False
        self.device_type = device_type
        self.use_loss_weighting = use_loss_weighting
        if use_loss_weighting:
            self.log_vars = nn.ParameterDict(
                {v: nn.Parameter(torch.zeros(1)) for v in itertools.chain(self.variables, ["mmd_loss"])})
        latent, f = config["latent_dim"], config["hidden_dim_factor"]
        self.input_dims = [len(dataset.features[k]) for k in self.input_layers]
        self.output_dims = [len(dataset.features[k]) for k in self.output_layers]
        self.encoders = nn.ModuleList([Encoder(d, [int(d * f)], latent) for d in self.input_dims])
        self.FC_mean = nn.Linear(len(self.input_layers) * latent, latent)
        self.FC_log_var = nn.Linear(len(self.input_layers) * latent, latent)
        self.decoders = nn.ModuleList([Decoder(latent, [int(d * f)], d) for d in self.output_dims])
        self.MLPs = nn.ModuleDict()
        for var in self.variables:
            classes = 1 if self.variable_types[var] == "numerical" else len(np.unique(self.ann[var]))
            self.MLPs[var] = MLP(latent, config["supervisor_hidden_dim"], classes)

    def _split_batch(self, batch):
        dat, y_dict = batch[0], batch[1]
        return [[dat[k] for k in self.input_layers], [dat[k] for k in self.output_layers]], y_dict

    def _batches(self, dataset, batch_size):
        n = len(dataset)
        for s in range(0, n, batch_size):
            yield [dataset.dat[k][s:s + batch_size] for k in self.input_layers], list(dataset.samples[s:s + batch_size])

    def forward(self, x_list_input):
        x_list = list(x_list_input)
        use_engine = all(x.is_cuda and not x.requires_grad for x in x_list) and not (
            self.training and torch.is_grad_enabled())
        if use_engine:
            eng = self.engine(x_list[0].device)
            B = x_list[0].shape[0]
            # the fused decoder epilogue always reads a reconstruction target; inference has none: zeros
            dummy = [torch.zeros(B, d, device=x_list[0].device) for d in self.output_dims]
            ws = eng.evaluate([x_list, dummy], None, train_mode=self.training, want_xhat=True)
            Lt = eng.latent
            return ([t.clone() for t in ws["xhat"]], ws["z"][:, :Lt].clone(), ws["mean"][:, :Lt].clone(),
                    ws["s"][:, :Lt].clone(), self._outputs_from_ws(eng, ws))
        mean, log_var = self.multi_encoder(x_list)
        z = self.reparameterization(mean, log_var)
        return [dec(z) for dec in self.decoders], z, mean, log_var, {v: mlp(z) for v, mlp in self.MLPs.items()}

    def decode(self, dataset):
        """{output layer: DataFrame [samples x features]} of reconstructions (crossmodal_pred.py:467-481)."""
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        outs, names = [[] for _ in self.output_layers], []
        with torch.no_grad():
            for xs, samples in self._batches(dataset, 4096 if device.type == "cuda" else 64):
                x_hat = self.forward([x.to(device, torch.float32) for x in xs])[0]
                for j, t in enumerate(x_hat):
                    outs[j].append(t.detach().cpu())
                names.extend(samples)
        # features x samples, as the reference lays the frames out
        return {k: pd.DataFrame(torch.cat(outs[j], 0).numpy().T, index=dataset.features[k], columns=names)
                for j, k in enumerate(self.output_layers)}


class GNN(_EngineModel):
    """Graph-convolutional early-fusion model (flexynesis/models/gnn_early.py:55-140): one flexGCN over node features
    [B, N, F] with a graph shared by all samples, supervisor heads on its embedding."""

    def __init__(self, config, dataset, target_variables, batch_variables=None, surv_event_var=None,
                 surv_time_var=None, use_loss_weighting=True, device_type=None, gnn_conv_type=None):
        super().__init__()
        self.config = config
        self.target_variables = target_variables
        self.surv_event_var, self.surv_time_var = surv_event_var, surv_time_var
        if surv_event_var is not None and surv_time_var is not None:
            self.target_variables = self.target_variables + [surv_event_var]
        self.batch_variables = batch_variables
        self.variables = self.target_variables + batch_variables if batch_variables else self.target_variables
        base = getattr(dataset, "multiomic_dataset", dataset)
        self.variable_types = base.variable_types
        self.ann = base.ann
        self.feature_importances = {}
        self.use_loss_weighting = use_loss_weighting
        self.device_type = device_type
        self.gnn_conv_type = gnn_conv_type
        self.edge_index = dataset.edge_index.to(_resolve_device(device_type))   # shared by all samples: kept on device
        if use_loss_weighting:
            self.log_vars = nn.ParameterDict({v: nn.Parameter(torch.zeros(1)) for v in self.variables})
        first = dataset[0][0]
        self.encoders = nn.ModuleList([flexGCN(
            node_count=first.shape[0], node_feature_count=first.shape[1],
            node_embedding_dim=int(config["node_embedding_dim"]), num_convs=int(config["num_convs"]),
            output_dim=config["latent_dim"], act=config["activation"], conv=gnn_conv_type)])
        self.MLPs = nn.ModuleDict()
        for var in self.variables:
            classes = 1 if self.variable_types[var] == "numerical" else len(np.unique(self.ann[var]))
            self.MLPs[var] = MLP(config["latent_dim"], config["supervisor_hidden_dim"], classes)

    def _make_engine(self, device):
        from .engine import GNNEngine
        return GNNEngine(self, device)

    def _split_batch(self, batch):
        return [[batch[0]]], batch[1]

    def _same_graph(self, edge_index) -> bool:
        return edge_index is None or edge_index is self.edge_index or (
            edge_index.shape == self.edge_index.shape and bool((edge_index.to(self.edge_index.device) == self.edge_index).all()))

    def _embed(self, x, edge_index=None):
        """flexGCN embedding [B, latent]; engine for CUDA inputs on the model's own graph, torch containers otherwise."""
        use_engine = x.is_cuda and not x.requires_grad and not (self.training and torch.is_grad_enabled()) \
            and self._same_graph(edge_index)
        if use_engine:
            eng = self.engine(x.device)
            ws = eng.evaluate([[x]], None, train_mode=self.training)
            return eng.embedding(ws).clone(), self._outputs_from_ws(eng, ws)
        ei = self.edge_index if edge_index is None else edge_index
        emb = self.encoders[0](x, ei.to(x.device))
        return emb, {v: mlp(emb) for v, mlp in self.MLPs.items()}

    def forward(self, x, edge_index=None):
        return self._embed(x, edge_index)[1]

    def forward_target(self, *args):
        """captum adaptor (gnn_early.py:428-438): args = (x[steps, B, N, F], target_var, steps)."""
        x, target_var, steps = args[0], args[-2], args[-1]
        return torch.cat([self.forward(x[i])[target_var] for i in range(steps)], dim=0)

    def compute_feature_importance(self, dataset, target_var, method="IntegratedGradients", steps_or_samples=5,
                                   batch_size=512):
        """Mean |attribution| per gene, class and omics layer with the reference's DataFrame layout (gnn_early.py:440-633:
        one row per (class, layer = node-feature column, gene)). The path integral runs through torch autograd on the
        container modules (contract safety; not an engine path), with captum's quadrature (`attribution_path`)."""
        device = _resolve_device(self.device_type)
        was_training = self.training
        self.to(device)
        self.eval()
        if dataset.variable_types[target_var] == "numerical":
            num_class = 1
        else:
            num_class = len(np.unique(np.asarray(dataset.ann[target_var], dtype=np.float64)))
        alphas, weights = attribution_path(method, steps_or_samples)
        sums, n = [None] * num_class, 0
        for x, _ in self._node_batches(dataset, batch_size):
            x = x.to(device, torch.float32)
            n += x.shape[0]
            for cls in range(num_class):
                G = torch.zeros_like(x)
                for al, w in zip(alphas, weights):
                    xk = (x * float(al)).detach().requires_grad_(True)
                    out = self.forward(xk)[target_var]
                    G += float(w) * torch.autograd.grad(out[:, cls].sum(), [xk])[0]
                red = (x * G).abs().sum(0).double().cpu()                      # [N, F]
                sums[cls] = red if sums[cls] is None else sums[cls] + red
        self.to("cpu")
        if was_training:
            self.train()
        layers = list(getattr(dataset, "multiomic_dataset", dataset).dat.keys()) if hasattr(
            getattr(dataset, "multiomic_dataset", dataset), "dat") else [f"layer{j}" for j in range(sums[0].shape[1])]
        genes = getattr(dataset, "common_features", None) or [f"node{j}" for j in range(sums[0].shape[0])]
        mappings = getattr(dataset, "label_mappings", {}) or {}
        frames = []
        for i in range(num_class):
            imp = (sums[i] / n).float().numpy()
            label = mappings[target_var].get(i) if target_var in mappings else ""
            for j, layer in enumerate(layers):
                frames.append(pd.DataFrame({"target_variable": target_var, "target_class": i, "target_class_label": label,
                                            "layer": layer, "name": genes,
                                            "importance": imp[:, j] if imp.shape[1] > 1 else imp[:, 0]}))
        df = pd.concat(frames, ignore_index=True)
        self.feature_importances[target_var] = df
        return df

    def _node_batches(self, dataset, batch_size):
        n = len(dataset)
        feats = getattr(dataset, "node_features_tensor", None)
        for s in range(0, n, batch_size):
            idx = range(s, min(s + batch_size, n))
            x = feats[s:s + batch_size] if feats is not None else torch.stack([dataset[i][0] for i in idx])
            yield x, [dataset[i][2] for i in idx] if feats is None else list(dataset.samples[s:s + batch_size])

    def _run_eval(self, dataset, want):
        self.eval()
        device = _resolve_device(self.device_type)
        self.to(device)
        self.edge_index = self.edge_index.to(device)
        vt = getattr(dataset, "variable_types", self.variable_types)
        out = {v: [] for v in self.variables}
        embs, names = [], []
        with torch.no_grad():
            for x, samples in self._node_batches(dataset, 4096 if device.type == "cuda" else 64):
                emb, outputs = self._embed(x.to(device, torch.float32), dataset.edge_index.to(device))
                embs.append(emb.detach().cpu())
                names.extend(samples)
                for v in self.variables:
                    o = outputs[v].detach().float().cpu()
                    out[v].append(torch.softmax(o, dim=1) if vt[v] == "categorical" else o)
        if want == "embedding":
            e = torch.cat(embs, 0)
            return pd.DataFrame(e.numpy(), index=names, columns=[f"E{i}" for i in range(e.shape[1])])
        return {v: torch.cat(p).numpy() for v, p in out.items()}

    def predict(self, dataset):
        return self._run_eval(dataset, "predict")

    def transform(self, dataset):
        return self._run_eval(dataset, "embedding")
