"""CPU oracle: a functional restatement of flexynesis's training hot path.

TEST INFRASTRUCTURE ONLY. Nothing in the product package (flexynesis_b200/) may import this module; only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do, and only as the
checker or as the timed CPU baseline.

Parity status: PINNED. The reference ships no golden vectors for this path, so the
restatement is pinned against the reference's own unmodified source executed on CPU through
oracle/ref_shim.py (see oracle/make_golden.py, which also writes tests/golden/*.pt). The only arithmetic that
is *not* under /root/reference is PyG's GCNConv (third-party torch_geometric, un-pinned in
pyproject.toml:40); `gcn_conv` restates its published algorithm and is anchored on the reference call
sites flexynesis/modules.py:221-226, :239-246, :254.

Style: pure functions over a flat {state_dict key: tensor} mapping `P` (the reference's own key names, see
SURVEY.md section 8b) instead of nn.Module classes. Noise (dropout masks, epsilon, the MMD prior) is an
explicit input so CPU and GPU runs can replay identical draws.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


# ----------------------------------------------------------------------------------------------------
# noise handling
# ----------------------------------------------------------------------------------------------------
class Noise:
    """Source of every random draw of a step. `record` maps a site name to the tensor that was used.

    mode "replay": take draws from `given` (KeyError if one is missing);
    mode "draw":   draw with torch's global CPU generator (what the reference does) and record them.
    """

    def __init__(self, given: Optional[Dict[str, Tensor]] = None):
        self.given = given
        self.record: Dict[str, Tensor] = {}

    def dropout_mask(self, site: str, like: Tensor, p: float) -> Tensor:
        if self.given is not None:
            m = self.given[site]
        else:
            m = (torch.rand_like(like) >= p).to(like.dtype)
        self.record[site] = m
        return m

    def normal(self, site: str, shape: Sequence[int], like: Tensor) -> Tensor:
        if self.given is not None:
            e = self.given[site]
        else:
            e = torch.randn(*shape, dtype=like.dtype, device=like.device)
        self.record[site] = e
        return e


# ----------------------------------------------------------------------------------------------------
# building blocks (flexynesis/modules.py)
# ----------------------------------------------------------------------------------------------------
def batchnorm(P: Dict[str, Tensor], prefix: str, x: Tensor, train: bool) -> Tensor:
    """nn.BatchNorm1d(eps=1e-5, momentum=0.1): batch stats + running-stat update in train mode
    (biased var for normalisation, unbiased for the running buffer), running stats in eval mode."""
    w, b = P[prefix + ".weight"], P[prefix + ".bias"]
    rm, rv = P[prefix + ".running_mean"], P[prefix + ".running_var"]
    if train:
        n = x.shape[0]
        if n <= 1:
            raise ValueError("Expected more than 1 value per channel when training")
        mean = x.mean(0)
        var = x.var(0, unbiased=False)
        with torch.no_grad():
            rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
            rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * var.detach() * n / (n - 1))
            P[prefix + ".num_batches_tracked"] += 1
    else:
        mean, var = rm, rv
    return (x - mean) * torch.rsqrt(var + BN_EPS) * w + b


def mlp(P, prefix: str, x: Tensor, train: bool, noise: Noise, p_drop: float = 0.1, site: str = "") -> Tensor:
    """MLP.forward, modules.py:135-150: Linear -> BatchNorm1d -> ReLU -> Dropout(0.1) -> Linear
    (layer_out has no bias when output_dim == 1, modules.py:126-130). `site` prefixes the noise-site name when
    the same module runs several times per step (triplet anchor/positive/negative)."""
    z = F.linear(x, P[prefix + ".layer_1.weight"], P[prefix + ".layer_1.bias"])
    y = batchnorm(P, prefix + ".batchnorm", z, train)
    a = torch.relu(y)
    if train:
        a = a * noise.dropout_mask(site + prefix + ".dropout", a, p_drop) / (1.0 - p_drop)
    return F.linear(a, P[prefix + ".layer_out.weight"], P.get(prefix + ".layer_out.bias"))


def vae_hidden(P, prefix: str, x: Tensor, train: bool) -> Tensor:
    """Encoder/Decoder trunk, modules.py:21-36 / :71-86 with a one-entry hidden_dims list:
    Linear -> LeakyReLU(0.2) -> BatchNorm1d  (activation BEFORE the norm, no dropout)."""
    z = F.linear(x, P[prefix + ".hidden_layers.0.weight"], P[prefix + ".hidden_layers.0.bias"])
    return batchnorm(P, prefix + ".hidden_layers.2", F.leaky_relu(z, 0.2), train)


def vae_encoder(P, prefix: str, x: Tensor, train: bool) -> Tuple[Tensor, Tensor]:
    """Encoder.forward, modules.py:43-57."""
    h = vae_hidden(P, prefix, x, train)
    return (F.linear(h, P[prefix + ".FC_mean.weight"], P[prefix + ".FC_mean.bias"]),
            F.linear(h, P[prefix + ".FC_var.weight"], P[prefix + ".FC_var.bias"]))


def vae_decoder(P, prefix: str, z: Tensor, train: bool) -> Tensor:
    """Decoder.forward, modules.py:91-103."""
    h = vae_hidden(P, prefix, z, train)
    return torch.sigmoid(F.linear(h, P[prefix + ".FC_output.weight"], P[prefix + ".FC_output.bias"]))


def cox_ph(outputs: Tensor, durations: Tensor, events: Tensor) -> Tensor:
    """cox_ph_loss, modules.py:265-305. No max-subtraction before exp (as in the reference); zero when there is
    no valid row or the result is not finite."""
    ok = ~torch.isnan(durations) & ~torch.isnan(events)
    if int(ok.sum()) == 0:
        return torch.zeros((), dtype=outputs.dtype, requires_grad=True)
    o, e, t = outputs[ok], events[ok], durations[ok]
    order = torch.argsort(t, descending=True)
    log_risk = torch.log(torch.cumsum(torch.exp(o)[order], dim=0))
    e_sorted = e[order]
    picked = e_sorted == 1
    total = -(o[order][picked].sum() - log_risk[picked].sum()) / e.sum()
    if not bool(torch.isfinite(total)):
        return torch.zeros((), dtype=outputs.dtype, requires_grad=True)
    return total


def supervised_loss(kind: str, y: Tensor, y_hat: Tensor) -> Tensor:
    """compute_loss, direct_pred.py:146-190 (identical copies in the other three models)."""
    if kind == "numerical":
        ok = ~torch.isnan(y)
        if int(ok.sum()) == 0:
            return torch.zeros((), dtype=y_hat.dtype, requires_grad=True)
        return F.mse_loss(torch.flatten(y_hat[ok]), y[ok].float())
    ok = (y != -1) & ~torch.isnan(y)
    if int(ok.sum()) == 0:
        return torch.zeros((), dtype=y_hat.dtype, requires_grad=True)
    return F.cross_entropy(y_hat[ok], y[ok].long())


def total_loss(P, losses: Dict[str, Tensor], weighting: bool) -> Tensor:
    """compute_total_loss, direct_pred.py:192-223: Kendall uncertainty weighting iff enabled and > 1 loss."""
    if weighting and len(losses) > 1:
        return sum(torch.exp(-P["log_vars." + k]) * v + P["log_vars." + k] for k, v in losses.items())
    return sum(losses.values())


def gaussian_kernel(x: Tensor, y: Tensor) -> Tensor:
    """compute_kernel, supervised_vae.py:494-513: exp(-mean_k (x_ik - y_jk)^2 / dim). Evaluated through the Gram
    identity in float64 so the oracle does not materialise the reference's [x, y, dim] tensor."""
    dim = x.shape[1]
    xd, yd = x.double(), y.double()
    d2 = (xd * xd).sum(1)[:, None] + (yd * yd).sum(1)[None, :] - 2.0 * xd @ yd.T
    return torch.exp(-(d2.clamp_min(0.0) / dim) / float(dim)).to(x.dtype)


def gaussian_kernel_literal(x: Tensor, y: Tensor) -> Tensor:
    """The reference's literal broadcast formulation (small sizes only); used to pin `gaussian_kernel`."""
    dim = x.shape[1]
    diff = x[:, None, :] - y[None, :, :]
    return torch.exp(-(diff.pow(2).mean(2) / float(dim)))


def mmd(prior: Tensor, z: Tensor, literal: bool = False) -> Tensor:
    """compute_mmd, supervised_vae.py:515-530."""
    k = gaussian_kernel_literal if literal else gaussian_kernel
    return k(prior, prior).mean() + k(z, z).mean() - 2 * k(prior, z).mean()


def triplet(anchor: Tensor, positive: Tensor, negative: Tensor, margin: float = 1.0) -> Tensor:
    """triplet_loss, triplet_encoder.py:178-194."""
    dp = (anchor - positive).pow(2).sum(1)
    dn = (anchor - negative).pow(2).sum(1)
    return torch.relu(dp - dn + margin).mean()


# ----------------------------------------------------------------------------------------------------
# GCNConv (torch_geometric, restated; A6 of SURVEY.md)
# ----------------------------------------------------------------------------------------------------
def gcn_norm(edge_index: Tensor, num_nodes: int, dtype=torch.float32) -> Tuple[Tensor, Tensor, Tensor]:
    """PyG gcn_norm(add_self_loops=True, improved=False, flow='source_to_target') with edge_weight=None: the edge list
    goes through add_remaining_self_loops(fill_value=1) -- every self loop present in the input is REMOVED and exactly one
    self loop per node (weight 1: the unit weight of an existing loop, or the fill value) is appended, so duplicated
    self loops collapse to one while duplicated ordinary edges stay (torch_geometric/utils/loop.py); then
    deg = in-degree over targets (edge_index[1]) on the directed list, w_e = deg^-1/2[src] * deg^-1/2[dst], inf -> 0."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    keep = src != dst
    loops = torch.arange(num_nodes)
    src = torch.cat([src[keep], loops])
    dst = torch.cat([dst[keep], loops])
    w = torch.ones(src.numel(), dtype=dtype)
    deg = torch.zeros(num_nodes, dtype=dtype).scatter_add_(0, dst, w)
    dinv = deg.pow(-0.5)
    dinv[torch.isinf(dinv)] = 0
    return src, dst, dinv[src] * w * dinv[dst]


def gcn_conv(P, prefix: str, x: Tensor, edge_index: Tensor) -> Tensor:
    """GCNConv(in, out) on batched dense input x [B, N, F] with a shared edge_index (node_dim = -2):
    out[b, v] = sum_{(u -> v)} w_uv * (x[b, u] @ W^T) + bias. Parameters: `lin.weight` [out, in], `bias` [out]."""
    n = x.shape[-2]
    src, dst, w = gcn_norm(edge_index, n, x.dtype)
    h = x @ P[prefix + ".lin.weight"].T
    msg = h[..., src, :] * w[:, None]
    out = torch.zeros_like(h).index_add_(-2, dst, msg)
    return out + P[prefix + ".bias"]


def graph_conv(P, prefix: str, x: Tensor, edge_index: Tensor) -> Tensor:
    """torch_geometric.nn.GraphConv(in, out, aggr='add', bias=True) as flexGCN builds it (modules.py:225, :239-246) on
    batched dense x [B, N, F]:  out[b, v] = lin_rel( sum_{(u -> v)} x[b, u] ) + lin_root( x[b, v] ). The edge list is used
    as given (directed, no self loops added, no edge weights). Parameters: `lin_rel.weight` [out, in], `lin_rel.bias`,
    `lin_root.weight` (no bias). Restated from PyG's published semantics (PyG is not installable here: parity against
    PyG itself is unpinned, as for gcn_conv)."""
    src, dst = edge_index[0].long(), edge_index[1].long()
    agg = torch.zeros_like(x).index_add_(-2, dst, x[..., src, :])
    return (F.linear(agg, P[prefix + ".lin_rel.weight"], P[prefix + ".lin_rel.bias"])
            + F.linear(x, P[prefix + ".lin_root.weight"]))


def sage_conv(P, prefix: str, x: Tensor, edge_index: Tensor) -> Tensor:
    """torch_geometric.nn.SAGEConv(in, out, aggr='mean', root_weight=True, bias=True, normalize=False, project=False):
    out[b, v] = lin_l( mean_{(u -> v)} x[b, u] ) + lin_r( x[b, v] ); the mean over an empty neighbourhood is 0.
    Parameters: `lin_l.weight`, `lin_l.bias`, `lin_r.weight` (no bias). Same pinning status as graph_conv."""
    n = x.shape[-2]
    src, dst = edge_index[0].long(), edge_index[1].long()
    deg = torch.zeros(n, dtype=x.dtype).index_add_(0, dst, torch.ones(dst.numel(), dtype=x.dtype))
    agg = torch.zeros_like(x).index_add_(-2, dst, x[..., src, :]) / deg.clamp_min(1.0)[:, None]
    return F.linear(agg, P[prefix + ".lin_l.weight"], P[prefix + ".lin_l.bias"]) + F.linear(x, P[prefix + ".lin_r.weight"])


CONVS = {"GCN": gcn_conv, "GC": graph_conv, "SAGE": sage_conv}

ACTS = {"relu": torch.relu, "sigmoid": torch.sigmoid, "leakyrelu": lambda t: F.leaky_relu(t, 0.01),
        "tanh": torch.tanh, "gelu": F.gelu}


def flexgcn(P, prefix: str, x: Tensor, edge_index: Tensor, num_convs: int, act: str, train: bool, noise: Noise,
            p_drop: float = 0.2, conv: str = "GCN") -> Tensor:
    """flexGCN.forward, modules.py:252-262 with conv in {'GCN', 'GC', 'SAGE'} (GAT is not restated)."""
    for k in range(num_convs):
        x = CONVS[conv](P, f"{prefix}.convs.{k}", x, edge_index)
        x = batchnorm(P, f"{prefix}.bns.{k}", x.reshape(-1, x.shape[2]), train).view_as(x)
        x = ACTS[act](x)
        if train:
            x = x * noise.dropout_mask(f"{prefix}.dropout.{k}", x, p_drop) / (1.0 - p_drop)
    return F.linear(x.reshape(x.shape[0], -1), P[prefix + ".fc.weight"], P[prefix + ".fc.bias"])


# ----------------------------------------------------------------------------------------------------
# model specification + parameter construction
# ----------------------------------------------------------------------------------------------------
@dataclass
class Spec:
    """Everything the reference constructors derive from (config, dataset, variables)."""
    model: str                       # "DirectPred" | "supervised_vae" | "MultiTripletNetwork" | "GNN"
    input_dims: List[int]            # features per layer (GNN: [node_feature_count])
    latent_dim: int
    hidden_dim_factor: float = 0.5
    supervisor_hidden_dim: int = 32
    variables: List[str] = field(default_factory=list)       # target (+ batch) variables, surv event var last
    variable_types: Dict[str, str] = field(default_factory=dict)
    num_classes: Dict[str, int] = field(default_factory=dict)  # categorical vars only
    surv_event_var: Optional[str] = None
    surv_time_var: Optional[str] = None
    use_loss_weighting: bool = True
    # GNN only
    node_count: int = 0
    node_embedding_dim: int = 0
    num_convs: int = 2
    activation: str = "relu"
    # CrossModalPred only: indices (into input_dims / the batch's layers) of the encoded and of the reconstructed layers
    in_idx: Optional[List[int]] = None
    out_idx: Optional[List[int]] = None
    conv: str = "GCN"                # flexGCN convolution: "GCN" | "GC" (GraphConv, the CLI default) | "SAGE"
    mmd_literal: bool = False        # evaluate the MMD kernels the reference's literal [x, y, dim] way (timing runs)

    def hidden(self, i: int) -> int:
        h = int(self.input_dims[i] * self.hidden_dim_factor)
        return max(h, 2)  # MLP clamps inside (modules.py:124); svae clamps at the call site (supervised_vae.py:92)

    def head_out(self, var: str) -> int:
        return 1 if self.variable_types[var] == "numerical" else self.num_classes[var]

    def loss_names(self) -> List[str]:
        extra = {"supervised_vae": ["mmd_loss"], "CrossModalPred": ["mmd_loss"],
                 "MultiTripletNetwork": ["triplet_loss"]}.get(self.model, [])
        return list(self.variables) + extra


def _linear(P, name: str, fan_in: int, fan_out: int, bias: bool = True, xavier: bool = False) -> None:
    lin = torch.nn.Linear(fan_in, fan_out, bias=bias)   # default init = kaiming_uniform(a=sqrt 5) + uniform bias
    if xavier:
        torch.nn.init.xavier_uniform_(lin.weight)
    P[name + ".weight"] = lin.weight.detach().clone()
    if bias:
        P[name + ".bias"] = lin.bias.detach().clone()


def _bn(P, name: str, n: int) -> None:
    P[name + ".weight"] = torch.ones(n)
    P[name + ".bias"] = torch.zeros(n)
    P[name + ".running_mean"] = torch.zeros(n)
    P[name + ".running_var"] = torch.ones(n)
    P[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def _mlp_params(P, prefix: str, d: int, h: int, o: int) -> None:
    h = max(h, 2)
    _linear(P, prefix + ".layer_1", d, h)
    _linear(P, prefix + ".layer_out", h, o, bias=(o > 1))
    _bn(P, prefix + ".batchnorm", h)


def init_params(spec: Spec) -> Dict[str, Tensor]:
    """Build a parameter/buffer mapping with the reference's state_dict keys, consuming torch's global RNG in
    the same order as the reference constructors (direct_pred.py:58-105, supervised_vae.py:77-130,
    triplet_encoder.py:79-123, gnn_early.py:105-140)."""
    P: Dict[str, Tensor] = {}
    if spec.use_loss_weighting:
        for name in spec.loss_names():
            P["log_vars." + name] = torch.zeros(1)
    L = spec.latent_dim
    n = len(spec.input_dims)
    if spec.model in ("DirectPred", "MultiTripletNetwork"):
        for i, d in enumerate(spec.input_dims):
            _mlp_params(P, f"encoders.{i}", d, int(d * spec.hidden_dim_factor), L)
        if n > 1:
            _linear(P, "fusion_block", L * n, L)
    elif spec.model == "supervised_vae":
        for i, d in enumerate(spec.input_dims):
            h = spec.hidden(i)
            _linear(P, f"encoders.{i}.hidden_layers.0", d, h, xavier=True)
            _bn(P, f"encoders.{i}.hidden_layers.2", h)
            _linear(P, f"encoders.{i}.FC_mean", h, L, xavier=True)
            _linear(P, f"encoders.{i}.FC_var", h, L, xavier=True)
        _linear(P, "FC_mean", n * L, L)
        _linear(P, "FC_log_var", n * L, L)
        for i, d in enumerate(spec.input_dims):
            h = spec.hidden(i)
            _linear(P, f"decoders.{i}.hidden_layers.0", L, h, xavier=True)
            _bn(P, f"decoders.{i}.hidden_layers.2", h)
            _linear(P, f"decoders.{i}.FC_output", h, d, xavier=True)
    elif spec.model == "CrossModalPred":
        # crossmodal_pred.py:80-121: Encoders over the input layers, FC_mean / FC_log_var, Decoders into the output
        # layers; hidden width int(d * factor) (no clamp)
        ins = spec.in_idx if spec.in_idx is not None else list(range(len(spec.input_dims)))
        outs = spec.out_idx if spec.out_idx is not None else list(range(len(spec.input_dims)))
        for i, li in enumerate(ins):
            d = spec.input_dims[li]
            h = int(d * spec.hidden_dim_factor)
            _linear(P, f"encoders.{i}.hidden_layers.0", d, h, xavier=True)
            _bn(P, f"encoders.{i}.hidden_layers.2", h)
            _linear(P, f"encoders.{i}.FC_mean", h, L, xavier=True)
            _linear(P, f"encoders.{i}.FC_var", h, L, xavier=True)
        _linear(P, "FC_mean", len(ins) * L, L)
        _linear(P, "FC_log_var", len(ins) * L, L)
        for i, li in enumerate(outs):
            d = spec.input_dims[li]
            h = int(d * spec.hidden_dim_factor)
            _linear(P, f"decoders.{i}.hidden_layers.0", L, h, xavier=True)
            _bn(P, f"decoders.{i}.hidden_layers.2", h)
            _linear(P, f"decoders.{i}.FC_output", h, d, xavier=True)
    elif spec.model == "GNN":
        emb = spec.node_embedding_dim
        fin = spec.input_dims[0]
        for k in range(spec.num_convs):
            # PyG GCNConv: lin = Linear(in, out, bias=False, weight_initializer='glorot'); bias = zeros(out)
            cin = fin if k == 0 else emb
            if spec.conv == "GCN":
                w = torch.empty(emb, cin)
                torch.nn.init.xavier_uniform_(w)
                P[f"encoders.0.convs.{k}.bias"] = torch.zeros(emb)
                P[f"encoders.0.convs.{k}.lin.weight"] = w
            else:
                # PyG GraphConv: lin_rel (bias) + lin_root (no bias); SAGEConv: lin_l (bias) + lin_r (no bias); PyG's
                # Linear default init = kaiming_uniform(a = sqrt(5)) / uniform(+-1/sqrt(fan_in)), i.e. nn.Linear's
                a, r = ("lin_rel", "lin_root") if spec.conv == "GC" else ("lin_l", "lin_r")
                _linear(P, f"encoders.0.convs.{k}.{a}", cin, emb)
                _linear(P, f"encoders.0.convs.{k}.{r}", cin, emb, bias=False)
            _bn(P, f"encoders.0.bns.{k}", emb)
        _linear(P, "encoders.0.fc", emb * spec.node_count, L)
    else:
        raise ValueError(spec.model)
    for var in spec.variables:
        _mlp_params(P, f"MLPs.{var}", L, spec.supervisor_hidden_dim, spec.head_out(var))
    return P


def trainable(P: Dict[str, Tensor]) -> List[str]:
    return [k for k in P if not (k.endswith("running_mean") or k.endswith("running_var")
                                 or k.endswith("num_batches_tracked"))]


# ----------------------------------------------------------------------------------------------------
# model forwards / training losses
# ----------------------------------------------------------------------------------------------------
def _fused_embedding(P, spec: Spec, x_list, train, noise, tag="") -> Tensor:
    """direct_pred.py:118-128 / triplet_encoder.py:125-138."""
    embs = [mlp(P, f"encoders.{i}", x, train, noise, site=tag) for i, x in enumerate(x_list)]
    cat = torch.cat(embs, dim=1)
    if len(x_list) > 1:
        return F.linear(cat, P["fusion_block.weight"], P["fusion_block.bias"])
    return cat


def _heads(P, spec: Spec, emb: Tensor, train, noise) -> Dict[str, Tensor]:
    return {var: mlp(P, f"MLPs.{var}", emb, train, noise) for var in spec.variables}


def _head_losses(spec: Spec, outputs, y_dict) -> Dict[str, Tensor]:
    """The per-variable loop shared by all training_steps (direct_pred.py:243-253)."""
    out = {}
    for var in spec.variables:
        if var == spec.surv_event_var:
            out[var] = cox_ph(outputs[var], y_dict[spec.surv_time_var], y_dict[spec.surv_event_var])
        else:
            out[var] = supervised_loss(spec.variable_types[var], y_dict[var], outputs[var])
    return out


def forward(P, spec: Spec, batch, train: bool, noise: Noise, edge_index: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Model forward + every loss term. `batch` is what the reference's training_step receives:
       DirectPred / supervised_vae: (dat: {layer: [B, d]}, y_dict, samples)
       MultiTripletNetwork:         (anchor, positive, negative, y_dict)
       GNN:                         (x [B, N, F], y_dict, samples) plus the model-level `edge_index` [2, E]
    Returns a dict with 'outputs' (per var), 'embedding', every loss term, 'total' (training objective) and
    'val_total' (validation objective = unweighted sum, direct_pred.py:290)."""
    res: Dict[str, Tensor] = {}
    if spec.model == "DirectPred":
        dat, y_dict = batch[0], batch[1]
        emb = _fused_embedding(P, spec, list(dat.values()), train, noise)
        outputs = _heads(P, spec, emb, train, noise)
        losses = _head_losses(spec, outputs, y_dict)
    elif spec.model == "MultiTripletNetwork":
        anchor, pos, neg, y_dict = batch
        emb = _fused_embedding(P, spec, list(anchor.values()), train, noise, "anchor.")
        emb_p = _fused_embedding(P, spec, list(pos.values()), train, noise, "positive.")
        emb_n = _fused_embedding(P, spec, list(neg.values()), train, noise, "negative.")
        outputs = _heads(P, spec, emb, train, noise)
        losses = {"triplet_loss": triplet(emb, emb_p, emb_n)}   # triplet_encoder.py:292-296 (first key)
        losses.update(_head_losses(spec, outputs, y_dict))
        res["embedding_positive"], res["embedding_negative"] = emb_p, emb_n
    elif spec.model == "supervised_vae":
        dat, y_dict = batch[0], batch[1]
        x_list = list(dat.values())
        means, logvars = zip(*[vae_encoder(P, f"encoders.{i}", x, train) for i, x in enumerate(x_list)])
        mean = F.linear(torch.cat(means, 1), P["FC_mean.weight"], P["FC_mean.bias"])
        log_var = F.linear(torch.cat(logvars, 1), P["FC_log_var.weight"], P["FC_log_var.bias"])
        z = mean + log_var * noise.normal("epsilon", log_var.shape, log_var)     # supervised_vae.py:198-200
        x_hat = [vae_decoder(P, f"decoders.{i}", z, train) for i in range(len(x_list))]
        outputs = _heads(P, spec, z, train, noise)
        per_layer = []
        for i, x in enumerate(x_list):                                           # MMD_loss, :532-550
            prior = noise.normal(f"mmd_prior.{i}", (200, z.shape[1]), z)
            per_layer.append(mmd(prior, z, literal=spec.mmd_literal) + (x_hat[i] - x).pow(2).mean())
        losses = {"mmd_loss": torch.mean(torch.stack(per_layer))}
        losses.update(_head_losses(spec, outputs, y_dict))
        emb = z
        res["mean"], res["log_var"], res["x_hat"] = mean, log_var, x_hat
    elif spec.model == "CrossModalPred":
        dat, y_dict = batch[0], batch[1]
        layers = list(dat.values())
        ins = spec.in_idx if spec.in_idx is not None else list(range(len(layers)))
        outs = spec.out_idx if spec.out_idx is not None else list(range(len(layers)))
        x_in, x_out = [layers[i] for i in ins], [layers[i] for i in outs]
        means, logvars = zip(*[vae_encoder(P, f"encoders.{i}", x, train) for i, x in enumerate(x_in)])
        mean = F.linear(torch.cat(means, 1), P["FC_mean.weight"], P["FC_mean.bias"])
        log_var = F.linear(torch.cat(logvars, 1), P["FC_log_var.weight"], P["FC_log_var.bias"])
        z = mean + log_var * noise.normal("epsilon", log_var.shape, log_var)     # crossmodal_pred.py:189-202
        x_hat = [vae_decoder(P, f"decoders.{i}", z, train) for i in range(len(x_out))]
        outputs = _heads(P, spec, z, train, noise)
        per_layer = []
        for i, x in enumerate(x_out):                                            # :321-328, MMD_loss per OUTPUT layer
            prior = noise.normal(f"mmd_prior.{i}", (200, z.shape[1]), z)
            per_layer.append(mmd(prior, z, literal=spec.mmd_literal) + (x_hat[i] - x).pow(2).mean())
        losses = {"mmd_loss": torch.mean(torch.stack(per_layer))}
        losses.update(_head_losses(spec, outputs, y_dict))
        emb = z
        res["mean"], res["log_var"], res["x_hat"] = mean, log_var, x_hat
    elif spec.model == "GNN":
        x, y_dict = batch[0], batch[1]
        emb = flexgcn(P, "encoders.0", x, edge_index, spec.num_convs, spec.activation, train, noise, conv=spec.conv)
        outputs = _heads(P, spec, emb, train, noise)
        losses = _head_losses(spec, outputs, y_dict)
    else:
        raise ValueError(spec.model)
    res["outputs"] = outputs
    res["embedding"] = emb
    res["losses"] = losses
    res["total"] = total_loss(P, losses, spec.use_loss_weighting)
    res["val_total"] = sum(losses.values())
    return res


# ----------------------------------------------------------------------------------------------------
# step policy (Lightning's automatic optimisation as configured in flexynesis/main.py:212-225)
# ----------------------------------------------------------------------------------------------------
class Trainer:
    """zero_grad -> training_step -> backward -> clip_grad_norm_(params, 1.0) -> Adam.step  (A7 of SURVEY.md)."""

    def __init__(self, P: Dict[str, Tensor], spec: Spec, lr: float, clip: float = 1.0,
                 edge_index: Optional[Tensor] = None):
        self.P, self.spec, self.clip, self.edge_index = P, spec, clip, edge_index
        self.names = trainable(P)
        for k in self.names:
            P[k].requires_grad_(True)
        self.opt = torch.optim.Adam([P[k] for k in self.names], lr=lr)   # direct_pred.py:143

    def step(self, batch, noise: Optional[Noise] = None) -> Dict[str, Tensor]:
        noise = noise or Noise()
        self.opt.zero_grad(set_to_none=True)
        res = forward(self.P, self.spec, batch, True, noise, self.edge_index)
        res["total"].backward()
        res["grads"] = {k: (None if self.P[k].grad is None else self.P[k].grad.detach().clone()) for k in self.names}
        res["grad_norm"] = torch.nn.utils.clip_grad_norm_([self.P[k] for k in self.names], self.clip)
        self.opt.step()
        return res


# ----------------------------------------------------------------------------------------------------
# synthetic data of SURVEY.md section 8d
# ----------------------------------------------------------------------------------------------------
def synthetic_batch(spec: Spec, n: int, seed: int = 0, missing: float = 0.05):
    """Seeded synthetic multi-omics batch: X_i = randn(n, d_i); numerical y = randn with `missing` NaN;
    categorical = randint(C) with `missing` NaN (every class present); survival t = 100*rand, e = rand > 0.3."""
    g = torch.Generator().manual_seed(seed)
    dat = {f"layer{i}": torch.randn(n, d, generator=g) for i, d in enumerate(spec.input_dims)}
    y: Dict[str, Tensor] = {}
    for var in spec.variables:
        if var == spec.surv_event_var:
            y[spec.surv_time_var] = 100.0 * torch.rand(n, generator=g)
            y[var] = (torch.rand(n, generator=g) > 0.3).float()
            continue
        if spec.variable_types[var] == "numerical":
            v = torch.randn(n, generator=g)
        else:
            c = spec.num_classes[var]
            v = torch.randint(0, c, (n,), generator=g).float()
            v[:c] = torch.arange(c).float()
        drop = torch.rand(n, generator=g) < missing
        if spec.variable_types[var] != "numerical":
            drop[: spec.num_classes[var]] = False
        v[drop] = float("nan")
        y[var] = v
    return dat, y


def synthetic_graph(num_nodes: int, num_edges: int, seed: int = 0) -> Tensor:
    """`num_edges` distinct unordered gene pairs, random orientation, stored once as protein1 -> protein2
    (the reference's MultiOmicDatasetNW keeps user graphs directed as given, data.py:1194-1207)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    seen = set()
    while len(seen) < num_edges:
        a, b = rng.integers(0, num_nodes, 2)
        if a == b:
            continue
        seen.add((min(a, b), max(a, b)))
    pairs = np.array(sorted(seen))
    flip = rng.random(len(pairs)) < 0.5
    pairs[flip] = pairs[flip][:, ::-1]
    return torch.from_numpy(pairs.T.copy()).long()


# ----------------------------------------------------------------------------------------------------
# attribution (DirectPred.compute_feature_importance, direct_pred.py:432-590)
# ----------------------------------------------------------------------------------------------------
def attribution_path(method: str, n: int, generator=None):
    """captum's path from the all-zero baseline to the input, restated from its published algorithm (captum is not
    installable here, so this is unpinned against captum itself): IntegratedGradients(method='gausslegendre',
    n_steps=n) evaluates the integrand at alphas = (1 + x_k) / 2 with step sizes w_k / 2, (x_k, w_k) the
    Gauss-Legendre rule of order n; GradientShap with zero baselines and stdevs = 0 draws alphas ~ U(0, 1) and
    averages."""
    import numpy as np
    if method == "IntegratedGradients":
        x, w = np.polynomial.legendre.leggauss(int(n))
        return list(0.5 * (1.0 + x)), list(0.5 * w)
    a = torch.rand(int(n), generator=generator).tolist()
    return a, [1.0 / n] * int(n)


def feature_importance_sums(P, spec: Spec, dat: Dict[str, Tensor], var: str, alphas, weights) -> List[List[Tensor]]:
    """[class][layer] -> sum over the batch of |x * sum_k w_k d out[var][:, class] / d x (alpha_k x)|, the quantity the
    reference accumulates per batch (`a.abs().sum(dim=1)`, direct_pred.py:527, :553) before dividing by the number of
    samples. Eval mode (running BatchNorm statistics, no dropout), torch autograd, DirectPred only."""
    assert spec.model == "DirectPred"
    xs = list(dat.values())
    C = spec.head_out(var)
    out = []
    for cls in range(C):
        G = [torch.zeros_like(x) for x in xs]
        for al, w in zip(alphas, weights):
            xk = [(x * float(al)).detach().requires_grad_(True) for x in xs]
            Pc = {k: v.detach().clone() for k, v in P.items()}
            emb = _fused_embedding(Pc, spec, xk, False, Noise())
            o = _heads(Pc, spec, emb, False, Noise())[var]
            o[:, cls].sum().backward()
            for g, x in zip(G, xk):
                g += float(w) * x.grad
        out.append([(x * g).abs().sum(0) for x, g in zip(xs, G)])
    return out
