"""Parameter containers with the reference's module names, shapes, init and state_dict keys.

These nn.Modules own the parameters (so optimizers, deepcopy, pickling, .to() and strict load_state_dict of
reference checkpoints all work -- SURVEY.md section 8b), but they are NOT the training path: on a CUDA device the
engine (flexynesis_b200.engine) reads their tensors through raw pointers and runs the hand-written kernels. Their
`forward` methods are the differentiable torch formulation kept for the captum adaptor (`forward_target`, which
needs d out / d input) and for CPU-resident inference of a saved model.

Key layout (checked against tests/golden/*.pt recorded from the reference):
  MLP       layer_1.{weight,bias}  batchnorm.*  layer_out.{weight[,bias]}        flexynesis/modules.py:106-150
  Encoder   hidden_layers.{0,2}.*  FC_mean.*  FC_var.*                            flexynesis/modules.py:10-57
  Decoder   hidden_layers.{0,2}.*  FC_output.*                                    flexynesis/modules.py:60-103
  flexGCN   convs.<k>.{bias,lin.weight}  bns.<k>.*  fc.*                          flexynesis/modules.py:153-262
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

ACTIVATIONS = {"relu": 1, "leakyrelu": 2, "sigmoid": 3, "tanh": 4, "gelu": 5}   # codes of fxn_bn_act_fwd
_TORCH_ACT = {"relu": torch.relu, "leakyrelu": lambda t: F.leaky_relu(t, 0.01), "sigmoid": torch.sigmoid,
              "tanh": torch.tanh, "gelu": F.gelu}


def _xavier_linear(fan_in: int, fan_out: int) -> nn.Linear:
    lin = nn.Linear(fan_in, fan_out)
    nn.init.xavier_uniform_(lin.weight)
    return lin


class MLP(nn.Module):
    """Linear -> BatchNorm1d -> ReLU -> Dropout(0.1) -> Linear; encoder of DirectPred / MultiTripletNetwork and
    every supervisor head. hidden_dim is clamped to >= 2; a single-output head has no output bias."""

    p_drop = 0.1

    def __init__(self, input_dim: int, hidden_dim: int, output_dim: int):
        super().__init__()
        hidden = max(int(hidden_dim), 2)
        # construction order matters: it fixes the RNG draws so that the same seed gives the reference's init
        self.layer_1 = nn.Linear(input_dim, hidden)
        self.layer_out = nn.Linear(hidden, output_dim, bias=output_dim > 1)
        self.relu, self.dropout = nn.ReLU(), nn.Dropout(p=self.p_drop)
        self.batchnorm = nn.BatchNorm1d(hidden)

    def forward(self, x):
        return self.layer_out(self.dropout(self.relu(self.batchnorm(self.layer_1(x)))))


class _VAETrunk(nn.Module):
    """Linear -> LeakyReLU(0.2) -> BatchNorm1d as Sequential indices 0, 1, 2 (one hidden layer in every shipped
    model: supervised_vae.py:92, :111)."""

    def __init__(self, input_dim: int, hidden_dims):
        super().__init__()
        if len(hidden_dims) != 1:
            raise ValueError("the B200 engine implements the one-hidden-layer Encoder/Decoder every flexynesis model uses")
        self.act = nn.LeakyReLU(0.2)
        self.hidden_layers = nn.Sequential(_xavier_linear(input_dim, hidden_dims[0]), self.act,
                                           nn.BatchNorm1d(hidden_dims[0]))


class Encoder(_VAETrunk):
    def __init__(self, input_dim, hidden_dims, latent_dim):
        super().__init__(input_dim, hidden_dims)
        self.FC_mean = _xavier_linear(hidden_dims[-1], latent_dim)
        self.FC_var = _xavier_linear(hidden_dims[-1], latent_dim)

    def forward(self, x):
        h = self.hidden_layers(x)
        return self.FC_mean(h), self.FC_var(h)


class Decoder(_VAETrunk):
    def __init__(self, latent_dim, hidden_dims, output_dim):
        super().__init__(latent_dim, hidden_dims)
        self.FC_output = _xavier_linear(hidden_dims[-1], output_dim)

    def forward(self, z):
        return torch.sigmoid(self.FC_output(self.hidden_layers(z)))


class GCNConv(nn.Module):
    """Parameter container + torch formulation of torch_geometric.nn.GCNConv (un-vendored third party; algorithm
    restated from its published semantics, see oracle/restatement.py:gcn_conv and SURVEY.md A6): add remaining self
    loops, symmetric normalisation by in-degree on the directed edge list, aggregate lin(x) over incoming edges,
    add bias. State keys `lin.weight` [out, in] (glorot) and `bias` [out] (zeros) match PyG's."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.lin = nn.Linear(in_channels, out_channels, bias=False)
        nn.init.xavier_uniform_(self.lin.weight)
        self.bias = nn.Parameter(torch.zeros(out_channels))

    @staticmethod
    def normalized_edges(edge_index: torch.Tensor, num_nodes: int):
        # torch_geometric's add_remaining_self_loops: input self loops are dropped, then ONE unit loop per node is appended
        src, dst = edge_index[0].long(), edge_index[1].long()
        keep = src != dst
        loops = torch.arange(num_nodes, device=edge_index.device)
        src, dst = torch.cat([src[keep], loops]), torch.cat([dst[keep], loops])
        deg = torch.zeros(num_nodes, device=edge_index.device).scatter_add_(0, dst, torch.ones_like(dst, dtype=torch.float32))
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        return src, dst, dinv[src] * dinv[dst]

    def forward(self, x, edge_index):
        src, dst, w = self.normalized_edges(edge_index, x.shape[-2])
        h = self.lin(x)
        return torch.zeros_like(h).index_add_(-2, dst, h[..., src, :] * w[:, None]) + self.bias


class _RootConv(nn.Module):
    """Shared torch formulation of torch_geometric's GraphConv / SAGEConv on batched dense x [B, N, F] with one edge_index
    (un-vendored third party; algorithms restated from their published semantics, oracle/restatement.py:graph_conv /
    sage_conv): neighbour aggregate through one Linear with bias, the node's own features through a bias-free Linear."""
    MEAN = False
    NEIGH, ROOT = "lin_rel", "lin_root"

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        setattr(self, self.NEIGH, nn.Linear(in_channels, out_channels, bias=True))
        setattr(self, self.ROOT, nn.Linear(in_channels, out_channels, bias=False))

    @property
    def neigh(self) -> nn.Linear:
        return getattr(self, self.NEIGH)

    @property
    def root(self) -> nn.Linear:
        return getattr(self, self.ROOT)

    @classmethod
    def edge_weights(cls, edge_index: torch.Tensor, num_nodes: int):
        """(src, dst, w): the directed list as given; w = 1 (sum) or 1 / in-degree of the destination (mean)."""
        src, dst = edge_index[0].long(), edge_index[1].long()
        w = torch.ones(dst.numel(), dtype=torch.float32, device=edge_index.device)
        if cls.MEAN:
            deg = torch.zeros(num_nodes, device=edge_index.device).scatter_add_(0, dst, w)
            w = 1.0 / deg.clamp_min(1.0)[dst]
        return src, dst, w

    def forward(self, x, edge_index):
        src, dst, w = self.edge_weights(edge_index, x.shape[-2])
        agg = torch.zeros_like(x).index_add_(-2, dst, x[..., src, :] * w[:, None])
        return self.neigh(agg) + self.root(x)


class GraphConv(_RootConv):
    """GraphConv(in, out, aggr='add'): state keys lin_rel.{weight,bias}, lin_root.weight (PyG's)."""


class SAGEConv(_RootConv):
    """SAGEConv(in, out, aggr='mean', root_weight=True): state keys lin_l.{weight,bias}, lin_r.weight (PyG's)."""
    MEAN = True
    NEIGH, ROOT = "lin_l", "lin_r"


CONVS = {"GCN": GCNConv, "GC": GraphConv, "SAGE": SAGEConv}


class flexGCN(nn.Module):
    """num_convs x (conv -> BatchNorm1d over B*N rows -> act -> Dropout(0.2)) -> flatten -> Linear(N*emb, out), conv in
    {GCN, GC, SAGE} (flexynesis/modules.py:195-262; the reference's default is GC)."""

    def __init__(self, node_count, node_feature_count, node_embedding_dim, output_dim, num_convs=2, dropout_rate=0.2,
                 conv="GC", act="relu"):
        super().__init__()
        if act not in ACTIVATIONS:
            raise ValueError("Invalid activation function string. Choose from ", list(ACTIVATIONS))
        if conv not in ("GCN", "GAT", "SAGE", "GC"):
            raise ValueError("Unknown convolution type. Choose one of: ", ["GCN", "GAT", "SAGE", "GC"])
        if conv == "GAT":
            raise NotImplementedError("conv='GAT': the B200 engine implements GCN, GC (GraphConv) and SAGE; attention "
                                      "convolutions are not built (DESIGN.md, out of scope)")
        self.conv_name = conv
        self.act_name, self.dropout_rate = act, dropout_rate
        self.act = {"relu": nn.ReLU(), "sigmoid": nn.Sigmoid(), "leakyrelu": nn.LeakyReLU(), "tanh": nn.Tanh(),
                    "gelu": nn.GELU()}[act]
        self.convs, self.bns = nn.ModuleList(), nn.ModuleList()
        self.dropout = nn.Dropout(dropout_rate)
        for k in range(num_convs):
            self.convs.append(CONVS[conv](node_feature_count if k == 0 else node_embedding_dim, node_embedding_dim))
            self.bns.append(nn.BatchNorm1d(node_embedding_dim))
        self.fc = nn.Linear(node_embedding_dim * node_count, output_dim)

    def forward(self, x, edge_index):
        for conv, bn in zip(self.convs, self.bns):
            x = conv(x, edge_index)
            x = self.dropout(self.act(bn(x.reshape(-1, x.size(2))).view_as(x)))
        return self.fc(x.reshape(x.size(0), -1))
