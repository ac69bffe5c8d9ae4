"""CPU cross-check of the two independent torch formulations of the flexGCN convolutions: the oracle's restatement
(oracle/restatement.py, what the GPU kernels are tested against) and the drop-in containers (flexynesis_b200/containers.py,
what CPU-side inference of a saved model runs). torch_geometric is not installable here, so neither is pinned against PyG
itself; they are pinned against each other and against a literal per-edge loop."""
import pytest
import torch

from oracle.restatement import Spec, graph_conv, gcn_conv, init_params, sage_conv, synthetic_graph

VT = {"y": "numerical"}


def _graph(n, e, seed):
    ei = synthetic_graph(n, e, seed)
    extra = torch.tensor([[0, 1, 1, 5, 5], [0, 2, 2, 5, 5]])    # self loops on 0 and (twice) 5, a duplicated edge 1 -> 2
    return torch.cat([ei, extra], 1)


def _literal(kind, P, prefix, x, ei):
    """per-edge python loop straight from the published definitions"""
    B, N, F = x.shape
    agg = torch.zeros_like(x)
    deg = torch.zeros(N)
    for s, d in ei.t().tolist():
        agg[:, d] += x[:, s]
        deg[d] += 1
    if kind == "SAGE":
        agg = agg / deg.clamp_min(1)[None, :, None]
        a, r = "lin_l", "lin_r"
    else:
        a, r = "lin_rel", "lin_root"
    return agg @ P[f"{prefix}.{a}.weight"].T + P[f"{prefix}.{a}.bias"] + x @ P[f"{prefix}.{r}.weight"].T


@pytest.mark.parametrize("kind", ["GC", "SAGE"])
def test_root_weight_convs_agree(kind):
    from flexynesis_b200.containers import GraphConv, SAGEConv
    torch.manual_seed(0)
    spec = Spec(model="GNN", input_dims=[3], latent_dim=8, variables=["y"], variable_types=VT, node_count=40,
                node_embedding_dim=12, num_convs=2, conv=kind)
    P = init_params(spec)
    ei = _graph(40, 90, 1)
    x = torch.randn(5, 40, 3)
    fn = graph_conv if kind == "GC" else sage_conv
    want = fn(P, "encoders.0.convs.0", x, ei)
    lit = _literal(kind, P, "encoders.0.convs.0", x, ei)
    assert torch.allclose(want, lit, atol=1e-5)
    mod = (GraphConv if kind == "GC" else SAGEConv)(3, 12)
    mod.load_state_dict({k[len("encoders.0.convs.0."):]: v for k, v in P.items() if k.startswith("encoders.0.convs.0.")},
                        strict=True)
    assert torch.allclose(mod(x, ei), want, atol=1e-5)


def test_gcn_conv_follows_add_remaining_self_loops():
    """GCNConv as torch_geometric defines it (gcn_norm -> add_remaining_self_loops -> symmetric normalisation -> sum over
    incoming edges), written out as a per-edge loop: input self loops -- also duplicated ones -- collapse to ONE unit loop
    per node, duplicated ordinary edges count twice, nodes without a loop get one. Checks the oracle's restatement, the
    drop-in container and the engine's CSR operator against it."""
    from flexynesis_b200.containers import GCNConv
    from flexynesis_b200.engine import build_gcn_csr
    torch.manual_seed(1)
    n, fin, emb = 30, 4, 6
    ei = _graph(n, 60, 2)
    x = torch.randn(3, n, fin)
    W, b = torch.randn(emb, fin), torch.randn(emb)
    # literal definition
    edges = [(s, d) for s, d in ei.t().tolist() if s != d] + [(v, v) for v in range(n)]
    deg = [0.0] * n
    for s, d in edges:
        deg[d] += 1.0
    h = x @ W.T
    want = torch.zeros(3, n, emb)
    for s, d in edges:
        want[:, d] += h[:, s] / (deg[s] ** 0.5 * deg[d] ** 0.5)
    want = want + b
    P = {"c.lin.weight": W, "c.bias": b}
    assert torch.allclose(gcn_conv(P, "c", x, ei), want, atol=1e-5)
    mod = GCNConv(fin, emb)
    mod.load_state_dict({"lin.weight": W, "bias": b})
    assert torch.allclose(mod(x, ei), want, atol=1e-5)
    (rp, col, w), _ = build_gcn_csr(ei, n, "cpu", "GCN")
    got = torch.zeros(3, n, emb)
    for v in range(n):
        for e in range(int(rp[v]), int(rp[v + 1])):
            got[:, v] += float(w[e]) * h[:, int(col[e])]
    assert torch.allclose(got + b, want, atol=1e-5)
    assert int(rp[-1]) == len(edges)


def test_state_dict_keys_follow_pyg():
    from flexynesis_b200.containers import flexGCN
    keys = {"GCN": {"convs.0.bias", "convs.0.lin.weight"},
            "GC": {"convs.0.lin_rel.weight", "convs.0.lin_rel.bias", "convs.0.lin_root.weight"},
            "SAGE": {"convs.0.lin_l.weight", "convs.0.lin_l.bias", "convs.0.lin_r.weight"}}
    for conv, want in keys.items():
        m = flexGCN(10, 2, 8, 4, num_convs=1, conv=conv)
        got = {k for k in m.state_dict() if k.startswith("convs.")}
        assert got == want, (conv, got)
    with pytest.raises(ValueError):
        flexGCN(10, 2, 8, 4, conv=None)
    with pytest.raises(NotImplementedError):
        flexGCN(10, 2, 8, 4, conv="GAT")


@pytest.mark.parametrize("conv", ["GCN", "GC", "SAGE"])
def test_engine_csr_matches_the_dense_operator(conv):
    """engine.build_gcn_csr (host logic that feeds the aggregate kernels): the CSR by destination and the CSR by source
    must both encode the dense operator A[dst, src] = sum of edge weights that the torch containers apply, including
    duplicate edges, self loops, isolated nodes and (GCN) the added self loops."""
    from flexynesis_b200.containers import CONVS
    from flexynesis_b200.engine import build_gcn_csr
    n = 23
    ei = _graph(n, 40, 3)
    ei = ei[:, (ei[0] < n) & (ei[1] < n)]
    (rp_in, col_in, w_in), (rp_out, col_out, w_out) = build_gcn_csr(ei, n, "cpu", conv)
    if conv == "GCN":
        src, dst, w = CONVS[conv].normalized_edges(ei, n)
    else:
        src, dst, w = CONVS[conv].edge_weights(ei, n)
    dense = torch.zeros(n, n).index_put_((dst, src), w, accumulate=True)
    a_in, a_out = torch.zeros(n, n), torch.zeros(n, n)
    for v in range(n):
        for e in range(int(rp_in[v]), int(rp_in[v + 1])):
            a_in[v, int(col_in[e])] += float(w_in[e])
        for e in range(int(rp_out[v]), int(rp_out[v + 1])):
            a_out[int(col_out[e]), v] += float(w_out[e])
    assert torch.allclose(a_in, dense, atol=1e-6) and torch.allclose(a_out, dense, atol=1e-6)
    assert int(rp_in[-1]) == int(rp_out[-1]) == src.numel()
