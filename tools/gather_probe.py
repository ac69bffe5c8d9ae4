"""Times / profiles fxn_graph_gather alone at the config-4 shape (B=4096, N=2000, C=32, 20000 edges + self loops).
usage: python tools/gather_probe.py [reps]        (under ncu: -k regex:graph_gather -c 2)"""
import sys
import torch
sys.path.insert(0, ".")
from flexynesis_b200 import _lib as L
from flexynesis_b200.engine import build_gcn_csr
from oracle.restatement import synthetic_graph

B, N, C = 4096, 2000, 32
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda:0")
ei = synthetic_graph(N, 20000, 0)
csr_in, csr_out = build_gcn_csr(ei, N, dev, "GCN")
def by_degree(csr):
    deg = (csr[0][1:] - csr[0][:-1]).to(torch.int64)
    return (*csr, torch.sort(deg, descending=True, stable=True).indices.to(torch.int32).contiguous())
import os
if not os.environ.get("NO_ORDER"):
    csr_in, csr_out = by_degree(csr_in), by_degree(csr_out)
x = torch.randn(B, N, C, device=dev)
planes = L.Planes.empty(B * N, C, dev)
out = torch.empty(B, N, C, device=dev)
for _ in range(2):
    L.graph_gather(x.data_ptr(), B, N, C, csr_in, out_planes=planes)
    L.graph_gather(x.data_ptr(), B, N, C, csr_out, out=out.data_ptr())
torch.cuda.synchronize()
for name, kw, csr in (("planes", dict(out_planes=planes), csr_in), ("fp32", dict(out=out.data_ptr()), csr_out)):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        L.graph_gather(x.data_ptr(), B, N, C, csr, **kw)
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) / reps * 1e3
    print(f"{name}: {us:.1f} us  {2 * B * N * C * 4 / us / 1e3:.0f} GB/s algorithmic")
# check against a dense reference on a few samples
rp, col, w = [t.cpu() for t in csr_in[:3]]
A = torch.zeros(N, N)
for v in range(N):
    for e in range(int(rp[v]), int(rp[v + 1])):
        A[v, int(col[e])] += float(w[e])
want = A.to(dev) @ x[:3]
L.graph_gather(x.data_ptr(), B, N, C, csr_in, out=out.data_ptr())
torch.cuda.synchronize()
print("max abs err vs dense:", float((out[:3] - want).abs().max()))
