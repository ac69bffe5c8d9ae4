"""Concurrency experiment for the two first-layer GEMMs of config 2 (forward and weight gradient): each alone, both as
parallel branches of one CUDA graph, with the tile widths given on the command line. Prints per-kernel start/duration from
CUPTI (torch.profiler) and the CUDA-event time of the whole graph.
Usage: python tools/gemm_pair_exp.py fwd 256 160 | wgrad 0 0     (not a benchmark: profiler attached for the timeline)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from flexynesis_b200 import _lib as L
from flexynesis_b200._lib import Planes, pad8

mode = sys.argv[1] if len(sys.argv) > 1 else "fwd"
bn0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bn1 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
order = sys.argv[4] if len(sys.argv) > 4 else "01"
g0 = int(sys.argv[5]) if len(sys.argv) > 5 else 0           # max_groups of the two launches
g1 = int(sys.argv[6]) if len(sys.argv) > 6 else 0
use_fix = int(sys.argv[7]) if len(sys.argv) > 7 else 0      # 1: give the forward GEMMs a fix-up workspace
dev = torch.device("cuda", 0)
B = 4096
dims, hs = [5000, 3000], [512, 307]
torch.manual_seed(0)


def planes_of(t):
    p = Planes.empty(t.shape[0], t.shape[1], dev)
    L.split_planes(t, p)
    return p


X = [planes_of(torch.randn(B, d, device=dev)) for d in dims]
W = [planes_of(torch.randn(h, d, device=dev) * 0.02) for h, d in zip(hs, dims)]
dZ = [planes_of(torch.randn(B, h, device=dev)) for h in hs]
Z = [torch.zeros(B, pad8(h), device=dev) for h in hs]
bias = [torch.zeros(h, device=dev) for h in hs]
part = [torch.zeros(L.stat_tiles(B) * 2 * h, device=dev) for h in hs]
dW = [torch.zeros(h, d, device=dev) for h, d in zip(hs, dims)]


fixws = [L.FixWorkspace(dev) for _ in range(2)]
both = False


def launch(i, bn):
    g = (g0, g1)[i] if both else 0
    if mode == "fwd":
        L.gemm(B, hs[i], dims[i], X[i], 0, W[i], 0, C_ptr=Z[i].data_ptr(), ldc=Z[i].stride(0), bias=bias[i].data_ptr(),
               colstats=part[i].data_ptr(), stats_mode=2, block_n=bn, max_groups=g, fix=fixws[i] if use_fix else None)
    else:
        L.gemm(hs[i], dims[i], B, dZ[i], 1, X[i], 1, C_ptr=dW[i].data_ptr(), ldc=dims[i], splitk=-1, block_n=bn, max_groups=g,
               prezeroed=True)


def capture(which):
    global both
    both = which == "both"
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        if which == "both":
            ev = torch.cuda.Event(); ev.record(torch.cuda.current_stream()); side.wait_event(ev)
            first, second = (0, 1) if order == "01" else (1, 0)
            launch(first, (bn0, bn1)[first])
            with torch.cuda.stream(side):
                launch(second, (bn0, bn1)[second])
            ev2 = torch.cuda.Event(); ev2.record(side); torch.cuda.current_stream().wait_event(ev2)
        elif which == "seq":
            launch(0, bn0); launch(1, bn1)
        else:
            launch(which, (bn0, bn1)[which])
    return g


for which in (0, 1, "seq", "both"):
    # warm up eagerly (tensor-map encode, attribute set)
    if which in (0, 1):
        launch(which, (bn0, bn1)[which])
    torch.cuda.synchronize()
    g = capture(which)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{mode} {which}: bn=({bn0},{bn1}) groups=({g0},{g1}) fix={use_fix} order={order}  {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per replay")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            g.replay()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 2
    last = evs[-n:]
    t0 = last[0].time_range.start
    for e in last:
        print(f"    {e.time_range.start - t0:8.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:60]}")
