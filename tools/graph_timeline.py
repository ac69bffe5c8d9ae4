"""Device timeline of ONE CUDA-graph replay of a training step (CUPTI through torch.profiler; no nsys in the image):
start offset, duration, stream and name of every kernel, plus the busy time and gaps of the critical stream.
Usage: python tools/graph_timeline.py [cfg2] > profiles/<name>.log     (not a benchmark: profiler attached)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from flexynesis_b200.fit import GraphedStep

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
prob = bench.build_problem(w, 0)
model = bench.build_model(w, prob, dev)
batch = bench.device_batch(prob, dev)
step = GraphedStep(model, batch)
for _ in range(10):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# split into replays by the largest gaps
n = len(evs) // 3
last = evs[-n:]
t0 = last[0].time_range.start
end = max(e.time_range.end for e in last)
print(f"{name}: {n} device activities per replay, span {end - t0:.1f} us")
for e in last:
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
