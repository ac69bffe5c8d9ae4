"""Device timeline of one replay of the data-parallel training step on rank 0 (CUPTI through torch.profiler).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
           tools/dp_timeline.py [cfg2] > profiles/<name>.log        (not a benchmark: profiler attached)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
import bench
from flexynesis_b200.fit import GraphedStep

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
w = bench.WORKLOADS[name]
prob = bench.build_problem(w, rank)
model, allreduce, mode, note = bench.build_data_parallel(w, prob, dev, world, False)
step = GraphedStep(model, bench.device_batch(prob, dev), allreduce=allreduce, grad_scale=1.0 / world)
for _ in range(20):
    step()
torch.cuda.synchronize()
dist.barrier()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(5):
        step()
    torch.cuda.synchronize()
dist.barrier()
if rank == 0:
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    n = len(evs) // 5
    last = evs[-n:]
    t0 = last[0].time_range.start
    end = max(e.time_range.end for e in last)
    prev_start = evs[-2 * n].time_range.start
    print(f"{name} x{world} ({mode}{' ' + note if note else ''}): {n} device activities per replay, span {end - t0:.1f} us, "
          f"replay period {t0 - prev_start:.1f} us")
    for e in last:
        print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:7.1f} us  {e.name[:90]}")
dist.destroy_process_group()
