"""Times the data-parallel reduce / Adam-broadcast kernels alone for several slice sizes (fixed vs per-byte cost).
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_probe.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem
from flexynesis_b200 import _lib as L

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 24
grad = symm_mem.empty(n, dtype=torch.float32, device=dev); grad.normal_()
flat = symm_mem.empty(n, dtype=torch.float32, device=dev); flat.normal_()
part = symm_mem.empty(64, dtype=torch.float32, device=dev); part.zero_()
hg, hf, hp = (symm_mem.rendezvous(t, dist.group.WORLD) for t in (grad, flat, part))
m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
scratch = torch.zeros(4, device=dev)
step = torch.zeros(1, dtype=torch.int64, device=dev)
norm = torch.zeros(1, device=dev)
peers = [int(p) for p in hg.buffer_ptrs]
torch.cuda.synchronize(); dist.barrier()


def timed(fn, reps=20):
    """per-call device time with the calls captured in one CUDA graph (no launch overhead between them)"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize(); dist.barrier()
    g.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for size in (1 << 10, 1 << 16, 1 << 18, 1 << 19, 1 << 20, 1 << 21):
    b, e = rank * size, (rank + 1) * size
    if e > n:
        break
    t_mm = timed(lambda: L.dp_reduce_sumsq(hg.multicast_ptr, grad.data_ptr(), b, e, 1.0, hp.multicast_ptr, rank, scratch.data_ptr(), step.data_ptr()))
    t_pp = timed(lambda: L.dp_reduce_sumsq(hg.multicast_ptr, grad.data_ptr(), b, e, 1.0, hp.multicast_ptr, rank, scratch.data_ptr(), step.data_ptr(), peers=peers))
    t_ad = timed(lambda: L.dp_adam_bcast(hf.multicast_ptr, flat.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), b, e, part.data_ptr(), world, 1e-3, 1.0, step.data_ptr(), norm.data_ptr()))
    t_loc = timed(lambda: torch.mul(grad[b:e], 1.0, out=m[b:e]))
    if rank == 0:
        print(f"slice {size:8d} floats ({size * 4 / 1e6:6.2f} MB): reduce multimem {t_mm:6.1f} us, reduce peer-loads {t_pp:6.1f} us, "
              f"adam+bcast {t_ad:6.1f} us, local copy {t_loc:6.1f} us", flush=True)
dist.barrier()
dist.destroy_process_group()
