"""N-GPU check of parallel.GlobalBatchSync over NCCL: W ranks, each with 1/W of a batch, against ONE engine step on the
concatenated batch (run on every rank's own GPU for comparison). Losses, logits and the rank-averaged gradient arena must
agree to fp32 rounding. (tests/test_global_batch.py proves the same statement over gloo on one GPU and against the CPU
oracle; this is the deployment transport.)
Run: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/global_batch_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import flexynesis_b200 as fx
from flexynesis_b200.parallel import GlobalBatchSync
from oracle.restatement import Spec, init_params, synthetic_batch

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
VT = {"y": "numerical", "c": "categorical", "e": "numerical", "t": "numerical"}
spec = Spec(model="DirectPred", input_dims=[700, 400], latent_dim=64, hidden_dim_factor=0.2, supervisor_hidden_dim=16,
            variables=["y", "c", "e"], variable_types=VT, num_classes={"c": 4}, surv_event_var="e", surv_time_var="t")
Bl = 160
B = Bl * world
torch.manual_seed(0)
P0 = init_params(spec)
dat, y = synthetic_batch(spec, B, 0)


class DS:
    pass


def build(d, yy):
    ds = DS()
    ds.dat, ds.variable_types = d, VT
    ds.ann = {k: torch.nan_to_num(v, nan=0.0) for k, v in y.items()}
    ds.features = {k: list(range(v.shape[1])) for k, v in d.items()}
    cfg = {"latent_dim": 64, "hidden_dim_factor": 0.2, "supervisor_hidden_dim": 16, "lr": 1e-3}
    m = fx.DirectPred(cfg, ds, ["y", "c"], surv_event_var="e", surv_time_var="t", device_type="gpu")
    m.load_state_dict(P0, strict=True)
    return m.to(dev).train()


g = torch.Generator().manual_seed(5)
h = [int(d * 0.2) for d in spec.input_dims]
masks = {f"encoders.{i}.dropout": (torch.rand(B, h[i], generator=g) > 0.1).to(torch.uint8) for i in range(2)}
for v in spec.variables:
    masks[f"MLPs.{v}.dropout"] = (torch.rand(B, 16, generator=g) > 0.1).to(torch.uint8)
sl = slice(rank * Bl, (rank + 1) * Bl)
# sharded step with global-batch semantics
mA = build({k: v[sl] for k, v in dat.items()}, y)
eA = mA.engine(dev)
eA.sync = GlobalBatchSync()
gA, yA = mA._split_batch(({k: v[sl].to(dev) for k, v in dat.items()}, {k: v[sl].to(dev) for k, v in y.items()}, None))
wsA = eA.forward_backward(gA, yA, {k: v[sl].to(dev).contiguous() for k, v in masks.items()})
gradA = eA.arena.grad.clone()
dist.all_reduce(gradA)
gradA /= world
lossA = eA.losses(wsA)["__total__"].clone().reshape(1)
dist.all_reduce(lossA)
lossA /= world
# one step on the concatenated batch
mB = build(dat, y)
eB = mB.engine(dev)
gB, yB = mB._split_batch(({k: v.to(dev) for k, v in dat.items()}, {k: v.to(dev) for k, v in y.items()}, None))
wsB = eB.forward_backward(gB, yB, {k: v.to(dev).contiguous() for k, v in masks.items()})
torch.cuda.synchronize()
gscale = float(eB.arena.grad.abs().max())
gerr = float((gradA - eB.arena.grad).abs().max()) / gscale
lerr = abs(float(lossA) - float(eB.losses(wsB)["__total__"]))
lg = max(float((wsA["heads"]["logits"][v] - wsB["heads"]["logits"][v][sl]).abs().max()) for v in spec.variables)
bn = max(float((a - b).abs().max()) for (ka, a), (kb, b) in zip(mA.named_buffers(), mB.named_buffers()) if "running" in ka)
print(f"rank {rank}/{world}: loss {float(lossA):.6f} vs {float(eB.losses(wsB)['__total__']):.6f} (|d| {lerr:.2e}); "
      f"max grad diff / max grad {gerr:.2e}; logits diff {lg:.2e}; running-stat diff {bn:.2e}; collectives {eA.sync.calls}")
ok = gerr < 2e-5 and lerr < 1e-5 and lg < 1e-4 and bn < 1e-5
# the same sharded step captured in ONE CUDA graph (the NCCL collectives of GlobalBatchSync are captured with the kernels)
# and replayed: its gradients must equal the eager step's
gerr_graph = float("nan")
try:
    grad_eager = eA.arena.grad.clone()
    mask_dev = {k: v[sl].to(dev).contiguous() for k, v in masks.items()}
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eA.forward_backward(gA, yA, mask_dev)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize(); dist.barrier()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        eA.forward_backward(gA, yA, mask_dev)
    torch.cuda.synchronize(); dist.barrier()
    eA.arena.grad.zero_()
    graph.replay()
    torch.cuda.synchronize()
    gerr_graph = float((eA.arena.grad - grad_eager).abs().max()) / float(grad_eager.abs().max())
    print(f"rank {rank}/{world}: captured step vs eager step: max grad diff / max grad {gerr_graph:.2e}")
    ok = ok and gerr_graph < 2e-6
    del graph                                            # (a live graph holding captured NCCL work stalls the teardown below)
    torch.cuda.synchronize()
except Exception as e:                                   # report, fail the check
    print(f"rank {rank}/{world}: graph capture of the global-batch step failed: {type(e).__name__}: {e}")
    ok = False
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("GLOBAL BATCH CHECK OK" if int(flag) else "GLOBAL BATCH CHECK FAILED")
rc = 0 if int(flag) else 1
sys.stdout.flush()
dist.barrier()
os._exit(rc)          # skip the process-group teardown: with captured NCCL collectives it waited until the launcher's timeout
