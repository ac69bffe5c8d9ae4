"""Worker of tests/test_global_batch.py (GPU part): `world` processes share cuda:0 and talk over gloo (NCCL refuses two
ranks on one device; the collectives of parallel.GlobalBatchSync are backend-agnostic). Each rank runs the engine on
its shard of the recorded global batch with GlobalBatchSync, the gradient arenas are averaged over ranks (what the
data-parallel step does) and rank 0 writes the result for the parent to compare with the CPU oracle."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def shard(t, rank, world):
    n = t.shape[0] // world
    return t[rank * n:(rank + 1) * n].contiguous()


def main():
    rank, world, port, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.cuda.set_device(0)
    from oracle.restatement import Spec
    from test_gpu_parity import build_model, masks_from_noise, to_cuda
    from flexynesis_b200.parallel import GlobalBatchSync

    rec = torch.load(path, weights_only=False)
    spec = Spec(**rec["spec"])
    dat, y = rec["batch"][0], rec["batch"][1]
    local = ({k: shard(v, rank, world) for k, v in dat.items()}, {k: shard(v, rank, world) for k, v in y.items()}, None)
    model = build_model(spec, local, rec["lr"], rec["P0"])
    model.train()
    eng = model.engine()
    eng.sync = GlobalBatchSync()
    out = {"steps": []}
    for st in rec["steps"]:
        model.load_state_dict({k: v for k, v in st["P_before"].items()}, strict=True)
        masks = masks_from_noise({k: shard(v, rank, world) for k, v in st["noise"].items()})
        groups, yy = model._split_batch(to_cuda(local))
        ws = eng.forward_backward(groups, yy, masks)
        g = eng.arena.grad.clone()
        dist.all_reduce(g)
        g /= world
        vals = eng.losses(ws)
        names = [k for k in vals]
        lv = torch.stack([vals[k].reshape(()) for k in names]).clone()
        dist.all_reduce(lv)
        lv /= world
        logits = {}
        for v in eng.heads.vars:
            parts = [torch.empty_like(ws["heads"]["logits"][v]) for _ in range(world)]
            dist.all_gather(parts, ws["heads"]["logits"][v].contiguous())
            logits[v] = torch.cat(parts, 0).cpu()
        bn = {k: b.detach().cpu().clone() for k, b in model.named_buffers()}
        out["steps"].append(dict(grads={k: eng.arena.view(k, g).cpu().clone() for k in eng.arena.names},
                                 losses={k: float(x) for k, x in zip(names, lv.cpu())}, logits=logits, buffers=bn))
    out["sync_calls"] = eng.sync.calls
    if rank == 0:
        torch.save(out, path + ".out")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
