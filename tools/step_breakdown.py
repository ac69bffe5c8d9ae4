"""Per-call device time of one eager training step (CUDA events around every C-ABI call). Usage:
    python tools/step_breakdown.py [cfg2|cfg1|cfg3|cfg4|cfg5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from flexynesis_b200 import _lib as L

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
prob = bench.build_problem(w, 0)
model = bench.build_model(w, prob, dev)
batch = bench.device_batch(prob, dev)
model.engine().parallel_encoders = False          # serialise the modality chains so per-call times are meaningful
for _ in range(3):
    model.fit_step(batch)
torch.cuda.synchronize()

records = []
def wrap(fname):
    orig = getattr(L, fname)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(*a, **k); e1.record()
        desc = fname
        if fname == "gemm":
            desc = f"gemm M={a[0]} N={a[1]} K={a[2]} a_mn={a[4]} b_mn={a[6]} splitk={k.get('splitk',0)} epi={k.get('epi_act',0)}"
        elif fname in ("bn_fwd", "bn_bwd"):
            desc = f"{fname} rows={k.get('rows')} cols={k.get('cols')}"
        elif fname in ("gcn_fwd", "gcn_bwd"):
            desc = f"{fname} Fin={a[3] if fname == 'gcn_fwd' else a[4]}"
        records.append((desc, e0, e1))
        return r
    setattr(L, fname, f)
for fn in ["gemm", "bn_fwd", "bn_bwd", "head_out_fwd", "head_out_bwd", "cox_fwd", "total_loss", "clip_adam",
           "split_planes_multi", "split_planes", "col_stats", "gcn_fwd", "gcn_bwd", "merge_col_stats", "reparam_fwd",
           "reparam_bwd", "row_sqnorm", "mmd_finish", "mmd_grad", "loss_weights", "randn", "triplet_fwd", "triplet_bwd"]:
    wrap(fn)
N = 5
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
per_iter = []
for it in range(N):
    records.clear()
    t0.record(); model.fit_step(batch); t1.record()
    torch.cuda.synchronize()
    per_iter.append([(d, a.elapsed_time(b) * 1e3) for d, a, b in records])
    tot = t0.elapsed_time(t1) * 1e3
last = per_iter[-1]
print(f"{name}: eager step wall (device) {tot:.1f} us; sum of calls {sum(t for _, t in last):.1f} us; {len(last)} calls")
for i, (d, t) in enumerate(last):
    best = min(p[i][1] for p in per_iter[1:])
    fl = ""
    if d.startswith("gemm"):
        parts = dict(kv.split("=") for kv in d.split()[1:])
        f = 2.0 * int(parts["M"]) * int(parts["N"]) * int(parts["K"])
        fl = f"  {f / (best * 1e-6) / 1e12:7.1f} TF/s algorithmic"
    print(f"{best:9.1f} us  {d}{fl}")
