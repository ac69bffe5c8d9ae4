"""Multi-step divergence report: engine vs CPU oracle, per parameter (debug aid)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_parity import CASES, oracle_reference, build_model, to_cuda, masks_from_noise
from oracle.restatement import Noise, Trainer, init_params, synthetic_batch

name = sys.argv[1]
spec, B = CASES[name]
lr = 1e-3
torch.manual_seed(0)
P = init_params(spec)
P0 = {k: v.clone() for k, v in P.items()}
dat, y = synthetic_batch(spec, B, 0)
batch = (dat, y, None)
tr = Trainer(P, spec, lr)
model = build_model(spec, batch, lr, P0)
model.train()
cb = to_cuda(batch)
eng = model.engine()
for s in range(3):
    torch.manual_seed(1000 + s)
    noise = Noise()
    res = tr.step(batch, noise)
    masks = masks_from_noise(noise.record)
    groups, yy = model._split_batch(cb)
    ws = eng.forward_backward(groups, yy, masks)
    print(f"== step {s}: total oracle {float(res['total']):.6f} engine {float(eng.losses(ws)['__total__']):.6f}")
    for k, g in res["grads"].items():
        got = eng.arena.view(k, eng.arena.grad).cpu()
        if g is None:
            continue
        print(f"   grad  {k:36s} err/max {float((got-g).abs().max())/max(float(g.abs().max()),1e-30):.2e}  max {float(g.abs().max()):.2e}")
    eng.optimizer_step(lr, 1.0)
    print(f"   grad_norm oracle {float(res['grad_norm']):.6f} engine {float(eng.arena.grad_norm):.6f}")
    sd = model.state_dict()
    for k in res["grads"]:
        a, b = sd[k].detach().cpu(), P[k].detach()
        d = (a - b).abs()
        print(f"   param {k:36s} max|diff|/lr {float(d.max())/lr:.3f}  frac(>0.5lr) {float((d > 0.5*lr).float().mean()):.4f}")
