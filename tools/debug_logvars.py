"""Repeat forward_backward of the multitask case and compare the log_vars gradients with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from test_gpu_parity import CASES, oracle_reference, build_model, to_cuda, masks_from_noise, sync_state
spec, B = CASES["multitask"]
P0, batch, steps, _ = oracle_reference(spec, B, 1e-3, steps=1)
model = build_model(spec, batch, 1e-3, P0); model.train()
cb = to_cuda(batch)
eng = model.engine()
eng.parallel_encoders = os.environ.get("SERIAL", "0") != "1"
st = steps[0]
a = eng.arena
groups, y = model._split_batch(cb)
masks = masks_from_noise(st["noise"])
want = {n: st["grads"][n] for n in a.names}
for it in range(6):
    sync_state(model, st["P_before"])
    ws = eng.forward_backward(groups, y, masks)
    torch.cuda.synchronize()
    bad = []
    for n in a.names:
        g = want[n]
        if g is None: continue
        got = a.view(n, a.grad).cpu()
        err = float((got - g).abs().max()) / max(float(g.abs().max()), 1e-12)
        if err > 2e-3 and float(g.abs().max()) > 1e-5: bad.append((n, round(err, 4)))
    print(it, "log_vars", a.grad[[0, 8, 16]].tolist(), "bad:", bad[:6])
print("oracle", [float(st["grads"][n]) for n in a.names[:3]])
