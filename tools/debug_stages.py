"""Stage-by-stage comparison of the engine's buffers against the torch formulation on the same GPU (debug aid)."""
import sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import torch.nn.functional as F
from oracle.restatement import Spec
from test_gpu_parity import build_model, to_cuda, masks_from_noise


def planes_f(pl, rows=None):
    rows = rows or pl.rows
    h = pl.hi.float() + pl.lo.float()
    idx = pl.off + torch.arange(rows, device=h.device)[:, None] * pl.ld + torch.arange(pl.cols, device=h.device)[None, :]
    return h[idx]


def rep(tag, a, b):
    err = float((a.double() - b.double()).abs().max())
    print(f"  {tag:28s} max|diff| {err:.3e}  scale {float(b.abs().max()):.3e}")


def main(path, train):
    g = torch.load(path, weights_only=False)
    spec = Spec(**g["spec"])
    model = build_model(spec, g["batch"], g["lr"], g["P0"])
    model.train(train)
    cb = to_cuda(g["batch"])
    eng = model.engine()
    groups, y = model._split_batch(cb)
    masks = masks_from_noise(g["steps"][0]["noise"]) if train else None
    if train:
        ws = eng.forward_backward(groups, y, masks)
    else:
        ws = eng.evaluate(groups, y)
    torch.cuda.synchronize()
    B = ws["B"]
    print(f"{os.path.basename(path)} train={train} B={B}")
    xs = groups[0]
    embs = []
    for i, enc in enumerate(model.encoders):
        h = eng.h[i]
        rep(f"X[{i}] planes", planes_f(ws["X"][i], B), xs[i])
        z = F.linear(xs[i], enc.layer_1.weight, enc.layer_1.bias)
        rep(f"Z[{i}]", ws["Z"][i][:B, :h], z)
        if train:
            mean, var = z.mean(0), z.var(0, unbiased=False)
        else:
            # running stats already updated in train mode -> only meaningful for eval
            mean, var = enc.batchnorm.running_mean, enc.batchnorm.running_var
        yv = (z - mean) * torch.rsqrt(var + 1e-5) * enc.batchnorm.weight + enc.batchnorm.bias
        d = torch.relu(yv)
        if train:
            d = d * masks[f"encoders.{i}.dropout"].float() / 0.9
        rep(f"D[{i}] planes", planes_f(ws["D"][i], B), d)
        e = F.linear(d, enc.layer_out.weight, enc.layer_out.bias)
        rep(f"E[{i}]", ws["Ecat"][:B, i * eng.Lp:i * eng.Lp + eng.latent], e)
        embs.append(e)
    cat = torch.cat(embs, 1)
    f = F.linear(cat, model.fusion_block.weight, model.fusion_block.bias) if eng.fused else cat
    rep("F", ws["F"][:B, :eng.latent], f)
    rep("F planes", planes_f(ws["F_p"], B), f)
    hw = ws["heads"]
    for i, v in enumerate(eng.heads.vars):
        mlp = model.MLPs[v]
        sh, c0 = eng.heads.sh, i * eng.heads.shp
        zh = F.linear(f, mlp.layer_1.weight, mlp.layer_1.bias)
        rep(f"Zh[{v}]", hw["Zh"][:B, c0:c0 + sh], zh)
        if train:
            mean, var = zh.mean(0), zh.var(0, unbiased=False)
        else:
            mean, var = mlp.batchnorm.running_mean, mlp.batchnorm.running_var
        dh = torch.relu((zh - mean) * torch.rsqrt(var + 1e-5) * mlp.batchnorm.weight + mlp.batchnorm.bias)
        if train:
            dh = dh * masks[f"MLPs.{v}.dropout"].float() / 0.9
        rep(f"Dh[{v}]", hw["Dh"][:B, c0:c0 + sh], dh)
        lo = F.linear(dh, mlp.layer_out.weight, mlp.layer_out.bias)
        rep(f"logits[{v}]", hw["logits"][v], lo)
    print("  losses:", {k: float(v) for k, v in eng.losses(ws).items()})
    if train:
        st = g["steps"][0]
        print("  ref   :", {k: float(v) for k, v in st["losses"].items()})
        for k, gr in st["grads"].items():
            got = eng.arena.view(k, eng.arena.grad).cpu()
            if gr is None:
                print(f"  grad {k:40s} ref None, got max {float(got.abs().max()):.2e}")
            else:
                print(f"  grad {k:40s} max|diff| {float((got - gr).abs().max()):.3e} scale {float(gr.abs().max()):.3e}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p, False)
        main(p, True)
