"""Diagnostic: how sensitive are the training gradients of the tiny-batch golden case to fp32 summation
order?  Compares (a) the oracle run on CUDA fp32, (b) the product fp32 path, against the CPU golden and a
float64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from util import O, load_golden, namespace, compare_grads
from adamml_b200.models import build_model
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
g = load_golden("resnet50_rgb_b2"); case = g["case"]
cfg = O.make_cfg(case["modality"], num_segments=1)
model, _ = build_model(namespace(case, compute_dtype=torch.float32))
shapes = {k: v.shape for k, v in model.state_dict().items()}
sd0 = O.fill_state_dict(shapes, seed=0)
xs, y = O.make_inputs(cfg, 2, 1, hw=224)
gen = torch.Generator(); gen.manual_seed(g["seed"])
mask = torch.empty(2, 2048).bernoulli_(0.5, generator=gen).div_(0.5)

def oracle(dev, dt):
    sd = {k: v.to(dev).to(dt) if v.is_floating_point() else v.to(dev) for k, v in O.clone_sd(sd0, False).items()}
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")): v.requires_grad_(True)
    lg = O.resnet_forward(sd, "", xs[0].to(dev).to(dt), cfg, True, mask.to(dev).to(dt))
    F.cross_entropy(lg, y.to(dev)).backward()
    return lg.detach().cpu(), {k: v.grad.detach().cpu() for k, v in sd.items() if v.requires_grad}

l64, g64 = oracle("cpu", torch.float64)
l32, g32 = oracle("cpu", torch.float32)
lc, gc = oracle("cuda", torch.float32)
model.load_state_dict(sd0); model = model.cuda().train()
lp = model(xs[0].cuda(), drop_mask=mask.cuda()); F.cross_entropy(lp, y.cuda()).backward()
gp = {k: p.grad.detach().cpu() for k, p in model.named_parameters()}
def worst(a, b):
    bad = compare_grads(a, b, tol=0.0)
    bad.sort(key=lambda t: -t[1]); return [(k, round(e, 5)) for k, e, _ in bad[:4]], sum(e for _, e, _ in bad) / len(bad)
print("logits: cpu32-f64 %.2e cuda32-f64 %.2e product-f64 %.2e" % tuple(((a.double() - l64).abs().max() / l64.abs().max()).item() for a in (l32, lc, lp.detach().cpu())))
print("cpu32  vs f64:", worst(g32, g64))
print("cuda32 vs f64:", worst(gc, g64))
print("product vs f64:", worst(gp, g64))
print("product vs cpu32:", worst(gp, g32))
print("cuda32 vs cpu32:", worst(gc, g32))
