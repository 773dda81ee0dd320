"""GPU parity: the CUDA path (through the C-ABI) against the golden vectors recorded from the
reference and against the CPU oracle run live on the same seeded inputs.

Tolerance (BASELINE.json north_star): logits within 1e-3 relative (max|a-b| / max|b|) in the
fp32 parity mode, policy selections bit-exact under a fixed seed.  The bf16 speed mode is
checked separately with its own (stated) tolerance.
"""
import pytest
import torch
import torch.nn.functional as F

from util import (O, assert_grads_as_good_as_reference, compare_grads, fingerprint, load_golden, namespace,
                  noise_for_model, oracle_run, rel, grad_errors)

pytestmark = pytest.mark.gpu

CASES = ["resnet50_rgb_b2", "adamml_rgb_sound_eval", "adamml_rgb_flow_train", "adamml_rgb_sound_flow_train",
         "adamml_rgb_sound_train", "adamml_rgb_sound_nocausal_train", "adamml_rgb_sound_train_s5",
         "adamml_rgb_sound_eval_s10"]
LOGIT_TOL = 1e-3


def build(case, dtype, cuda):
    from adamml_b200.models import build_model
    model, arch = build_model(namespace(case, compute_dtype=dtype))
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    model.load_state_dict(O.fill_state_dict(shapes, seed=0), strict=True)
    return model.to(cuda), arch


def run_product(model, case, g_seed, cuda):
    cfg = O.make_cfg(case["modality"], num_segments=case["S"], causality_modeling=case.get("causality", "lstm"))
    N, S_run, training = case["N"], case.get("S_run", case["S"]), case["training"]
    xs, y = O.make_inputs(cfg, N, S_run, hw=case["hw"])
    xs = [x.to(cuda) for x in xs]
    y = y.to(cuda)
    model.train(training)
    if case["kind"] == "resnet":
        gen = torch.Generator(); gen.manual_seed(g_seed)
        mask = torch.empty(N, 2048).bernoulli_(0.5, generator=gen).div_(0.5).to(cuda)
        logits = model(xs[0], drop_mask=mask)
        return logits, None, F.cross_entropy(logits, y)
    noise = noise_for_model(O.draw_noise(g_seed, cfg, N, S_run, training), cuda)
    with torch.set_grad_enabled(training):
        logits, dec = model(xs, num_segments=S_run, noise=noise)
    loss = F.cross_entropy(logits, y)
    if training:  # utils/utils.py:362,380-382 — the loss tail stays in torch
        correct = (logits.detach().argmax(-1) == y).float()
        sel = dec.mean(1) ** 2
        pl = sum(1.0 * torch.mean(correct * c) for c in sel.chunk(sel.shape[-1], dim=-1))
        loss = loss + pl + torch.mean((1 - correct) * 10.0)
    return logits, dec, loss


@pytest.mark.parametrize("name", CASES)
def test_fp32_mode_matches_reference_golden(cuda, name):
    g = load_golden(name)
    case = g["case"]
    model, arch = build(case, torch.float32, cuda)
    assert arch == g["arch"]
    logits, dec, loss = run_product(model, case, g["seed"], cuda)
    assert rel(logits, g["logits"]) < LOGIT_TOL
    if dec is not None:
        assert torch.equal(dec.detach().cpu(), g["decisions"]), "policy selections must be bit-exact"
    assert abs(loss.item() - g["loss"].item()) < 1e-3 * max(1.0, abs(g["loss"].item()))
    if not case["training"]:
        return
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(g["grad_fp"]) <= set(grads), sorted(set(g["grad_fp"]) - set(grads))[:5]
    # gradients: same size of error against the float64 oracle as the reference's own fp32 gradients
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    sd0 = O.fill_state_dict(shapes, seed=0)
    _, _, g32, _ = oracle_run(case, g["seed"], torch.float32, sd0)
    # the fp32 oracle IS the reference (bit-exact on the machine that made the goldens; on another host the
    # CPU threading changes the summation order, so only closeness is asserted here)
    m32, _, _ = grad_errors(g32, g["grad_small"])
    assert m32 < 2e-2, m32
    _, dec64, g64, _ = oracle_run(case, g["seed"], torch.float64, sd0)
    if dec is not None:
        assert torch.equal(dec64.float(), g["decisions"])
    assert_grads_as_good_as_reference(grads, g32, g64)
    sd = model.state_dict()
    for k, fp in g["running_fp"].items():
        assert rel(fingerprint(sd[k])[1], fp[1]) < 1e-4, k
    for k, v in g["num_batches_tracked"].items():
        assert int(sd[k]) == v, k
    print(f"{name}: logits rel {rel(logits, g['logits']):.2e}")


def test_fp32_mode_well_conditioned_all_grads(cuda):
    """A better-conditioned case (8 videos => >= 32 values per BN channel even in layer4): every parameter
    gradient element-wise against the fp32 oracle, tight tolerance."""
    case = dict(kind="adamml", modality=["rgb", "sound"], N=8, S=2, hw=64, training=True)
    model, _ = build(case, torch.float32, cuda)
    logits, dec, loss = run_product(model, case, 5, cuda)
    loss.backward()
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    sd0 = O.fill_state_dict(shapes, seed=0)
    o_logits, o_dec, g32, sd = oracle_run(case, 5, torch.float32, sd0)
    _, _, g64, _ = oracle_run(case, 5, torch.float64, sd0)
    assert rel(logits, o_logits) < LOGIT_TOL
    assert torch.equal(dec.detach().cpu(), o_dec)
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert set(grads) == set(g32)
    assert_grads_as_good_as_reference(grads, g32, g64)
    new = model.state_dict()
    for k in new:
        if k.endswith(("running_mean", "running_var")):
            assert rel(new[k], sd[k]) < 1e-4, k


def test_frozen_phases(cuda):
    """Warm-up / policy phases (train_adamml.py:344-345,410-411): frozen halves get no grads, the
    other half's grads are unchanged."""
    case = dict(kind="adamml", modality=["rgb", "sound"], N=1, S=2, hw=64, training=True)
    model, _ = build(case, torch.float32, cuda)
    _, _, loss = run_product(model, case, 7, cuda)
    loss.backward()
    ref = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    sd0 = O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0)
    for phase in ("policy_frozen", "main_frozen"):
        model.load_state_dict(sd0)
        model.unfreeze_policy_net(); model.unfreeze_main_net()
        (model.freeze_policy_net if phase == "policy_frozen" else model.freeze_main_net)()
        _, _, loss = run_product(model, case, 7, cuda)
        loss.backward()
        for k, p in model.named_parameters():
            frozen = k.startswith("policy_net." if phase == "policy_frozen" else "main_net.")
            assert (p.grad is None) == frozen, k
        live = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
        bad = compare_grads(live, {k: ref[k] for k in live}, tol=1e-3)
        assert not bad, (phase, bad[:8])
        model.zero_grad(set_to_none=True)


def test_bf16_mode_close_to_oracle(cuda):
    """Speed mode: bf16 activations/operands (tcgen05 where the shape fits), fp32 accumulation, fp32/fp64 BN
    statistics, fp32 policy head.

    Stated tolerances (bf16 has 8 mantissa bits and every one of the 53+52+52 conv outputs is rounded to it):
      * eval mode (running-stat BN, the golden 'adamml_rgb_sound_eval' case): logits within 5e-2;
      * train mode on a 2-video batch: batch-stat BN over as few as 18 values per channel amplifies the
        rounding noise, so only finiteness, gradient presence and a <= 25 % selection mismatch rate are
        asserted, and the logits error is asserted only when all selections agree (< 0.3).
    """
    g = load_golden("adamml_rgb_sound_eval")
    model, _ = build(g["case"], torch.bfloat16, cuda)
    logits, dec, _ = run_product(model, g["case"], g["seed"], cuda)
    mism = (dec.detach().cpu() != g["decisions"]).float().mean().item()
    e = rel(logits, g["logits"])
    print(f"bf16 eval: logits rel {e:.3e}, selection mismatch {mism:.3f}")
    assert mism <= 0.25
    if mism == 0:
        assert e < 5e-2

    case = dict(kind="adamml", modality=["rgb", "sound"], N=2, S=2, hw=96, training=True)
    model, _ = build(case, torch.bfloat16, cuda)
    logits, dec, loss = run_product(model, case, 5, cuda)
    loss.backward()
    torch.cuda.synchronize()
    cfg = O.make_cfg(case["modality"], num_segments=2)
    sd = O.clone_sd(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0), False)
    xs, _ = O.make_inputs(cfg, 2, 2, hw=96)
    with torch.no_grad():
        o_logits, o_dec = O.adamml_forward(sd, xs, cfg, True, O.draw_noise(5, cfg, 2, 2, True))
    assert torch.isfinite(logits).all()
    mism = (dec.detach().cpu() != o_dec).float().mean().item()
    e = rel(logits, o_logits)
    print(f"bf16 train: logits rel {e:.3e}, selection mismatch {mism:.3f}")
    assert mism <= 0.25
    if mism == 0:
        assert e < 0.3
    for k, p in model.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_bf16_mode_vs_fp32_mode_same_decisions(cuda):
    """bf16 speed mode against the (oracle-verified) fp32 parity mode on a batch large enough for stable BatchNorm
    statistics (16 clips x 2 segments, 112^2).  rng_policy=True draws the gating decisions from torch's RNG instead
    of the policy net (adamml.py:76-78), so both precisions gate identically and the comparison isolates the
    numerical error of the main path (ResNet-50 + sound MobileNetV2 + fusion): train-mode logits and the
    classifier gradients must agree to bf16 noise."""
    from adamml_b200.models import build_model
    case = dict(kind="adamml", modality=["rgb", "sound"], N=16, S=2, hw=112, training=True)
    cfg = O.make_cfg(case["modality"], num_segments=2)
    gen = torch.Generator(device=cuda).manual_seed(7)
    rgb = torch.randn(16, 2 * 8 * 3, 112, 112, device=cuda, generator=gen)
    snd = torch.randn(16, 2, 256, 256, device=cuda, generator=gen) * 3 - 5
    y = torch.randint(0, 31, (16,), device=cuda, generator=gen)
    res = {}
    sd0 = None
    for tag, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        model, _ = build_model(namespace(case, compute_dtype=dt, rng_policy=True, rng_threshold=0.5))
        if sd0 is None:
            sd0 = O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0)
        model.load_state_dict(sd0)
        model = model.to(cuda).train()
        torch.manual_seed(11)  # same decisions and dropout masks for both precisions
        logits, dec = model([rgb, snd])
        F.cross_entropy(logits, y).backward()
        torch.cuda.synchronize()
        res[tag] = (logits.detach(), dec.detach(), model.main_net.nets[0].fc.weight.grad.clone(),
                    model.main_net.nets[1].classifier[1].weight.grad.clone())
    assert torch.equal(res["fp32"][1], res["bf16"][1])
    e_logits = rel(res["bf16"][0], res["fp32"][0])
    e_g0, e_g1 = rel(res["bf16"][2], res["fp32"][2]), rel(res["bf16"][3], res["fp32"][3])
    print(f"bf16 vs fp32 mode (N=16): logits rel {e_logits:.3e}, fc grad rel {e_g0:.3e} / {e_g1:.3e}")
    # measured 0.12 / ... on randomly initialised weights: this is the price of bf16 STORAGE through 53 stacked
    # conv+BN layers, not of the kernels — test_bf16_resnet_matches_bf16_storage_emulation pins the kernels
    assert e_logits < 0.3 and e_g0 < 0.3 and e_g1 < 0.3


def test_bf16_resnet_tracks_bf16_storage_emulation(cuda):
    """The bf16 engine (s2d tcgen05 stem, tcgen05 GEMM/conv with fused BN statistics, row-streaming BN) against a
    torch fp32 computation that rounds to bf16 exactly where the engine stores bf16 (conv operands, conv outputs z,
    block outputs), stage by stage through ResNet-50 (8 clips x 8 frames, 112^2, train-mode BN).

    On randomly initialised weights the network amplifies any perturbation by ~1.25x per block (the emulation itself
    drifts to ~50 % rms from pure fp32 at layer4 — scripts/diag_bf16_drift.py), so end-to-end closeness proves
    nothing.  What pins the kernels: (1) the first stages agree to far better than one bf16 ulp of noise
    (stem 3e-5, layer1.0 7e-4 rms measured), (2) at EVERY stage the engine is closer to the emulation than the
    emulation is to fp32, i.e. it never adds error beyond bf16 storage noise, and the growth has no jump."""
    from adamml_b200.engine import Exec
    from adamml_b200.models.resnet import ResNet
    from adamml_b200.ops import ACT_RELU
    torch.manual_seed(3)
    net = ResNet(50, 8, num_classes=31, dropout=0.5, input_channels=3, compute_dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.5 + 0.75)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
    net = net.to(cuda).train()
    N = 8
    x = torch.randn(N, 24, 112, 112, generator=g).to(cuda)

    def rms(a, b):
        return ((a.float() - b.float()).pow(2).mean().sqrt() / b.float().pow(2).mean().sqrt()).item()

    def run_torch(rounding):
        q = (lambda t: t.bfloat16().float()) if rounding else (lambda t: t)
        conv = lambda m, a: q(F.conv2d(a, q(m.weight), None, m.stride, m.padding))  # noqa: E731
        bn = lambda m, z: F.batch_norm(z, None, None, m.weight, m.bias, True, 0.1, m.eps)  # noqa: E731
        outs = []
        with torch.no_grad():
            a = q(F.relu(bn(net.bn1, conv(net.conv1, q(x.view(N * 8, 3, 112, 112))))))
            outs.append(a)
            a = F.max_pool2d(a, 3, 2, 1)
            frames = 8
            for li in range(4):
                for blk in getattr(net, f"layer{li + 1}"):
                    idn = a
                    o = q(F.relu(bn(blk.bn1, conv(blk.conv1, a))))
                    o = q(F.relu(bn(blk.bn2, conv(blk.conv2, o))))
                    o = bn(blk.bn3, conv(blk.conv3, o))
                    if blk.downsample is not None:
                        idn = bn(blk.downsample[1], conv(blk.downsample[0], a))
                    a = q(F.relu(o + idn))
                    outs.append(a)
                if li < 3:
                    nt, c, h, w = a.shape
                    v = a.view(-1, frames, c, h, w).transpose(1, 2)
                    a = F.max_pool3d(v, (3, 1, 1), (2, 1, 1), (1, 0, 0)).transpose(1, 2).contiguous().view(-1, c, h, w)
                    frames //= 2
        return outs

    eng = []
    with torch.no_grad():
        ex = Exec(torch.bfloat16, True, 1, save=False)
        a = ex.cba(net.pack_input(x, 1), net.conv1, net.bn1, ACT_RELU)
        eng.append(a.permute(0, 3, 1, 2).float())
        a = ex.maxpool(a)
        frames = 8
        for li in range(4):
            for blk in getattr(net, f"layer{li + 1}"):
                a = ex.bottleneck(a, blk)
                eng.append(a.permute(0, 3, 1, 2).float())
            if li < 3:
                a = ex.tpool(a, frames, False)
                frames //= 2
    t16, t32 = run_torch(True), run_torch(False)
    d_eng = [rms(a, b) for a, b in zip(eng, t16)]
    d_sto = [rms(b, c) for b, c in zip(t16, t32)]
    print("engine~emulation rms:", " ".join(f"{v:.1e}" for v in d_eng))
    print("emulation~fp32  rms:", " ".join(f"{v:.1e}" for v in d_sto))
    assert d_eng[0] < 1e-3 and d_eng[1] < 5e-3, d_eng[:2]
    for i, (de, ds) in enumerate(zip(d_eng, d_sto)):
        assert de <= ds, (i, de, ds)                      # never worse than bf16 storage noise itself
        if i:
            assert de < 3.0 * max(d_eng[i - 1], 1e-3), (i, de, d_eng[i - 1])  # smooth growth, no jump at a block


def test_unimodal_sound_mobilenet_matches_oracle(cuda):
    from adamml_b200.models import build_model
    ns = namespace(dict(kind="resnet", modality=["sound"], S=1), backbone_net="sound_mobilenet_v2", modality="sound",
                   input_channels=1, compute_dtype=torch.float32)
    model, arch = build_model(ns)
    assert arch.startswith("kinetics-sounds-sound-sound_mobilenet_v2")
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    sd0 = O.fill_state_dict(shapes, seed=0)
    model.load_state_dict(sd0)
    model = model.to(cuda).train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 1, 128, 128, generator=g)
    mask = torch.empty(3, 1280).bernoulli_(0.5, generator=g).div_(0.5)
    y = model(x.to(cuda), drop_mask=mask.to(cuda))
    y.sum().backward()
    def run(dt):
        sd = {k: (v.detach().clone().to(dt).requires_grad_(not k.endswith(("running_mean", "running_var")))
                  if v.is_floating_point() else v.clone()) for k, v in sd0.items()}
        yo = O.sound_mobilenet_forward(sd, "", x.to(dt), True, mask.to(dt))
        yo.sum().backward()
        return yo.detach(), {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}
    yo, g32 = run(torch.float32)
    _, g64 = run(torch.float64)
    assert rel(y, yo) < LOGIT_TOL
    assert_grads_as_good_as_reference({k: p.grad for k, p in model.named_parameters()}, g32, g64)
