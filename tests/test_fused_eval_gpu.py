"""Inference-mode fused epilogue (NS2 / SURVEY §2.5 K5-K7): conv + BatchNorm (+ residual) + ReLU/ReLU6 as ONE kernel
(adamml_tc_*_bn_act_*, adamml_dwconv_bn_act_fwd*) against the unfused conv -> bn_apply pair on the same operands, and
the whole model with the fusion on / off."""
import pytest
import torch

from util import O, load_golden, namespace, noise_for_model, rel

pytestmark = pytest.mark.gpu


def split(t):
    from adamml_b200 import ops
    hi = t.bfloat16()
    return ops.X2(hi.contiguous(), (t - hi.float()).half().contiguous())


def val(t):
    return t.float()


def err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


CASES = [
    # IMGS,H,W,Cin,Cout,R,stride,pad
    (4, 14, 14, 64, 256, 1, 1, 0),     # Bottleneck conv3 (+ identity)
    (3, 9, 9, 144, 24, 1, 1, 0),       # MobileNetV2 project layer (+ identity), linear-store width
    (3, 9, 9, 96, 160, 1, 1, 0),       # partial last column block
    (2, 14, 14, 64, 64, 3, 1, 1),      # Bottleneck conv2
    (2, 28, 28, 128, 128, 3, 2, 1),
    (2, 14, 14, 128, 256, 1, 2, 0),    # downsample
    (2, 7, 7, 512, 512, 3, 1, 1),
]


@pytest.mark.parametrize("mode", ["x2", "bf16"])
@pytest.mark.parametrize("case", CASES)
def test_conv_bn_act_matches_unfused(cuda, case, mode):
    from adamml_b200 import ops
    IMGS, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x32 = torch.randn(IMGS, H, W, Cin, generator=g).to(cuda)
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(cuda)
    Ho, Wo = ops.conv_out_hw(H, W, R, R, stride, pad)
    r32 = torch.randn(IMGS, Ho, Wo, Cout, generator=g).to(cuda)
    ss = torch.stack([torch.rand(Cout, generator=g) + 0.5, torch.randn(Cout, generator=g)], -1).to(cuda).contiguous()
    if mode == "x2":
        x, res, prec, tol = split(x32), split(r32), ops.PREC_X2, 3e-5
    else:
        x, res, prec, tol = x32.bfloat16(), r32.bfloat16(), torch.bfloat16, 1.5e-2
    wp = ops.pack_weight(w.contiguous(), prec)
    for act in (ops.ACT_NONE, ops.ACT_RELU, ops.ACT_RELU6):
        for r in (None, res):
            fused = ops.conv_bn_act_fwd(x, wp, stride, pad, ss, act, res=r)
            assert fused is not None
            z, _ = ops.conv_fwd(x, wp, stride, pad)
            ref = ops.bn_apply(z, ss.view(1, Cout, 2), 1, act, res=r)
            e = err(val(fused), val(ref))
            assert e < tol, (mode, act, r is not None, e)


@pytest.mark.parametrize("mode", ["x2", "bf16"])
def test_first_conv_and_depthwise_bn_act(cuda, mode):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(11)
    prec = ops.PREC_X2 if mode == "x2" else torch.bfloat16
    tol = 3e-5 if mode == "x2" else 1.5e-2
    # 7x7 stem on the s2d operand
    x = torch.randn(2, 2 * 2 * 3, 64, 64, generator=g).to(cuda)
    w = (torch.randn(64, 3, 7, 7, generator=g) / 147 ** 0.5).to(cuda).contiguous()
    ss = torch.stack([torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g)], -1).to(cuda).contiguous()
    xs = ops.pack_frames_s2d(x, 2, 2, 3, x2=(mode == "x2"))
    fused = ops.stem_conv_bn_act_fwd(xs, w, ss, ops.ACT_RELU)
    ref = ops.bn_apply(ops.stem_conv_fwd(xs, w), ss.view(1, 64, 2), 1, ops.ACT_RELU)
    assert err(val(fused), val(ref)) < tol
    # depthwise 3x3, stride 1 and 2
    C = 96
    xd32 = torch.randn(3, 19, 22, C, generator=g).to(cuda)
    xd = split(xd32) if mode == "x2" else xd32.bfloat16()
    wd = ops.pack_weight_dw((torch.randn(C, 1, 3, 3, generator=g) / 3).to(cuda).contiguous())
    ssd = torch.stack([torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)], -1).to(cuda).contiguous()
    for stride in (1, 2):
        fused = ops.dwconv_bn_act_fwd(xd, wd, stride, ssd, ops.ACT_RELU6)
        ref = ops.bn_apply(ops.dwconv_fwd(xd, wd, stride), ssd.view(1, C, 2), 1, ops.ACT_RELU6)
        assert err(val(fused), val(ref)) < tol, stride


@pytest.mark.parametrize("mode", ["x2", "bf16"])
def test_model_eval_fused_vs_unfused(cuda, mode):
    """whole AdaMML inference pass (S_run = 3 on an S = 2 model) with the fused epilogue on and off: same selections,
    logits equal to rounding; and the default mode still meets the golden."""
    from adamml_b200 import engine, ops
    from adamml_b200.models import build_model
    g = load_golden("adamml_rgb_sound_eval")
    case = g["case"]
    prec = ops.PREC_X2 if mode == "x2" else torch.bfloat16
    model, _ = build_model(namespace(case, compute_dtype=prec))
    model.load_state_dict(O.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=0))
    model = model.to(cuda).eval()
    cfg = O.make_cfg(case["modality"], num_segments=case["S"])
    xs, _ = O.make_inputs(cfg, case["N"], case["S_run"], hw=case["hw"])
    xs = [x.to(cuda) for x in xs]
    noise = noise_for_model(O.draw_noise(g["seed"], cfg, case["N"], case["S_run"], False), cuda)
    outs = {}
    old = engine.FUSE_EVAL
    try:
        for fuse in (True, False):
            engine.FUSE_EVAL = fuse
            from adamml_b200 import _lib
            n0 = _lib.launch_count()
            with torch.no_grad():
                outs[fuse] = model(xs, num_segments=case["S_run"], noise=noise)
            outs[fuse] += (_lib.launch_count() - n0,)
    finally:
        engine.FUSE_EVAL = old
    e = rel(outs[True][0], outs[False][0])
    print(f"{mode}: fused vs unfused logits rel {e:.2e}; launches {outs[True][2]} vs {outs[False][2]}")
    assert outs[True][2] < outs[False][2]   # every bn_apply launch is gone (bn_finalize and weight packing remain)
    if mode == "x2":
        assert torch.equal(outs[True][1], outs[False][1])
        assert e < 2e-4
        assert rel(outs[True][0], g["logits"]) < 1e-3
        assert torch.equal(outs[True][1].cpu(), g["decisions"])
    else:
        assert e < 0.2
