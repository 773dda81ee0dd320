"""x2 precision mode (the DEFAULT mode, the one bench.py measures): two-plane forward activations
(hi bf16 + lo fp16 remainder), 4-product tcgen05 GEMMs, bf16 backward on the hi planes.

Bar (BASELINE.json north_star): logits within 1e-3 relative (max|a-b| / max|b|) of the reference's fp32 path on every
golden, policy selections bit-exact under a fixed seed.  Kernel-level checks compare every x2 entry point with a
float64 torch computation on the SAME (plane-rounded) inputs, so only accumulation and output rounding differ."""
import pytest
import torch
import torch.nn.functional as F

from util import O, load_golden, namespace, noise_for_model, oracle_run, rel

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
CASES = ["resnet50_rgb_b2", "adamml_rgb_sound_eval", "adamml_rgb_flow_train", "adamml_rgb_sound_flow_train",
         "adamml_rgb_sound_train", "adamml_rgb_sound_nocausal_train", "adamml_rgb_sound_train_s5",
         "adamml_rgb_sound_eval_s10"]


def split(t):
    """fp32 tensor -> ops.X2 planes (torch restatement of x2_split, csrc/common.cuh)"""
    from adamml_b200 import ops
    hi = t.bfloat16()
    lo = (t - hi.float()).half()
    return ops.X2(hi.contiguous(), lo.contiguous())


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# one x2 store rounds to ~2^-20 relative (1e-6).  The fp32 accumulation in TMEM TRUNCATES (no round-to-nearest), so the
# error of a K-term sum grows ~linearly with K: measured 6e-6 at K = 1152, 1.8e-5 at K = 4608 (layer4's 3x3).
X2_TOL = 4e-6


def x2_tol(K):
    return max(X2_TOL, 8e-9 * K)


# (the last four run the shallow kernel with ONE column block per CTA -- single block or pinned scheduling -- i.e. the
# register-resident statistics: group boundaries inside a tile, > 32 tiles per CTA and group, ragged last tile)
@pytest.mark.parametrize("M,N,K", [(1000, 64, 64), (777, 16, 96), (4096, 96, 16), (513, 128, 256), (2048, 256, 64),
                                   (300, 2048, 512), (129, 24, 144), (640, 1280, 320), (40000, 96, 16),
                                   (80002, 256, 64), (76800, 64, 128), (1515520, 64, 64),
                                   # several column blocks whose weight block does NOT fit the ring: pinned for the
                                   # statistics alone (one flush per BatchNorm group instead of one per tile)
                                   (20002, 1024, 256), (9000, 2048, 512)])
def test_tc_gemm_x2(cuda, M, N, K):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(M + N + K)
    a = split(torch.randn(M, 1, 1, K, generator=g).to(cuda))
    w = torch.randn(N, K, 1, 1, generator=g).to(cuda) / K ** 0.5
    wp = ops.pack_weight(w.contiguous(), ops.PREC_X2)
    G = 2 if M % 2 == 0 else 1
    sums = torch.empty((G, N, 2), device=cuda, dtype=torch.float64)
    z, fused = ops.conv_fwd(a, wp, 1, 0, stats=sums, rows_per_group=M // G)
    assert fused
    # exactly what the kernel multiplies: hi * (b1 + b2 + b3) + lo * fp16(w), accumulated in fp32
    ref = (a.hi.double().view(M, K) @ wp.cascade().view(N, K).t()
           + a.lo.double().view(M, K) @ wp.f16().double().view(N, K).t())
    assert err(z.float().view(M, N), ref) < X2_TOL
    # ... and that is the product of the full-precision operands to ~2^-20
    assert err(z.float().view(M, N), a.float().double().view(M, K) @ w.double().view(N, K).t()) < 1e-5
    zz = z.float().double().view(G, M // G, N)
    assert err(sums[..., 0], zz.sum(1)) < 1e-9 + 1e-6
    assert err(sums[..., 1], (zz * zz).sum(1)) < 1e-6
    # the bf16 cascade carries the parameter to 24 bits, plane 3 is its fp16 rounding
    assert err(wp.cascade().view(N, K), w.view(N, K)) < 2e-7
    assert torch.equal(wp.f16().view(N, K), w.view(N, K).half())


@pytest.mark.parametrize("case", [
    # IMGS,H,W,Cin,Cout,R,stride,pad
    (4, 14, 14, 64, 64, 3, 1, 1),
    (3, 15, 13, 32, 48, 3, 2, 1),
    (2, 28, 28, 128, 128, 3, 2, 1),
    (2, 14, 14, 128, 256, 1, 2, 0),
    (2, 7, 7, 512, 512, 3, 1, 1),
    # enough row blocks for column-block pinning (non-resident weights): 3x3 with 2 column blocks, 1x1/s2 with 8
    (64, 28, 28, 128, 128, 3, 1, 1),
    (48, 28, 28, 128, 512, 1, 2, 0),
])
def test_tc_conv_x2(cuda, case):
    from adamml_b200 import ops
    IMGS, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(sum(case))
    x = split(nhwc(torch.randn(IMGS, Cin, H, W, generator=g)).to(cuda))
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(cuda)
    wp = ops.pack_weight(w.contiguous(), ops.PREC_X2)
    G = 2 if IMGS % 2 == 0 else 1
    sums = torch.empty((G, Cout, 2), device=cuda, dtype=torch.float64)
    Ho, Wo = ops.conv_out_hw(H, W, R, R, stride, pad)
    z, fused = ops.conv_fwd(x, wp, stride, pad, stats=sums, rows_per_group=IMGS * Ho * Wo // G)
    assert fused
    ref = (F.conv2d(nchw(x.hi.double()), wp.cascade().permute(0, 3, 1, 2), None, stride, pad)  # OHWI -> OIHW
           + F.conv2d(nchw(x.lo.double()), wp.f16().double().permute(0, 3, 1, 2), None, stride, pad))
    assert err(nchw(z.float()), ref) < x2_tol(Cin * R * R)
    assert err(nchw(z.float()), F.conv2d(nchw(x.float().double()), w.double(), None, stride, pad)) < 4e-5
    zz = z.float().double().view(G, -1, Cout)
    assert err(sums[..., 0], zz.sum(1)) < 1e-6
    assert err(sums[..., 1], (zz * zz).sum(1)) < 1e-6


@pytest.mark.parametrize("C,R,hw", [(3, 7, 64), (10, 7, 32), (1, 3, 64), (3, 3, 40), (15, 3, 32)])
def test_first_conv_x2(cuda, C, R, hw):
    """ResNet stem (7x7/s2/p3 on the s2d operand written by the data layer) and MobileNetV2 first conv (3x3/s2/p1
    through nhwc_to_s2d) on x2 planes against F.conv2d in float64."""
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(C * R)
    N, S, Fr = 2, 2, 2
    x = torch.randn(N, S * Fr * C, hw, hw, generator=g).to(cuda)
    Cout = 64 if R == 7 else 32
    w = (torch.randn(Cout, C, R, R, generator=g) / (C * R * R) ** 0.5).to(cuda)
    if R == 7:
        xs = ops.pack_frames_s2d(x, S, Fr, C, x2=True)
    else:
        xs = ops.nhwc_to_s2d(ops.pack_frames(x, S, Fr, C, ops.PREC_X2), 3)
    sums = torch.empty((S, Cout, 2), device=cuda, dtype=torch.float64)
    z = ops.stem_conv_fwd(xs, w.contiguous(), stats=sums, imgs_per_group=N * Fr)
    frames = x.view(N, S, Fr, C, hw, hw).transpose(0, 1).reshape(S * N * Fr, C, hw, hw)
    fr = split(frames)
    ref = (F.conv2d(fr.hi.double(), w.double(), None, 2, R // 2)
           + F.conv2d(fr.lo.double(), w.half().double(), None, 2, R // 2))
    assert err(nchw(z.float()), ref) < X2_TOL
    zz = z.float().double().view(S, -1, Cout)
    assert err(sums[..., 0], zz.sum(1)) < 1e-6


@pytest.mark.parametrize("case", [(4, 16, 16, 32), (6, 8, 8, 960), (4, 10, 10, 576), (10, 5, 5, 960), (2, 40, 40, 144),
                                  (2, 128, 128, 32), (4, 33, 19, 16), (6, 21, 35, 48), (12, 16, 16, 384),
                                  (40, 8, 8, 64), (600, 16, 16, 128), (720, 10, 10, 384)])
def test_dwconv_fwd_stats_x2(cuda, case):
    """x2 depthwise stride-1 training forward on TMA tiles with fused BatchNorm statistics: z against a float64
    depthwise conv of the joined planes and against the register-window kernel; sums against float64 sums of z."""
    from adamml_b200 import ops
    IMGS, H, W, C = case
    G = 2
    g = torch.Generator().manual_seed(sum(case))
    x = split(nhwc(torch.randn(IMGS, C, H, W, generator=g) + 0.25).to(cuda))
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(cuda)
    wd = ops.pack_weight_dw(w.contiguous())
    sums = torch.full((G, C, 2), float("nan"), device=cuda, dtype=torch.float64)
    z, fused = ops.dwconv_fwd_stats(x, wd, 1, sums, IMGS // G)
    assert fused
    ref = F.conv2d(nchw(x.float().double()), w.double(), None, 1, 1, groups=C)
    assert err(nchw(z.float()), ref) < X2_TOL
    z_old = ops.dwconv_fwd(x, wd, 1)
    assert err(z.float(), z_old.float()) < X2_TOL
    zz = z.float().double().view(G, -1, C)
    assert err(sums[..., 0], zz.sum(1)) < 1e-6
    assert err(sums[..., 1], (zz * zz).sum(1)) < 1e-6


@pytest.mark.parametrize("case", [(4, 16, 16, 64, 1), (6, 15, 17, 24, 2), (2, 112, 112, 64, 1)])
def test_bn_act_maxpool_x2(cuda, case):
    """training stem: maxpool(act(bn(z))) straight from the pre-BN planes == bn_apply_x2 followed by the x2 max-pool
    (rounding to the planes is monotonic, so the pooled VALUES are identical; only ties may pick another position)"""
    from adamml_b200 import ops
    IMGS, H, W, C, act = case
    G = 2
    g = torch.Generator().manual_seed(sum(case))
    z = split(nhwc(torch.randn(IMGS, C, H, W, generator=g) * 1.5).to(cuda))
    ss = torch.stack([torch.randn(G, C, generator=g), torch.randn(G, C, generator=g)], dim=-1).to(cuda)
    y, pos = ops.bn_act_maxpool_fwd(z, ss, G, act)
    y0, pos0 = ops.maxpool_fwd(ops.bn_apply(z, ss, G, act), want_pos=True)
    # (same values; a value half a bf16 ulp above its hi plane may be re-split as (hi + ulp, -ulp / 2) by the second pass)
    assert torch.equal(y.float(), y0.float())
    assert (pos != pos0).float().mean().item() < 0.02   # ties after rounding (ReLU zeros): any maximal position is valid
    # positions really point at a maximum of the window
    o = nchw(ops.bn_apply(z, ss, G, act).float())
    ref = F.max_pool2d(o, 3, 2, 1)
    assert err(nchw(y.float()), ref) < X2_TOL


def test_resnet_stem_pool_fusion_same_step(cuda):
    """whole ResNet in the default mode with the stem's BN + ReLU applied inside the max-pool kernel vs bn_apply + pool:
    identical logits, gradients equal up to tie-breaking among equal maxima"""
    import importlib
    from adamml_b200 import engine, ops
    ResNet = importlib.import_module("adamml_b200.models.resnet").ResNet
    g = torch.Generator().manual_seed(3)
    net = ResNet(18, 8, num_classes=17, dropout=0.5, input_channels=3, compute_dtype=ops.PREC_X2).to(cuda).train()
    x = torch.randn(6, 24, 64, 64, generator=g).to(cuda)
    mask = torch.empty(6, 512).bernoulli_(0.5, generator=g).div_(0.5).to(cuda)
    dy = torch.randn(6, 17, generator=g).to(cuda)
    res = {}
    old = engine.FUSE_STEM_POOL
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    try:
        for fuse in (True, False):
            engine.FUSE_STEM_POOL = fuse
            net.load_state_dict(sd)
            net.zero_grad(set_to_none=True)
            y = net(x, drop_mask=mask)
            y.backward(dy)
            res[fuse] = (y.detach().clone(), {k: p.grad.clone() for k, p in net.named_parameters()})
    finally:
        engine.FUSE_STEM_POOL = old
    # (equal pooled values can come as different plane pairs -- (hi, +ulp/2) vs (hi + ulp, -ulp/2) -- and the lo plane
    # multiplies the fp16 copy of the weights: logits agree to the x2 rounding level, not bit for bit)
    assert err(res[True][0], res[False][0]) < 2e-4
    for k, gr in res[True][1].items():
        a, b = gr.double(), res[False][1][k].double()
        assert ((a - b).norm() / b.norm().clamp_min(1e-30)).item() < 5e-2, k


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("C", [64, 24, 1024])
def test_bn_mask_bits(cuda, act, C):
    """residual layers: bn_apply_x2 writes a 1-bit activation mask per output element; bn_bwd_reduce with the bits
    (instead of the saved output) gives the same sums and the same in-place masked gradient"""
    from adamml_b200 import ops
    G, rows = 2, 515
    g = torch.Generator().manual_seed(C + act)
    z = split((torch.randn(G * rows, 1, 1, C, generator=g) * 2).to(cuda))
    res = split((torch.randn(G * rows, 1, 1, C, generator=g) * 2).to(cuda))
    ss = torch.stack([torch.rand(G, C, generator=g) + 0.5, torch.randn(G, C, generator=g)], dim=-1).to(cuda)
    mi = torch.stack([torch.randn(G, C, generator=g) * 0.3, torch.rand(G, C, generator=g) + 0.5], dim=-1).to(cuda)
    bits = torch.full((G * rows * C // 8,), 0xAA, device=cuda, dtype=torch.uint8)
    out = ops.bn_apply(z, ss, G, act, res=res, mask_bits=bits)
    out0 = ops.bn_apply(z, ss, G, act, res=res)
    assert torch.equal(out.hi, out0.hi) and torch.equal(out.lo, out0.lo)
    o = out.float().view(-1, C // 8, 8)
    keep = ((o > 0) & (o < 6)) if act == 2 else (o > 0)
    want = (keep.to(torch.int32) << torch.arange(8, device=cuda, dtype=torch.int32)).sum(-1).to(torch.uint8).view(-1)
    assert torch.equal(bits, want)
    dout = torch.randn(G * rows, 1, 1, C, generator=g).to(cuda).bfloat16()
    d1, d2 = dout.clone(), dout.clone()
    # (reference: the mask taken from the x2 value, as the bits are; the hi plane alone differs only where it rounds to 6)
    s2 = ops.bn_bwd_reduce(d2, out.hi, z.hi, mi, G, act, gm_inplace=True)
    s1 = ops.bn_bwd_reduce(d1, None, z.hi, mi, G, act, gm_inplace=True, mask_bits=bits)
    gm_want = torch.where(keep.view(dout.shape), dout, torch.zeros_like(dout))
    assert torch.equal(d1, gm_want)
    same = (d1 == d2).float().mean().item()
    assert same > 0.995
    if same == 1.0:
        assert err(s1, s2) < 1e-12


def test_bn_apply_and_stats_x2(cuda):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(3)
    G, rows, C = 2, 500, 96
    z = split(torch.randn(G * rows, 1, 1, C, generator=g).to(cuda) * 3 + 1)
    res = split(torch.randn(G * rows, 1, 1, C, generator=g).to(cuda))
    rz = split(torch.randn(G * rows, 1, 1, C, generator=g).to(cuda))
    ss = torch.randn(G, C, 2, generator=g).to(cuda)
    rss = torch.randn(G, C, 2, generator=g).to(cuda)
    zf, rf, rzf = (t.float().double().view(G, rows, C) for t in (z, res, rz))
    lin = zf * ss[:, None, :, 0].double() + ss[:, None, :, 1].double()
    for act, fn in ((ops.ACT_NONE, lambda t: t), (ops.ACT_RELU, torch.relu), (ops.ACT_RELU6, lambda t: t.clamp(0, 6))):
        out = ops.bn_apply(z, ss, G, act)
        assert err(out.float().view(G, rows, C), fn(lin)) < 2e-6
        out = ops.bn_apply(z, ss, G, act, res=res)
        assert err(out.float().view(G, rows, C), fn(lin + rf)) < 2e-6
        out = ops.bn_apply(z, ss, G, act, res_z=rz, res_ss=rss)
        ref = fn(lin + rzf * rss[:, None, :, 0].double() + rss[:, None, :, 1].double())
        assert err(out.float().view(G, rows, C), ref) < 2e-6
    sums = ops.bn_stats(z, G)
    assert err(sums[..., 0], zf.sum(1)) < 1e-6
    assert err(sums[..., 1], (zf * zf).sum(1)) < 1e-6


@pytest.mark.parametrize("stride", [1, 2])
def test_dwconv_fwd_x2(cuda, stride):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(stride)
    IMGS, H, W, C = 3, 19, 22, 96
    x = split(nhwc(torch.randn(IMGS, C, H, W, generator=g)).to(cuda))
    w = torch.randn(C, 1, 3, 3, generator=g).to(cuda) / 3
    y = ops.dwconv_fwd(x, ops.pack_weight_dw(w.contiguous()), stride)
    ref = F.conv2d(nchw(x.float().double()), w.double(), None, stride, 1, 1, C)
    assert err(nchw(y.float()), ref) < 2e-6


def test_pools_x2(cuda):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(9)
    IMGS, H, W, C = 8, 14, 14, 64
    xf = nhwc(torch.randn(IMGS, C, H, W, generator=g)).to(cuda)
    x = split(xf)
    xr = nchw(x.float())
    y, pos = ops.maxpool_fwd(x, want_pos=True)
    assert torch.equal(nchw(y.float()), F.max_pool2d(xr, 3, 2, 1))
    # recorded positions drive the bf16 backward: same as the positions of a bf16 pool wherever the hi planes decide
    dy = torch.randn(y.shape, generator=g).to(cuda).bfloat16()
    dx = ops.maxpool_bwd(None, dy, pos=pos, x_shape=tuple(x.shape))
    xr2 = xr.clone().requires_grad_(True)
    F.max_pool2d(xr2, 3, 2, 1).backward(nchw(dy.float()))
    assert err(nchw(dx.float()), xr2.grad) < 1e-2
    for T in (2, 4, 8):
        yt = ops.tpool_fwd(x, T)
        v = xr.view(IMGS // T, T, C, H, W).transpose(1, 2)
        ref = F.max_pool3d(v, (3, 1, 1), (2, 1, 1), (1, 0, 0)).transpose(1, 2).reshape(-1, C, H, W)
        assert torch.equal(nchw(yt.float()), ref)
    f = ops.avgpool_fwd(x)
    assert err(f, xr.double().mean((2, 3))) < 1e-6


def test_data_layer_x2(cuda):
    from adamml_b200 import ops
    g = torch.Generator().manual_seed(4)
    N, S, Fr, C, hw = 2, 2, 4, 3, 48
    x = torch.randn(N, S * Fr * C, hw, hw, generator=g).to(cuda)
    p = ops.pack_frames(x, S, Fr, C, ops.PREC_X2)
    frames = x.view(N, S, Fr, C, hw, hw).transpose(0, 1).reshape(S * N * Fr, C, hw, hw)
    assert err(nchw(p.float()), frames) < 2e-6
    r = ops.resize_frames(x, S, Fr, C, 32, 32, 2, ops.PREC_X2)
    ref = F.interpolate(x, size=(32, 32), mode="bilinear").view(N, S, Fr, C, 32, 32)[:, :, ::2]
    ref = ref.transpose(0, 1).reshape(-1, C, 32, 32)
    assert err(nchw(r.float()), ref) < 2e-6
    # the planes are exactly the x2 split of the fp32 result
    r32 = ops.resize_frames(x, S, Fr, C, 32, 32, 2, torch.float32)
    s = split(r32)
    assert torch.equal(r.hi, s.hi) and torch.equal(r.lo, s.lo)


# ------------------------------------------------------------------------------------------ whole-model parity
def build(case, cuda, dtype=None):
    from adamml_b200 import ops
    from adamml_b200.models import build_model
    model, arch = build_model(namespace(case, compute_dtype=dtype or ops.PREC_X2))
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    model.load_state_dict(O.fill_state_dict(shapes, seed=0), strict=True)
    return model.to(cuda), arch


def run_product(model, case, g_seed, cuda):
    from test_parity_gpu import run_product as rp
    return rp(model, case, g_seed, cuda)


def cosine(a, b):
    a, b = a.double().flatten().cpu(), b.double().flatten().cpu()
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


@pytest.mark.parametrize("name", CASES)
def test_default_mode_matches_reference_golden(cuda, name):
    """THE parity gate of the mode bench.py measures: logits within 1e-3 of the reference's recorded logits and
    policy selections bit-exact, on every golden (incl. the S=5 train / S=10 eval goldens at the benchmark's S)."""
    from adamml_b200.models.resnet import default_compute_dtype
    from adamml_b200 import ops
    assert default_compute_dtype() == ops.PREC_X2, "the default precision mode must be the parity-passing one"
    g = load_golden(name)
    case = g["case"]
    model, arch = build(case, cuda)
    assert arch == g["arch"]
    logits, dec, loss = run_product(model, case, g["seed"], cuda)
    e = rel(logits, g["logits"])
    print(f"x2 {name}: logits rel {e:.2e}")
    assert e < LOGIT_TOL
    if dec is not None:
        assert torch.equal(dec.detach().cpu(), g["decisions"]), "policy selections must be bit-exact"
    assert abs(loss.item() - g["loss"].item()) < 1e-3 * max(1.0, abs(g["loss"].item()))
    if not case["training"]:
        return
    loss.backward()
    torch.cuda.synchronize()
    sd = model.state_dict()
    from util import fingerprint
    for k, fp in g["running_fp"].items():
        assert rel(fingerprint(sd[k])[1], fp[1]) < 1e-4, k
    for k, v in g["num_batches_tracked"].items():
        assert int(sd[k]) == v, k


def test_default_mode_gradients_track_oracle(cuda):
    """Backward of the default mode = bf16 engine on the hi planes of an fp32-class forward.  On a reasonably
    conditioned case (8 videos) every parameter gradient must point the same way as the fp32 oracle's
    (per-tensor cosine), and the direction of the whole gradient must agree to bf16 noise."""
    case = dict(kind="adamml", modality=["rgb", "sound"], N=8, S=2, hw=64, training=True)
    model, _ = build(case, cuda)
    logits, dec, loss = run_product(model, case, 5, cuda)
    loss.backward()
    shapes = {k: v.shape for k, v in model.state_dict().items()}
    sd0 = O.fill_state_dict(shapes, seed=0)
    o_logits, o_dec, g32, _ = oracle_run(case, 5, torch.float32, sd0)
    assert rel(logits, o_logits) < LOGIT_TOL
    assert torch.equal(dec.detach().cpu(), o_dec)
    grads = {k: p.grad for k, p in model.named_parameters()}
    assert set(grads) == set(g32)
    # tensors whose true gradient is mathematically zero (e.g. the beta of a linear-bottleneck BN that feeds a
    # train-mode BN, see util.compare_grads) hold round-off noise on both sides: leave them out of the per-tensor check
    import statistics as st

    def gkey(k):
        parts = k.split(".")
        return ".".join(parts[:parts.index("nets") + 2]) if "nets" in parts else parts[0]
    med = {}
    for k, v in g32.items():
        med.setdefault(gkey(k), []).append(v.abs().max().item())
    med = {k: st.median(v) for k, v in med.items()}
    cos = {k: cosine(grads[k], g32[k]) for k in g32
           if g32[k].numel() > 8 and g32[k].abs().max().item() > 0.02 * med[gkey(k)]}
    worst = sorted(cos.items(), key=lambda kv: kv[1])[:5]
    flat_p = torch.cat([grads[k].double().flatten().cpu() for k in sorted(g32)])
    flat_o = torch.cat([g32[k].double().flatten() for k in sorted(g32)])
    total = cosine(flat_p, flat_o)
    import statistics
    print(f"x2 gradient cosine vs fp32 oracle: whole model {total:.5f}, median tensor "
          f"{statistics.median(cos.values()):.5f}, worst {worst}")
    assert total > 0.99
    assert statistics.median(cos.values()) > 0.99
    assert sum(1 for v in cos.values() if v < 0.9) <= len(cos) // 20, worst


def test_x2_eval_skip_bit_identical(cuda):
    """decision-driven skipping (AdaMML._forward_selected) on x2 planes: bit-identical to run-everything"""
    case = dict(kind="adamml", modality=["rgb", "sound"], N=4, S=2, S_run=3, hw=64, training=False)
    model, _ = build(case, cuda)
    model.eval()
    cfg = O.make_cfg(case["modality"], num_segments=case["S"])
    xs, _ = O.make_inputs(cfg, 4, 3, hw=64)
    xs = [x.to(cuda) for x in xs]
    noise = noise_for_model(O.draw_noise(3, cfg, 4, 3, False), cuda)
    outs = []
    with torch.no_grad():
        for skip in (True, False):
            model.skip_unselected = skip
            outs.append(model(xs, num_segments=3, noise=noise))
    assert torch.equal(outs[0][1], outs[1][1])
    assert torch.equal(outs[0][0], outs[1][0])


@pytest.mark.parametrize("mode", ["x2", "bf16"])
def test_recompute_mode_same_gradients_less_memory(cuda, mode):
    """ADAMML_B200_RECOMPUTE: outputs of layers without residual input are not kept for backward; the consumer's
    weight gradient rebuilds them from the saved pre-BN tensor.  Same logits / selections (forward is untouched),
    gradients equal to bf16 rounding, and a clearly smaller peak."""
    from adamml_b200 import engine, ops
    case = dict(kind="adamml", modality=["rgb", "sound"], N=4, S=2, hw=64, training=True)
    res = {}
    old, old_fuse = engine.RECOMPUTE, engine.DW_FUSE_PRE
    # (the recompute machinery is compared with the fused depthwise-backward BN reduction off: that kernel takes the
    # ReLU6 mask of the expand layer from its saved output, which recompute mode rebuilds from the bf16 pre-BN tensor --
    # a different, equally legitimate mask for the ~0.1 % of elements next to a threshold; the fused path has its
    # own tests in test_blocks_gpu.py / test_kernels_gpu.py and runs in every whole-model test)
    engine.DW_FUSE_PRE = False
    try:
        for rc in (False, True):
            engine.RECOMPUTE = rc
            model, _ = build(case, cuda, dtype=ops.PREC_X2 if mode == "x2" else torch.bfloat16)
            torch.cuda.synchronize()
            torch.cuda.reset_peak_memory_stats()
            base = torch.cuda.memory_allocated()
            logits, dec, loss = run_product(model, case, 5, cuda)
            fwd_peak = torch.cuda.max_memory_allocated() - base
            held = torch.cuda.memory_allocated() - base          # what the tape keeps alive for backward
            loss.backward()
            torch.cuda.synchronize()
            res[rc] = (logits.detach(), dec.detach(), {k: p.grad.clone() for k, p in model.named_parameters()}, held,
                       fwd_peak)
            del model, logits, dec, loss
    finally:
        engine.RECOMPUTE, engine.DW_FUSE_PRE = old, old_fuse
    assert torch.equal(res[True][0], res[False][0]) and torch.equal(res[True][1], res[False][1])
    from util import compare_grads
    bad = compare_grads(res[True][2], res[False][2], tol=3e-2)
    assert len(bad) <= len(res[True][2]) // 50, bad[:5]
    print(f"{mode}: tape memory {res[False][3] / 2**20:.0f} -> {res[True][3] / 2**20:.0f} MiB")
    assert res[True][3] < 0.8 * res[False][3]


def test_weight_pack_cache_same_training_trajectory(cuda):
    """AdaMML.forward refreshes every weight operand with ONE multi-tensor launch (ops.WeightPackCache,
    adamml_pack_weights_multi) instead of one launch per layer and kind.  Over three SGD steps (weights change between
    the passes): the first pass runs and records the per-layer launches, later passes launch ONE pack kernel, every
    operand in the arena is bit-identical to the per-layer conversion of the CURRENT parameter, and the trajectory
    matches the per-layer path (first pass to rounding noise of the statistics atomics, later passes to the run-to-run
    divergence of the atomically accumulated weight gradients -- a stale operand would be off by tens of percent)."""
    import copy
    import importlib
    from adamml_b200 import _lib, ops
    adamml_mod = importlib.import_module("adamml_b200.models.adamml")
    g = load_golden("adamml_rgb_sound_train")
    case = g["case"]
    model0, _ = build(case, cuda)
    cfg = O.make_cfg(case["modality"], num_segments=case["S"])
    xs, y = O.make_inputs(cfg, case["N"], case["S"], hw=case["hw"])
    xs, y = [t.to(cuda) for t in xs], y.to(cuda)
    noise = noise_for_model(O.draw_noise(1, cfg, case["N"], case["S"], True), cuda)
    results = {}
    old = adamml_mod.PACK_CACHE
    try:
        for on in (False, True):
            adamml_mod.PACK_CACHE = on
            model = copy.deepcopy(model0).train()
            opt = torch.optim.SGD(model.parameters(), 0.05)
            trace, packs = [], []
            for step in range(3):
                _lib.PROFILE = []
                opt.zero_grad(set_to_none=True)
                logits, dec = model(xs, noise=noise)
                F.cross_entropy(logits, y).backward()
                opt.step()
                names = [n for n, *_ in _lib.PROFILE]
                _lib.PROFILE = None
                packs.append((sum(n in ("pack_weight", "pack_weight_x2", "pack_weight_dgrad", "pack_weight_dw")
                                  for n in names), names.count("pack_weights_multi")))
                trace.append((logits.detach().clone(), dec.detach().clone()))
            results[on] = (trace, packs)
            if on:  # the arena against the per-layer conversions of the parameters as they are NOW (after 3 updates)
                cache = model.__dict__["_pack_cache"]
                cache.begin()
                kinds = set()
                for e in cache.ent.values():
                    assert e.in_table
                    if e.kind == ops.PK_OHWI_X2:
                        ref = ops.pack_weight(e.w, ops.PREC_X2).planes
                    elif e.kind in (ops.PK_OHWI_BF16, ops.PK_OHWI_F32):
                        ref = ops.pack_weight(e.w, e.dtype)
                    elif e.kind in (ops.PK_DGRAD_BF16, ops.PK_DGRAD_F32):
                        ref = ops.pack_weight_dgrad(e.w, e.dtype)
                    else:
                        ref = ops.pack_weight_dw(e.w)
                    assert torch.equal(e.out.view(torch.int16 if e.out.element_size() == 2 else torch.int32),
                                       ref.view(torch.int16 if ref.element_size() == 2 else torch.int32)), e.kind
                    kinds.add(e.kind)
                assert {ops.PK_OHWI_X2, ops.PK_DGRAD_BF16, ops.PK_DW} <= kinds
                # operands nobody asks for any more (here: the data-gradient operands during forward-only passes) are
                # pruned, and come back with the next backward pass
                cache.PRUNE_EVERY = 2
                n_all = len(cache.ent)
                with torch.no_grad():
                    for _ in range(6):
                        lg, _ = model(xs, noise=noise)
                assert 0 < len(cache.ent) < n_all
                assert not any(e.kind in (ops.PK_DGRAD_BF16, ops.PK_DGRAD_F32) for e in cache.ent.values())
                opt.zero_grad(set_to_none=True)
                lg2, _ = model(xs, noise=noise)
                assert rel(lg2, lg) < 1e-5
                F.cross_entropy(lg2, y).backward()
                assert any(e.kind == ops.PK_DGRAD_BF16 for e in cache.ent.values())
                assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    finally:
        adamml_mod.PACK_CACHE = old
        _lib.PROFILE = None
    (t0, p0), (t1, p1) = results[False], results[True]
    assert rel(t1[0][0], t0[0][0]) < 1e-5 and torch.equal(t0[0][1], t1[0][1])
    for (l0, d0), (l1, d1) in zip(t0[1:], t1[1:]):
        assert rel(l1, l0) < 2e-2
    assert all(m == 0 for _, m in p0) and p0[0][0] > 50
    assert p1[0] == (p0[0][0], 0), (p1, p0)          # first pass: per-layer launches, recorded
    # then ONE launch per pass (+ the space-to-depth stem operands of the first convolutions, packed by stem_conv_fwd)
    assert p1[1] == p1[2] and p1[1][1] == 1 and p1[1][0] <= 4, p1
