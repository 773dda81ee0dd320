"""Per-kernel numerics: every C-ABI entry point against the torch fp32 op it replaces."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


TOL = {torch.float32: 2e-5, torch.bfloat16: 2e-2}

CONV_CASES = [
    # IMGS,H,W,Cin,Cout,R,stride,pad
    (2, 14, 14, 64, 64, 1, 1, 0),
    (2, 14, 14, 64, 128, 3, 1, 1),
    (3, 15, 13, 32, 48, 3, 2, 1),
    (2, 20, 20, 3, 64, 7, 2, 3),
    (2, 16, 16, 1, 32, 3, 2, 1),
    (2, 14, 14, 128, 256, 1, 2, 0),
    (5, 1, 1, 2052, 31, 1, 1, 0),
    (4, 9, 9, 24, 144, 1, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_simt_conv_fwd_dgrad_wgrad(cuda, case, dtype):
    from adamml_b200 import ops
    ops.TC_MODE = "simt"
    IMGS, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(IMGS, Cin, H, W, generator=g).to(cuda)
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(cuda)
    if dtype == torch.bfloat16:  # make inputs exactly representable so only accumulation differs
        x = x.bfloat16().float()
        w = w.bfloat16().float()
    x.requires_grad_(True)
    w.requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride, pad)
    dy = torch.randn(y_ref.shape, generator=g).to(cuda)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y_ref.backward(dy)

    xn = nhwc(x.detach()).to(dtype)
    wp = ops.pack_weight(w.detach().contiguous(), dtype)
    y, _ = ops.conv_fwd(xn, wp, stride, pad)
    assert relerr(nchw(y), y_ref) < TOL[dtype]
    dyn = nhwc(dy).to(dtype)
    dx = ops.conv_dgrad(dyn, wp, tuple(xn.shape), stride, pad)
    assert relerr(nchw(dx), x.grad) < TOL[dtype]
    add = torch.randn(xn.shape, generator=g).to(cuda).to(dtype)
    dx2 = ops.conv_dgrad(dyn, wp, tuple(xn.shape), stride, pad, addend=add)
    assert relerr(dx2.float(), dx.float() + add.float()) < TOL[dtype]
    dw = ops.conv_wgrad(xn, dyn, tuple(wp.shape), stride, pad)
    dw_oihw = ops.unpack_wgrad(dw, Cin)
    assert relerr(dw_oihw, w.grad) < 5e-5 if dtype == torch.float32 else relerr(dw_oihw, w.grad) < 1e-3
    ops.TC_MODE = "auto"


TC_CASES = [
    # M, N, K
    (128, 64, 64),
    (256, 128, 64),
    (1000, 256, 128),
    (3136 * 2, 64, 256),
    (777, 24, 144),
    # narrow / shallow MobileNetV2 pointwise shapes: linear bulk-store epilogue, two-CTA shallow-K config
    (1000, 16, 32),
    (5000, 96, 16),
    (300, 32, 16),
    (777, 144, 24),
    (129, 8, 8),
    (513, 1280, 320),
    (4096, 512, 1024),
    (130, 16, 32),
    (50000, 256, 64),
    (300, 2048, 512),
]


@pytest.mark.parametrize("case", TC_CASES)
def test_tc_gemm(cuda, case):
    from adamml_b200 import _lib
    M, N, K = case
    g = torch.Generator(device="cpu").manual_seed(2)
    A = torch.randn(M, K, generator=g).to(cuda).bfloat16()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).bfloat16()
    D = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    _lib.call("tc_gemm_bf16", A, B, D, M, N, K, 0, 0, 0, _lib.BF16, None, 0)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    assert torch.isfinite(D.float()).all()
    assert relerr(D, ref) < 1e-2
    # fp32 output is outside the envelope (the epilogue stages bf16 tiles): reported, never silently converted
    Df = torch.empty((M, N), device=cuda, dtype=torch.float32)
    assert _lib.call("tc_gemm_bf16", A, B, Df, M, N, K, 0, 0, 0, _lib.F32, None, 0, allow_unsupported=True) == 3


@pytest.mark.parametrize("case", [(3136 * 4, 64, 64, 3136 * 2), (1000, 256, 128, 250), (6 * 98, 512, 256, 98),
                                  (777, 24, 144, 777), (1000, 16, 32, 250), (5000, 96, 16, 1000), (768, 144, 24, 256),
                                  (6 * 640, 192, 32, 640)])
def test_tc_gemm_bn_stats(cuda, case):
    from adamml_b200 import _lib
    M, N, K, rpg = case
    G = M // rpg
    g = torch.Generator(device="cpu").manual_seed(3)
    A = (torch.randn(M, K, generator=g) + 0.3).to(cuda).bfloat16()
    B = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda).bfloat16()
    D = torch.empty((M, N), device=cuda, dtype=torch.bfloat16)
    stats = torch.full((G, N, 2), float("nan"), device=cuda, dtype=torch.float64)
    _lib.call("tc_gemm_bf16", A, B, D, M, N, K, 0, 0, 0, _lib.BF16, stats, rpg)
    torch.cuda.synchronize()
    Dd = D.double().view(G, rpg, N)
    ref = torch.stack([Dd.sum(1), (Dd * Dd).sum(1)], dim=-1)
    assert relerr(stats, ref) < 1e-5


DW_CASES = [(2, 16, 16, 32, 1), (3, 15, 17, 96, 2), (2, 8, 8, 960, 1), (1, 10, 10, 144, 2)]


@pytest.mark.parametrize("case", DW_CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_dwconv(cuda, case, dtype):
    from adamml_b200 import ops
    IMGS, H, W, C, stride = case
    g = torch.Generator(device="cpu").manual_seed(4)
    x = torch.randn(IMGS, C, H, W, generator=g).to(cuda)
    w = torch.randn(C, 1, 3, 3, generator=g).to(cuda) / 3
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    x.requires_grad_(True); w.requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride, 1, groups=C)
    dy = torch.randn(y_ref.shape, generator=g).to(cuda)
    if dtype == torch.bfloat16:
        dy = dy.bfloat16().float()
    y_ref.backward(dy)
    xn = nhwc(x.detach()).to(dtype)
    wd = ops.pack_weight_dw(w.detach().contiguous())   # tap-major [9, C]
    assert torch.equal(wd, w.detach().view(C, 9).t().contiguous())
    y = ops.dwconv_fwd(xn, wd, stride)
    assert relerr(nchw(y), y_ref) < TOL[dtype]
    dyn = nhwc(dy).to(dtype)
    dx = ops.dwconv_dgrad(dyn, wd, tuple(xn.shape), stride)
    assert relerr(nchw(dx), x.grad) < TOL[dtype]
    dw = ops.dwconv_wgrad(xn, dyn, stride)
    assert relerr(dw, w.grad) < 5e-5


# fused depthwise backward on TMA tiles (csrc/dwconv_tma.cu): every channel-chunk width (64 | 48 | 32 | 16), both
# strides, row-tile heights 8 | 5 (stride 1) and 4 | 5 (stride 2), ragged widths / heights, several tiles per CTA,
# images-per-tile > 1 and image counts that do not fill the last tile
DW_BWD_CASES = [(2, 16, 16, 32, 1), (3, 15, 17, 96, 2), (2, 8, 8, 960, 1), (1, 10, 10, 144, 2), (3, 10, 10, 576, 1),
                (5, 5, 5, 960, 1), (2, 40, 40, 144, 1), (1, 128, 128, 32, 1), (2, 64, 64, 96, 2), (7, 20, 20, 192, 2),
                (3, 33, 19, 16, 1), (2, 21, 35, 48, 2), (9, 16, 16, 384, 1), (40, 8, 8, 64, 1), (37, 10, 10, 64, 2),
                (2, 80, 80, 96, 2), (300, 16, 16, 128, 1), (720, 10, 10, 384, 1), (90, 10, 10, 576, 2)]


@pytest.mark.parametrize("case", DW_BWD_CASES)
def test_dwconv_bwd_fused(cuda, case):
    """dx and dw of ONE adamml_dwconv_bwd launch vs torch autograd (fp32 math on the bf16-rounded operands) and vs
    the separate dgrad / wgrad kernels."""
    from adamml_b200 import ops
    IMGS, H, W, C, stride = case
    g = torch.Generator(device="cpu").manual_seed(sum(case))
    x = torch.randn(IMGS, C, H, W, generator=g).to(cuda).bfloat16().float()
    w = torch.randn(C, 1, 3, 3, generator=g).to(cuda) / 3
    x.requires_grad_(True); w.requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride, 1, groups=C)
    dy = torch.randn(y_ref.shape, generator=g).to(cuda).bfloat16().float()
    y_ref.backward(dy)
    xn, dyn = nhwc(x.detach()).bfloat16(), nhwc(dy).bfloat16()
    wd = ops.pack_weight_dw(w.detach().contiguous())
    assert ops.dwconv_bwd_ok(xn, dyn, stride)
    dx, dw = ops.dwconv_bwd(xn, dyn, wd, stride)
    assert relerr(nchw(dx), x.grad) < TOL[torch.bfloat16]
    assert relerr(dw, w.grad) < 5e-5
    dx2 = ops.dwconv_dgrad(dyn, wd, tuple(xn.shape), stride)
    # same products, different summation order: equal up to one bf16 rounding of a few elements
    assert relerr(dx.float(), dx2.float()) < 2 ** -7
    assert (dx != dx2).float().mean().item() < 0.02


@pytest.mark.parametrize("case", [(4, 16, 16, 32, 1, 2), (6, 15, 17, 96, 2, 2), (4, 8, 8, 960, 1, 1),
                                  (10, 10, 10, 144, 2, 2), (12, 10, 10, 576, 1, 2), (10, 5, 5, 960, 1, 1),
                                  (4, 40, 40, 144, 1, 2), (6, 64, 64, 96, 2, 2), (80, 8, 8, 64, 1, 2),
                                  (600, 16, 16, 128, 1, 1), (330, 12, 12, 48, 2, 2)])
def test_dwconv_bwd_fused_producer_reduce(cuda, case):
    """adamml_dwconv_bwd with the fused BatchNorm-backward reduction of the layer that produced x = act(bn(z)):
    dx comes back masked, and bn_sums_from_out of the raw sums equals what bn_bwd_reduce computes from (dx, z) in
    a separate pass (same mask source: the saved output)."""
    from adamml_b200 import ops
    IMGS, H, W, C, stride, act = case
    G = 2
    g = torch.Generator(device="cpu").manual_seed(sum(case))
    z = (torch.randn(IMGS, H, W, C, generator=g) * 1.5 + 0.3).to(cuda).bfloat16()
    ss = torch.stack([torch.rand(G, C, generator=g) + 0.5, torch.randn(G, C, generator=g)], dim=-1).to(cuda)
    mi = torch.stack([torch.randn(G, C, generator=g) * 0.3, torch.rand(G, C, generator=g) + 0.5], dim=-1).to(cuda)
    x = ops.bn_apply(z, ss, G, act)                       # the producer's saved output (bf16)
    Ho, Wo = ops.conv_out_hw(H, W, 3, 3, stride, 1)
    dy = torch.randn(IMGS, Ho, Wo, C, generator=g).to(cuda).bfloat16()
    wd = ops.pack_weight_dw((torch.randn(C, 1, 3, 3, generator=g) / 3).to(cuda).contiguous())
    dx0, dw0 = ops.dwconv_bwd(x, dy, wd, stride)
    raw = torch.full((G, C, 2), float("nan"), device=cuda, dtype=torch.float64)
    gm, dw1 = ops.dwconv_bwd(x, dy, wd, stride, pre=(raw, IMGS // G, act))
    assert relerr(dw1, dw0) < 1e-5   # (fp32 atomics across CTAs: same products, order not fixed)
    xf = x.float()
    keep = (xf > 0) & (xf < 6) if act == 2 else (xf > 0)
    assert torch.equal(gm, torch.where(keep, dx0, torch.zeros_like(dx0)))
    # raw sums: fp32 products of the UNROUNDED gradient, so compare against the rounded one at bf16-sum accuracy
    gmd, xd = gm.double().view(G, -1, C), xf.double().view(G, -1, C)
    assert relerr(raw[..., 0], gmd.sum(1)) < 5e-3
    assert relerr(raw[..., 1], (gmd * xd).sum(1)) < 5e-3
    sums = ops.bn_sums_from_out(raw.clone(), ss, mi)
    want = ops.bn_bwd_reduce(dx0.clone(), x, z, mi, G, act)
    assert relerr(sums[..., 0], want[..., 0]) < 5e-3
    # sum gm * xhat: z is only known through the bf16 output here (one extra rounding per element)
    scale = (gmd.abs() * ((z.double().view(G, -1, C) - mi[..., 0].double().unsqueeze(1)) * mi[..., 1].double().unsqueeze(1)).abs()).sum(1)
    assert ((sums[..., 1] - want[..., 1]).abs() / scale.clamp_min(1e-30)).max().item() < 4e-3


@pytest.mark.parametrize("act", [0, 1, 2])
@pytest.mark.parametrize("residual", ["none", "plain", "bn"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_bn_train_fwd_bwd(cuda, act, residual, dtype):
    """G per-segment groups == G separate nn.BatchNorm2d calls in segment order."""
    from adamml_b200 import ops
    G, ipg, H, W, C = 3, 2, 6, 5, 40
    g = torch.Generator(device="cpu").manual_seed(5)
    z = (torch.randn(G * ipg, C, H, W, generator=g) * 2 + 0.5).to(cuda)
    z2 = (torch.randn(G * ipg, C, H, W, generator=g) * 1.5 - 0.2).to(cuda)
    res = torch.randn(G * ipg, C, H, W, generator=g).to(cuda)
    if dtype == torch.bfloat16:
        z, z2, res = z.bfloat16().float(), z2.bfloat16().float(), res.bfloat16().float()
    bn = torch.nn.BatchNorm2d(C).to(cuda)
    bn2 = torch.nn.BatchNorm2d(C).to(cuda)
    with torch.no_grad():
        for b in (bn, bn2):
            b.weight.copy_(torch.rand(C, generator=g) + 0.5)
            b.bias.copy_(torch.randn(C, generator=g) * 0.3)
            b.running_mean.copy_(torch.randn(C, generator=g) * 0.1)
            b.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    rm20, rv20 = bn2.running_mean.clone(), bn2.running_var.clone()
    z.requires_grad_(True); z2.requires_grad_(True); res.requires_grad_(True)
    outs = []
    for s in range(G):
        sl = slice(s * ipg, (s + 1) * ipg)
        o = bn(z[sl])
        if residual == "plain":
            o = o + res[sl]
        elif residual == "bn":
            o = o + bn2(z2[sl])
        if act == 1:
            o = F.relu(o)
        elif act == 2:
            o = F.relu6(o)
        outs.append(o)
    out_ref = torch.cat(outs)
    dout = torch.randn(out_ref.shape, generator=g).to(cuda)
    if dtype == torch.bfloat16:
        dout = dout.bfloat16().float()
    out_ref.backward(dout)

    zn, z2n, resn = nhwc(z.detach()).to(dtype), nhwc(z2.detach()).to(dtype), nhwc(res.detach()).to(dtype)
    count = ipg * H * W
    rm, rv = rm0.clone(), rv0.clone()
    sums = ops.bn_stats(zn, G)
    mi, ss = ops.bn_finalize(sums, bn.weight.detach(), bn.bias.detach(), rm, rv, count, 0.1, 1e-5, C, G, True, True)
    mi2 = ss2 = None
    if residual == "bn":
        rm2, rv2 = rm20.clone(), rv20.clone()
        sums2 = ops.bn_stats(z2n, G)
        mi2, ss2 = ops.bn_finalize(sums2, bn2.weight.detach(), bn2.bias.detach(), rm2, rv2, count, 0.1, 1e-5, C, G,
                                   True, True)
        assert relerr(rm2, bn2.running_mean) < 1e-5 and relerr(rv2, bn2.running_var) < 1e-5
    out = ops.bn_apply(zn, ss, G, act, res=resn if residual == "plain" else None,
                       res_z=z2n if residual == "bn" else None, res_ss=ss2)
    tol = TOL[dtype]
    assert relerr(nchw(out), out_ref) < tol
    assert relerr(rm, bn.running_mean) < 1e-5 and relerr(rv, bn.running_var) < 1e-5
    if dtype == torch.bfloat16:
        return  # backward masks depend on rounded outputs; fp32 covers the math
    doutn = nhwc(dout)
    bs = ops.bn_bwd_reduce(doutn, out, zn, mi, G, act)
    dz, dres = ops.bn_bwd_apply(doutn, out, zn, mi, bn.weight.detach(), bs, G, count, act, True, True, True)
    dgamma, dbeta = ops.bn_param_grad(bs, C, G)
    assert relerr(nchw(dz), z.grad) < 1e-4
    assert relerr(dgamma, bn.weight.grad) < 1e-4 and relerr(dbeta, bn.bias.grad) < 1e-4
    if residual == "plain":
        assert relerr(nchw(dres), res.grad) < 1e-5
    if residual != "none" and act != 0:
        # in-place masking: the reduce pass turns dout into gm = dout * act'(out); apply then needs no mask / out
        gm = doutn.clone()
        bs_i = ops.bn_bwd_reduce(gm, out, zn, mi, G, act, gm_inplace=True)
        assert relerr(bs_i, bs) < 1e-12 and torch.equal(gm, dres)
        dz_i, _ = ops.bn_bwd_apply(gm, None, zn, mi, bn.weight.detach(), bs_i, G, count, 0, True, True, False)
        assert torch.equal(dz_i, dz)
    if residual == "none" and act != 0:
        # mask recomputed from z with the forward scale/shift: `out` is never read (None)
        bs_z = ops.bn_bwd_reduce(doutn, None, zn, mi, G, act, mask_ss=ss)
        assert relerr(bs_z, bs) < 1e-12
        dz_z, _ = ops.bn_bwd_apply(doutn, None, zn, mi, bn.weight.detach(), bs_z, G, count, act, True, True, False,
                                   mask_ss=ss)
        assert torch.equal(dz_z, dz)
    if residual == "bn":
        bs2 = ops.bn_bwd_reduce(doutn, out, z2n, mi2, G, act)
        dz2, _ = ops.bn_bwd_apply(doutn, out, z2n, mi2, bn2.weight.detach(), bs2, G, count, act, True, True, False)
        assert relerr(nchw(dz2), z2.grad) < 1e-4


def test_bn_eval(cuda):
    from adamml_b200 import ops
    G, ipg, H, W, C = 2, 3, 4, 4, 24
    g = torch.Generator(device="cpu").manual_seed(6)
    z = torch.randn(G * ipg, C, H, W, generator=g).to(cuda)
    bn = torch.nn.BatchNorm2d(C).to(cuda).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5); bn.bias.copy_(torch.randn(C, generator=g))
        bn.running_mean.copy_(torch.randn(C, generator=g)); bn.running_var.copy_(torch.rand(C, generator=g) + 0.5)
    ref = F.relu6(bn(z))
    mi, ss = ops.bn_finalize(None, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, 1, 0.1, 1e-5,
                             C, G, False, False)
    out = ops.bn_apply(nhwc(z), ss, G, 2)
    assert relerr(nchw(out), ref) < 1e-5


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pools(cuda, dtype):
    from adamml_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(7)
    x = torch.randn(4, 16, 13, 12, generator=g).to(cuda)
    x = F.relu(x)  # many exact ties at zero, like the post-ReLU stem
    if dtype == torch.bfloat16:
        x = x.bfloat16().float()
    x.requires_grad_(True)
    y_ref = F.max_pool2d(x, 3, 2, 1)
    dy = torch.randn(y_ref.shape, generator=g).to(cuda)
    y_ref.backward(dy)
    xn = nhwc(x.detach()).to(dtype)
    y = ops.maxpool_fwd(xn)
    assert relerr(nchw(y), y_ref) == 0
    dx = ops.maxpool_bwd(xn, nhwc(dy).to(dtype))
    # ties at zero may pick a different element than cuDNN; compare where x > 0
    m = (x.detach() > 0)
    assert relerr(nchw(dx).float() * m, x.grad * m) < (1e-6 if dtype == torch.float32 else 1e-2)
    # recorded-position variant: forward stores the argmax code, backward is a gather that never reads x
    y_p, pos = ops.maxpool_fwd(xn, want_pos=True)
    assert pos is not None and torch.equal(y_p, y)
    dx_p = ops.maxpool_bwd(None, nhwc(dy).to(dtype), pos=pos, x_shape=tuple(xn.shape))
    assert torch.equal(dx_p, dx)

    for T in (8, 4, 2, 3):
        for mode in ("max", "avg"):
            if mode == "avg" and T < 3:
                continue  # torch's CUDA AvgPool3d refuses T < kernel; the max path covers T=2
            V = 3
            xt = torch.randn(V * T, 6, 5, 4, generator=g).to(cuda)
            if dtype == torch.bfloat16:
                xt = xt.bfloat16().float()
            xt.requires_grad_(True)
            pool = (torch.nn.MaxPool3d if mode == "max" else torch.nn.AvgPool3d)((3, 1, 1), (2, 1, 1), (1, 0, 0))
            v = xt.view(V, T, 6, 5, 4).transpose(1, 2)
            yt = pool(v).transpose(1, 2).contiguous().view(-1, 6, 5, 4)
            dyt = torch.randn(yt.shape, generator=g).to(cuda)
            yt.backward(dyt)
            xtn = nhwc(xt.detach()).to(dtype)
            y2 = ops.tpool_fwd(xtn, T, mode == "avg")
            assert relerr(nchw(y2), yt) < (1e-6 if dtype == torch.float32 else 1e-2)
            dx2 = ops.tpool_bwd(xtn, nhwc(dyt).to(dtype), T, mode == "avg")
            assert relerr(nchw(dx2), xt.grad) < (1e-6 if dtype == torch.float32 else 1e-2)

    xa = torch.randn(5, 70, 7, 7, generator=g).to(cuda)
    ya = ops.avgpool_fwd(nhwc(xa).to(dtype))
    assert relerr(ya, xa.mean((2, 3))) < TOL[dtype]
    dya = torch.randn(5, 70, generator=g).to(cuda)
    dxa = ops.avgpool_bwd(dya, (5, 7, 7, 70), dtype)
    assert relerr(nchw(dxa), (dya / 49)[:, :, None, None].expand(5, 70, 7, 7)) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_data_layer(cuda, dtype):
    from adamml_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(8)
    N, S, Fr, C, H, W = 2, 3, 8, 3, 32, 28
    x = torch.randn(N, S * Fr * C, H, W, generator=g).to(cuda)
    ref = x.view(N, S, Fr * C, H, W).transpose(0, 1).contiguous()  # adamml.py:65
    ref = ref.view(S * N * Fr, C, H, W)
    out = ops.pack_frames(x, S, Fr, C, dtype)
    assert relerr(nchw(out), ref) < (1e-7 if dtype == torch.float32 else 1e-2)
    out4 = ops.pack_frames(x, S, Fr, C, dtype, cpad=4)
    assert relerr(nchw(out4)[:, :3], ref) < (1e-7 if dtype == torch.float32 else 1e-2)
    assert (out4[..., 3] == 0).all()
    OH, OW = 20, 18
    tmp = F.interpolate(x, size=(OH, OW), mode="bilinear")  # adamml.py:59-62
    tmp = tmp.view(N, S, Fr, -1, OH, OW)[:, :, range(0, Fr, 2), ...]
    tmp = tmp.reshape(N, S, -1, OH, OW).transpose(0, 1).contiguous().view(S * N * (Fr // 2), C, OH, OW)
    out = ops.resize_frames(x, S, Fr, C, OH, OW, 2, dtype)
    assert relerr(nchw(out), tmp) < (2e-6 if dtype == torch.float32 else 1e-2)
    # the real geometry 224 -> 160
    x2 = torch.randn(1, 8 * 3, 224, 224, generator=g).to(cuda)
    ref2 = F.interpolate(x2, size=(160, 160), mode="bilinear").view(1, 1, 8, 3, 160, 160)[:, :, range(0, 8, 2)]
    ref2 = ref2.reshape(4, 3, 160, 160)
    o2 = ops.resize_frames(x2, 1, 8, 3, 160, 160, 2, torch.float32)
    assert relerr(nchw(o2), ref2) < 2e-6


def test_linear_helpers(cuda):
    from adamml_b200 import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    x = torch.randn(37, 2052, generator=g).to(cuda)
    w = torch.randn(1024, 2052, generator=g).to(cuda) / 45
    b = torch.randn(1024, generator=g).to(cuda)
    y = ops.linear_fwd(x, w)
    assert relerr(y, x @ w.t()) < 2e-5
    # strided: only the first 2048 input features
    y2 = ops.linear_fwd(x, w, K=2048)
    assert relerr(y2, x[:, :2048] @ w[:, :2048].t()) < 2e-5
    yb = ops.bias_act_(y.clone(), b, 1)
    assert relerr(yb, F.relu(x @ w.t() + b)) < 2e-5
    dy = torch.randn(37, 1024, generator=g).to(cuda)
    dz = ops.act_bwd(dy, yb, 1)
    assert relerr(dz, dy * (yb > 0)) == 0
    assert relerr(ops.colsum(dz), dz.sum(0)) < 1e-5
    dx = ops.linear_dgrad(dz, w)
    assert relerr(dx, dz @ w) < 2e-5
    dx2 = ops.linear_dgrad(dz, w, K=2048)
    assert relerr(dx2, dz @ w[:, :2048]) < 2e-5
    dw = ops.linear_wgrad(x, dz)
    assert relerr(dw, dz.t() @ x) < 5e-5
    a, c = torch.randn(1000, generator=g).to(cuda), torch.randn(1000, generator=g).to(cuda)
    assert relerr(ops.mul(a, c), a * c) == 0
    xm = torch.randn(12, 31, generator=g).to(cuda)
    assert relerr(ops.frame_mean(xm, 4), xm.view(3, 4, 31).mean(1)) < 1e-6
    dm = torch.randn(3, 31, generator=g).to(cuda)
    assert relerr(ops.frame_mean_bwd(dm, 4), (dm / 4)[:, None, :].expand(3, 4, 31).reshape(12, 31)) < 1e-6


TC_CONV_CASES = [
    # IMGS,H,W,Cin,Cout,R,stride,pad
    (4, 56, 56, 64, 64, 3, 1, 1),
    (8, 28, 28, 128, 128, 3, 1, 1),
    (32, 14, 14, 256, 256, 3, 1, 1),
    (20, 7, 7, 512, 512, 3, 1, 1),
    (3, 56, 56, 128, 128, 3, 2, 1),
    (5, 28, 28, 256, 256, 3, 2, 1),
    (6, 14, 14, 512, 512, 3, 2, 1),
    (3, 56, 56, 256, 512, 1, 2, 0),
    (5, 14, 14, 1024, 2048, 1, 2, 0),
    (3, 13, 17, 40, 72, 3, 1, 1),
    (2, 30, 30, 32, 64, 5, 2, 2),
    (2, 24, 24, 64, 64, 7, 2, 3),
    (3, 20, 20, 144, 24, 1, 1, 0),   # MobileNetV2 project layer: its dgrad reduces over K = Cout = 24 (ragged K block)
    (3, 20, 20, 32, 16, 1, 1, 0),
]


@pytest.mark.parametrize("case", TC_CONV_CASES)
def test_tc_conv_fwd_stats_dgrad(cuda, case):
    """tcgen05 implicit-GEMM conv (4D TMA taps) vs F.conv2d; fused BN stats; stride-1 dgrad with addend."""
    from adamml_b200 import _lib, ops
    IMGS, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(IMGS, Cin, H, W, generator=g).to(cuda).bfloat16().float()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(cuda).bfloat16().float()
    x.requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride, pad)
    xn = nhwc(x.detach()).bfloat16()
    wp = ops.pack_weight(w.contiguous(), torch.bfloat16)
    Ho, Wo = y_ref.shape[2], y_ref.shape[3]
    G = 1 if IMGS % 2 else 2
    y = torch.full((IMGS, Ho, Wo, Cout), float("nan"), device=cuda, dtype=torch.bfloat16)
    stats = torch.full((G, Cout, 2), float("nan"), device=cuda, dtype=torch.float64)
    _lib.call("tc_conv_bf16", xn, wp, y, None, IMGS, H, W, Cin, Cout, R, R, stride, pad, Ho, Wo, stats, IMGS // G, 0)
    torch.cuda.synchronize()
    assert torch.isfinite(y.float()).all()
    assert relerr(nchw(y), y_ref) < 1e-2
    yd = y.double().view(G, -1, Cout)
    ref_stats = torch.stack([yd.sum(1), (yd * yd).sum(1)], -1)
    assert relerr(stats, ref_stats) < 1e-5
    dy = torch.randn(y_ref.shape, generator=g).to(cuda).bfloat16().float()
    y_ref.backward(dy)
    w_rot = ops.pack_weight_dgrad(w.contiguous(), torch.bfloat16)
    if stride == 2 and R >= 2:
        # stride-2 data gradient = four parity-class stride-1 implicit GEMMs writing a strided output lattice
        dx = torch.full(xn.shape, float("nan"), device=cuda, dtype=torch.bfloat16)
        _lib.call("tc_dgrad_s2_bf16", nhwc(dy).bfloat16(), w_rot, dx, IMGS, H, W, Cin, Cout, R, R, pad, Ho, Wo)
        assert torch.isfinite(dx.float()).all()
        assert relerr(nchw(dx), x.grad) < 1e-2
        dx_b = ops.conv_dgrad(nhwc(dy).bfloat16(), wp, tuple(xn.shape), 2, pad, w_rot=w_rot)
        assert torch.equal(dx_b, dx)
        return
    if stride == 2:
        # stride-2 1x1: compact GEMM, scattered onto the even pixels by the epilogue of a stride-1 1x1 dgrad
        dxc = ops.conv_dgrad_compact(nhwc(dy).bfloat16(), w_rot)
        full = torch.zeros(IMGS, Cin, H, W, device=cuda)
        full[:, :, ::2, ::2] = nchw(dxc).float()
        assert relerr(full, x.grad) < 1e-2
        C1 = 64
        w1 = (torch.randn(C1, Cin, 1, 1, generator=g) / Cin ** 0.5).to(cuda).bfloat16().float()
        dy1 = torch.randn(IMGS, C1, H, W, generator=g).to(cuda).bfloat16().float()
        ref = F.conv_transpose2d(dy1, w1) + full
        dx1 = ops.conv_dgrad(nhwc(dy1).bfloat16(), ops.pack_weight(w1.contiguous(), torch.bfloat16), tuple(xn.shape), 1, 0,
                             addend=dxc, w_rot=ops.pack_weight_dgrad(w1.contiguous(), torch.bfloat16), addend_sub=2)
        assert relerr(nchw(dx1), ref) < 1e-2
        return
    add = torch.randn(xn.shape, generator=g).to(cuda).bfloat16()
    dx = ops.conv_dgrad(nhwc(dy).bfloat16(), wp, tuple(xn.shape), 1, pad, addend=None, w_rot=w_rot)
    assert relerr(nchw(dx), x.grad) < 1e-2
    dx2 = ops.conv_dgrad(nhwc(dy).bfloat16(), wp, tuple(xn.shape), 1, pad, addend=add, w_rot=w_rot)
    assert relerr(nchw(dx2), x.grad + nchw(add).float()) < 1e-2


TC_WGRAD_CASES = TC_CONV_CASES + [
    (6, 20, 20, 16, 96, 1, 1, 0),
    (4, 16, 16, 96, 24, 1, 1, 0),
    (10, 8, 8, 960, 160, 1, 1, 0),
    (3, 10, 10, 320, 1280, 1, 1, 0),
    (64, 56, 56, 64, 256, 1, 1, 0),
]


@pytest.mark.parametrize("case", TC_WGRAD_CASES)
def test_tc_wgrad(cuda, case):
    """tcgen05 MN-major split-K weight gradient vs torch autograd."""
    from adamml_b200 import _lib, ops
    IMGS, H, W, Cin, Cout, R, stride, pad = case
    g = torch.Generator(device="cpu").manual_seed(12)
    x = torch.randn(IMGS, Cin, H, W, generator=g).to(cuda).bfloat16().float()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / (Cin * R * R) ** 0.5).to(cuda).requires_grad_(True)
    y_ref = F.conv2d(x, w, None, stride, pad)
    dy = torch.randn(y_ref.shape, generator=g).to(cuda).bfloat16().float()
    y_ref.backward(dy)
    Ho, Wo = y_ref.shape[2], y_ref.shape[3]
    dw = torch.full((Cout, R, R, Cin), float("nan"), device=cuda, dtype=torch.float32)
    _lib.call("tc_wgrad_bf16", nhwc(x).bfloat16(), nhwc(dy).bfloat16(), dw, IMGS, H, W, Cin, Cout, R, R, stride, pad,
              Ho, Wo)
    torch.cuda.synchronize()
    assert torch.isfinite(dw).all()
    assert relerr(ops.unpack_wgrad(dw, Cin), w.grad) < 1e-3


@pytest.mark.parametrize("case", [(6, 3, 64, 64, 64, 3, 7), (4, 10, 32, 48, 64, 2, 7), (2, 3, 224, 224, 64, 1, 7),
                                  (3, 3, 20, 36, 64, 3, 7), (6, 3, 40, 40, 32, 3, 3), (4, 1, 64, 64, 32, 2, 3),
                                  (2, 15, 32, 48, 32, 1, 3)])
def test_tc_stem_s2d(cuda, case):
    """Stride-2 first convs (7x7/p3 ResNet stem, 3x3/p1 MobileNetV2) on the space-to-depth operand (overlapping-stride
    TMA view): fwd + BN stats + wgrad."""
    from adamml_b200 import ops
    IMGS, C, H, W, Cout, G, R = case
    g = torch.Generator(device="cpu").manual_seed(13)
    x = torch.randn(IMGS, C, H, W, generator=g).to(cuda).bfloat16().float()
    w = (torch.randn(Cout, C, R, R, generator=g) / (C * R * R) ** 0.5).to(cuda).bfloat16().float().requires_grad_(True)
    y_ref = F.conv2d(x, w, None, 2, R // 2)
    conv = torch.nn.Conv2d(C, Cout, R, 2, R // 2, bias=False)
    assert ops.first_conv_s2d_ok(conv, C, H, W, torch.bfloat16)
    if R == 7:
        xs = ops.pack_frames_s2d(x.view(IMGS, 1 * 1 * C, H, W).contiguous(), 1, 1, C)   # N=IMGS, S=F=1
        padl = 2
    else:
        xs = ops.nhwc_to_s2d(nhwc(x).bfloat16(), 3)
        padl = 1
    # the s2d operand holds exactly the input pixels
    t = xs.t[:, :, padl:padl + W // 2, :4 * C].float().view(IMGS, H // 2, W // 2, 2, 2, C)
    assert torch.equal(t.permute(0, 5, 1, 3, 2, 4).reshape(IMGS, C, H, W), x)
    assert xs.t[:, :, :padl].abs().max() == 0
    if xs.t.shape[2] > padl + W // 2:
        assert xs.t[:, :, padl + W // 2:].abs().max() == 0
    stats = torch.full((G, Cout, 2), float("nan"), device=cuda, dtype=torch.float64)
    z = ops.stem_conv_fwd(xs, w.detach().contiguous(), stats=stats, imgs_per_group=IMGS // G)
    torch.cuda.synchronize()
    assert relerr(nchw(z), y_ref) < 1e-2
    zd = z.double().view(G, -1, Cout)
    assert relerr(stats, torch.stack([zd.sum(1), (zd * zd).sum(1)], -1)) < 1e-5
    dy = torch.randn(y_ref.shape, generator=g).to(cuda).bfloat16().float()
    y_ref.backward(dy)
    dw = ops.stem_wgrad(xs, nhwc(dy).bfloat16(), Cout)
    assert relerr(dw, w.grad) < 1e-3
