#!/usr/bin/env python
"""Micro-benchmark of single C-ABI ops at the shapes of the N=72 RGB+Audio step (CUDA events, L2 flushed by
rotating over several operand sets larger than L2).  Usage: python scripts/bench_ops.py [filter]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from adamml_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda:0")
BF = torch.bfloat16
HBM = 6463.0


def timeit(fn, reps=8, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def report(name, ms, by, fl=0.0):
    print(f"{name:58s} {ms:8.3f} ms {by / ms / 1e6:8.0f} GB/s ({by / ms / 1e6 / HBM * 100:5.1f}% hbm) {fl / ms / 1e9:8.1f} TF/s",
          flush=True)


def rnd(*shape, dtype=BF):
    return torch.randn(*shape, device=dev, dtype=torch.float32).to(dtype)


def gemm(M, N, K, stats, G=5):
    A, B, D = rnd(M, K), rnd(N, K), torch.empty(M, N, device=dev, dtype=BF)
    st = torch.empty(G, N, 2, device=dev, dtype=torch.float64) if stats else None
    ms = timeit(lambda: _lib.call("tc_gemm_bf16", A, B, D, M, N, K, 0, 0, 0, _lib.BF16, st, M // G if stats else 0))
    report(f"tc_gemm M={M} N={N} K={K} stats={int(stats)}", ms, 2.0 * M * (N + K), 2.0 * M * N * K)


def x2rand(*shape):
    t = torch.randn(*shape, device=dev, dtype=torch.float32)
    hi = t.bfloat16()
    return ops.X2(hi, (t - hi.float()).half())


def gemm_x2(M, N, K, stats=True, G=5):
    a = x2rand(M, 1, 1, K)
    w = ops.pack_weight(torch.randn(N, K, 1, 1, device=dev) / K ** 0.5, ops.PREC_X2)
    st = torch.empty(G, N, 2, device=dev, dtype=torch.float64) if stats else None
    ms = timeit(lambda: ops.conv_fwd(a, w, 1, 0, stats=st, rows_per_group=M // G))
    report(f"tc_gemm_x2 M={M} N={N} K={K} stats={int(stats)}", ms, 4.0 * M * (N + K), 2.0 * M * N * K)


def conv_x2(I, H, W, Ci, Co, R, stride, pad, stats=True, G=5):
    x = x2rand(I, H, W, Ci)
    w = ops.pack_weight(torch.randn(Co, Ci, R, R, device=dev) / (Ci * R * R) ** 0.5, ops.PREC_X2)
    Ho, Wo = ops.conv_out_hw(H, W, R, R, stride, pad)
    st = torch.empty(G, Co, 2, device=dev, dtype=torch.float64) if stats else None
    ms = timeit(lambda: ops.conv_fwd(x, w, stride, pad, stats=st, rows_per_group=I * Ho * Wo // G))
    report(f"tc_conv_x2 I={I} {H}x{W} {Ci}->{Co} {R}x{R} s{stride} stats={int(stats)}", ms,
           4.0 * (I * H * W * Ci + I * Ho * Wo * Co), 2.0 * I * Ho * Wo * Co * R * R * Ci)


def bn_x2(rows, C, G=5, res=False):
    z = x2rand(rows * G, 1, 1, C)
    r = x2rand(rows * G, 1, 1, C) if res else None
    ss = torch.rand(G, C, 2, device=dev)
    ms = timeit(lambda: ops.bn_apply(z, ss, G, 1, res=r))
    report(f"bn_apply_x2 rows={rows}x{G} C={C} res={int(res)}", ms, (3.0 if res else 2.0) * 4 * rows * G * C)
    ms = timeit(lambda: ops.bn_stats(z, G))
    report(f"bn_stats_x2 rows={rows}x{G} C={C}", ms, 4.0 * rows * G * C)


def dw_x2(I, H, C, stride):
    x = x2rand(I, H, H, C)
    w = ops.pack_weight_dw(torch.randn(C, 1, 3, 3, device=dev))
    y = ops.dwconv_fwd(x, w, stride)
    report(f"dwconv_fwd_x2 I={I} {H}x{H} C={C} s{stride}", timeit(lambda: ops.dwconv_fwd(x, w, stride)),
           4.0 * (x.numel() + y.numel()), 18.0 * y.numel())


def conv(I, H, W, Ci, Co, R, stride, pad, stats, addend=False, G=5):
    x, w = rnd(I, H, W, Ci), rnd(Co, R, R, Ci)
    Ho, Wo = ops.conv_out_hw(H, W, R, R, stride, pad)
    y = torch.empty(I, Ho, Wo, Co, device=dev, dtype=BF)
    ad = rnd(I, Ho, Wo, Co) if addend else None
    st = torch.empty(G, Co, 2, device=dev, dtype=torch.float64) if stats else None
    ms = timeit(lambda: _lib.call("tc_conv_bf16", x, w, y, ad, I, H, W, Ci, Co, R, R, stride, pad, Ho, Wo, st,
                                  I // G if stats else 0, 1 if addend else 0))
    report(f"tc_conv I={I} {H}x{W} {Ci}->{Co} {R}x{R} s{stride} stats={int(stats)} add={int(addend)}", ms,
           2.0 * (I * H * W * Ci + I * Ho * Wo * Co * (2 if addend else 1)), 2.0 * I * Ho * Wo * Co * R * R * Ci)


def wgrad(I, H, W, Ci, Co, R, stride, pad):
    Ho, Wo = ops.conv_out_hw(H, W, R, R, stride, pad)
    x, dy = rnd(I, H, W, Ci), rnd(I, Ho, Wo, Co)
    dw = torch.empty(Co, R, R, Ci, device=dev, dtype=torch.float32)
    ms = timeit(lambda: _lib.call("tc_wgrad_bf16", x, dy, dw, I, H, W, Ci, Co, R, R, stride, pad, Ho, Wo))
    report(f"tc_wgrad I={I} {H}x{W} {Ci}->{Co} {R}x{R} s{stride}", ms, 2.0 * (I * H * W * Ci + I * Ho * Wo * Co),
           2.0 * I * Ho * Wo * Co * R * R * Ci)


def bn(rows, C, G=5):
    z, dout, out = rnd(rows * G, C), rnd(rows * G, C), rnd(rows * G, C)
    mi = torch.rand(G, C, 2, device=dev) + 0.5
    gamma = torch.rand(C, device=dev) + 0.5
    ms = timeit(lambda: ops.bn_bwd_reduce(dout, out, z, mi, G, 1))
    report(f"bn_bwd_reduce rows={rows}x{G} C={C}", ms, 3.0 * 2 * rows * G * C)
    sums = ops.bn_bwd_reduce(dout, out, z, mi, G, 1)
    ms = timeit(lambda: ops.bn_bwd_apply(dout, out, z, mi, gamma, sums, G, rows, 1, True))
    report(f"bn_bwd_apply rows={rows}x{G} C={C}", ms, 4.0 * 2 * rows * G * C)
    ss = torch.rand(G, C, 2, device=dev)
    ms = timeit(lambda: ops.bn_apply(z, ss, G, 1))
    report(f"bn_apply rows={rows}x{G} C={C}", ms, 2.0 * 2 * rows * G * C)


def dw(I, H, C, stride):
    x = rnd(I, H, H, C)
    w = ops.pack_weight_dw(torch.randn(C, 1, 3, 3, device=dev))
    y = ops.dwconv_fwd(x, w, stride)
    dy = torch.randn_like(y)
    by = 2.0 * (x.numel() + y.numel())
    report(f"dwconv_fwd I={I} {H}x{H} C={C} s{stride}", timeit(lambda: ops.dwconv_fwd(x, w, stride)), by, 18.0 * y.numel())
    report(f"dwconv_dgrad I={I} {H}x{H} C={C} s{stride}", timeit(lambda: ops.dwconv_dgrad(dy, w, tuple(x.shape), stride)), by)
    report(f"dwconv_wgrad I={I} {H}x{H} C={C} s{stride}", timeit(lambda: ops.dwconv_wgrad(x, dy, stride)), by)


_FLUSH = None


def timeit_cold(fn, reps=6, warm=2):
    """per-launch CUDA-event time with the 126 MB L2 flushed (256 MB written) before every launch: the small late
    layers would otherwise be timed L2-resident"""
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
    for _ in range(warm):
        fn()
    tot = 0.0
    for _ in range(reps):
        _FLUSH.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


# every depthwise layer shape of the N=72 RGB+Audio step: (IMGS, H, C, stride, launches per step)
DW_SHAPES = [(360, 128, 96, 2, 2), (1440, 80, 96, 2, 1), (360, 64, 144, 1, 2), (360, 128, 32, 1, 2),
             (1440, 40, 144, 1, 1), (360, 16, 384, 1, 8), (360, 32, 192, 1, 4), (1440, 80, 32, 1, 1),
             (1440, 20, 192, 1, 2), (360, 16, 576, 1, 4), (360, 64, 144, 2, 2), (1440, 40, 144, 2, 1),
             (360, 8, 960, 1, 6), (720, 10, 384, 1, 4), (720, 10, 576, 1, 2), (360, 32, 192, 2, 2),
             (360, 16, 576, 2, 2), (360, 5, 960, 1, 3), (720, 20, 192, 2, 1), (360, 10, 576, 2, 1)]


def dw_bwd_all():
    """separate dgrad + wgrad launches vs the fused TMA-tile backward, L2 flushed, summed over one step"""
    sep_tot = fus_tot = pre_tot = 0.0
    for I, H, C, stride, n in DW_SHAPES:
        x = rnd(I, H, H, C)
        w = ops.pack_weight_dw(torch.randn(C, 1, 3, 3, device=dev))
        Ho = (H - 1) // stride + 1
        dy = rnd(I, Ho, Ho, C)
        by = 2.0 * (2 * x.numel() + dy.numel())

        def sep():
            ops.dwconv_dgrad(dy, w, tuple(x.shape), stride)
            ops.dwconv_wgrad(x, dy, stride)
        t_sep = timeit_cold(sep)
        t_fus = timeit_cold(lambda: ops.dwconv_bwd(x, dy, w, stride))
        raw = torch.empty(5, C, 2, device=dev, dtype=torch.float64)
        t_pre = timeit_cold(lambda: ops.dwconv_bwd(x, dy, w, stride, pre=(raw, I // 5, 2)))
        sep_tot += n * t_sep
        fus_tot += n * t_fus
        pre_tot += n * t_pre
        print(f"dw_bwd I={I} {H}x{H} C={C} s{stride} x{n}: separate {t_sep:7.3f} ms  fused {t_fus:7.3f} ms "
              f"({by / t_fus / 1e6:6.0f} GB/s, {by / t_fus / 1e6 / HBM * 100:5.1f}% hbm)  + producer BN reduce "
              f"{t_pre:7.3f} ms", flush=True)
        del x, dy
    print(f"dw_bwd per step: separate {sep_tot:.2f} ms -> fused {fus_tot:.2f} ms -> with the producer's BN reduction "
          f"{pre_tot:.2f} ms", flush=True)


def dw_fwd_all():
    """x2 training forward of the stride-1 depthwise layers: register-window kernel + bn_stats_x2 pass vs the TMA-tile
    kernel with fused statistics, L2 flushed, summed over one step"""
    old_tot = new_tot = 0.0
    for I, H, C, stride, n in DW_SHAPES:
        if stride != 1:
            continue
        x = x2rand(I, H, H, C)
        w = ops.pack_weight_dw(torch.randn(C, 1, 3, 3, device=dev))
        sums = torch.empty(5, C, 2, device=dev, dtype=torch.float64)
        by = 4.0 * 2 * x.hi.numel()

        def old():
            z = ops.dwconv_fwd(x, w, 1)
            ops.bn_stats(z, 5, out=sums)
        t_old = timeit_cold(old)
        t_new = timeit_cold(lambda: ops.dwconv_fwd_stats(x, w, 1, sums, I // 5))
        old_tot += n * t_old
        new_tot += n * t_new
        print(f"dw_fwd_x2 I={I} {H}x{H} C={C} s1 x{n}: fwd + bn_stats {t_old:7.3f} ms  fused {t_new:7.3f} ms "
              f"({by / t_new / 1e6:6.0f} GB/s, {by / t_new / 1e6 / HBM * 100:5.1f}% hbm)", flush=True)
        del x
    print(f"dw_fwd_x2 (stride 1) per step: fwd + bn_stats {old_tot:.2f} ms -> fused {new_tot:.2f} ms", flush=True)


def bn_bwd_reduce_cases():
    """the two shapes of the BatchNorm backward reduction: residual layers (dout, out, z -> masked gradient written in
    place) and layers whose ReLU mask is recomputed from z (dout, z)"""
    for rows, C in [(1806336, 256), (225792, 512), (28224, 1024), (1806336, 64), (225792, 128), (294912, 144),
                    (1179648, 96)]:
        G = 5
        z, dout, out = rnd(rows * G, C), rnd(rows * G, C), rnd(rows * G, C)
        mi = torch.rand(G, C, 2, device=dev) + 0.5
        ss = torch.rand(G, C, 2, device=dev)
        ms = timeit_cold(lambda: ops.bn_bwd_reduce(dout, out, z, mi, G, 1, gm_inplace=True))
        report(f"bn_bwd_reduce residual (3 reads + gm write) rows={rows}x{G} C={C}", ms, 4.0 * 2 * rows * G * C)
        ms = timeit_cold(lambda: ops.bn_bwd_reduce(dout, None, z, mi, G, 1, mask_ss=ss))
        report(f"bn_bwd_reduce mask from z (2 reads)         rows={rows}x{G} C={C}", ms, 2.0 * 2 * rows * G * C)
        del z, dout, out


CASES = {
    "bnbwdred": bn_bwd_reduce_cases,
    "dwfwd": dw_fwd_all,
    "dwbwd": dw_bwd_all,
    "gemm": lambda: [gemm(9031680, 256, 64, True), gemm(9031680, 256, 64, False), gemm(9031680, 64, 256, True),
                     gemm(9031680, 64, 64, True), gemm(4515840, 128, 256, True), gemm(1128960, 512, 128, True),
                     gemm(9216000, 96, 16, True), gemm(2304000, 24, 144, True), gemm(282240, 1024, 256, True),
                     gemm(282240, 256, 1024, True)],
    "shallow": lambda: [gemm(9216000, 96, 16, True), gemm(5898240, 96, 16, True), gemm(2304000, 144, 24, True),
                        gemm(1474560, 144, 24, True), gemm(368640, 192, 32, True), gemm(5898240, 16, 96, True),
                        gemm(9216000, 16, 32, True), gemm(5898240, 32, 16, False), gemm(9031680, 64, 64, True),
                        gemm(1474560, 24, 96, True), gemm(92160, 64, 64, True)],
    "conv": lambda: [conv(2880, 56, 56, 64, 64, 3, 1, 1, True), conv(2880, 56, 56, 64, 64, 3, 1, 1, False),
                     conv(2880, 56, 56, 64, 256, 1, 1, 0, False, addend=True),
                     conv(1440, 28, 28, 128, 128, 3, 1, 1, True), conv(720, 14, 14, 256, 256, 3, 1, 1, True),
                     conv(1440, 56, 56, 128, 128, 3, 2, 1, True)],
    "wgrad": lambda: [wgrad(2880, 56, 56, 64, 64, 3, 1, 1), wgrad(2880, 56, 56, 256, 64, 1, 1, 0),
                      wgrad(2880, 56, 56, 64, 256, 1, 1, 0), wgrad(1440, 28, 28, 128, 128, 3, 1, 1),
                      wgrad(720, 14, 14, 256, 256, 3, 1, 1), wgrad(360, 7, 7, 512, 2048, 1, 1, 0)],
    "x2gemm": lambda: [gemm_x2(9031680, 256, 64), gemm_x2(9031680, 64, 256), gemm_x2(9031680, 64, 64),
                       gemm_x2(1128960, 512, 128), gemm_x2(1128960, 128, 512), gemm_x2(141120, 1024, 256),
                       gemm_x2(1474560, 144, 24), gemm_x2(5898240, 96, 16), gemm_x2(368640, 192, 32),
                       gemm_x2(92160, 576, 96), gemm_x2(5898240, 16, 32), gemm_x2(1474560, 24, 144),
                       gemm_x2(17640, 2048, 512)],
    "x2conv": lambda: [conv_x2(2880, 56, 56, 64, 64, 3, 1, 1), conv_x2(1440, 28, 28, 128, 128, 3, 1, 1),
                       conv_x2(720, 14, 14, 256, 256, 3, 1, 1), conv_x2(1440, 56, 56, 256, 512, 1, 2, 0)],
    "x2bn": lambda: [bn_x2(1806336, 256, res=True), bn_x2(1806336, 64), bn_x2(1843200, 96)],
    "x2dw": lambda: [dw_x2(1440, 40, 144, 1), dw_x2(360, 128, 96, 2), dw_x2(1440, 80, 32, 1)],
    "bn": lambda: [bn(1806336, 256), bn(1806336, 64), bn(225792, 512), bn(1843200, 96)],
    "dw": lambda: [dw(1440, 40, 144, 1), dw(360, 128, 96, 2), dw(1440, 80, 32, 1)],
}

if __name__ == "__main__":
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    for k, f in CASES.items():
        if flt in k:
            f()
            torch.cuda.empty_cache()
