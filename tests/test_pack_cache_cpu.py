"""Host logic of ops.WeightPackCache (one weight-conversion launch per pass) on CPU: the C-ABI calls are replaced by
a Python interpreter of the SAME job table / chunk lists the device kernel receives (csrc/data_layer.cu
pack_weights_multi_kernel: job = {src, dst, Cout, Cin, R, S, CinPad, kind}), reading and writing the tensors through
their addresses.  Checks: the table covers every destination element exactly once and reproduces the per-layer
conversions bit for bit, operands follow in-place weight updates, an operand is never handed out stale, unused
operands are pruned, tables a CUDA graph captured are retained, a new operand during capture is an error."""
import ctypes

import numpy as np
import pytest
import torch

from adamml_b200 import _lib, ops

CHUNK = 4096


def bf16_bits(t):
    return t.to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)


def x2_planes(v):
    """the four 2-byte planes of adamml_pack_weight_x2 for fp32 values v (flat): b1, b2, b3 (bf16 cascade), fp16(v)"""
    b1 = v.to(torch.bfloat16)
    r1 = v - b1.float()
    b2 = r1.to(torch.bfloat16)
    b3 = (r1 - b2.float()).to(torch.bfloat16)
    return [bf16_bits(b1), bf16_bits(b2), bf16_bits(b3), v.to(torch.float16).view(torch.int16).numpy().astype(np.uint16)]


def ref_ohwi(w):
    return w.permute(0, 2, 3, 1).contiguous()


def ref_dgrad(w):  # [Cin][R][S][Cout], taps rotated by 180 degrees
    return w.flip(2, 3).permute(1, 2, 3, 0).contiguous()


def ref_dw(w):     # [C,1,3,3] -> [9, C]
    return w.reshape(w.shape[0], 9).t().contiguous()


def as_np(ptr, n, ctype, dtype):
    return np.ctypeslib.as_array((ctype * n).from_address(ptr)).view(dtype)


class FakeABI:
    """per-layer entry points computed with torch, pack_weights_multi interpreted from its device tables"""

    def __init__(self):
        self.calls = []

    def __call__(self, name, *a):
        self.calls.append(name)
        if name == "pack_weight_x2":
            w, out = a[0], a[1]
            planes = x2_planes(ref_ohwi(w).reshape(-1))
            out.view(torch.int16).view(4, -1).copy_(torch.from_numpy(np.stack(planes).astype(np.int16)))
        elif name == "pack_weight":
            a[1].copy_(ref_ohwi(a[0]).to(a[1].dtype))
        elif name == "pack_weight_dgrad":
            a[1].copy_(ref_dgrad(a[0]).to(a[1].dtype))
        elif name == "pack_weight_dw":
            a[1].copy_(ref_dw(a[0]))
        elif name == "pack_weights_multi":
            self.multi(*a)
        else:
            raise AssertionError(name)
        return 0

    def multi(self, jobs, cj, cs, n_jobs, n_chunks):
        assert jobs.shape == (n_jobs, 8) and cj.numel() == n_chunks == cs.numel()
        covered = {}
        for b in range(n_chunks):
            j = int(cj[b])
            src_p, dst_p, Cout, Cin, R, S, CinPad, kind = (int(v) for v in jobs[j])
            total = Cout * 9 if kind == ops.PK_DW else (Cin * R * S * Cout if kind in (ops.PK_DGRAD_F32, ops.PK_DGRAD_BF16)
                                                        else Cout * R * S * CinPad)
            e0 = int(cs[b])
            e1 = min(e0 + CHUNK, total)
            assert 0 <= e0 < total
            covered.setdefault(j, []).append((e0, e1))
            idx = np.arange(e0, e1)
            nsrc = Cout * 9 if kind == ops.PK_DW else Cout * Cin * R * S
            src = as_np(src_p, nsrc, ctypes.c_float, np.float32)
            if kind == ops.PK_DW:
                c, t = idx // 9, idx % 9
                as_np(dst_p, total, ctypes.c_float, np.float32)[t * Cout + c] = src[idx]
                continue
            if kind in (ops.PK_DGRAD_F32, ops.PK_DGRAD_BF16):
                co, s_, r, ci = idx % Cout, (idx // Cout) % S, (idx // (Cout * S)) % R, idx // (Cout * S * R)
                v = src[((co * Cin + ci) * R + (R - 1 - r)) * S + (S - 1 - s_)]
            else:
                ci, s_, r, co = idx % CinPad, (idx // CinPad) % S, (idx // (CinPad * S)) % R, idx // (CinPad * S * R)
                v = np.where(ci < Cin, src[((co * Cin + np.minimum(ci, Cin - 1)) * R + r) * S + s_], 0.0).astype(np.float32)
            if kind in (ops.PK_OHWI_F32, ops.PK_DGRAD_F32):
                as_np(dst_p, total, ctypes.c_float, np.float32)[idx] = v
            elif kind in (ops.PK_OHWI_BF16, ops.PK_DGRAD_BF16):
                as_np(dst_p, total, ctypes.c_uint16, np.uint16)[idx] = bf16_bits(torch.from_numpy(v))
            else:
                dst = as_np(dst_p, 4 * total, ctypes.c_uint16, np.uint16)
                for pl, bits in enumerate(x2_planes(torch.from_numpy(v))):
                    dst[pl * total + idx] = bits
        for j, spans in covered.items():  # every destination element exactly once
            spans.sort()
            assert spans[0][0] == 0 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert len(covered) == n_jobs


@pytest.fixture
def fake(monkeypatch):
    abi = FakeABI()
    monkeypatch.setattr(ops, "call", abi)
    monkeypatch.setattr(ops, "_chk", lambda t, dtype=None: t)
    monkeypatch.setattr(ops, "dtype_code", lambda dt: 0)
    state = {"capturing": False}
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: state["capturing"])

    class _L:
        class cdll:
            @staticmethod
            def adamml_pack_chunk():
                return CHUNK
    monkeypatch.setattr(_lib, "lib", lambda: _L)
    abi.state = state
    return abi


def weights(seed=0):
    g = torch.Generator().manual_seed(seed)
    return dict(c3=torch.randn(48, 32, 3, 3, generator=g), c1=torch.randn(70, 24, 1, 1, generator=g),
                big=torch.randn(256, 64, 3, 3, generator=g), dw=torch.randn(40, 1, 3, 3, generator=g))


def one_pass(cache, W, backward=True):
    """what the engine asks for in one training pass"""
    out = dict(c3=ops.pack_weight(W["c3"], ops.PREC_X2, cache=cache), c1=ops.pack_weight(W["c1"], torch.bfloat16, cache=cache),
               big=ops.pack_weight(W["big"], ops.PREC_X2, cache=cache), dw=ops.pack_weight_dw(W["dw"], cache=cache))
    if backward:
        out["c3_d"] = ops.pack_weight_dgrad(W["c3"], torch.bfloat16, cache=cache)
        out["big_d"] = ops.pack_weight_dgrad(W["big"], torch.float32, cache=cache)
    return out


def check(got, W):
    for k in ("c3", "big"):
        want = torch.from_numpy(np.stack(x2_planes(ref_ohwi(W[k]).reshape(-1))).astype(np.int16))
        assert torch.equal(got[k].planes.view(torch.int16).view(4, -1), want), k
        assert got[k].shape == (W[k].shape[0], W[k].shape[2], W[k].shape[3], W[k].shape[1])
    assert torch.equal(got["c1"], ref_ohwi(W["c1"]).bfloat16())
    assert torch.equal(got["dw"], ref_dw(W["dw"]))
    if "c3_d" in got:
        assert torch.equal(got["c3_d"], ref_dgrad(W["c3"]).bfloat16())
        assert torch.equal(got["big_d"], ref_dgrad(W["big"]))


def test_first_pass_records_then_one_launch_per_pass(fake):
    cache, W = ops.WeightPackCache(), weights()
    cache.begin()
    check(one_pass(cache, W), W)
    assert fake.calls.count("pack_weights_multi") == 0 and len(fake.calls) == 6 and cache.dirty
    for step in range(3):
        for w in W.values():               # optimizer step: in place, same storage
            w.mul_(0.9).add_(0.01 * (step + 1))
        fake.calls.clear()
        cache.begin()
        got = one_pass(cache, W)
        assert fake.calls == ["pack_weights_multi"], fake.calls
        check(got, W)
        assert all(e.in_table for e in cache.ent.values()) and not cache.dirty
    # operands live in ONE arena
    lo, hi = cache.arena.data_ptr(), cache.arena.data_ptr() + cache.arena.numel()
    assert all(lo <= e.out.data_ptr() < hi and (e.out.data_ptr() - lo) % 256 == 0 for e in cache.ent.values())


def test_never_stale_without_begin_and_new_operands_join(fake):
    cache, W = ops.WeightPackCache(), weights(1)
    cache.begin(); one_pass(cache, W, backward=False)
    cache.begin(); one_pass(cache, W, backward=False)
    # a pass that asks without begin() having refreshed THIS epoch must not get the arena contents: simulate by
    # bumping the epoch the way begin() does, but skipping its launch
    W["c1"].add_(1.0)
    cache.epoch += 1
    fake.calls.clear()
    got = ops.pack_weight(W["c1"], torch.bfloat16, cache=cache)
    assert fake.calls == ["pack_weight"] and torch.equal(got, ref_ohwi(W["c1"]).bfloat16())
    # operands first seen later (the first backward after forward-only passes) are converted per layer, then join
    cache.begin()
    check(one_pass(cache, W, backward=True), W)
    assert cache.dirty
    fake.calls.clear()
    cache.begin()
    check(one_pass(cache, W, backward=True), W)
    assert fake.calls == ["pack_weights_multi"]


def test_prune_and_graph_retention(fake):
    cache, W = ops.WeightPackCache(), weights(2)
    cache.PRUNE_EVERY = 4
    cache.begin(); one_pass(cache, W)
    cache.begin(); one_pass(cache, W)
    n_all = len(cache.ent)
    first = (cache.table, cache.arena)
    for _ in range(10):                     # forward-only phase: the data-gradient operands fall out
        cache.begin()
        check(one_pass(cache, W, backward=False), W)
    assert len(cache.ent) == n_all - 2 and not cache._retired and cache.arena is not first[1]
    # a captured pass pins its table and arena when a later rebuild replaces them
    fake.state["capturing"] = True
    cache.begin(); one_pass(cache, W, backward=False)
    with pytest.raises(RuntimeError, match="during CUDA-graph capture"):
        ops.pack_weight_dgrad(W["c3"], torch.bfloat16, cache=cache)   # new operand while capturing: recorded ...
        cache.begin()                                                # ... and the rebuild refuses
    fake.state["capturing"] = False
    pinned = (cache.table, cache.arena)
    cache.begin()
    assert cache._retired and cache._retired[-1][1] is pinned[1] and cache.arena is not pinned[1]
    check(one_pass(cache, W, backward=False), W)
    import copy
    assert not copy.deepcopy(cache).ent
