"""Tensor-level wrappers over the C-ABI (include/adamml_b200.h).

Every function takes/returns torch CUDA tensors whose storage comes from torch's caching
allocator; all arithmetic happens inside libadamml_b200.so.  Activations are NHWC
`[IMGS, H, W, C]` tensors (fp32 or bf16); weights for the dense engines are OHWI.
"""
import os

import torch

from . import _lib
from ._lib import ACT_NONE, ACT_RELU, ACT_RELU6, call, dtype_code  # noqa: F401

# Engine selection for bf16 dense GEMM-shaped convs: "auto" = tcgen05 when the shape fits,
# "simt" = always the exact CUDA-core engine.
TC_MODE = "auto"

# Precision modes (the `compute_dtype` of a model):
#   PREC_X2        default: forward activations / operands as two planes (hi bf16 + lo fp16 remainder, ~20 mantissa
#                  bits), 4-product tcgen05 GEMMs with fp32 accumulation -> logits within 1e-3 of the reference's fp32
#                  path and bit-exact policy selections; the backward pass runs on the hi planes in bf16.
#   torch.bfloat16 speed mode: bf16 storage forward and backward (does NOT meet the 1e-3 bar).
#   torch.float32  exact CUDA-core engine (fp32 math everywhere).
PREC_X2 = "x2"


class X2:
    """Two-plane activation (include/adamml_b200.h "x2"): value = hi (bf16) + lo (fp16), both NHWC [IMGS, H, W, C]."""
    __slots__ = ("hi", "lo", "_adamml_src")
    dtype = PREC_X2

    def __init__(self, hi, lo):
        self.hi, self.lo = hi, lo
        self._adamml_src = None

    @staticmethod
    def empty(shape, device):
        return X2(torch.empty(shape, device=device, dtype=torch.bfloat16),
                  torch.empty(shape, device=device, dtype=torch.float16))

    @property
    def shape(self):
        return self.hi.shape

    @property
    def device(self):
        return self.hi.device

    def numel(self):
        return self.hi.numel()

    def record_stream(self, s):
        self.hi.record_stream(s)
        self.lo.record_stream(s)

    def float(self):
        """fp32 value (tests / fallbacks only)"""
        return self.hi.float() + self.lo.float()


class X2W:
    """Four-plane weight operand of the x2 tensor-core path (adamml_pack_weight_x2): `planes` is one 2-byte tensor
    [4, Cout, R, S, Cin]: b1 = bf16(w), b2 = bf16(w - b1), b3 = bf16(w - b1 - b2) and f = fp16(w) (plane 3 holds fp16
    bits).  planes[0] is the bf16 OHWI operand of the backward pass."""
    __slots__ = ("planes",)

    def __init__(self, planes):
        self.planes = planes

    @property
    def shape(self):
        return self.planes.shape[1:]

    @property
    def hi(self):
        return self.planes[0]

    def cascade(self):
        """fp64 value of b1 + b2 + b3 (what multiplies the hi plane)"""
        return self.planes[:3].double().sum(0)

    def f16(self):
        """fp16(w) (what multiplies the lo plane)"""
        return self.planes[3].view(torch.float16)


def hi_plane(t):
    """what the backward pass keeps of a forward tensor: the bf16 hi plane of an x2 activation"""
    if isinstance(t, (X2, X2W)):
        return t.hi
    if isinstance(t, S2D) and t.lo is not None:
        return S2D(t.t, t.C, t.H, t.W, t.R)
    return t


def _chk(t, dtype=None):
    assert t.is_cuda and t.is_contiguous(), "adamml_b200 ops need contiguous CUDA tensors"
    if dtype is not None:
        assert t.dtype == dtype, f"expected {dtype}, got {t.dtype}"
    return t


def conv_out_hw(H, W, R, S, stride, pad):
    return (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1


# ---------------------------------------------------------------- data layer
def _u8_norm(x, C, norm):
    """uint8 clips carry their normalisation onto the device: norm = (mean [C], std [C]) fp32 device tensors."""
    if x.dtype != torch.uint8:
        _chk(x, torch.float32)
        return None
    if norm is None:
        raise ValueError("uint8 frames need norm=(mean, std) per channel")
    mean, std = norm
    _chk(x, torch.uint8)
    _chk(mean, torch.float32)
    _chk(std, torch.float32)
    if mean.numel() != C or std.numel() != C:
        raise ValueError("norm must hold one mean/std per frame channel (%d)" % C)
    return mean, std


def pack_frames(x, S, F, C, dtype, cpad=None, norm=None):
    """x NCHW fp32 (or uint8 + norm) [N, S*F*C, H, W] -> NHWC [(s*N+n)*F+f, H, W, cpad]."""
    nm = _u8_norm(x, C, norm)
    N, SFC, H, W = x.shape
    assert SFC == S * F * C, (x.shape, S, F, C)
    cpad = cpad or C
    if dtype == PREC_X2:
        out = X2.empty((S * N * F, H, W, cpad), x.device)
        call("pack_frames_x2", x, nm[0] if nm else None, nm[1] if nm else None, out.hi, out.lo, N, S, F, C, H, W, cpad,
             int(nm is not None))
        return out
    out = torch.empty((S * N * F, H, W, cpad), device=x.device, dtype=dtype)
    if nm is None:
        call("pack_frames", x, out, N, S, F, C, H, W, cpad, dtype_code(dtype))
    else:
        call("pack_frames_u8", x, nm[0], nm[1], out, N, S, F, C, H, W, cpad, dtype_code(dtype))
    return out


def resize_frames(x, S, F, C, OH, OW, fstep, dtype, cpad=None, norm=None):
    nm = _u8_norm(x, C, norm)
    N, SFC, H, W = x.shape
    assert SFC == S * F * C
    cpad = cpad or C
    Fk = (F + fstep - 1) // fstep
    if dtype == PREC_X2:
        out = X2.empty((S * N * Fk, OH, OW, cpad), x.device)
        call("resize_frames_x2", x, nm[0] if nm else None, nm[1] if nm else None, out.hi, out.lo, N, S, F, C, H, W, OH,
             OW, fstep, cpad, int(nm is not None))
        return out
    out = torch.empty((S * N * Fk, OH, OW, cpad), device=x.device, dtype=dtype)
    if nm is None:
        call("resize_frames", x, out, N, S, F, C, H, W, OH, OW, fstep, cpad, dtype_code(dtype))
    else:
        call("resize_frames_u8", x, nm[0], nm[1], out, N, S, F, C, H, W, OH, OW, fstep, cpad, dtype_code(dtype))
    return out


class S2D:
    """Space-to-depth operand of a stride-2 first convolution on the tensor-core path: `t` is bf16
    [IMGS, H/2, W/2 + pads, Cs] (csrc/data_layer.cu); `R` = filter size (7: ResNet stem, 3: MobileNetV2 first conv),
    `taps` = (R+1)/2 s2d taps per axis; shape reports the logical NHWC input."""

    def __init__(self, t, C, H, W, R, lo=None):
        self.t, self.C, self.H, self.W, self.R = t, C, H, W, R
        self.lo = lo  # x2 mode: fp16 remainder plane of the same shape
        self.Cs = t.shape[-1]
        self.taps = (R + 1) // 2

    @property
    def shape(self):
        return (self.t.shape[0], self.H, self.W, self.C)

    @property
    def device(self):
        return self.t.device

    @property
    def dtype(self):
        return PREC_X2 if self.lo is not None else self.t.dtype

    def record_stream(self, s):
        self.t.record_stream(s)
        if self.lo is not None:
            self.lo.record_stream(s)


def _select_rows(t, idx, clips):
    if t.shape[0] % clips:
        raise ValueError("select_clips: %d images do not split into %d clips" % (t.shape[0], clips))
    T = t.shape[0] // clips
    return t.view(clips, -1).index_select(0, idx).view((idx.numel() * T,) + tuple(t.shape[1:]))


def select_clips(x, idx, clips):
    """x: NHWC image batch (or S2D operand) holding `clips` (segment, video) pairs of T consecutive frames each;
    -> the same layout restricted to the pairs listed in idx (int64, ascending)."""
    if isinstance(x, X2):
        return X2(_select_rows(x.hi, idx, clips), _select_rows(x.lo, idx, clips))
    if isinstance(x, S2D):
        return S2D(_select_rows(x.t, idx, clips), x.C, x.H, x.W, x.R,
                   lo=_select_rows(x.lo, idx, clips) if x.lo is not None else None)
    return _select_rows(x, idx, clips)


# ---------------------------------------------------------------- device-side gating (inference with skipping)
def select_compact(decisions, m):
    """decisions fp32 [S, M, N] (0/1) -> (idx int32 [S*N] ascending pair indices s*N+n of modality m, count int32 [1]),
    both on the device: nothing is read back."""
    S, M, N = decisions.shape
    _chk(decisions, torch.float32)
    idx = torch.empty(S * N, device=decisions.device, dtype=torch.int32)
    count = torch.empty(1, device=decisions.device, dtype=torch.int32)
    call("select_compact", decisions, S, M, N, m, idx, count)
    return idx, count


def _gather_plane(t, idx, count, clips):
    if t.shape[0] % clips:
        raise ValueError("gather_clips: %d images do not split into %d clips" % (t.shape[0], clips))
    out = torch.empty_like(t)
    call("gather_rows", t, out, idx, count, t.numel() // clips * t.element_size(), clips)
    return out


def gather_clips(x, idx, count, clips):
    """device-side counterpart of select_clips: the clips listed in idx[:count] move to the front of a buffer of the
    SAME static shape (the tail is left uninitialised: the live limit keeps every kernel away from it)."""
    if isinstance(x, X2):
        return X2(_gather_plane(x.hi, idx, count, clips), _gather_plane(x.lo, idx, count, clips))
    if isinstance(x, S2D):
        return S2D(_gather_plane(x.t, idx, count, clips), x.C, x.H, x.W, x.R,
                   lo=_gather_plane(x.lo, idx, count, clips) if x.lo is not None else None)
    return _gather_plane(x, idx, count, clips)


def scatter_rows(y, idx, count, rows_out):
    """y fp32 [K, C] (rows >= count are garbage) -> [rows_out, C] with out[idx[j]] = y[j], zeros elsewhere"""
    _chk(y, torch.float32)
    out = torch.empty((rows_out, y.shape[1]), device=y.device, dtype=torch.float32)
    call("scatter_rows_f32", y, idx, count, out, rows_out, y.shape[1])
    return out


def set_live_clips(count, capacity):
    """arm / disarm (count=None) the device-side work limit of the following inference launches of this thread"""
    _lib.lib().cdll.adamml_set_live_clips(count.data_ptr() if count is not None else None, int(capacity))


def first_conv_s2d_ok(conv, C, H, W, dtype):
    """stride-2 first convolutions that run on tcgen05 through the space-to-depth view: 7x7/p3 (ResNet stem) and
    3x3/p1 (MobileNetV2 first conv), even-sized frames, bf16 mode."""
    k = conv.kernel_size
    ok_geom = (k == (7, 7) and conv.padding == (3, 3)) or (k == (3, 3) and conv.padding == (1, 1))
    return (TC_MODE == "auto" and dtype in (torch.bfloat16, PREC_X2) and ok_geom and conv.stride == (2, 2)
            and conv.groups == 1 and H % 2 == 0 and W % 2 == 0 and 4 * C <= 64 and conv.out_channels % 8 == 0)


stem_s2d_ok = first_conv_s2d_ok


def pack_frames_s2d(x, S, F, C, norm=None, x2=False):
    """NCHW fp32 (or uint8 + norm) clip -> S2D operand of the 7x7 ResNet stem (two zero columns on either side)."""
    nm = _u8_norm(x, C, norm)
    N, SFC, H, W = x.shape
    assert SFC == S * F * C, (x.shape, S, F, C)
    Cs = ((4 * C + 15) // 16) * 16
    out = torch.empty((S * N * F, H // 2, W // 2 + 4, Cs), device=x.device, dtype=torch.bfloat16)
    if x2:
        lo = torch.empty_like(out, dtype=torch.float16)
        call("pack_frames_s2d_x2", x, nm[0] if nm else None, nm[1] if nm else None, out, lo, N, S, F, C, H, W, Cs,
             int(nm is not None))
        return S2D(out, C, H, W, 7, lo=lo)
    if nm is None:
        call("pack_frames_s2d", x, out, N, S, F, C, H, W, Cs)
    else:
        call("pack_frames_s2d_u8", x, nm[0], nm[1], out, N, S, F, C, H, W, Cs)
    return S2D(out, C, H, W, 7)


def nhwc_to_s2d(x, R):
    """NHWC bf16 (or x2) image batch -> S2D operand of an RxR stride-2 first conv (pad columns: T/2 left, T/2-1
    right).  The kernel is a pure 2-byte re-layout, so an x2 input is converted plane by plane."""
    IMGS, H, W, C = x.shape
    T = (R + 1) // 2
    Cs = ((4 * C + 7) // 8) * 8
    padl, padr = T // 2, T // 2 - 1
    planes = []
    for t in ((x.hi, x.lo) if isinstance(x, X2) else (_chk(x, torch.bfloat16),)):
        out = torch.empty((IMGS, H // 2, W // 2 + padl + padr, Cs), device=t.device, dtype=t.dtype)
        call("nhwc_to_s2d", t, out, IMGS, C, H, W, Cs, padl, padr)
        planes.append(out)
    return S2D(planes[0], C, H, W, R, lo=planes[1] if len(planes) > 1 else None)


def stem_conv_fwd(xs, w_oihw, stats=None, imgs_per_group=0):
    """xs: S2D, w_oihw: fp32 [Cout, C, R, R] parameter -> z bf16 [IMGS, H/2, W/2, Cout] (+ fused BN statistics)."""
    Cout = w_oihw.shape[0]
    T = xs.taps
    IMGS, Hs, Wp, Cs = xs.t.shape
    Ho, Wo = xs.H // 2, xs.W // 2
    if xs.lo is not None:  # x2 planes
        wp = torch.empty((4, Cout, T, T, xs.Cs), device=xs.device, dtype=torch.bfloat16)
        call("pack_weight_x2", w_oihw, wp, Cout, xs.C, xs.R, xs.R, xs.Cs, 1)
        z = X2.empty((IMGS, Ho, Wo, Cout), xs.device)
        call("tc_stem_conv_x2", xs.t, xs.lo, wp, z.hi, z.lo, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, T, stats, imgs_per_group)
        return z
    wp = torch.empty((Cout, T, T, xs.Cs), device=xs.device, dtype=torch.bfloat16)
    call("pack_weight_stem", w_oihw, wp, Cout, xs.C, xs.Cs, xs.R)
    z = torch.empty((IMGS, Ho, Wo, Cout), device=xs.device, dtype=torch.bfloat16)
    call("tc_stem_conv_bf16", xs.t, wp, z, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, T, stats, imgs_per_group)
    return z


def stem_wgrad(xs, dy, Cout):
    """-> fp32 OIHW gradient [Cout, C, R, R] of the first-conv weight."""
    IMGS, Hs, Wp, Cs = xs.t.shape
    T = xs.taps
    dwp = torch.empty((Cout, T, T, Cs), device=xs.device, dtype=torch.float32)
    call("tc_stem_wgrad_bf16", xs.t, dy, dwp, IMGS, Hs, Wp, Cs, Cout, dy.shape[1], dy.shape[2], T)
    dw = torch.empty((Cout, xs.C, xs.R, xs.R), device=xs.device, dtype=torch.float32)
    call("unpack_wgrad_stem", dwp, dw, Cout, xs.C, Cs, xs.R)
    return dw


# kinds of adamml_pack_weights_multi
PK_OHWI_F32, PK_OHWI_BF16, PK_OHWI_X2, PK_DGRAD_F32, PK_DGRAD_BF16, PK_DW = range(6)


class WeightPackCache:
    """Weight operands of ONE model, refreshed by ONE launch per forward pass (adamml_pack_weights_multi).

    The engine derives every weight operand from the fp32 OIHW parameter each step (OHWI in the compute precision, the
    rotated data-gradient operand, tap-major depthwise weights): ~360 small launches per RGB+Audio training step.  With
    a cache the FIRST step runs those per-layer launches as before and records them as jobs; from the second step on
    `begin()` -- called by AdaMML.forward before any backbone runs -- converts all of them in one launch into one
    persistent arena and the per-layer calls (`pack_weight(..., cache=)`) return views of it.  Entries are keyed by the
    parameter's storage address and hold a reference to it; an operand is handed out only in the pass whose `begin()`
    refreshed it (or that packed it itself), so a weight update between passes can never be missed.  Backbones run
    without a cache (unimodal models, direct engine calls) keep the per-layer launches."""

    class _Entry:
        __slots__ = ("w", "out", "dims", "kind", "in_table", "wrap", "shape", "dtype", "mk", "used")

    PRUNE_EVERY = 16  # passes

    def __init__(self):
        self.ent = {}
        self.epoch = 0
        self.packed_epoch = -1
        self.dirty = False
        self.table = None
        self.arena = None
        self._captured = False
        self._retired = []  # tables / arenas a captured CUDA graph may still read

    def __deepcopy__(self, memo):   # a copied / pickled model starts with an empty cache (entries are keyed by the
        return WeightPackCache()    # ORIGINAL parameters' addresses)

    def __reduce__(self):
        return (WeightPackCache, ())

    def begin(self):
        """start of a forward pass: refresh every recorded operand from the current parameter values"""
        self.epoch += 1
        if not self.ent:
            return
        if self.epoch % self.PRUNE_EVERY == 0 and not torch.cuda.is_current_stream_capturing():
            # operands nobody asked for lately (parameters replaced by .to() / re-assignment, the data-gradient operands
            # of a phase that no longer runs backward) stop being converted and release their parameter reference
            stale = [k for k, e in self.ent.items() if e.used < self.epoch - self.PRUNE_EVERY]
            for k in stale:
                del self.ent[k]
            if stale:
                self.dirty = True
                if not self.ent:
                    self._retire()
                    return
        if self.dirty:
            self._build()
        if torch.cuda.is_current_stream_capturing():
            self._captured = True
        jobs, cj, cs, n_jobs, n_chunks = self.table
        call("pack_weights_multi", jobs, cj, cs, n_jobs, n_chunks)
        self.packed_epoch = self.epoch

    def _retire(self):
        """drop the current table / arena -- unless a CUDA graph captured a launch that reads them"""
        if self.table is not None and self._captured:
            self._retired.append((self.table, self.arena))
        self.table = self.arena = None
        self._captured = False

    def _build(self):
        """job table + ONE arena holding every operand (a single allocation made between two passes: persistent
        per-layer buffers allocated in the middle of a forward pass pin the caching allocator's segments and, at
        119 GiB of activations, send eager steps into its free-and-retry path)"""
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("WeightPackCache: a new weight operand appeared during CUDA-graph capture; run one eager "
                               "step in the same mode before capturing")
        chunk = int(_lib.lib().cdll.adamml_pack_chunk())
        ents = list(self.ent.values())
        dev = ents[0].w.device
        offs, size = [], 0
        for e in ents:
            offs.append(size)
            nbytes = torch.empty((), dtype=e.dtype).element_size()
            for d in e.shape:
                nbytes *= d
            size += (nbytes + 255) // 256 * 256
        arena = torch.empty(size, dtype=torch.uint8, device=dev)
        rows, cj, cs = [], [], []
        for i, (e, off) in enumerate(zip(ents, offs)):
            n = 1
            for d in e.shape:
                n *= d
            nbytes = n * torch.empty((), dtype=e.dtype).element_size()
            e.out = arena[off:off + nbytes].view(e.dtype).view(e.shape)
            e.wrap = e.mk(e.out) if e.mk else e.out
            e.in_table = True
            Cout, Cin, R, S, cin_pad = e.dims
            rows.append([e.w.data_ptr(), e.out.data_ptr(), Cout, Cin, R, S, cin_pad, e.kind])
            total = Cout * 9 if e.kind == PK_DW else (Cin * R * S * Cout if e.kind in (PK_DGRAD_F32, PK_DGRAD_BF16)
                                                      else Cout * R * S * cin_pad)
            for o in range(0, total, chunk):
                cj.append(i)
                cs.append(o)
        self._retire()
        self.arena = arena
        self.table = (torch.tensor(rows, dtype=torch.int64, device=dev), torch.tensor(cj, dtype=torch.int32, device=dev),
                      torch.tensor(cs, dtype=torch.int64, device=dev), len(rows), len(cj))
        self.dirty = False

    def get(self, w, kind, dims, shape, dtype, launch, wrap=None):
        key = (w.data_ptr(), kind)
        e = self.ent.get(key)
        if e is None:
            e = WeightPackCache._Entry()
            e.w, e.dims, e.kind, e.in_table, e.shape, e.dtype, e.mk = w, dims, kind, False, tuple(shape), dtype, wrap
            e.out = e.wrap = None
            self.ent[key] = e
            self.dirty = True
        e.used = self.epoch
        if e.in_table and self.packed_epoch == self.epoch:
            return e.wrap
        # not (yet) refreshed by begin() in this pass: an ordinary per-layer launch into a buffer of its own
        out = torch.empty(e.shape, device=w.device, dtype=dtype)
        launch(out)
        return wrap(out) if wrap else out


def pack_weight(w, dtype, cin_pad=None, cache=None):
    """OIHW fp32 parameter -> OHWI operand [Cout, R, S, cin_pad] in `dtype`."""
    _chk(w, torch.float32)
    Cout, Cin, R, S = w.shape
    cin_pad = cin_pad or Cin
    if dtype == PREC_X2:
        def launch(out):
            call("pack_weight_x2", w, out, Cout, Cin, R, S, cin_pad, 0)
        if cache is not None:
            return cache.get(w, PK_OHWI_X2, (Cout, Cin, R, S, cin_pad), (4, Cout, R, S, cin_pad), torch.bfloat16, launch,
                             wrap=X2W)
        out = torch.empty((4, Cout, R, S, cin_pad), device=w.device, dtype=torch.bfloat16)
        launch(out)
        return X2W(out)
    if R == 1 and S == 1 and cin_pad == Cin and dtype == torch.float32:
        return w.view(Cout, 1, 1, Cin)

    def launch(out):
        call("pack_weight", w, out, Cout, Cin, R, S, cin_pad, dtype_code(dtype))
    if cache is not None:
        return cache.get(w, PK_OHWI_F32 if dtype == torch.float32 else PK_OHWI_BF16, (Cout, Cin, R, S, cin_pad),
                         (Cout, R, S, cin_pad), dtype, launch)
    out = torch.empty((Cout, R, S, cin_pad), device=w.device, dtype=dtype)
    launch(out)
    return out


def pack_weight_dgrad(w, dtype, cache=None):
    """OIHW fp32 parameter -> rotated dgrad operand [Cin, R, S, Cout] in `dtype`."""
    _chk(w, torch.float32)
    Cout, Cin, R, S = w.shape

    def launch(out):
        call("pack_weight_dgrad", w, out, Cout, Cin, R, S, dtype_code(dtype))
    if cache is not None:
        return cache.get(w, PK_DGRAD_F32 if dtype == torch.float32 else PK_DGRAD_BF16, (Cout, Cin, R, S, Cin),
                         (Cin, R, S, Cout), dtype, launch)
    out = torch.empty((Cin, R, S, Cout), device=w.device, dtype=dtype)
    launch(out)
    return out


def unpack_wgrad(dw_ohwi, Cin):
    Cout, R, S, cin_pad = dw_ohwi.shape
    _chk(dw_ohwi, torch.float32)
    if R == 1 and S == 1 and cin_pad == Cin:
        return dw_ohwi.view(Cout, Cin, 1, 1)
    out = torch.empty((Cout, Cin, R, S), device=dw_ohwi.device, dtype=torch.float32)
    call("unpack_wgrad", dw_ohwi, out, Cout, Cin, R, S, cin_pad, 0)
    return out


def cast(x, dtype):
    _chk(x)
    out = torch.empty_like(x, dtype=dtype)
    call("cast", x, out, x.numel(), dtype_code(x.dtype), dtype_code(dtype))
    return out


# ---------------------------------------------------------------- dense conv
def _tc_ok(x, Cin, Cout, R, S, stride, pad):
    """plain tcgen05 GEMM: 1x1, stride 1"""
    return (TC_MODE == "auto" and x.dtype == torch.bfloat16 and R == 1 and S == 1 and stride == 1 and pad == 0
            and Cin % 8 == 0 and Cout % 8 == 0)


def _tc_conv_ok(x, Cin, Cout, R, S, stride):
    """tcgen05 implicit GEMM (4D TMA taps): any RxS <= 49 taps, stride 1|2.  Cin >= 32 keeps the 64-wide K
    block at least half full (the 3-channel stem stays on the exact engine)."""
    return (TC_MODE == "auto" and x.dtype == torch.bfloat16 and Cin % 8 == 0 and Cout % 8 == 0 and Cin >= 16
            and R * S <= 49 and stride in (1, 2))


def conv_fwd(x, w, stride, pad, out=None, stats=None, rows_per_group=0):
    """x [IMGS,H,W,Cin], w OHWI [Cout,R,S,Cin] -> y [IMGS,Ho,Wo,Cout].

    If `stats` (double [G,Cout,2]) is given and the tcgen05 engine takes the layer, the BN
    batch statistics are produced by the GEMM epilogue and True is returned as second value.
    """
    if isinstance(x, X2):
        return _conv_fwd_x2(x, w, stride, pad, stats, rows_per_group)
    _chk(x); _chk(w, x.dtype)
    IMGS, H, W, Cin = x.shape
    Cout, R, S, Cw = w.shape
    assert Cw == Cin, (w.shape, x.shape)
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    if out is None:
        out = torch.empty((IMGS, Ho, Wo, Cout), device=x.device, dtype=x.dtype)
    if _tc_ok(x, Cin, Cout, R, S, stride, pad):
        call("tc_gemm_bf16", x, w, out, IMGS * H * W, Cout, Cin, 0, 0, 0, _lib.BF16, stats, rows_per_group)
        return out, stats is not None
    if _tc_conv_ok(x, Cin, Cout, R, S, stride):
        ipg = rows_per_group // (Ho * Wo) if stats is not None else 0
        call("tc_conv_bf16", x, w, out, None, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stats, ipg, 0)
        return out, stats is not None
    call("simt_conv_fwd", x, w, out, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, 0, 0, 0, dtype_code(x.dtype))
    return out, False


def _conv_fwd_x2(x, w, stride, pad, stats, rows_per_group):
    """x2 forward convolution: x, w are X2 (w from pack_weight(.., PREC_X2)) -> (X2 z, stats fused?)."""
    IMGS, H, W, Cin = x.shape
    Cout, R, S, Cw = w.shape
    assert Cw == Cin and isinstance(w, X2W), (w.shape, x.shape)
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    if Cin % 8 or Cout % 8 or R * S > 49 or stride not in (1, 2) or TC_MODE != "auto":
        raise NotImplementedError("x2 precision needs tcgen05-shaped dense convolutions (Cin, Cout multiples of 8); "
                                  "got Cin=%d Cout=%d %dx%d stride %d" % (Cin, Cout, R, S, stride))
    z = X2.empty((IMGS, Ho, Wo, Cout), x.device)
    if R == 1 and S == 1 and stride == 1 and pad == 0:
        call("tc_gemm_x2", x.hi, x.lo, w.planes, z.hi, z.lo, IMGS * H * W, Cout, Cin, stats, rows_per_group)
    else:
        ipg = rows_per_group // (Ho * Wo) if stats is not None else 0
        call("tc_conv_x2", x.hi, x.lo, w.planes, z.hi, z.lo, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, stats,
             ipg)
    return z, stats is not None


# ---------------------------------------------------------------- inference: conv + BN (+ residual) + act, one kernel
def conv_bn_act_fwd(x, w, stride, pad, ss, act, res=None):
    """out = act(conv(x, w) * scale + shift (+ res)) in ONE tcgen05 kernel (fused inference epilogue, see the header).
    x: bf16 NHWC or X2; w from pack_weight in the same mode; ss: fp32 [Cout, 2]; res: tensor of the output's shape.
    -> None when the layer is outside the tensor-core envelope (the caller then runs the unfused ops)."""
    IMGS, H, W, Cin = x.shape
    Cout, R, S, Cw = w.shape
    assert Cw == Cin, (w.shape, x.shape)
    Ho, Wo = conv_out_hw(H, W, R, S, stride, pad)
    _chk(ss, torch.float32)
    if TC_MODE != "auto" or Cin % 8 or Cout % 8 or R * S > 49 or stride not in (1, 2):
        return None
    gemm = R == 1 and S == 1 and stride == 1 and pad == 0
    if isinstance(x, X2):
        out = X2.empty((IMGS, Ho, Wo, Cout), x.device)
        rh, rl = (res.hi, res.lo) if res is not None else (None, None)
        if gemm:
            call("tc_gemm_bn_act_x2", x.hi, x.lo, w.planes, out.hi, out.lo, IMGS * H * W, Cout, Cin, ss, act, rh, rl)
        else:
            call("tc_conv_bn_act_x2", x.hi, x.lo, w.planes, out.hi, out.lo, IMGS, H, W, Cin, Cout, R, S, stride, pad,
                 Ho, Wo, ss, act, rh, rl)
        return out
    if x.dtype != torch.bfloat16 or (not gemm and Cin < 16):
        return None
    _chk(x); _chk(w, torch.bfloat16)
    out = torch.empty((IMGS, Ho, Wo, Cout), device=x.device, dtype=torch.bfloat16)
    if gemm:
        call("tc_gemm_bn_act_bf16", x, w, out, IMGS * H * W, Cout, Cin, ss, act, res)
    else:
        call("tc_conv_bn_act_bf16", x, w, out, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, ss, act, res)
    return out


def stem_conv_bn_act_fwd(xs, w_oihw, ss, act):
    """fused inference epilogue on the space-to-depth first convolutions (see stem_conv_fwd)"""
    Cout = w_oihw.shape[0]
    T = xs.taps
    IMGS, Hs, Wp, Cs = xs.t.shape
    Ho, Wo = xs.H // 2, xs.W // 2
    if xs.lo is not None:
        wp = torch.empty((4, Cout, T, T, xs.Cs), device=xs.device, dtype=torch.bfloat16)
        call("pack_weight_x2", w_oihw, wp, Cout, xs.C, xs.R, xs.R, xs.Cs, 1)
        out = X2.empty((IMGS, Ho, Wo, Cout), xs.device)
        call("tc_stem_conv_bn_act_x2", xs.t, xs.lo, wp, out.hi, out.lo, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, T, ss, act)
        return out
    wp = torch.empty((Cout, T, T, xs.Cs), device=xs.device, dtype=torch.bfloat16)
    call("pack_weight_stem", w_oihw, wp, Cout, xs.C, xs.Cs, xs.R)
    out = torch.empty((IMGS, Ho, Wo, Cout), device=xs.device, dtype=torch.bfloat16)
    call("tc_stem_conv_bn_act_bf16", xs.t, wp, out, IMGS, Hs, Wp, Cs, Cout, Ho, Wo, T, ss, act)
    return out


def dwconv_bn_act_fwd(x, w, stride, ss, act):
    """inference: y = act(dwconv(x) * scale + shift) in one pass; -> None for ragged channel counts"""
    IMGS, H, W, C = x.shape
    Ho, Wo = conv_out_hw(H, W, 3, 3, stride, 1)
    if isinstance(x, X2):
        y = X2.empty((IMGS, Ho, Wo, C), x.device)
        call("dwconv_bn_act_fwd_x2", x.hi, x.lo, w, y.hi, y.lo, IMGS, H, W, C, stride, Ho, Wo, ss, act)
        return y
    if not vec_channels(x):
        return None
    y = torch.empty((IMGS, Ho, Wo, C), device=x.device, dtype=x.dtype)
    call("dwconv_bn_act_fwd", x, w, y, IMGS, H, W, C, stride, Ho, Wo, ss, act, dtype_code(x.dtype))
    return y


def tc_dgrad_ok(dtype, Cout, Cin, R, S, stride):
    """data gradients that run on the tcgen05 engine (bf16): stride 1, stride-2 RxS (parity classes),
    stride-2 1x1 (compact GEMM, see conv_dgrad_compact)."""
    return (TC_MODE == "auto" and dtype == torch.bfloat16 and Cout % 8 == 0 and Cin % 8 == 0 and Cout >= 16
            and R * S <= 49 and stride in (1, 2))


def conv_dgrad_compact(dy, w_rot):
    """Stride-2 1x1 conv: the non-zero part of dx, [IMGS, Ho, Wo, Cin] = dy . w (one plain GEMM).  The caller
    scatters it onto the even pixels of the input grid (conv_dgrad(..., addend_sub=2))."""
    IMGS, Ho, Wo, Cout = dy.shape
    Cin = w_rot.shape[0]
    dxc = torch.empty((IMGS, Ho, Wo, Cin), device=dy.device, dtype=dy.dtype)
    call("tc_gemm_bf16", dy, w_rot, dxc, IMGS * Ho * Wo, Cin, Cout, 0, 0, 0, _lib.BF16, None, 0)
    return dxc


def conv_dgrad(dy, w, x_shape, stride, pad, addend=None, w_rot=None, addend_sub=1):
    """dx = conv_transpose(dy, w) (+ addend).  w_rot: optional rotated operand [Cin, R, S, Cout] from
    pack_weight_dgrad (bf16): the layer then runs as forward conv(s) of dy on the tcgen05 engine.
    addend_sub=2: addend is a compact stride-2 gradient (conv_dgrad_compact) added at even pixels."""
    _chk(dy); _chk(w, dy.dtype)
    IMGS, H, W, Cin = x_shape
    Cout, R, S, _ = w.shape
    Ho, Wo = dy.shape[1], dy.shape[2]
    dx = torch.empty(x_shape, device=dy.device, dtype=dy.dtype)
    if w_rot is not None and stride == 1 and _tc_conv_ok(dy, Cout, Cin, R, S, 1):
        if addend is None and R == 1 and S == 1:
            call("tc_gemm_bf16", dy, w_rot, dx, IMGS * H * W, Cin, Cout, 0, 0, 0, _lib.BF16, None, 0)
        else:
            call("tc_conv_bf16", dy, w_rot, dx, addend, IMGS, Ho, Wo, Cout, Cin, R, S, 1, R - 1 - pad, H, W, None, 0,
                 addend_sub if addend is not None else 0)
        return dx
    assert addend_sub == 1 or addend is None, "compact addends need the tcgen05 stride-1 path"
    if w_rot is not None and stride == 2 and R >= 2 and S >= 2 and addend is None:
        call("tc_dgrad_s2_bf16", dy, w_rot, dx, IMGS, H, W, Cin, Cout, R, S, pad, Ho, Wo)
        return dx
    call("simt_conv_dgrad", dy, w, dx, addend, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, 0, 0, 0,
         dtype_code(dy.dtype))
    return dx


def conv_wgrad(x, dy, w_shape, stride, pad):
    """-> dw fp32 OHWI [Cout,R,S,Cin]."""
    _chk(x); _chk(dy, x.dtype)
    IMGS, H, W, Cin = x.shape
    Cout, R, S, _ = w_shape
    Ho, Wo = dy.shape[1], dy.shape[2]
    dw = torch.empty((Cout, R, S, Cin), device=x.device, dtype=torch.float32)
    if (TC_MODE == "auto" and x.dtype == torch.bfloat16 and Cin % 8 == 0 and Cout % 8 == 0 and R * S <= 49
            and stride in (1, 2)):
        call("tc_wgrad_bf16", x, dy, dw, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo)
        return dw
    call("simt_conv_wgrad", x, dy, dw, IMGS, H, W, Cin, Cout, R, S, stride, pad, Ho, Wo, 0, 0, 0,
         dtype_code(x.dtype))
    return dw


# ---------------------------------------------------------------- fp32 linear algebra (H=W=1 convs)
def linear_fwd(x, w, out=None, x_ld=0, y_ld=0, w_ld=0, K=None):
    """y[rows, Nout] = x[rows, :K] . w[Nout, :K]^T   (fp32, strided)."""
    rows = x.shape[0]
    Nout = w.shape[0]
    K = K or w.shape[1]
    if out is None:
        out = torch.empty((rows, Nout), device=x.device, dtype=torch.float32)
    call("simt_conv_fwd", x, w, out, rows, 1, 1, K, Nout, 1, 1, 1, 0, 1, 1, x_ld or x.stride(0), y_ld or out.stride(0),
         w_ld or w.stride(0), _lib.F32)
    return out


def linear_dgrad(dy, w, K=None, out=None, w_ld=0, x_ld=0, y_ld=0):
    """dx[rows, K] = dy[rows, Nout] . w[Nout, :K]."""
    rows, Nout = dy.shape[0], w.shape[0]
    K = K or w.shape[1]
    if out is None:
        out = torch.empty((rows, K), device=dy.device, dtype=torch.float32)
    call("simt_conv_dgrad", dy, w, out, None, rows, 1, 1, K, Nout, 1, 1, 1, 0, 1, 1, x_ld or out.stride(0),
         y_ld or dy.stride(0), w_ld or w.stride(0), _lib.F32)
    return out


def linear_wgrad(x, dy, K=None, x_ld=0, y_ld=0):
    """dw[Nout, K] = dy[rows, Nout]^T . x[rows, :K]."""
    rows, Nout = dy.shape
    K = K or x.shape[1]
    dw = torch.empty((Nout, K), device=x.device, dtype=torch.float32)
    call("simt_conv_wgrad", x, dy, dw, rows, 1, 1, K, Nout, 1, 1, 1, 0, 1, 1, x_ld or x.stride(0),
         y_ld or dy.stride(0), 0, _lib.F32)
    return dw


# ---------------------------------------------------------------- depthwise
def pack_weight_dw(w, cache=None):
    """nn.Conv2d(groups=C).weight [C,1,3,3] fp32 -> tap-major operand [9, C] fp32."""
    _chk(w, torch.float32)
    C = w.shape[0]

    def launch(out):
        call("pack_weight_dw", w, out, C)
    if cache is not None:
        return cache.get(w, PK_DW, (C, 1, 3, 3, 1), (9, C), torch.float32, launch)
    out = torch.empty((9, C), device=w.device, dtype=torch.float32)
    launch(out)
    return out


def dwconv_fwd(x, w, stride):
    """w: tap-major [9, C] (pack_weight_dw)."""
    _chk(w, torch.float32)
    assert w.shape == (9, x.shape[-1]), "depthwise weights must be tap-major [9, C] (ops.pack_weight_dw)"
    IMGS, H, W, C = x.shape
    Ho, Wo = conv_out_hw(H, W, 3, 3, stride, 1)
    if isinstance(x, X2):
        y = X2.empty((IMGS, Ho, Wo, C), x.device)
        call("dwconv_fwd_x2", x.hi, x.lo, w, y.hi, y.lo, IMGS, H, W, C, stride, Ho, Wo)
        return y
    _chk(x)
    y = torch.empty((IMGS, Ho, Wo, C), device=x.device, dtype=x.dtype)
    call("dwconv_fwd", x, w, y, IMGS, H, W, C, stride, Ho, Wo, dtype_code(x.dtype))
    return y


DW_TMA_FWD = os.environ.get("ADAMML_B200_DW_TMA_FWD", "1") != "0"


def dwconv_fwd_stats(x, w, stride, stats, imgs_per_group):
    """Training forward of a depthwise conv with the BatchNorm statistics of its output: -> (z, fused).  x2 planes /
    stride 1 / C % 16 == 0 run the TMA-tile kernel that fills `stats` ([G, C, 2] float64: sum, sum of squares per
    group of imgs_per_group images) in the same pass (csrc/dwconv_tma.cu); otherwise fused is False and the caller
    runs bn_stats over z."""
    IMGS, H, W, C = x.shape
    if DW_TMA_FWD and isinstance(x, X2) and stride == 1 and C % 16 == 0 and stats is not None:
        _chk(w, torch.float32)
        assert w.shape == (9, C) and stats.dtype == torch.float64 and IMGS % imgs_per_group == 0
        y = X2.empty((IMGS, H, W, C), x.device)
        call("dwconv_fwd_stats_x2", x.hi, x.lo, w, y.hi, y.lo, stats, IMGS, H, W, C, imgs_per_group)
        return y, True
    return dwconv_fwd(x, w, stride), False


def dwconv_dgrad(dy, w, x_shape, stride, addend=None):
    IMGS, H, W, C = x_shape
    dx = torch.empty(x_shape, device=dy.device, dtype=dy.dtype)
    call("dwconv_dgrad", dy, w, dx, addend, IMGS, H, W, C, stride, dy.shape[1], dy.shape[2], dtype_code(dy.dtype))
    return dx


DW_FUSED_BWD = os.environ.get("ADAMML_B200_DW_FUSED_BWD", "1") != "0"


def dwconv_bwd_ok(x, dy, stride):
    """the fused TMA-tile backward (csrc/dwconv_tma.cu) handles bf16 tensors with C % 16 == 0"""
    return (DW_FUSED_BWD and isinstance(x, torch.Tensor) and x.dtype == torch.bfloat16 and dy.dtype == torch.bfloat16
            and x.shape[-1] % 16 == 0 and stride in (1, 2))


def dwconv_bwd(x, dy, w, stride, pre=None):
    """-> (dx, dw): data gradient (bf16 NHWC) and fp32 weight gradient in torch's [C,1,3,3] layout from ONE pass
    over dy and x.  w: tap-major [9, C] (pack_weight_dw).
    pre = (raw_sums [G, C, 2] float64, imgs_per_group, act): also fuse the BatchNorm-backward reduction of the layer
    that produced x = act(bn(z)): dx comes back masked (dx * act'(x)) and raw_sums holds (sum gm, sum gm * x) per
    group and channel (-> bn_sums_from_out)."""
    _chk(x); _chk(dy, x.dtype); _chk(w, torch.float32)
    IMGS, H, W, C = x.shape
    dx = torch.empty((IMGS, H, W, C), device=x.device, dtype=x.dtype)
    dwt = torch.empty((9, C), device=x.device, dtype=torch.float32)
    raw, ipg, act = pre if pre is not None else (None, 0, ACT_NONE)
    if raw is not None:
        assert raw.dtype == torch.float64 and raw.is_contiguous() and IMGS % ipg == 0
        assert tuple(raw.shape) == (IMGS // ipg, C, 2)
    call("dwconv_bwd", x, dy, w, dx, dwt, raw, ipg, act, IMGS, H, W, C, stride, dy.shape[1], dy.shape[2])
    dw = torch.empty((C, 1, 3, 3), device=x.device, dtype=torch.float32)
    call("unpack_wgrad_dw", dwt, dw, C)
    return dx, dw


def bn_sums_from_out(raw, scale_shift, mean_invstd, out=None):
    """(sum gm, sum gm * out) of dwconv_bwd's fused reduction -> (sum gm, sum gm * xhat) as bn_bwd_reduce returns"""
    G, C, _ = raw.shape
    sums = out if out is not None else raw
    call("bn_sums_from_out", raw, scale_shift, mean_invstd, sums, C, G)
    return sums


def dwconv_wgrad(x, dy, stride):
    """-> fp32 gradient in torch's [C,1,3,3] layout."""
    IMGS, H, W, C = x.shape
    dwt = torch.empty((9, C), device=x.device, dtype=torch.float32)
    call("dwconv_wgrad", x, dy, dwt, IMGS, H, W, C, stride, dy.shape[1], dy.shape[2], dtype_code(x.dtype))
    dw = torch.empty((C, 1, 3, 3), device=x.device, dtype=torch.float32)
    call("unpack_wgrad_dw", dwt, dw, C)
    return dw


# ---------------------------------------------------------------- batch norm
def bn_stats(z, G, out=None):
    C = z.shape[-1]
    rows = z.numel() // C
    sums = out if out is not None else torch.empty((G, C, 2), device=z.device, dtype=torch.float64)
    if isinstance(z, X2):
        call("bn_stats_x2", z.hi, z.lo, sums, rows // G, C, G)
    else:
        call("bn_stats", z, sums, rows // G, C, G, dtype_code(z.dtype))
    return sums


def bn_finalize(sums, gamma, beta, running_mean, running_var, count, momentum, eps, C, G, training, update_running):
    dev = gamma.device
    mean_invstd = torch.empty((G, C, 2), device=dev, dtype=torch.float32)
    scale_shift = torch.empty((G, C, 2), device=dev, dtype=torch.float32)
    call("bn_finalize", sums, gamma, beta, running_mean, running_var, mean_invstd, scale_shift, float(count),
         float(momentum), float(eps), C, G, int(training), int(update_running))
    return mean_invstd, scale_shift


def bn_apply(z, scale_shift, G, act, res=None, res_z=None, res_ss=None, out=None, mask_bits=None):
    """mask_bits (x2 only): uint8 [rows, C // 8] that receives the 1-bit activation mask of every output element
    (read by bn_bwd_reduce instead of the saved output)."""
    C = z.shape[-1]
    rows = z.numel() // C
    if isinstance(z, X2):
        if out is None:
            out = X2.empty(z.shape, z.device)
        if mask_bits is not None:
            assert mask_bits.dtype == torch.uint8 and mask_bits.numel() == rows * (C // 8) and C % 8 == 0
        call("bn_apply_x2", z.hi, z.lo, scale_shift, res.hi if res is not None else None,
             res.lo if res is not None else None, res_z.hi if res_z is not None else None,
             res_z.lo if res_z is not None else None, res_ss, out.hi, out.lo, rows // G, C, G, act, mask_bits)
        return out
    assert mask_bits is None
    if out is None:
        out = torch.empty_like(z)
    call("bn_apply", z, scale_shift, res, res_z, res_ss, out, rows // G, C, G, act, dtype_code(z.dtype))
    return out


def _mask_ss(z, mask_ss):
    """the z-recomputed activation mask needs the vectorised kernels (C % 16 bytes == 0)"""
    if mask_ss is None or z.shape[-1] % (8 if z.dtype == torch.bfloat16 else 4):
        return None
    return mask_ss


def vec_channels(t):
    """True when the row-streaming (16-byte vector) kernels take this tensor."""
    return t.shape[-1] % (8 if t.dtype == torch.bfloat16 else 4) == 0


def bn_bwd_reduce(dout, out, z, mean_invstd, G, act, mask_ss=None, gm_inplace=False, sums_out=None, mask_bits=None):
    """mask_ss: forward scale/shift [G,C,2] of a layer without residual input -> the ReLU/ReLU6 mask is recomputed
    from z and `out` is not read.  gm_inplace: dout is overwritten with the masked gradient dout * act'(out), which
    is also the gradient of a residual input; the following bn_bwd_apply then runs with act = NONE."""
    C = z.shape[-1]
    rows = z.numel() // C
    sums = sums_out if sums_out is not None else torch.empty((G, C, 2), device=z.device, dtype=torch.float64)
    call("bn_bwd_reduce", dout, None if mask_bits is not None else out, z, mean_invstd, _mask_ss(z, mask_ss), sums,
         dout if gm_inplace else None, rows // G, C, G, act, dtype_code(z.dtype), mask_bits)
    return sums


def bn_bwd_apply(dout, out, z, mean_invstd, gamma, sums, G, count, act, training, want_dz=True, want_dres=False,
                 mask_ss=None):
    C = dout.shape[-1]
    rows = dout.numel() // C
    dz = torch.empty_like(dout) if want_dz else None
    dres = torch.empty_like(dout) if want_dres else None
    call("bn_bwd_apply", dout, out, z, mean_invstd, gamma, _mask_ss(z, mask_ss), sums, dz, dres, rows // G, C, G,
         float(count), act, int(training), dtype_code(dout.dtype))
    return dz, dres


def bn_param_grad(sums, C, G):
    dgamma = torch.empty(C, device=sums.device, dtype=torch.float32)
    dbeta = torch.empty(C, device=sums.device, dtype=torch.float32)
    call("bn_param_grad", sums, dgamma, dbeta, C, G, 0)
    return dgamma, dbeta


# ---------------------------------------------------------------- pools
def maxpool_fwd(x, want_pos=False):
    """-> y, or (y, pos) with pos = uint8 window position of every maximum (for the gather backward)."""
    IMGS, H, W, C = x.shape
    Ho, Wo = conv_out_hw(H, W, 3, 3, 2, 1)
    if isinstance(x, X2):
        y = X2.empty((IMGS, Ho, Wo, C), x.device)
        pos = torch.empty((IMGS, Ho, Wo, C), device=x.device, dtype=torch.uint8) if want_pos else None
        call("maxpool3x3s2_fwd_x2", x.hi, x.lo, y.hi, y.lo, pos, IMGS, H, W, C, Ho, Wo)
        return (y, pos) if want_pos else y
    y = torch.empty((IMGS, Ho, Wo, C), device=x.device, dtype=x.dtype)
    pos = None
    if want_pos and C % (8 if x.dtype == torch.bfloat16 else 4) == 0:
        pos = torch.empty((IMGS, Ho, Wo, C), device=x.device, dtype=torch.uint8)
    call("maxpool3x3s2_fwd", x, y, pos, IMGS, H, W, C, Ho, Wo, dtype_code(x.dtype))
    return (y, pos) if want_pos else y


def bn_act_maxpool_fwd(z, scale_shift, G, act):
    """x2 training stem: (maxpool3x3s2(act(z * scale + shift)), pos) from the pre-BN planes in one pass"""
    assert isinstance(z, X2)
    IMGS, H, W, C = z.shape
    Ho, Wo = conv_out_hw(H, W, 3, 3, 2, 1)
    y = X2.empty((IMGS, Ho, Wo, C), z.device)
    pos = torch.empty((IMGS, Ho, Wo, C), device=z.device, dtype=torch.uint8)
    call("bn_act_maxpool3x3s2_fwd_x2", z.hi, z.lo, scale_shift, IMGS // G, act, y.hi, y.lo, pos, IMGS, H, W, C, Ho, Wo)
    return y, pos


def maxpool_bwd(x, dy, pos=None, x_shape=None):
    """dx from either the forward input x or the recorded positions (then x may be None, pass x_shape)."""
    IMGS, H, W, C = x_shape if x is None else x.shape
    dx = torch.empty((IMGS, H, W, C), device=dy.device, dtype=dy.dtype)
    call("maxpool3x3s2_bwd", x if pos is None else None, pos, dy, dx, IMGS, H, W, C, dy.shape[1], dy.shape[2],
         dtype_code(dy.dtype))
    return dx


def tpool_fwd(x, T, mode_avg=False):
    IMGS, H, W, C = x.shape
    V = IMGS // T
    To = (T + 2 - 3) // 2 + 1
    if isinstance(x, X2):
        y = X2.empty((V * To, H, W, C), x.device)
        call("tpool_fwd_x2", x.hi, x.lo, y.hi, y.lo, V, T, H * W * C, int(mode_avg))
        return y
    y = torch.empty((V * To, H, W, C), device=x.device, dtype=x.dtype)
    call("tpool_fwd", x, y, V, T, H * W * C, int(mode_avg), dtype_code(x.dtype))
    return y


def tpool_bwd(x, dy, T, mode_avg=False):
    IMGS, H, W, C = x.shape
    dx = torch.empty_like(x)
    call("tpool_bwd", x, dy, dx, IMGS // T, T, H * W * C, int(mode_avg), dtype_code(x.dtype))
    return dx


def avgpool_fwd(x, out=None, out_ld=0):
    IMGS, H, W, C = x.shape
    if out is None:
        out = torch.empty((IMGS, C), device=x.device, dtype=torch.float32)
    if isinstance(x, X2):
        call("avgpool_fwd_x2", x.hi, x.lo, out, IMGS, H * W, C, out_ld or out.stride(0))
    else:
        call("avgpool_fwd", x, out, IMGS, H * W, C, out_ld or out.stride(0), dtype_code(x.dtype))
    return out


def avgpool_bwd(dy, x_shape, dtype, dy_ld=0):
    IMGS, H, W, C = x_shape
    dx = torch.empty(x_shape, device=dy.device, dtype=dtype)
    call("avgpool_bwd", dy, dx, IMGS, H * W, C, dy_ld or dy.stride(0), dtype_code(dtype))
    return dx


def frame_mean(x, T, out=None, out_ld=0):
    rows, C = x.shape
    V = rows // T
    if out is None:
        out = torch.empty((V, C), device=x.device, dtype=torch.float32)
    call("frame_mean", x, out, V, T, C, out_ld or out.stride(0))
    return out


def frame_mean_bwd(dy, T, dy_ld=0):
    V, C = dy.shape
    dx = torch.empty((V * T, C), device=dy.device, dtype=torch.float32)
    call("frame_mean_bwd", dy, dx, V, T, C, dy_ld or dy.stride(0))
    return dx


# ---------------------------------------------------------------- small fp32 helpers
def bias_act_(y, bias, act, rows=None, cols=None, ld=0):
    rows = rows if rows is not None else y.shape[0]
    cols = cols if cols is not None else y.shape[1]
    call("bias_act", y, bias, rows, cols, ld or y.stride(0), act)
    return y


def act_bwd(dy, y, act, rows=None, cols=None, ld_dy=0, ld_y=0):
    rows = rows if rows is not None else dy.shape[0]
    cols = cols if cols is not None else dy.shape[1]
    dz = torch.empty((rows, cols), device=dy.device, dtype=torch.float32)
    call("act_bwd", dy, y, dz, rows, cols, ld_dy or dy.stride(0), ld_y or y.stride(0), cols, act)
    return dz


def colsum(x, rows=None, cols=None, ld=0):
    rows = rows if rows is not None else x.shape[0]
    cols = cols if cols is not None else x.shape[1]
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    call("colsum", x, out, rows, cols, ld or x.stride(0), 0)
    return out


def mul(a, b):
    out = torch.empty_like(a)
    call("mul", a, b, out, a.numel())
    return out
