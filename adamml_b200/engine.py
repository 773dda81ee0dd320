"""Execution engine: runs a backbone as a flat sequence of fused CUDA ops with a manual tape.

A backbone (ResNet / sound MobileNetV2 / policy MobileNetV2) is executed for ALL segments
of ALL videos in one batched pass (images ordered segment-major, one BatchNorm group per
segment, see csrc/bn.cu) instead of the reference's Python loop over segments
(models/adamml.py:84-86, models/policy_net.py:323-326).  Forward pushes records on a tape,
backward pops them; both only call the C-ABI (adamml_b200.ops).  torch is used for memory,
streams, autograd plumbing and (sync-BN) torch.distributed collectives.
"""
import torch

from . import ops
from .dist_utils import P2PStats, allreduce_stats, sync_bn_group
from .ops import ACT_NONE, ACT_RELU, ACT_RELU6


# inference: conv + BN (+ residual) + act fused into one kernel (Exec._cba_fused_eval); ADAMML_B200_FUSE_EVAL=0 keeps the
# conv -> bn_apply pair (tests compare the two)
import os as _os
FUSE_EVAL = _os.environ.get("ADAMML_B200_FUSE_EVAL", "1") != "0"
# training memory: ADAMML_B200_RECOMPUTE=1 does not keep the post-activation output of layers WITHOUT a residual input
# for the backward pass; the consumer's weight gradient rebuilds it from the saved pre-BN tensor (one extra bf16
# bn_apply pass per such layer in backward).  Measured RGB+Audio N=72: see DESIGN.md §4.
RECOMPUTE = _os.environ.get("ADAMML_B200_RECOMPUTE", "0") != "0"
# backward of a depthwise conv also reduces the BatchNorm gradient sums of the layer that produced its input
# (csrc/dwconv_tma.cu PreReduce); ADAMML_B200_DW_FUSE_PRE=0 keeps the separate bn_bwd_reduce pass (tests compare)
DW_FUSE_PRE = _os.environ.get("ADAMML_B200_DW_FUSE_PRE", "1") != "0"
# training stem of the ResNets: BN + ReLU applied inside the max-pool kernel (Exec.cba_maxpool); =0 keeps bn_apply + pool
FUSE_STEM_POOL = _os.environ.get("ADAMML_B200_FUSE_STEM_POOL", "1") != "0"
# residual layers keep a 1-bit ReLU mask of their output for the BatchNorm backward reduction (=0: it re-reads `out`)
MASK_BITS = _os.environ.get("ADAMML_B200_MASK_BITS", "1") != "0"


class _ShapeOnly:
    """stand-in for a forward tensor that was not kept and is needed only for its shape (frozen weights)"""

    def __init__(self, shape):
        self.shape = torch.Size(shape)


def _bn_key(bn):
    """slot key of a BatchNorm layer in the peer-memory arena: its index in the model's module order (assigned by
    assign_bn_keys, identical on every rank) rather than a per-process id()"""
    return getattr(bn, "_adamml_bn_key", None) or ("id", id(bn))


def assign_bn_keys(model):
    """number the BatchNorm layers of `model` in module order (after convert_sync_batchnorm has replaced them)"""
    k = 0
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            k += 1
            m._adamml_bn_key = ("bn", k)


class Exec:
    """State of one backbone pass."""

    def __init__(self, dtype, training, groups, save, param_needs_grad=True, lane=0, packs=None):
        self.lane = lane  # stream / backbone index: the sync-BN peer-memory exchange keeps one flag row per lane
        self.packs = packs  # ops.WeightPackCache of the model this pass belongs to (None: per-layer pack launches)
        # x2 precision: the FORWARD pass runs on two-plane activations (ops.X2); the tape keeps their bf16 hi planes
        # and the backward pass is the bf16 engine
        self.x2 = dtype == ops.PREC_X2
        self.dtype = torch.bfloat16 if self.x2 else dtype
        self.training = training
        self.G = groups
        self.save = save
        self.param_needs_grad = param_needs_grad
        self.tape = []
        self.grads = {}  # parameter -> gradient tensor (filled by bwd)
        self.nbt = []    # num_batches_tracked buffers of the BatchNorm layers this pass ran in train mode
        self.recompute = RECOMPUTE and dtype != torch.float32

    def finish_forward(self):
        """num_batches_tracked += S for every train-mode BatchNorm of the pass (the reference's S sequential segment
        calls each add 1) as one multi-tensor launch instead of one tiny kernel per layer"""
        if self.nbt:
            torch._foreach_add_(self.nbt, self.G)
            self.nbt = []

    # ------------------------------------------------------------------ helpers
    def _acc(self, p, g):
        if p is None or not p.requires_grad:
            return
        if p in self.grads:
            self.grads[p] = self.grads[p] + g
        else:
            self.grads[p] = g

    @staticmethod
    def _sync_group(bn):
        return sync_bn_group(bn)

    # ------------------------------------------------------------------ conv + BN + act (+ residual)
    def cba(self, x, conv, bn, act, res=None, res_rec=None):
        """out = act(bn(conv(x)) [+ res] [+ bn_ds(z_ds) from res_rec]).

        res_rec: record returned by `conv_bn_stats` for the downsample branch (resnet.py:164-167).
        """
        if not self.training and not self.save and FUSE_EVAL:
            out = self._cba_fused_eval(x, conv, bn, act, res, res_rec)
            if out is not None:
                return out
        x_src = getattr(x, "_adamml_src", None) if self.save else None  # x is a recomputable activation (see below)
        rec = self.conv_bn_stats(x, conv, bn)
        # residual layers: the backward reduction needs the activation mask of the OUTPUT (it cannot be recomputed from
        # z alone); bn_apply writes it as 1 bit per element so that pass does not stream the 16-bit output again
        bits = None
        z_ = rec["z"]
        if (MASK_BITS and self.save and self.x2 and act != ACT_NONE and (res is not None or res_rec is not None)
                and isinstance(z_, ops.X2) and z_.shape[-1] % 8 == 0):
            bits = torch.empty((z_.hi.numel() // 8,), device=z_.hi.device, dtype=torch.uint8)
        out = ops.bn_apply(rec["z"], rec["ss"], self.G, act, res=res,
                           res_z=res_rec["z"] if res_rec else None, res_ss=res_rec["ss"] if res_rec else None,
                           mask_bits=bits)
        rec["mask_bits"] = bits
        # recompute mode: the output of a layer without residual input is act(z * scale + shift) of tensors the tape
        # keeps anyway, so it is not saved; a consumer that needs it for its weight gradient rebuilds it (bf16)
        lazy = (self.save and self.recompute and res is None and res_rec is None
                and ops.vec_channels(ops.hi_plane(rec["z"])))
        # the forward scale/shift of a layer without residual input is kept (tiny): backward recomputes the
        # ReLU/ReLU6 mask from z with it instead of streaming `out` again
        if res is not None or res_rec is not None or (act == ACT_NONE and not lazy):
            rec["ss"] = None
        if res_rec:
            res_rec["ss"] = None
            res_rec["x"], res_rec["z"] = ops.hi_plane(res_rec["x"]), ops.hi_plane(res_rec["z"])
        if self.save:
            if x_src is not None and rec["x"] is x:   # (not re-laid out into an s2d operand)
                rec.update(x=None, x_lazy=x_src, x_shape=tuple(x.shape))
            rec.update(x=ops.hi_plane(rec["x"]), z=ops.hi_plane(rec["z"]), out=None if lazy else ops.hi_plane(out),
                       act=act, has_res=res is not None, res_rec=res_rec)
            if lazy:
                out._adamml_src = dict(z=rec["z"], ss=rec["ss"], act=act, G=self.G)
            self.tape.append(rec)
        return out

    def cba_maxpool(self, x, conv, bn, act):
        """ResNet stem: maxpool3x3s2(act(bn(conv(x)))) (resnet.py:197-200).  Default-mode training: BN + ReLU are
        applied inside the pooling kernel on the pre-BN planes, so the full-resolution post-activation tensor (which
        only the pool reads; backward takes the mask from z and the pool's recorded positions) is never written."""
        Cout = conv.out_channels
        if not (FUSE_STEM_POOL and self.training and self.save and self.x2 and act != ACT_NONE and Cout % 8 == 0):
            return self.maxpool(self.cba(x, conv, bn, act))
        rec = self.conv_bn_stats(x, conv, bn)
        z = rec["z"]
        y, pos = ops.bn_act_maxpool_fwd(z, rec["ss"], self.G, act)
        rec.update(x=ops.hi_plane(rec["x"]), z=ops.hi_plane(z), out=None, act=act, has_res=False, res_rec=None)
        self.tape.append(rec)
        self.tape.append(dict(x=None, pos=pos, shape=tuple(z.shape)))
        return y

    def _cba_fused_eval(self, x, conv, bn, act, res, res_rec):
        """Inference (running statistics, no tape): conv + BN (+ residual) + ReLU/ReLU6 as ONE kernel — the folded
        scale / shift, the residual add and the activation run in the epilogue of the convolution (tcgen05 for dense
        layers, the stencil kernel for depthwise ones); the pre-BN tensor z never exists (resnet.py:96-111).
        -> None when the layer is outside that envelope (the caller runs conv -> bn_apply)."""
        w = conv.weight
        Cout, _, R, S = w.shape
        stride, pad = conv.stride[0], conv.padding[0]
        if conv.groups > 1 and (res is not None or res_rec is not None):
            return None
        if res_rec is not None:  # BN(downsample) as residual: that branch was not pre-reduced to a plain tensor
            return None
        _, ss = ops.bn_finalize(None, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, 1, 0.1,
                                bn.eps, Cout, 1, False, False)
        ss = ss[0]
        if (not isinstance(x, ops.S2D) and conv.groups == 1 and w.shape[1] < 16 and (R, S) == (3, 3)
                and ops.first_conv_s2d_ok(conv, w.shape[1], x.shape[1], x.shape[2], x.dtype)):
            x = ops.nhwc_to_s2d(x, 3)
        if isinstance(x, ops.S2D):
            return ops.stem_conv_bn_act_fwd(x, w.detach(), ss, act) if res is None else None
        if conv.groups > 1:
            if not (conv.groups == conv.in_channels == Cout and R == 3 and pad == 1):
                return None
            return ops.dwconv_bn_act_fwd(x, ops.pack_weight_dw(w.detach(), cache=self.packs), stride, ss, act)
        if not (self.x2 or self.dtype == torch.bfloat16):
            return None
        wp = ops.pack_weight(w.detach(), ops.PREC_X2 if self.x2 else self.dtype, cache=self.packs)
        return ops.conv_bn_act_fwd(x, wp, stride, pad, ss, act, res=res)

    def conv_bn_stats(self, x, conv, bn):
        """z = conv(x); BN statistics (train) or folded running stats (eval) -> record with z, mi, ss."""
        G = self.G
        w = conv.weight
        Cout, Cin_g, R, S = w.shape
        stride, pad = conv.stride[0], conv.padding[0]
        depthwise = conv.groups > 1
        sums = None
        fused = False
        pg = self._sync_group(bn) if self.training else None
        p2p = P2PStats.get(pg, w.device) if pg is not None else None
        slot_off = None
        if p2p is not None:  # this rank's partial sums are produced straight into the layer's symmetric slot
            slot_off, flat = p2p.slot((_bn_key(bn), "f"), G * Cout * 2)
            p2p_sums = flat.view(G, Cout, 2)
        if (not isinstance(x, ops.S2D) and not depthwise and w.shape[1] < 16 and (R, S) == (3, 3)
                and ops.first_conv_s2d_ok(conv, w.shape[1], x.shape[1], x.shape[2], x.dtype)):
            # 1- / 3-channel stride-2 first conv of a MobileNetV2: tensor cores through the space-to-depth view
            x = ops.nhwc_to_s2d(x, 3)
        stem = isinstance(x, ops.S2D)
        if stem:
            wp = None
            if self.training:
                sums = p2p_sums if p2p is not None else torch.empty((G, Cout, 2), device=x.device, dtype=torch.float64)
            z = ops.stem_conv_fwd(x, w.detach(), stats=sums, imgs_per_group=x.shape[0] // G)
            fused = sums is not None
        elif depthwise:
            assert conv.groups == conv.in_channels == Cout and R == 3 and pad == 1
            wp = ops.pack_weight_dw(w.detach(), cache=self.packs)
            if self.training:
                sums = p2p_sums if p2p is not None else torch.empty((G, Cout, 2), device=x.device, dtype=torch.float64)
                z, fused = ops.dwconv_fwd_stats(x, wp, stride, sums, x.shape[0] // G)
            else:
                z = ops.dwconv_fwd(x, wp, stride)
        else:
            wp = ops.pack_weight(w.detach(), ops.PREC_X2 if self.x2 else self.dtype, cache=self.packs)
            if self.training:
                sums = p2p_sums if p2p is not None else torch.empty((G, Cout, 2), device=x.device, dtype=torch.float64)
            rows = x.shape[0] * ((x.shape[1] + 2 * pad - R) // stride + 1) * ((x.shape[2] + 2 * pad - S) // stride + 1)
            z, fused = ops.conv_fwd(x, wp, stride, pad, stats=sums, rows_per_group=rows // G)
            wp = ops.hi_plane(wp)  # the backward pass multiplies by the bf16 weights
        C = z.shape[-1]
        count = z.numel() // C // G
        if self.training:
            if not fused:
                sums = ops.bn_stats(z, G, out=sums)
            if p2p is not None:
                sums = p2p.allreduce(slot_off, G * C * 2, self.lane).view(G, C, 2)
                count = count * p2p.world
            elif pg is not None:
                count = allreduce_stats(sums, count, pg)
            if bn.momentum is None:
                raise NotImplementedError("BatchNorm momentum=None (cumulative moving average) is not on the AdaMML "
                                          "path (every reference BN uses the default momentum 0.1)")
            mi, ss = ops.bn_finalize(sums, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var,
                                     count, bn.momentum, bn.eps, C, G, True, bn.track_running_stats)
            if bn.num_batches_tracked is not None:
                self.nbt.append(bn.num_batches_tracked)  # advanced by G in ONE multi-tensor launch (finish_forward)
        else:
            mi, ss = ops.bn_finalize(None, bn.weight.detach(), bn.bias.detach(), bn.running_mean, bn.running_var, 1,
                                     0.1, bn.eps, C, G, False, False)
        return dict(x=x, z=z, mi=mi, ss=ss, w=wp, conv=conv, bn=bn, count=count, pg=pg, depthwise=depthwise, stem=stem)

    def cba_bwd(self, dout, need_dx=True, addend=None, addend_sub=1):
        """Pops one cba record.  Returns (dx or None, dres or None)."""
        rec = self.tape.pop()
        shared = (rec["has_res"] or rec["res_rec"] is not None) and rec["act"] != ACT_NONE
        dx, dres, _ = self._bn_conv_bwd(rec, dout, rec["out"], rec["act"], need_dx, addend,
                                        want_dres=rec["has_res"] and rec["act"] != ACT_NONE, addend_sub=addend_sub,
                                        mask_inplace=shared)
        return dx, dres

    def _bn_conv_bwd(self, rec, dout, out, act, need_dx, addend, want_dres, addend_sub=1, compact_ok=False,
                     mask_inplace=False):
        """-> (dx, dres, dx_is_compact)"""
        G = self.G
        conv, bn, z, mi = rec["conv"], rec["bn"], rec["z"], rec["mi"]
        C = z.shape[-1]
        mask_ss = rec.get("ss") if out is rec.get("out") else None
        # residual layers: the reduce pass masks dout IN PLACE (gm = dout * relu'(out)); gm is the gradient of the
        # residual input and lets the apply pass (and a downsample branch sharing dout) skip `out` entirely
        inplace = mask_inplace and act != ACT_NONE and ops.vec_channels(z)
        had_dres = want_dres
        p2p = P2PStats.get(rec["pg"], z.device) if rec["pg"] is not None else None
        sums_out = slot_off = None
        if p2p is not None:
            slot_off, flat = p2p.slot((_bn_key(bn), "b"), G * C * 2)
            sums_out = flat.view(G, C, 2)
        pre = rec.pop("pre_reduced", None)
        if pre is not None:
            # the depthwise backward kernel that produced dout already masked it and reduced (sum gm, sum gm * out)
            # (ops.dwconv_bwd pre=...): no pass over (dout, z) here
            sums = ops.bn_sums_from_out(pre, rec["ss"], mi, out=sums_out)
            out, act, want_dres, mask_ss = None, ACT_NONE, False, None
        else:
            bits = rec.get("mask_bits") if (out is not None and out is rec.get("out")) else None
            sums = ops.bn_bwd_reduce(dout, out, z, mi, G, act, mask_ss=mask_ss, gm_inplace=inplace, sums_out=sums_out,
                                     mask_bits=bits)
        if inplace:
            rec["dout_masked"] = True
            out, act, want_dres = None, ACT_NONE, False
        if bn.weight.requires_grad or bn.bias.requires_grad:
            dgamma, dbeta = ops.bn_param_grad(sums, C, G)
            self._acc(bn.weight, dgamma)
            self._acc(bn.bias, dbeta)
        if p2p is not None:
            sums = p2p.allreduce(slot_off, G * C * 2, self.lane).view(G, C, 2)
        elif rec["pg"] is not None:
            allreduce_stats(sums, 0, rec["pg"])
        need_w = conv.weight.requires_grad
        dz, dres = ops.bn_bwd_apply(dout, out, z, mi, bn.weight.detach(), sums, G, rec["count"], act, self.training,
                                    want_dz=(need_w or need_dx), want_dres=want_dres, mask_ss=mask_ss)
        if inplace and had_dres:
            dres = dout
        dx = None
        x = rec["x"]
        if x is None and rec.get("x_lazy") is not None:
            if need_w:  # rebuild the producer's output from its saved pre-BN tensor (recompute mode)
                src = rec["x_lazy"]
                x = ops.bn_apply(src["z"], src["ss"], src["G"], src["act"])
            else:
                x = _ShapeOnly(rec["x_shape"])
        stride, pad = conv.stride[0], conv.padding[0]
        if rec["stem"]:
            assert not need_dx, "the s2d stem has no data gradient (its input is data)"
            if need_w:
                self._acc(conv.weight, ops.stem_wgrad(x, dz, conv.out_channels))
        elif rec["depthwise"]:
            if need_w and need_dx and addend is None and ops.dwconv_bwd_ok(x, dz, stride):
                # both gradients from one pass over dz and x (TMA-staged tiles, csrc/dwconv_tma.cu); when x is the
                # output of a plain conv + BN + act layer, that layer's BN-backward reduction rides along
                prev = self._dw_producer(rec, x)
                pre = None
                if prev is not None:
                    raw = torch.empty((G, C, 2), device=x.device, dtype=torch.float64)
                    pre = (raw, x.shape[0] // G, prev["act"])
                dx, dw = ops.dwconv_bwd(x, dz, rec["w"], stride, pre=pre)
                if prev is not None:
                    prev["pre_reduced"] = raw
                self._acc(conv.weight, dw)
            else:
                if need_w:
                    self._acc(conv.weight, ops.dwconv_wgrad(x, dz, stride))
                if need_dx:
                    dx = ops.dwconv_dgrad(dz, rec["w"], tuple(x.shape), stride, addend=addend)
        else:
            w = rec["w"]
            if need_w:
                dw = ops.conv_wgrad(x, dz, tuple(w.shape), stride, pad)
                self._acc(conv.weight, ops.unpack_wgrad(dw, conv.weight.shape[1]))
            if need_dx:
                w_rot = None
                R, S = w.shape[1], w.shape[2]
                if ops.tc_dgrad_ok(self.dtype, w.shape[0], w.shape[3], R, S, stride):
                    w_rot = ops.pack_weight_dgrad(conv.weight.detach(), self.dtype, cache=self.packs)
                if w_rot is not None and stride == 2 and R == 1 and S == 1 and addend is None and compact_ok:
                    # stride-2 1x1 (Bottleneck downsample): only the even pixels receive gradient -> compact GEMM,
                    # scattered by the consumer's epilogue (conv_dgrad addend_sub=2)
                    return ops.conv_dgrad_compact(dz, w_rot), dres, True
                dx = ops.conv_dgrad(dz, w, tuple(x.shape), stride, pad, addend=addend, w_rot=w_rot,
                                    addend_sub=addend_sub)
        return dx, dres, False

    def _dw_producer(self, rec, x):
        """The tape record of the layer whose output is the input x of the depthwise layer `rec`, when its
        BatchNorm-backward reduction can be fused into the depthwise backward kernel: a conv + BN + ReLU/ReLU6 layer
        without residual input whose forward scale / shift was kept (cba), directly below `rec` on the tape."""
        if not DW_FUSE_PRE or not self.tape:
            return None
        prev = self.tape[-1]
        if prev.get("bn") is None or prev.get("has_res") or prev.get("res_rec") is not None:
            return None
        if prev.get("act", ACT_NONE) == ACT_NONE or prev.get("ss") is None or prev.get("pre_reduced") is not None:
            return None
        if prev.get("out") is not None:
            same = prev["out"] is rec["x"]
        else:  # recompute mode: x was rebuilt from the producer's saved pre-BN tensor
            lz = rec.get("x_lazy")
            same = lz is not None and lz["z"] is prev["z"]
        if not same or tuple(prev["z"].shape) != tuple(x.shape) or x.shape[0] % self.G:
            return None
        return prev

    # ------------------------------------------------------------------ ResNet blocks
    def bottleneck(self, x, blk):
        """resnet.py:93-113."""
        a = self.cba(x, blk.conv1, blk.bn1, ACT_RELU)
        a = self.cba(a, blk.conv2, blk.bn2, ACT_RELU)
        if blk.downsample is not None:
            if not self.training and not self.save and FUSE_EVAL:
                # inference: the downsample branch is a fused conv + BN kernel of its own, its output the residual
                idt = self._cba_fused_eval(x, blk.downsample[0], blk.downsample[1], ACT_NONE, None, None)
                if idt is not None:
                    return self.cba(a, blk.conv3, blk.bn3, ACT_RELU, res=idt)
            ds = self.conv_bn_stats(x, blk.downsample[0], blk.downsample[1])
            return self.cba(a, blk.conv3, blk.bn3, ACT_RELU, res_rec=ds)
        return self.cba(a, blk.conv3, blk.bn3, ACT_RELU, res=x)

    def bottleneck_bwd(self, dout, need_dx=True):
        rec3 = self.tape[-1]
        ds = rec3["res_rec"]
        d_id = None
        dx_ds = None
        compact = False
        c1 = self.tape[-3]["conv"]
        out3, act3 = rec3["out"], rec3["act"]
        da, d_id = self.cba_bwd(dout)          # conv3/bn3 (+identity grad); may mask dout in place
        if ds is not None:
            # downsample branch shares (dout, out, act mask) with bn3; conv1 of the block is a stride-1 1x1 whose
            # tcgen05 dgrad epilogue can scatter a compact stride-2 downsample gradient
            ok = ops.tc_dgrad_ok(self.dtype, c1.out_channels, c1.in_channels, 1, 1, 1) and c1.kernel_size == (1, 1)
            if rec3.get("dout_masked"):
                out3, act3 = None, ACT_NONE
            dx_ds, _, compact = self._bn_conv_bwd(ds, dout, out3, act3, need_dx, None, want_dres=False,
                                                  compact_ok=ok)
        da, _ = self.cba_bwd(da)               # conv2/bn2
        addend = dx_ds if ds is not None else d_id
        dx, _ = self.cba_bwd(da, need_dx=need_dx, addend=addend if need_dx else None,
                             addend_sub=2 if compact else 1)  # conv1/bn1
        return dx

    def basicblock(self, x, blk):
        """resnet.py:59-74."""
        a = self.cba(x, blk.conv1, blk.bn1, ACT_RELU)
        if blk.downsample is not None:
            if not self.training and not self.save and FUSE_EVAL:
                idt = self._cba_fused_eval(x, blk.downsample[0], blk.downsample[1], ACT_NONE, None, None)
                if idt is not None:
                    return self.cba(a, blk.conv2, blk.bn2, ACT_RELU, res=idt)
            ds = self.conv_bn_stats(x, blk.downsample[0], blk.downsample[1])
            return self.cba(a, blk.conv2, blk.bn2, ACT_RELU, res_rec=ds)
        return self.cba(a, blk.conv2, blk.bn2, ACT_RELU, res=x)

    def basicblock_bwd(self, dout, need_dx=True):
        rec2 = self.tape[-1]
        ds = rec2["res_rec"]
        dx_ds = None
        out2, act2 = rec2["out"], rec2["act"]
        da, d_id = self.cba_bwd(dout)          # may mask dout in place
        if ds is not None:
            if rec2.get("dout_masked"):
                out2, act2 = None, ACT_NONE
            dx_ds, _, _ = self._bn_conv_bwd(ds, dout, out2, act2, need_dx, None, want_dres=False)
        addend = dx_ds if ds is not None else d_id
        dx, _ = self.cba_bwd(da, need_dx=need_dx, addend=addend if need_dx else None)
        return dx

    # ------------------------------------------------------------------ MobileNetV2 block
    def inverted_residual(self, x, layers, use_res):
        """layers: list of (conv, bn, act) triples: [pw] + dw + pw-linear
        (sound_mobilenet_v2.py:52-69, policy_net.py:63-95)."""
        a = x
        for i, (conv, bn, act) in enumerate(layers):
            last = i == len(layers) - 1
            a = self.cba(a, conv, bn, act, res=x if (last and use_res) else None)
        if self.save:
            self.tape.append(dict(kind="ir", n=len(layers), use_res=use_res))
        return a

    def inverted_residual_bwd(self, dout, need_dx=True):
        meta = self.tape.pop()
        n, use_res = meta["n"], meta["use_res"]
        d = dout
        for i in range(n):
            first = i == n - 1
            addend = dout if (first and use_res) else None  # linear bottleneck: d(res) = dout itself
            d, _ = self.cba_bwd(d, need_dx=(need_dx or not first), addend=addend if (need_dx or not first) else None)
        return d

    # ------------------------------------------------------------------ pools
    def maxpool(self, x):
        if not self.save:
            return ops.maxpool_fwd(x)
        y, pos = ops.maxpool_fwd(x, want_pos=True)
        self.tape.append(dict(x=ops.hi_plane(x) if pos is None else None, pos=pos, shape=tuple(x.shape)))
        return y

    def maxpool_bwd(self, dy):
        rec = self.tape.pop()
        return ops.maxpool_bwd(rec["x"], dy, pos=rec["pos"], x_shape=rec["shape"])

    def tpool(self, x, frames, mode_avg=False):
        y = ops.tpool_fwd(x, frames, mode_avg)
        if self.save:
            self.tape.append(dict(x=ops.hi_plane(x), frames=frames, avg=mode_avg))
        return y

    def tpool_bwd(self, dy):
        rec = self.tape.pop()
        return ops.tpool_bwd(rec["x"], dy, rec["frames"], rec["avg"])

    # ------------------------------------------------------------------ heads
    def avgpool(self, x):
        y = ops.avgpool_fwd(x)
        if self.save:
            self.tape.append(dict(shape=tuple(x.shape)))
        return y

    def avgpool_bwd(self, dy):
        rec = self.tape.pop()
        return ops.avgpool_bwd(dy, rec["shape"], self.dtype)

    def classifier(self, feat, fc, drop_mask, frames):
        """Dropout (mask supplied by the caller, torch-RNG order) + Linear + mean over the remaining
        frames (resnet.py:214-221, sound_mobilenet_v2.py:157)."""
        xin = ops.mul(feat, drop_mask) if drop_mask is not None else feat
        y = ops.linear_fwd(xin, fc.weight.detach())
        ops.bias_act_(y, fc.bias.detach() if fc.bias is not None else None, ACT_NONE)
        if frames > 1:
            y = ops.frame_mean(y, frames)
        if self.save:
            self.tape.append(dict(xin=xin, fc=fc, mask=drop_mask, frames=frames))
        return y

    def classifier_bwd(self, dy):
        rec = self.tape.pop()
        fc = rec["fc"]
        if rec["frames"] > 1:
            dy = ops.frame_mean_bwd(dy, rec["frames"])
        if fc.weight.requires_grad:
            self._acc(fc.weight, ops.linear_wgrad(rec["xin"], dy))
        if fc.bias is not None and fc.bias.requires_grad:
            self._acc(fc.bias, ops.colsum(dy))
        dx = ops.linear_dgrad(dy, fc.weight.detach())
        if rec["mask"] is not None:
            dx = ops.mul(dx, rec["mask"])
        return dx


class BackboneFunction(torch.autograd.Function):
    """autograd bridge: (x_nhwc, *params) -> head output; backward replays the tape in reverse."""

    @staticmethod
    def forward(ctx, net, x, groups, extra, *params):
        need = bool(extra.get("_save")) and any(ctx.needs_input_grad[4:])
        ex = Exec(net.compute_dtype, net.training, groups, save=need, lane=int(extra.get("_lane", 0)),
                  packs=extra.get("_packs"))
        live = extra.get("_live")  # (count tensor, clip capacity): inference with device-side skipping
        if live is not None:
            assert not need and not net.training, "the device-side work limit is an inference-only feature"
            ops.set_live_clips(live[0], live[1])
        try:
            y = net.run_forward(ex, x, extra)
            ex.finish_forward()
        finally:
            if live is not None:
                ops.set_live_clips(None, 0)
        ctx.ex = ex if need else None
        ctx.net = net
        ctx.params = params
        return y

    @staticmethod
    def backward(ctx, dy):
        ex = ctx.ex
        if ex is None:
            if getattr(ctx, "consumed", False):
                raise RuntimeError("Trying to backward through the backbone a second time: its tape (saved "
                                   "activations) was freed by the first backward; run the forward pass again")
            return (None,) * (4 + len(ctx.params))
        dy = dy.contiguous()
        ctx.net.run_backward(ex, dy)
        grads = tuple(ex.grads.get(p) if need else None for p, need in zip(ctx.params, ctx.needs_input_grad[4:]))
        ctx.ex = None
        ctx.consumed = True
        return (None, None, None, None) + grads


def run_backbone(net, x, groups, extra=None):
    params = tuple(net.parameters())
    extra = dict(extra or {})
    # grad mode is invisible inside Function.forward, so capture it here
    extra["_save"] = torch.is_grad_enabled()
    return BackboneFunction.apply(net, x, groups, extra, *params)


# ---------------------------------------------------------------------------------------------- multi-stream
_SIDE_STREAMS = {}


def _side_stream(device, i):
    key = (device.index, i)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def run_backbones_parallel(jobs):
    """jobs: list of (net, x, groups, extra).  The backbones of one AdaMML step (policy MobileNetV2s, main ResNets /
    sound MobileNetV2) are independent until the policy head / late fusion, so each runs on its own CUDA stream:
    inside the captured graph they become parallel branches and the latency-bound kernels of one net (narrow
    MobileNetV2 GEMMs, late ResNet layers) fill the gaps of another.  autograd replays every backward on its
    forward stream, so the backward passes overlap the same way."""
    import os
    if len(jobs) <= 1 or os.environ.get("ADAMML_B200_STREAMS", "1") == "0":
        return [run_backbone(net, x, g, extra) for net, x, g, extra in jobs]
    cur = torch.cuda.current_stream()
    outs = []
    for i, (net, x, g, extra) in enumerate(jobs):
        s = _side_stream(cur.device, i)
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            x.record_stream(s)           # allocated on the caller's stream, consumed (and kept on the tape) here
            for v in (extra or {}).values():
                if isinstance(v, torch.Tensor):
                    v.record_stream(s)
            y = run_backbone(net, x, g, dict(extra or {}, _lane=i))
        outs.append((y, s))
    for y, s in outs:
        cur.wait_stream(s)
        y.record_stream(cur)
    return [y for y, _ in outs]
