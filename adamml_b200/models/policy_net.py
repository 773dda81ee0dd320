"""Policy network: per-modality MobileNetV2 -> joint MLP -> LSTMCell -> per-modality FC -> hard
Gumbel-softmax (drop-in for reference models/policy_net.py).

Backbones run batched over all segments on the conv engine; the joint MLP / LSTM / FC /
Gumbel head runs in fp32 through the exact GEMM engine and the fused per-step warp kernel
(csrc/policy.cu).  The Exp(1) noise of F.gumbel_softmax is drawn by torch (same order as the
reference: one [M*N, 2] draw per segment) or injected by the caller for parity tests.
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import call
from ..engine import run_backbone
from ..ops import ACT_NONE, ACT_RELU, ACT_RELU6
from .resnet import default_compute_dtype
from .sound_mobilenet_v2 import MBV2_SETTING


def _conv_bn_relu6(cin, cout, k, stride, groups=1):
    return [nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(cout),
            nn.ReLU6(inplace=True)]


class _InvertedResidual(nn.Module):
    """policy_net.py:54-95: `conv` indices 0..4 (t == 1) or 0..7, optional temporal max-pool in front."""

    def __init__(self, inp, oup, stride, t, num_frames=None):
        super().__init__()
        self.pool_frames = num_frames if num_frames else None
        hid = round(inp * t)
        self.identity = stride == 1 and inp == oup
        mods = []
        if t != 1:
            mods += _conv_bn_relu6(inp, hid, 1, 1)
        mods += _conv_bn_relu6(hid, hid, 3, stride, groups=hid)
        mods += [nn.Conv2d(hid, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*mods)

    def layers(self):
        m = list(self.conv)
        out = []
        i = 0
        while i < len(m):
            act = ACT_RELU6 if (i + 2 < len(m) and isinstance(m[i + 2], nn.ReLU6)) else ACT_NONE
            out.append((m[i], m[i + 1], act))
            i += 3 if act == ACT_RELU6 else 2
        return out


class MobileNetV2(nn.Module):
    def __init__(self, num_classes=1000, num_frames=4, input_channels=3, compute_dtype=None):
        super().__init__()
        self.input_channels = input_channels
        self.num_frames = self.orig_num_frames = num_frames
        self.compute_dtype = compute_dtype or default_compute_dtype()
        layers = [nn.Sequential(*_conv_bn_relu6(input_channels, 32, 3, 2))]
        cin = 32
        frames = num_frames
        for t, c, n, s in MBV2_SETTING:
            has_tp = c in (64, 160)  # policy_net.py:121
            for i in range(n):
                nf = frames if (i == 0 and has_tp and frames != 1) else None  # 0 frames -> no pool (falsy)
                layers.append(_InvertedResidual(cin, c, s if i == 0 else 1, t, num_frames=nf))
                cin = c
            if has_tp:
                frames //= 2
        self.num_frames = frames
        self.features = nn.Sequential(*layers)
        self.last_channel = 1280
        self.conv = nn.Sequential(*_conv_bn_relu6(cin, 1280, 1, 1))
        self.classifier = nn.Linear(1280, num_classes)
        for m in self.modules():  # policy_net.py:169-181
            if isinstance(m, nn.Conv2d):
                n_ = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, (2.0 / n_) ** 0.5)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
            elif isinstance(m, nn.Linear):
                m.weight.data.normal_(0, 0.01)
                m.bias.data.zero_()

    @property
    def network_name(self):
        return "mobilenet_v2"

    def run_forward(self, ex, x, extra):
        """x NHWC [G*N*T, h, w, C] -> features [G*N*T', 1280] (policy_net.py:142-149)."""
        f = list(self.features)
        a = ex.cba(x, f[0][0], f[0][1], ACT_RELU6)
        for blk in f[1:]:
            if blk.pool_frames:
                a = ex.tpool(a, blk.pool_frames, False)
            a = ex.inverted_residual(a, blk.layers(), blk.identity)
        a = ex.cba(a, self.conv[0], self.conv[1], ACT_RELU6)
        return ex.avgpool(a)

    def run_backward(self, ex, dy):
        d = ex.avgpool_bwd(dy)
        d, _ = ex.cba_bwd(d)
        for blk in reversed(list(self.features)[1:]):
            d = ex.inverted_residual_bwd(d)
            if blk.pool_frames:
                d = ex.tpool_bwd(d)
        ex.cba_bwd(d, need_dx=False)


class JointMobileNetV2(nn.Module):
    """policy_net.py:206-258 (fc/dropout are deleted by PolicyNet, so they are not created)."""

    def __init__(self, num_frames, modality, num_classes=1000, dropout=0.5, input_channels=None, compute_dtype=None):
        super().__init__()
        self.num_frames = num_frames
        self.modality = modality
        self.nets = nn.ModuleList()
        for i, m in enumerate(modality):
            net = MobileNetV2(num_classes, num_frames=1 if m == "sound" else num_frames,
                              input_channels=input_channels[i], compute_dtype=compute_dtype)
            del net.classifier
            self.nets.append(net)
        self.last_channels = 2048
        self.joint = nn.Sequential(nn.Linear(1280 * len(modality), 2048), nn.ReLU(True), nn.Linear(2048, 2048),
                                   nn.ReLU(True))


class _PolicyHead(torch.autograd.Function):
    """joint MLP + S-step LSTM + FCs + hard Gumbel (policy_net.py:243-245, 345-365) with manual BPTT."""

    @staticmethod
    def forward(ctx, net, expo, tau, save, nfeat, *args):
        feats, params = args[:nfeat], args[nfeat:]
        j0w, j0b, j2w, j2b, w_ih, w_hh, b_ih, b_hh = params[:8]
        fcp = params[8:]
        M = len(fcp) // 2
        S = expo.shape[0]
        SN = feats[0].shape[0]
        N = SN // S
        Hd = w_hh.shape[1]
        Fd = j2w.shape[0]
        dev = feats[0].device
        f = torch.cat(feats, 1) if nfeat > 1 else feats[0]
        h1 = ops.linear_fwd(f, j0w)
        ops.bias_act_(h1, j0b, ACT_RELU)
        xin = torch.zeros((SN, Fd + 2 * M), device=dev, dtype=torch.float32)
        ops.linear_fwd(h1, j2w, out=xin)
        ops.bias_act_(xin, j2b, ACT_RELU, rows=SN, cols=Fd)
        gx = ops.linear_fwd(xin, w_ih, K=Fd)
        fc_w = torch.stack([fcp[2 * m] for m in range(M)]).contiguous()      # [M,2,Hd]
        fc_b = torch.stack([fcp[2 * m + 1] for m in range(M)]).contiguous()  # [M,2]
        gates = torch.empty((S, N, 4 * Hd), device=dev)
        hs = torch.zeros((S + 1, N, Hd), device=dev)
        cs = torch.zeros((S + 1, N, Hd), device=dev)
        logits = torch.empty((S, M, N, 2), device=dev)
        ysoft = torch.empty((S, M, N, 2), device=dev)
        dec = torch.empty((S, M, N), device=dev)
        gx3 = gx.view(S, N, 4 * Hd)
        xin3 = xin.view(S, N, Fd + 2 * M)
        for s in range(S):
            call("policy_step_fwd", gx3[s], logits[s - 1] if s > 0 else None, hs[s] if s > 0 else None,
                 cs[s] if s > 0 else None, w_ih, w_ih.stride(0), Fd, w_hh, b_ih, b_hh, fc_w, fc_b, expo[s],
                 float(tau), gates[s], hs[s + 1], cs[s + 1], logits[s], ysoft[s], dec[s], xin3[s][:, Fd:],
                 xin3.stride(1), N, M, Hd)
        if save:
            ctx.saved = (f, h1, xin, gates, hs, cs, ysoft, fc_w, params)
            ctx.dims = (S, N, M, Hd, Fd, nfeat, float(tau))
        else:
            ctx.saved = None
        ctx.mark_non_differentiable(logits)
        return dec, logits

    @staticmethod
    def backward(ctx, d_dec, _d_logits):
        n_in = 5
        if ctx.saved is None:
            raise RuntimeError("policy head was run without a tape (no_grad)")
        f, h1, xin, gates, hs, cs, ysoft, fc_w, params = ctx.saved
        S, N, M, Hd, Fd, nfeat, tau = ctx.dims
        j0w, j0b, j2w, j2b, w_ih, w_hh, b_ih, b_hh = params[:8]
        dev = f.device
        d_dec = d_dec.contiguous()
        dgates = torch.empty((S, N, 4 * Hd), device=dev)
        dl = torch.empty((M, S, N, 2), device=dev)
        dh = torch.empty((2, N, Hd), device=dev)
        dc = torch.empty((2, N, Hd), device=dev)
        dfb = torch.empty((2, M, N, 2), device=dev)
        for s in reversed(range(S)):
            last, first = s == S - 1, s == 0
            cur, nxt = s % 2, (s + 1) % 2
            call("policy_step_bwd", d_dec[s], None if last else dfb[nxt], None if last else dh[nxt],
                 None if last else dc[nxt], gates[s], cs[s + 1], None if first else cs[s], ysoft[s], w_ih,
                 w_ih.stride(0), Fd, w_hh, fc_w, tau, dl[:, s], dl.stride(0), dgates[s],
                 None if first else dh[cur], None if first else dc[cur], None if first else dfb[cur], N, M, Hd)
        SN = S * N
        dg2 = dgates.view(SN, 4 * Hd)
        need = ctx.needs_input_grad[n_in + nfeat:]
        grads = [None] * len(params)
        if need[4]:
            grads[4] = ops.linear_wgrad(xin, dg2)
        if need[5]:
            grads[5] = ops.linear_wgrad(hs[:S].reshape(SN, Hd), dg2)
        if need[6] or need[7]:
            db = ops.colsum(dg2)
            grads[6], grads[7] = db, db.clone()
        h_all = hs[1:].reshape(SN, Hd)
        for m in range(M):
            dlm = dl[m].reshape(SN, 2)
            if need[8 + 2 * m]:
                grads[8 + 2 * m] = ops.linear_wgrad(h_all, dlm)
            if need[9 + 2 * m]:
                grads[9 + 2 * m] = ops.colsum(dlm)
        d_x = ops.linear_dgrad(dg2, w_ih, K=Fd)                       # [SN, Fd]
        dz2 = ops.act_bwd(d_x, xin, ACT_RELU, rows=SN, cols=Fd)        # mask from relu output in xin[:, :Fd]
        if need[2]:
            grads[2] = ops.linear_wgrad(h1, dz2)
        if need[3]:
            grads[3] = ops.colsum(dz2)
        dh1 = ops.linear_dgrad(dz2, j2w)
        dz1 = ops.act_bwd(dh1, h1, ACT_RELU)
        if need[0]:
            grads[0] = ops.linear_wgrad(f, dz1)
        if need[1]:
            grads[1] = ops.colsum(dz1)
        dfeats = [None] * nfeat
        if any(ctx.needs_input_grad[n_in:n_in + nfeat]):
            df = ops.linear_dgrad(dz1, j0w)
            w = df.shape[1] // nfeat
            dfeats = [df[:, i * w:(i + 1) * w].contiguous() for i in range(nfeat)]
        ctx.saved = None
        return (None, None, None, None, None) + tuple(dfeats) + tuple(grads)


class _PolicyHeadNoCausal(torch.autograd.Function):
    """causality_modeling=None (policy_net.py:330-339): joint MLP -> per-modality Linear(2048, 2) on every
    (segment, video) feature -> ONE hard Gumbel-softmax over all (modality, segment, video) rows."""

    @staticmethod
    def forward(ctx, expo, tau, save, nfeat, S, *args):
        feats, params = args[:nfeat], args[nfeat:]
        j0w, j0b, j2w, j2b = params[:4]
        fcp = params[4:]
        M = len(fcp) // 2
        SN = feats[0].shape[0]
        dev = feats[0].device
        f = torch.cat(feats, 1) if nfeat > 1 else feats[0]
        h1 = ops.linear_fwd(f, j0w)
        ops.bias_act_(h1, j0b, ACT_RELU)
        o = ops.linear_fwd(h1, j2w)
        ops.bias_act_(o, j2b, ACT_RELU)
        logits = torch.empty((M, SN, 2), device=dev)
        for m in range(M):
            ops.linear_fwd(o, fcp[2 * m], out=logits[m])
            ops.bias_act_(logits[m], fcp[2 * m + 1], ACT_NONE)
        ysoft = torch.empty_like(logits)
        dec = torch.empty((M, SN), device=dev)
        call("gumbel_hard_fwd", logits, expo, float(tau), ysoft, dec, M * SN)
        ctx.saved = (f, h1, o, ysoft, params) if save else None
        ctx.dims = (M, SN, nfeat, float(tau))
        N = SN // S
        ctx.mark_non_differentiable(logits)
        # (MSN) -> [M,S,N] -> [S,M,N]   /   logits [M,S,N,2] -> [S,M,N,2]
        return dec.view(M, S, N).transpose(0, 1).contiguous(), logits.view(M, S, N, 2).transpose(0, 1).contiguous()

    @staticmethod
    def backward(ctx, d_dec, _d_logits):
        n_in = 5
        if ctx.saved is None:
            raise RuntimeError("policy head was run without a tape (no_grad)")
        f, h1, o, ysoft, params = ctx.saved
        M, SN, nfeat, tau = ctx.dims
        j0w, j0b, j2w, j2b = params[:4]
        fcp = params[4:]
        dd = d_dec.transpose(0, 1).contiguous().view(M, SN)
        dl = torch.empty((M, SN, 2), device=f.device)
        call("gumbel_hard_bwd", dd, ysoft, tau, dl, M * SN)
        need = ctx.needs_input_grad[n_in + nfeat:]
        grads = [None] * len(params)
        d_o = None
        for m in range(M):
            if need[4 + 2 * m]:
                grads[4 + 2 * m] = ops.linear_wgrad(o, dl[m])
            if need[5 + 2 * m]:
                grads[5 + 2 * m] = ops.colsum(dl[m])
            dm = ops.linear_dgrad(dl[m], fcp[2 * m])
            d_o = dm if d_o is None else d_o + dm
        dz2 = ops.act_bwd(d_o, o, ACT_RELU)
        if need[2]:
            grads[2] = ops.linear_wgrad(h1, dz2)
        if need[3]:
            grads[3] = ops.colsum(dz2)
        dh1 = ops.linear_dgrad(dz2, j2w)
        dz1 = ops.act_bwd(dh1, h1, ACT_RELU)
        if need[0]:
            grads[0] = ops.linear_wgrad(f, dz1)
        if need[1]:
            grads[1] = ops.colsum(dz1)
        dfeats = [None] * nfeat
        if any(ctx.needs_input_grad[n_in:n_in + nfeat]):
            df = ops.linear_dgrad(dz1, j0w)
            w = df.shape[1] // nfeat
            dfeats = [df[:, i * w:(i + 1) * w].contiguous() for i in range(nfeat)]
        ctx.saved = None
        return (None,) * n_in + tuple(dfeats) + tuple(grads)


class PolicyNet(nn.Module):
    def __init__(self, joint_net, modality, causality_modeling="lstm"):
        super().__init__()
        self.joint_net = joint_net
        self.modality = modality
        self.causality_modeling = causality_modeling
        self.num_modality = len(modality)
        self.temperature = 5.0
        feature_dim = joint_net.last_channels
        if causality_modeling == "lstm":
            self.lstm = nn.LSTMCell(feature_dim + 2 * self.num_modality, 256)
            self.fcs = nn.ModuleList([nn.Linear(256, 2) for _ in range(self.num_modality)])
        elif causality_modeling is None:  # policy_net.py:281
            self.fcs = nn.ModuleList([nn.Linear(feature_dim, 2) for _ in range(self.num_modality)])
        else:
            raise ValueError("unknown mode")

    def set_temperature(self, temperature):
        self.temperature = temperature

    def decay_temperature(self, decay_ratio=None):
        if decay_ratio:
            self.temperature *= decay_ratio
        print("Current temperature: {}".format(self.temperature), flush=True)

    @property
    def network_name(self):
        return "j_mobilenet_v2{}".format("-" + self.causality_modeling if self.causality_modeling else "")

    def draw_gumbel_noise(self, S, N, device):
        """Exp(1) samples in the reference's draw order (F.gumbel_softmax, policy_net.py:288): one [M*N, 2] draw per
        segment for the LSTM policy, a single [M*S*N, 2] draw for causality_modeling=None."""
        if self.causality_modeling is None:
            return torch.empty((1, self.num_modality * S * N, 2), device=device).exponential_()
        return torch.stack([torch.empty((self.num_modality * N, 2), device=device).exponential_() for _ in range(S)])

    def backbone_jobs(self, p_x, S):
        """(net, x, groups, extra) per policy modality, for engine.run_backbones_parallel"""
        return [(net, x, S, None) for net, x in zip(self.joint_net.nets, p_x)]

    def forward(self, p_x, S, N, expo=None, feats=None):
        """p_x: list over policy modalities of NHWC image batches (segment-major); feats: backbone features when the
        caller already ran them (AdaMML runs all backbones of the step on parallel streams).
        -> decisions [S, M, N] (float 0/1, straight-through grad), logits [S, M, N, 2]."""
        if feats is None:
            feats = [run_backbone(net, x, S) for net, x in zip(self.joint_net.nets, p_x)]
        for f in feats:
            if f.shape[0] != S * N:
                raise ValueError("policy backbone must reduce every clip to one frame (groups in {2,4,8})")
        if expo is None:
            expo = self.draw_gumbel_noise(S, N, feats[0].device)
        j = self.joint_net.joint
        if self.causality_modeling is None:
            expo = expo.reshape(self.num_modality * S * N, 2).contiguous().float()
            params = [j[0].weight, j[0].bias, j[2].weight, j[2].bias]
            for fc in self.fcs:
                params += [fc.weight, fc.bias]
            return _PolicyHeadNoCausal.apply(expo, self.temperature, torch.is_grad_enabled(), len(feats), S, *feats,
                                             *params)
        expo = expo.reshape(S, self.num_modality * N, 2).contiguous().float()
        params = [j[0].weight, j[0].bias, j[2].weight, j[2].bias, self.lstm.weight_ih, self.lstm.weight_hh,
                  self.lstm.bias_ih, self.lstm.bias_hh]
        for fc in self.fcs:
            params += [fc.weight, fc.bias]
        dec, logits = _PolicyHead.apply(self, expo, self.temperature, torch.is_grad_enabled(), len(feats), *feats,
                                        *params)
        return dec, logits


def p_joint_mobilenet(num_frames, modality, input_channels, causality_modeling, compute_dtype=None):
    joint_net = JointMobileNetV2(num_frames=num_frames, modality=modality, input_channels=input_channels,
                                 compute_dtype=compute_dtype)
    return PolicyNet(joint_net, modality, causality_modeling=causality_modeling)
