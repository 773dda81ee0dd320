"""AdaMML wrapper (drop-in for reference models/adamml.py): data layer -> policy -> gated main
nets -> late fusion, with the segment x modality loop packed into batched launches."""
import os

import torch
import torch.nn as nn

from .. import ops
from ..engine import run_backbones_parallel
from .joint_resnet_mobilenetv2 import joint_resnet_mobilenetv2
from .policy_net import p_joint_mobilenet
from .resnet import default_compute_dtype

PACK_CACHE = os.environ.get("ADAMML_B200_PACK_CACHE", "1") != "0"


class AdaMML(nn.Module):
    def __init__(self, policy_net, main_net, num_frames, num_segments, modality, rng_policy, rng_threshold,
                 num_classes, compute_dtype=None):
        super().__init__()
        self.rng_policy = rng_policy
        self.policy_net = policy_net
        self.main_net = main_net
        self.num_segments = num_segments
        self.num_frames = num_frames * num_segments
        self.num_frames_per_segment = num_frames
        self.modality = modality
        self.num_modality = len(modality) - 1 if ("rgbdiff" in modality and "flow" in modality) else len(modality)
        self.p_data_idx = [modality.index(x) for x in policy_net.modality]
        self.m_data_idx = [modality.index(x) for x in main_net.modality]
        self.rng_threshold = rng_threshold
        self.decay_ratio = 0.965
        self.update_policy_net = True
        self.update_main_net = True
        self.compute_dtype = compute_dtype or default_compute_dtype()
        # inference: run the main backbones only on the (segment, video) pairs the policy selected (see forward)
        # ADAMML_B200_EVAL_SKIP: "device" (default; compaction and work limit on the device, no host sync, the pass is
        # CUDA-graph capturable), "host" (decisions read back, data-dependent batch shape), "0" (run everything)
        mode = os.environ.get("ADAMML_B200_EVAL_SKIP", "device")
        self.skip_unselected = mode != "0"
        self.skip_mode = "host" if mode in ("host", "1") else "device"
        self._sel_state = None
        if rng_policy:
            self.freeze_policy_net()
            del self.policy_net.fcs

    # ------------------------------------------------------------------ data layer (adamml.py:42-67)
    def _input_norm(self, m, c, device):
        """(mean, std) per frame channel for decoded uint8 frames of modality m, as GroupNormalize repeats them
        (utils/video_transforms.py:77-78; values from mean()/std() below, adamml.py:93-109)."""
        key = (m, c, device)
        cache = self.__dict__.setdefault("_norm_cache", {})
        if key not in cache:
            mean, std = self.mean(m), self.std(m)
            if c % len(mean):
                raise ValueError("%d channels per frame cannot repeat a %d-entry mean" % (c, len(mean)))
            cache[key] = (torch.tensor(mean * (c // len(mean)), dtype=torch.float32, device=device),
                          torch.tensor(std * (c // len(std)), dtype=torch.float32, device=device))
        return cache[key]

    def data_layer(self, x, num_segments, p_rgb_size=(160, 160)):
        """-> (p_x, m_x): NHWC image batches ordered (segment, video, frame) in the compute dtype.
        A visual modality may arrive as decoded uint8 frames (same [N, S*F*C, H, W] shape): the scaling to [0,1] and
        the mean/std normalisation of the loader (ToTorchFormatTensor + GroupNormalize) then run inside the
        re-layout kernels, bit-identical to normalising on the host first."""
        p_x, m_x = [], []
        S, F = num_segments, self.num_frames_per_segment
        dt = self.compute_dtype
        for idx, (x_, m) in enumerate(zip(x, self.modality)):
            norm = None
            if x_.dtype == torch.uint8 and m != "sound":
                norm = self._input_norm(m, x_.size(1) // (S * F), x_.device)
            else:
                x_ = x_.float()
            if m == "sound":
                if x_.size(-1) != x_.size(-2):  # segments stacked along the last dim
                    x_ = torch.stack(x_.chunk(S, dim=-1), dim=1).reshape(x_.size(0), -1, x_.size(-2),
                                                                         x_.size(-1) // S)
                x_ = x_.reshape(x_.size(0), -1, x_.size(-2), x_.size(-1)).contiguous()
                c = x_.size(1) // S
                t = ops.pack_frames(x_, S, 1, c, dt)
                p_x.append(t)
                m_x.append(t)
                continue
            x_ = x_.contiguous()
            c = x_.size(1) // (S * F)
            if idx in self.p_data_idx:
                p_x.append(ops.resize_frames(x_, S, F, c, p_rgb_size[0], p_rgb_size[1], 2, dt, norm=norm))
            if idx in self.m_data_idx:
                net = self.main_net.nets[self.m_data_idx.index(idx)]
                m_x.append(net.pack_input(x_, S, norm=norm) if hasattr(net, "pack_input")
                           else ops.pack_frames(x_, S, F, c, dt, norm=norm))
        return p_x, m_x, S

    def forward(self, x, num_segments=None, noise=None):
        """x: list over modalities of [N, S*F*C, H, W] (sound: [N, S, 256, 256]).
        -> (logits [N, classes], decisions [N, S, M]).  `noise` (tests only): dict(expo=[S,M*N,2],
        drop=[per main modality [S*N*T', feat]]) replacing the torch RNG draws."""
        S = num_segments if num_segments else self.num_segments
        N = x[0].size(0)
        if not self.__dict__.get("_bn_keys_done"):  # after a possible convert_sync_batchnorm (train_adamml.py:125-127)
            from ..engine import assign_bn_keys
            assign_bn_keys(self)
            self.__dict__["_bn_keys_done"] = True
        # weight operands of every backbone: one multi-tensor launch here instead of ~360 per-layer ones (first pass:
        # per-layer launches that record the jobs); ADAMML_B200_PACK_CACHE=0 keeps the per-layer launches
        cache = self.__dict__.get("_pack_cache")
        if cache is None and PACK_CACHE:
            cache = self.__dict__["_pack_cache"] = ops.WeightPackCache()
        if cache is not None:
            cache.begin()
        p_x, m_x, S = self.data_layer(x, S)
        dev = x[0].device
        expo = noise["expo"] if noise else None
        drop_masks = noise.get("drop") if noise else None
        # RNG draws in the reference's order (policy Gumbel noise | rng decisions first, then the dropout masks)
        if not self.rng_policy:
            if expo is None:
                expo = self.policy_net.draw_gumbel_noise(S, N, dev)
            p_jobs = self.policy_net.backbone_jobs(p_x, S)
        else:  # adamml.py:76-78
            decisions = (torch.rand((S, self.num_modality, N), dtype=torch.float32, device=dev)
                         > self.rng_threshold).float()
            p_jobs = []
        if self._can_skip():
            return self._forward_selected(p_jobs, m_x, S, N, expo, None if not self.rng_policy else decisions)
        if drop_masks is None:
            drop_masks = self.main_net.draw_drop_masks(S, N, dev)
        m_jobs = self.main_net.backbone_jobs(m_x, S, drop_masks)
        # every backbone of the step (policy + main) is independent until the policy head / late fusion.  The main
        # backbones are enqueued FIRST: their millisecond-scale kernels keep the device busy while the host issues the
        # hundreds of microsecond-scale launches of the policy nets (eager mode; inside a captured graph the order
        # is irrelevant)
        outs = self._run_backbones(m_jobs + p_jobs)
        n_main = len(m_jobs)
        del p_x, m_x, p_jobs, m_jobs
        if not self.rng_policy:
            decisions, _ = self.policy_net(None, S, N, expo=expo, feats=outs[n_main:])
        outs = outs[:n_main]
        logits = self.main_net(None, decisions, S, N, per_mod=outs)
        return logits, decisions.permute(2, 0, 1)

    def _run_backbones(self, jobs):
        cache = self.__dict__.get("_pack_cache")
        return run_backbones_parallel([(n, x_, g, dict(e or {}, _packs=cache)) for n, x_, g, e in jobs])

    # ------------------------------------------------------------------ inference with decision-driven skipping
    def _can_skip(self):
        """The reference always runs every main backbone and multiplies its logits by the 0/1 decision
        (adamml.py:81-86, joint_resnet_mobilenetv2.py:94).  Without a tape and with every main BN on running
        statistics (utils/utils.py:427-507, validate_adamml) a (segment, video) pair is independent of the rest
        of the batch, so an unselected pair contributes exactly 0 and its backbone pass can be dropped."""
        return (self.skip_unselected and not torch.is_grad_enabled()
                and not any(mod.training for mod in self.main_net.modules()))

    @property
    def last_selected_fraction(self):
        """share of (segment, video, modality) pairs the last skipping pass ran (reads the device counts lazily)"""
        st = self._sel_state
        if st is None:
            return None
        if isinstance(st, tuple):
            counts, total = st
            self._sel_state = st = sum(int(c.item()) for c in counts) / float(total)
        return st

    @last_selected_fraction.setter
    def last_selected_fraction(self, v):
        self._sel_state = v

    def _forward_selected_device(self, m_x, S, N, decisions):
        """Skipping without a host round trip: compaction, gather, work limit and scatter all run on the device
        (csrc/gating.cu); shapes are static, so the pass can be captured into a CUDA graph."""
        dec = decisions.detach().float().contiguous()           # [S, M, N]
        SN = S * N
        jobs, meta = [], []
        for m, (net, x) in enumerate(zip(self.main_net.nets, m_x)):
            idx, count = ops.select_compact(dec, m)
            jobs.append((net, ops.gather_clips(x, idx, count, SN), 1, dict(_live=(count, SN))))
            meta.append((idx, count))
        outs = self._run_backbones(jobs)
        per_mod = [ops.scatter_rows(y.contiguous(), idx, count, SN) for y, (idx, count) in zip(outs, meta)]
        self._sel_state = ([c for _, c in meta], SN * len(m_x))
        logits = self.main_net(None, decisions, S, N, per_mod=per_mod)
        return logits, decisions.permute(2, 0, 1)

    def _forward_selected(self, p_jobs, m_x, S, N, expo, decisions):
        if decisions is None:
            feats = self._run_backbones(p_jobs)
            decisions, _ = self.policy_net(None, S, N, expo=expo, feats=feats)
            del feats
        if self.skip_mode == "device":
            return self._forward_selected_device(m_x, S, N, decisions)
        dev = decisions.device
        dec_host = decisions.detach().to("cpu", torch.float32)            # [S, M, N]; the one D2H sync of the pass
        SN = S * N
        jobs, slots, per_mod = [], [], [None] * len(m_x)
        chosen = 0
        for m, (net, x) in enumerate(zip(self.main_net.nets, m_x)):
            idx = torch.nonzero(dec_host[:, m, :].reshape(SN) > 0).squeeze(1)   # packed order is (segment, video)
            K = idx.numel()
            chosen += K
            if K == 0:
                per_mod[m] = torch.zeros((SN, self.main_net.num_classes), device=dev, dtype=torch.float32)
                continue
            if K < SN:
                idx = idx.to(dev, non_blocking=True)
                x = ops.select_clips(x, idx, SN)
            else:
                idx = None
            jobs.append((net, x, 1, None))
            slots.append((m, idx))
        outs = self._run_backbones(jobs)
        for (m, idx), y in zip(slots, outs):
            if idx is None:
                per_mod[m] = y
            else:
                full = torch.zeros((SN, y.shape[1]), device=dev, dtype=y.dtype)
                per_mod[m] = full.index_copy_(0, idx, y)
        self.last_selected_fraction = chosen / float(SN * len(m_x))
        logits = self.main_net(None, decisions, S, N, per_mod=per_mod)
        return logits, decisions.permute(2, 0, 1)

    def mean(self, modality="rgb"):
        return [0.485, 0.456, 0.406] if modality in ("rgb", "rgbdiff") else [0.5]

    def std(self, modality="rgb"):
        return [0.229, 0.224, 0.225] if modality in ("rgb", "rgbdiff") else [sum([0.229, 0.224, 0.225]) / 3]

    @property
    def network_name(self):
        name = "adamml"
        if self.rng_policy:
            name += "-rng-{:.1f}".format(self.rng_threshold)
        else:
            name += "-{}".format(self.policy_net.network_name)
        return name + "-{}".format(self.main_net.network_name)

    def decay_temperature(self, decay_ratio=None):
        self.policy_net.decay_temperature(decay_ratio if decay_ratio else self.decay_ratio)

    def _set_grad(self, net, flag):
        for p in net.parameters():
            p.requires_grad = flag

    def freeze_policy_net(self):
        self.update_policy_net = False
        self._set_grad(self.policy_net, False)

    def unfreeze_policy_net(self):
        self.update_policy_net = True
        self._set_grad(self.policy_net, True)

    def freeze_main_net(self):
        self.update_main_net = False
        self._set_grad(self.main_net, False)

    def unfreeze_main_net(self):
        self.update_main_net = True
        self._set_grad(self.main_net, True)


def adamml(groups, modality, input_channels, num_segments, rng_policy, rng_threshold, causality_modeling,
           num_classes, depth, without_t_stride, dropout, pooling_method, fusion_point, unimodality_pretrained,
           learnable_lf_weights, **kwargs):
    """Factory with the reference's kwargs (adamml.py:134-171)."""
    cd = kwargs.get("compute_dtype")
    if "rgbdiff" in modality and "flow" in modality:
        p_mod = [m for m in modality if m != "flow"]
        m_mod = [m for m in modality if m != "rgbdiff"]
        p_ch = [c for c, m in zip(input_channels, modality) if m != "flow"]
        m_ch = [c for c, m in zip(input_channels, modality) if m != "rgbdiff"]
    else:
        p_mod, m_mod, p_ch, m_ch = modality, modality, input_channels, input_channels
    policy_net = p_joint_mobilenet(num_frames=max(1, groups // 2), modality=p_mod, input_channels=p_ch,
                                   causality_modeling=causality_modeling, compute_dtype=cd)
    main_net = joint_resnet_mobilenetv2(depth=depth, num_classes=num_classes, without_t_stride=without_t_stride,
                                        groups=groups, dropout=dropout, pooling_method=pooling_method,
                                        input_channels=m_ch, fusion_point=fusion_point, modality=m_mod,
                                        unimodality_pretrained=unimodality_pretrained,
                                        learnable_lf_weights=learnable_lf_weights, compute_dtype=cd)
    return AdaMML(policy_net, main_net, num_frames=groups, num_segments=num_segments, modality=modality,
                  rng_policy=rng_policy, rng_threshold=rng_threshold, num_classes=num_classes, compute_dtype=cd)
