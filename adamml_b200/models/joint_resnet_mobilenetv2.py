"""Main network: one backbone per modality, decision-gated late fusion at the logits
(drop-in for reference models/joint_resnet_mobilenetv2.py, fusion_point='logits')."""
import torch
import torch.nn as nn

from .._lib import call
from ..engine import run_backbone
from .resnet import ResNet
from .sound_mobilenet_v2 import MobileNetV2


class _GateFuse(torch.autograd.Function):
    """out[n,c] = 1/S sum_s sum_m w_m d[s,m,n] logits[m,s,n,c]
    (joint_resnet_mobilenetv2.py:92-97,112-127 + adamml.py:88)."""

    @staticmethod
    def forward(ctx, logits, dec, lf):
        M, S, N, C = logits.shape
        logits = logits.contiguous()
        dec = dec.contiguous() if dec is not None else None  # the kernel indexes [S][M][N] densely
        out = torch.empty((N, C), device=logits.device, dtype=torch.float32)
        call("fuse_fwd", logits, dec, lf, out, M, S, N, C)
        ctx.save_for_backward(logits, dec, lf)
        return out

    @staticmethod
    def backward(ctx, g):
        logits, dec, lf = ctx.saved_tensors
        M, S, N, C = logits.shape
        g = g.contiguous()
        dlogits = torch.empty_like(logits) if ctx.needs_input_grad[0] else None
        ddec = torch.empty_like(dec) if (dec is not None and ctx.needs_input_grad[1]) else None
        dlf = torch.empty_like(lf) if (lf is not None and ctx.needs_input_grad[2]) else None
        call("fuse_bwd", g, logits, dec, lf, dlogits, ddec, dlf, M, S, N, C)
        return dlogits, ddec, dlf


class JointResNetMobileNetV2(nn.Module):
    def __init__(self, depth, num_frames, modality, num_classes=1000, dropout=0.5, zero_init_residual=False,
                 without_t_stride=False, pooling_method="max", input_channels=None, fusion_point="logits",
                 learnable_lf_weights=False, compute_dtype=None):
        super().__init__()
        if fusion_point != "logits":
            raise NotImplementedError("fusion_point='fc2' raises inside AdaMML in the reference too "
                                      "(joint_resnet_mobilenetv2.py:96); only 'logits' is on the accelerated path")
        self.depth = depth
        self.num_frames = num_frames
        self.without_t_stride = without_t_stride
        self.pooling_method = pooling_method
        self.fusion_point = fusion_point
        self.modality = modality
        self.learnable_lf_weights = learnable_lf_weights
        self.num_classes = num_classes
        self.nets = nn.ModuleList()
        for i, m in enumerate(modality):
            if m != "sound":
                self.nets.append(ResNet(depth, num_frames, num_classes, dropout, zero_init_residual, without_t_stride,
                                        pooling_method, input_channels[i], compute_dtype=compute_dtype))
            else:
                self.nets.append(MobileNetV2(num_classes, dropout=dropout, input_channels=input_channels[i],
                                             compute_dtype=compute_dtype))
        self.lf_weights = None
        if learnable_lf_weights:
            self.lf_weights = nn.Parameter(torch.tensor([1.0 / len(modality)] * (len(modality) - 1)))

    def mean(self, modality="rgb"):
        return [0.485, 0.456, 0.406] if modality in ("rgb", "rgbdiff") else [0.5]

    def std(self, modality="rgb"):
        return [0.229, 0.224, 0.225] if modality in ("rgb", "rgbdiff") else [sum([0.229, 0.224, 0.225]) / 3]

    @property
    def network_name(self):
        name = "joint_resnet-{}_mobilenet_v2-{}".format(self.depth, self.fusion_point)
        if self.lf_weights is not None:
            name += "-llf" if self.learnable_lf_weights else "-llfc"
        if not self.without_t_stride:
            name += "-ts-{}".format(self.pooling_method)
        return name

    def draw_drop_masks(self, S, N, device):
        """Dropout masks in the reference's RNG order: per segment, per modality."""
        if not self.training:
            return None
        per_seg = [[net.draw_drop_mask(N * getattr(net, "out_frames", 1), device) for net in self.nets]
                   for _ in range(S)]
        if per_seg[0][0] is None:
            return None
        return [torch.cat([per_seg[s][m] for s in range(S)], 0) for m in range(len(self.nets))]

    def backbone_jobs(self, m_x, S, drop_masks):
        """(net, x, groups, extra) per main modality, for engine.run_backbones_parallel"""
        return [(net, x, S, dict(drop_mask=drop_masks[i] if drop_masks else None))
                for i, (net, x) in enumerate(zip(self.nets, m_x))]

    def forward(self, m_x, decisions, S, N, drop_masks=None, per_mod=None):
        """m_x: NHWC image batches (segment-major) per main modality; decisions [S, M, N] or None; per_mod: per-modality
        logits when the caller already ran the backbones.  -> fused logits [N, classes]."""
        if per_mod is None:
            if drop_masks is None:
                drop_masks = self.draw_drop_masks(S, N, m_x[0].device)
            per_mod = [run_backbone(*job) for job in self.backbone_jobs(m_x, S, drop_masks)]
        logits = torch.stack(per_mod, 0).view(len(per_mod), S, N, -1)
        return _GateFuse.apply(logits, decisions, self.lf_weights)


def joint_resnet_mobilenetv2(depth, num_classes, without_t_stride, groups, dropout, pooling_method, input_channels,
                             fusion_point, modality, unimodality_pretrained, learnable_lf_weights, **kwargs):
    model = JointResNetMobileNetV2(depth, num_frames=groups, num_classes=num_classes,
                                   without_t_stride=without_t_stride, dropout=dropout, pooling_method=pooling_method,
                                   input_channels=input_channels, fusion_point=fusion_point, modality=modality,
                                   learnable_lf_weights=learnable_lf_weights,
                                   compute_dtype=kwargs.get("compute_dtype"))
    if len(unimodality_pretrained) > 0:  # joint_resnet_mobilenetv2.py:141-155
        if len(unimodality_pretrained) != len(model.nets):
            raise ValueError("the number of pretrained models is incorrect.")
        for i, _ in enumerate(modality):
            print("Loading unimodality pretrained model from: {}".format(unimodality_pretrained[i]))
            sd = torch.load(unimodality_pretrained[i], map_location="cpu")["state_dict"]
            model.nets[i].load_state_dict({k.replace("module.", ""): v for k, v in sd.items()}, strict=True)
    return model
