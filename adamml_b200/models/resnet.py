"""ResNet with temporal max-pooling (drop-in for reference models/resnet.py).

The nn.Module tree only HOLDS parameters/buffers under the reference's state_dict names
(conv1, bn1, layer{1..4}.{i}.conv{1,2,3}/bn{1,2,3}/downsample.{0,1}, fc); the arithmetic is
executed by adamml_b200.engine on the CUDA kernels, for all segments in one batched pass.
"""
import torch
import torch.nn as nn

from .. import ops
from ..engine import run_backbone
from ..ops import ACT_RELU

_LAYERS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def default_compute_dtype():
    """ADAMML_B200_PRECISION = x2 (default: two-plane forward, meets the 1e-3 / bit-exact-selection bar) | bf16 (speed
    mode) | fp32 (exact CUDA-core engine); see ops.PREC_X2."""
    import os
    return {"x2": ops.PREC_X2, "bf16": torch.bfloat16, "fp32": torch.float32}[
        os.environ.get("ADAMML_B200_PRECISION", "x2")]


class _Block(nn.Module):
    """Parameter holder for BasicBlock (resnet.py:46-74) / Bottleneck (resnet.py:77-113)."""

    def __init__(self, inplanes, planes, stride, bottleneck, with_downsample):
        super().__init__()
        self.bottleneck = bottleneck
        out = planes * (4 if bottleneck else 1)
        if bottleneck:
            self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
            self.conv3 = nn.Conv2d(planes, out, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(out)
        else:
            self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = None
        if with_downsample:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, out, 1, stride, bias=False), nn.BatchNorm2d(out))


class ResNet(nn.Module):
    expansion = {True: 4, False: 1}

    def __init__(self, depth, num_frames, num_classes=1000, dropout=0.5, zero_init_residual=False,
                 without_t_stride=False, pooling_method="max", input_channels=3, compute_dtype=None):
        super().__init__()
        self.depth = depth
        self.num_frames = self.orig_num_frames = num_frames
        self.num_classes = num_classes
        self.without_t_stride = without_t_stride
        self.pooling_method = pooling_method.lower()
        if self.pooling_method not in ("max", "avg"):
            raise ValueError("only support avg or max")
        self.input_channels = input_channels
        self.dropout_p = dropout
        self.compute_dtype = compute_dtype or default_compute_dtype()
        bott = depth >= 50
        self.conv1 = nn.Conv2d(input_channels, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), _LAYERS[depth])):
            blocks = []
            for bi in range(n):
                stride = 2 if (li > 0 and bi == 0) else 1
                need_ds = bi == 0 and (stride != 1 or inplanes != planes * self.expansion[bott])
                blocks.append(_Block(inplanes, planes, stride, bott, need_ds))
                inplanes = planes * self.expansion[bott]
            setattr(self, f"layer{li + 1}", nn.Sequential(*blocks))
        self.dropout = nn.Dropout(dropout)
        self.fc = nn.Linear(inplanes, num_classes)
        self.feature_dim = inplanes

    # frames left after the three temporal pools (resnet.py:144-154)
    @property
    def out_frames(self):
        t = self.orig_num_frames
        if not self.without_t_stride:
            for _ in range(3):
                t = max(1, t // 2)
        return t

    def mean(self, modality="rgb"):
        return [0.485, 0.456, 0.406] if modality in ("rgb", "rgbdiff") else [0.5]

    def std(self, modality="rgb"):
        return [0.229, 0.224, 0.225] if modality in ("rgb", "rgbdiff") else [sum([0.229, 0.224, 0.225]) / 3]

    # ------------------------------------------------------------------ engine program
    def run_forward(self, ex, x, extra):
        """x: NHWC [G*videos*frames, H, W, C] -> logits [G*videos, classes] (resnet.py:195-223)."""
        frames = self.orig_num_frames
        a = ex.cba_maxpool(x, self.conv1, self.bn1, ACT_RELU)
        for li in range(4):
            for blk in getattr(self, f"layer{li + 1}"):
                a = ex.bottleneck(a, blk) if blk.bottleneck else ex.basicblock(a, blk)
            if li < 3 and not self.without_t_stride:
                a = ex.tpool(a, frames, self.pooling_method == "avg")
                frames = max(1, frames // 2)
        feat = ex.avgpool(a)
        if extra.get("features_only"):
            return feat
        return ex.classifier(feat, self.fc, extra.get("drop_mask") if self.training else None, frames)

    def run_backward(self, ex, dy):
        d = ex.classifier_bwd(dy)
        d = ex.avgpool_bwd(d)
        for li in reversed(range(4)):
            if li < 3 and not self.without_t_stride:
                d = ex.tpool_bwd(d)
            blocks = list(getattr(self, f"layer{li + 1}"))
            for blk in reversed(blocks):
                d = ex.bottleneck_bwd(d) if blk.bottleneck else ex.basicblock_bwd(d)
        d = ex.maxpool_bwd(d)
        ex.cba_bwd(d, need_dx=False)

    def pack_input(self, x, S, norm=None):
        """NCHW fp32 (or uint8 + norm) [N, S*F*C, H, W] -> the stem operand of the engine: space-to-depth bf16 for
        the tensor-core stem, plain NHWC otherwise (adamml.py:53,65 / resnet.py:197)."""
        f = self.orig_num_frames
        c = x.shape[1] // (S * f)
        if ops.stem_s2d_ok(self.conv1, c, x.shape[2], x.shape[3], self.compute_dtype):
            return ops.pack_frames_s2d(x, S, f, c, norm=norm, x2=self.compute_dtype == ops.PREC_X2)
        return ops.pack_frames(x, S, f, c, self.compute_dtype, norm=norm)

    # ------------------------------------------------------------------ unimodal API (resnet.py:195)
    def draw_drop_mask(self, rows, device):
        if not self.training or self.dropout_p <= 0:
            return None
        keep = 1.0 - self.dropout_p
        return torch.empty((rows, self.feature_dim), device=device).bernoulli_(keep).div_(keep)

    def forward(self, x, drop_mask=None):
        n, ct, h, w = x.shape
        if ct == 1:
            raise ValueError("single-channel (audio) input is served by sound_mobilenet_v2, not ResNet")
        xn = self.pack_input(x.contiguous().float(), 1)
        if drop_mask is None:
            drop_mask = self.draw_drop_mask(n * self.out_frames, x.device)
        return run_backbone(self, xn, 1, dict(drop_mask=drop_mask))


def resnet(depth, num_classes, without_t_stride, groups, dropout, pooling_method, input_channels,
           imagenet_pretrained=True, **kwargs):
    """Factory with the reference's kwargs (resnet.py:244-259).  ImageNet weights cannot be downloaded
    here (no network): imagenet_pretrained=True raises instead of silently training from scratch."""
    if imagenet_pretrained:
        raise RuntimeError("imagenet_pretrained=True needs torchvision weights from the network; load a checkpoint "
                           "with load_state_dict() (state_dict keys match the reference) and pass "
                           "imagenet_pretrained=False")
    return ResNet(depth, num_frames=groups, num_classes=num_classes, without_t_stride=without_t_stride,
                  dropout=dropout, pooling_method=pooling_method, input_channels=input_channels,
                  compute_dtype=kwargs.get("compute_dtype"))
