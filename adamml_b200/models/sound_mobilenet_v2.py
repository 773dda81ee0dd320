"""MobileNetV2 for 1x256x256 log-spectrograms (drop-in for reference models/sound_mobilenet_v2.py).

Parameter holder with the reference's state_dict names (features.N.0 / features.N.conv.K,
classifier.1); executed by adamml_b200.engine on the CUDA kernels.
"""
import torch
import torch.nn as nn

from .. import ops
from ..engine import run_backbone
from ..ops import ACT_NONE, ACT_RELU6
from .resnet import default_compute_dtype

# t (expansion), c (channels), n (repeats), s (stride) — sound_mobilenet_v2.py:101-110
MBV2_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                (6, 320, 1, 1))


def _cbr(cin, cout, k=3, stride=1, groups=1):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, groups=groups, bias=False),
                         nn.BatchNorm2d(cout), nn.ReLU6(inplace=True))


class _InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, t):
        super().__init__()
        hid = int(round(inp * t))
        self.use_res_connect = stride == 1 and inp == oup
        mods = []
        if t != 1:
            mods.append(_cbr(inp, hid, 1))
        mods += [_cbr(hid, hid, 3, stride, groups=hid), nn.Conv2d(hid, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*mods)

    def layers(self):
        m = list(self.conv)
        out = [(s[0], s[1], ACT_RELU6) for s in m[:-2]]
        out.append((m[-2], m[-1], ACT_NONE))
        return out


class MobileNetV2(nn.Module):
    def __init__(self, num_classes=1000, input_channels=3, dropout=0.5, compute_dtype=None):
        super().__init__()
        self.input_channels = input_channels
        self.dropout_p = dropout
        self.last_channel = 1280
        self.compute_dtype = compute_dtype or default_compute_dtype()
        feats = [_cbr(input_channels, 32, 3, 2)]
        cin = 32
        for t, c, n, s in MBV2_SETTING:
            for i in range(n):
                feats.append(_InvertedResidual(cin, c, s if i == 0 else 1, t))
                cin = c
        feats.append(_cbr(cin, self.last_channel, 1))
        self.features = nn.Sequential(*feats)
        self.classifier = nn.Sequential(nn.Dropout(dropout), nn.Linear(self.last_channel, num_classes))
        for m in self.modules():  # sound_mobilenet_v2.py:137-147
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)
            elif isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.zeros_(m.bias)

    def mean(self, modality="rgb"):
        return [0.485, 0.456, 0.406] if modality in ("rgb", "rgbdiff") else [0.5]

    def std(self, modality="rgb"):
        return [0.229, 0.224, 0.225] if modality in ("rgb", "rgbdiff") else [sum([0.229, 0.224, 0.225]) / 3]

    # ------------------------------------------------------------------ engine program
    def run_forward(self, ex, x, extra):
        """x NHWC [G*N, 256, 256, C] -> logits [G*N, classes] (sound_mobilenet_v2.py:152-162)."""
        f = list(self.features)
        a = ex.cba(x, f[0][0], f[0][1], ACT_RELU6)
        for blk in f[1:-1]:
            a = ex.inverted_residual(a, blk.layers(), blk.use_res_connect)
        a = ex.cba(a, f[-1][0], f[-1][1], ACT_RELU6)
        feat = ex.avgpool(a)
        if extra.get("features_only"):
            return feat
        return ex.classifier(feat, self.classifier[1], extra.get("drop_mask") if self.training else None, 1)

    def run_backward(self, ex, dy):
        d = ex.classifier_bwd(dy)
        d = ex.avgpool_bwd(d)
        d, _ = ex.cba_bwd(d)
        for _ in range(len(self.features) - 2):
            d = ex.inverted_residual_bwd(d)
        ex.cba_bwd(d, need_dx=False)

    def draw_drop_mask(self, rows, device):
        if not self.training or self.dropout_p <= 0:
            return None
        keep = 1.0 - self.dropout_p
        return torch.empty((rows, self.last_channel), device=device).bernoulli_(keep).div_(keep)

    def forward(self, x, drop_mask=None):
        n, c, h, w = x.shape
        xn = ops.pack_frames(x.contiguous().float(), 1, 1, c, self.compute_dtype)
        if drop_mask is None:
            drop_mask = self.draw_drop_mask(n, x.device)
        return run_backbone(self, xn, 1, dict(drop_mask=drop_mask))


def sound_mobilenet_v2(num_classes, input_channels, dropout, imagenet_pretrained=True, **kwargs):
    """Factory with the reference's kwargs (sound_mobilenet_v2.py:177-198)."""
    if imagenet_pretrained:
        raise RuntimeError("imagenet_pretrained=True needs torchvision weights from the network; load a checkpoint "
                           "with load_state_dict() and pass imagenet_pretrained=False")
    return MobileNetV2(num_classes=num_classes, input_channels=input_channels, dropout=dropout,
                       compute_dtype=kwargs.get("compute_dtype"))
