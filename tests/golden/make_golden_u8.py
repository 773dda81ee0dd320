#!/usr/bin/env python
"""Golden vector for the uint8 input path: runs the REFERENCE's own loader transforms
(utils/video_transforms.py: Stack -> ToTorchFormatTensor -> GroupNormalize) on small random uint8 frames.
Run in the build container only (needs /root/reference); the fixture travels, the reference does not.
Usage: python tests/golden/make_golden_u8.py"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference")
from utils.video_transforms import GroupNormalize, Stack, ToTorchFormatTensor  # noqa: E402


def main():
    rng = np.random.RandomState(7)
    out = {}
    # rgb: 4 frames of 6x5 RGB; flow: 10 single-channel planes (x/y pairs) -> mean [0.5], std [mean of rgb stds]
    cases = {"rgb": ("RGB", 4, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]),
             "flow": ("L", 10, [0.5], [float(np.mean([0.229, 0.224, 0.225]))])}
    for name, (mode, n, mean, std) in cases.items():
        frames = [Image.fromarray(rng.randint(0, 256, (6, 5, 3) if mode == "RGB" else (6, 5)).astype(np.uint8), mode)
                  for _ in range(n)]
        stacked = Stack()(frames)                                   # HW(FC) uint8
        u8_chw = torch.from_numpy(stacked).permute(2, 0, 1).contiguous()
        t = ToTorchFormatTensor()(stacked)                          # float / 255
        t = GroupNormalize(mean, std)(t)
        out[name] = dict(u8=u8_chw, normalized=t.clone(), mean=mean, std=std)
        print(name, tuple(u8_chw.shape), t.dtype, float(t.abs().max()))
    torch.save(out, os.path.join(HERE, "u8_normalize.pt"))


if __name__ == "__main__":
    main()
