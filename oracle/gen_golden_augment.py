"""Generates tests/golden/augment.npz by running the reference's own augmentation classes (imported from /root/reference;
they cannot travel to the GPU box) on a seeded image batch.

    python oracle/gen_golden_augment.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from wmar.augmentations.geometric import (HorizontalFlip, Rotate, UpperLeftCropWithPadBack,  # noqa: E402
                                          UpperLeftCropWithResizeBack)
from wmar.augmentations.valuemetric import Brightness, GaussianBlur, GaussianNoise  # noqa: E402


def main():
    torch.manual_seed(0)
    # smooth-ish random image (blocks + noise) so that interpolation differences are visible but bounded
    img = torch.rand(2, 3, 8, 8).repeat_interleave(8, 2).repeat_interleave(8, 3) * 0.7 + torch.rand(2, 3, 64, 64) * 0.3
    out = {"image": img.numpy()}
    for k in (3, 9, 19):
        out[f"blur/{k}"] = GaussianBlur()(img, k).numpy()
    for f in (1.25, 2.5):
        out[f"brightness/{f}"] = Brightness()(img, f).numpy()
    torch.manual_seed(5)
    noise = torch.randn_like(img)
    torch.manual_seed(5)
    out["noise/0.1"] = GaussianNoise()(img, 0.1).numpy()
    out["noise/draw"] = noise.numpy()
    for a in (-20, -5, 10, 20):
        out[f"rotate/{a}"] = Rotate()(img, a).numpy()
    out["hflip"] = HorizontalFlip()(img).numpy()
    for f in (0.95, 0.75, 0.5):
        out[f"crop_resize/{f}"] = UpperLeftCropWithResizeBack()(img, f).numpy()
        out[f"crop_pad/{f}"] = UpperLeftCropWithPadBack()(img, f).numpy()
    path = os.path.join(ROOT, "tests", "golden", "augment.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
