"""TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's evaluation augmentations (wmar/augmentations/
valuemetric.py:76-137, geometric.py:26-117): each function makes the same torchvision.transforms.functional calls as the
reference class (torchvision is the third-party library that holds the arithmetic; 0.26 here, `torchvision` unpinned in
the reference's requirements).  Pinned by tests/golden/augment.npz, produced by the imported reference classes
(oracle/gen_golden_augment.py)."""
import torch
import torchvision.transforms.functional as F


def gaussian_blur(image, kernel_size):                 # valuemetric.py:88-94
    if kernel_size == 0:
        return image
    return F.gaussian_blur(image, kernel_size).clamp(0, 1)


def brightness(image, factor):                         # valuemetric.py:111-115
    return F.adjust_brightness(image, factor).clamp(0, 1)


def gaussian_noise(image, std, noise):                 # valuemetric.py:132-137 (noise = torch.randn_like(image))
    return (image + noise * std).clamp(0, 1)


def rotate(image, angle):                              # geometric.py:42-51
    base = angle // 90 * 90
    rest = angle - base
    image = F.rotate(image, base, expand=True)
    return F.rotate(image, rest)


def hflip(image):                                      # geometric.py:112-114
    return F.hflip(image)


def crop_resize_back(image, factor):                   # geometric.py:72-92
    H, W = image.shape[-2:]
    h, w = int(factor * H), int(factor * W)
    return F.resize(F.crop(image, 0, 0, h, w), (H, W), antialias=True)


def crop_pad_back(image, factor):                      # geometric.py:100-105
    H, W = image.shape[-2:]
    h, w = int(factor * H), int(factor * W)
    pad = H - h
    return F.pad(F.crop(image, 0, 0, h, w), (0, 0, pad, pad), padding_mode="constant")
