"""Host-side mirror of ``wmar.augmentations`` for the evaluation round trip: same class
names and call signatures as valuemetric.py:43-140 / geometric.py:15-117, the arithmetic runs in csrc/augment.cu
(``wmar_augment``).  All augmentations expect CUDA images [B,3,H,W] in [0,1] (augmentation_manager.py:24).

Host work is limited to parameters: the k x k Gaussian weights (torchvision ``_get_gaussian_kernel2d``: 25 floats..361
floats), the 2x3 inverse affine matrix of ``F.rotate`` (``_get_inverse_affine_matrix``), and drawing the N(0,1) noise
with ``torch.randn_like`` so the CUDA generator advances exactly like the reference's (valuemetric.py:134).  JPEG is the
reference's own PIL round trip on the host (valuemetric.py:18-40).
"""
import ctypes
import io
import math

import torch

from . import _lib

OP = dict(brightness=0, noise=1, blur=2, hflip=3, affine=4, crop_resize=5, crop_pad=6)


def _check(img):
    assert isinstance(img, torch.Tensor) and img.is_cuda and img.ndim == 4 and img.shape[1] == 3, "images must be CUDA [B,3,H,W]"
    return img.to(torch.float32).contiguous()


def _run(op, img, params=(), aux=None, out_hw=None):
    img = _check(img)
    B, _, H, W = img.shape
    OH, OW = out_hw if out_hw is not None else (H, W)
    out = torch.empty((B, 3, OH, OW), dtype=torch.float32, device=img.device)
    arr = (ctypes.c_float * max(1, len(params)))(*[float(p) for p in params])
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().wmar_augment(OP[op], _lib.ptr(img), _lib.ptr(out), B, H, W, OH, OW, arr, len(params),
                                           _lib.ptr(aux) if aux is not None else None, _lib.current_stream()))
    return out


def gaussian_kernel2d(kernel_size):
    """torchvision.transforms._functional_tensor._get_gaussian_kernel2d with F.gaussian_blur's default sigma."""
    sigma = kernel_size * 0.15 + 0.35
    half = (kernel_size - 1) * 0.5
    x = torch.linspace(-half, half, steps=kernel_size)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    k1 = pdf / pdf.sum()
    return torch.mm(k1[:, None], k1[None, :])


def inverse_rotation_matrix(angle):
    """torchvision F.rotate: _get_inverse_affine_matrix([0, 0], -angle, [0, 0], 1.0, [0, 0])."""
    rot = math.radians(-angle)
    a, b, c, d = math.cos(rot), -math.sin(rot), math.sin(rot), math.cos(rot)
    return [d, -b, 0.0, -c, a, 0.0]


class Identity:
    def __call__(self, image, *args, **kwargs):
        return image

    def __repr__(self):
        return "Identity"


class GaussianBlur:
    def __call__(self, image, kernel_size=None):
        if kernel_size == 0:
            return image
        assert kernel_size is not None and kernel_size % 2 == 1, "kernel size must be given and odd"
        k2 = gaussian_kernel2d(int(kernel_size)).to(image.device).contiguous()
        return _run("blur", image, (kernel_size,), k2)

    def __repr__(self):
        return "GaussianBlur"


class Brightness:
    def __call__(self, image, factor=None):
        assert factor is not None
        return _run("brightness", image, (factor,))

    def __repr__(self):
        return "Brightness"


class GaussianNoise:
    def __call__(self, image, std=None):
        assert std is not None
        image = _check(image)
        noise = torch.randn_like(image)          # same generator calls as the reference
        return _run("noise", image, (std,), noise)

    def __repr__(self):
        return "GaussianNoise"


class HorizontalFlip:
    def __call__(self, image, *args, **kwargs):
        return _run("hflip", image)

    def __repr__(self):
        return "HorizontalFlip"


class Rotate:
    """geometric.py:42-51: rotate by the multiple of 90 below the angle with expand=True, then by the rest."""

    def __call__(self, image, angle=None):
        assert angle is not None
        base = angle // 90 * 90
        rest = angle - base
        H, W = image.shape[-2:]
        if base != 0:
            oh, ow = (W, H) if (base // 90) % 2 else (H, W)
            image = _run("affine", image, inverse_rotation_matrix(base), out_hw=(oh, ow))
        if rest != 0:
            image = _run("affine", image, inverse_rotation_matrix(rest))
        return image

    def __repr__(self):
        return "Rotate"


class UpperLeftCropWithResizeBack:
    def __call__(self, image, crop_size=None):
        H, W = image.shape[-2:]
        return _run("crop_resize", image, (int(crop_size * H), int(crop_size * W)))


class UpperLeftCropWithPadBack:
    def __call__(self, image, crop_size=None):
        H, W = image.shape[-2:]
        return _run("crop_pad", image, (int(crop_size * H), int(crop_size * W)))


class JPEG:
    """The reference's PIL round trip on the host (valuemetric.py:18-70), image by image."""

    def __call__(self, image, quality=None):
        import numpy as np
        from PIL import Image
        image = torch.clamp(image, 0, 1)
        out = []
        for im in image:
            arr = im.mul(255).byte().permute(1, 2, 0).cpu().numpy()      # ToPILImage: mul(255).byte()
            buf = io.BytesIO()
            Image.fromarray(arr).save(buf, format="JPEG", quality=int(quality))
            buf.seek(0)
            dec = np.asarray(Image.open(buf).convert("RGB"), dtype=np.uint8)
            out.append(torch.from_numpy(dec.copy()).permute(2, 0, 1).float().div(255))
        return torch.stack(out).to(image.device)

    def __repr__(self):
        return "JPEG"


def default_augmentations():
    """The (name, fn, params) list of AugmentationManager (augmentation_manager.py:38-71) without the neural compressors."""
    return [
        ("gaussian-blur", lambda x, k: GaussianBlur()(x, k), [0, 1, 3, 5, 7, 9, 11, 13, 15, 17, 19]),
        ("gaussian-noise", lambda x, s: GaussianNoise()(x, s), [0, 0.025, 0.05, 0.075, 0.1, 0.125, 0.15, 0.175, 0.2]),
        ("jpeg", lambda x, q: JPEG()(x, q), [100, 95, 85, 75, 65, 55, 45, 35, 25, 15, 5]),
        ("brightness", lambda x, b: Brightness()(x, b), [1, 1.25, 1.5, 1.75, 2, 2.25, 2.5, 2.75, 3]),
        ("rotation", lambda x, a: Rotate()(x, a), [-20, -15, -10, -5, 0, 5, 10, 15, 20]),
        ("flip-h", lambda x, do: HorizontalFlip()(x) if do else x, [0, 1]),
        ("upperleft-crop", lambda x, f: UpperLeftCropWithResizeBack()(x, f), [1.0, 0.95, 0.9, 0.85, 0.8, 0.75, 0.7, 0.65, 0.6, 0.55, 0.5]),
    ]
