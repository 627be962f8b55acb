"""CPU: the augmentation restatement (oracle/augment.py) vs goldens produced by the reference's own classes."""
import os

import numpy as np
import torch

from helpers import G


def test_augment_oracle_matches_reference_classes():
    from oracle import augment as oa
    g = np.load(os.path.join(G, "augment.npz"))
    img = torch.from_numpy(g["image"])
    for k in (3, 9, 19):
        np.testing.assert_array_equal(oa.gaussian_blur(img, k).numpy(), g[f"blur/{k}"])
    for f in (1.25, 2.5):
        np.testing.assert_array_equal(oa.brightness(img, f).numpy(), g[f"brightness/{f}"])
    np.testing.assert_array_equal(oa.gaussian_noise(img, 0.1, torch.from_numpy(g["noise/draw"])).numpy(), g["noise/0.1"])
    for a in (-20, -5, 10, 20):
        np.testing.assert_array_equal(oa.rotate(img, a).numpy(), g[f"rotate/{a}"])
    np.testing.assert_array_equal(oa.hflip(img).numpy(), g["hflip"])
    for f in (0.95, 0.75, 0.5):
        np.testing.assert_array_equal(oa.crop_resize_back(img, f).numpy(), g[f"crop_resize/{f}"])
        np.testing.assert_array_equal(oa.crop_pad_back(img, f).numpy(), g[f"crop_pad/{f}"])


def test_gaussian_kernel_and_rotation_matrix_match_torchvision():
    """The two host-side parameter computations of wmar_b200.augmentations (no GPU needed)."""
    import math
    from torchvision.transforms import functional as F
    from torchvision.transforms import _functional_tensor as FT
    import importlib.util
    import sys
    spec = importlib.util.spec_from_file_location("aug_params", os.path.join(os.path.dirname(G), "..", "wmar_b200", "augmentations.py"))
    src = open(spec.origin).read()
    ns = {}
    # only the two pure functions (the module itself imports the CUDA library lazily but needs the package context)
    start = src.index("def gaussian_kernel2d")
    end = src.index("class Identity")
    exec("import math\nimport torch\n" + src[start:end], ns)
    for k in (3, 9, 19):
        want = FT._get_gaussian_kernel2d([k, k], [k * 0.15 + 0.35] * 2, torch.float32, torch.device("cpu"))
        assert torch.equal(ns["gaussian_kernel2d"](k), want)
    for a in (-90, -20, 5, 85, 90):
        want = F._get_inverse_affine_matrix([0.0, 0.0], -a, [0.0, 0.0], 1.0, [0.0, 0.0])
        got = ns["inverse_rotation_matrix"](a)
        assert all(math.isclose(x, y, rel_tol=0, abs_tol=1e-15) for x, y in zip(got, want)), (a, got, want)
