"""Golden values of the evaluation log: the REFERENCE's fill_batch_log (generate.py:111-164) and compute_metric
(wmar/utils/metrics.py:25-45) run here on the CPU over a toy tokenizer (tests/helpers.py ToyTokenizerModel) with the
reference's own augmentation classes and GentimeWatermark.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_evallog.py        -> tests/golden/evallog.npz
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")


def _stub_missing(names):
    """generate.py imports its optional attack stacks (DiffPure, neural compressors) at module level; they are not
    installed offline and are not on the path under test: give them empty stand-in modules."""
    import types

    class _Any(types.ModuleType):
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return type(k, (), {})

    import importlib.abc
    import importlib.machinery

    class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
        def find_spec(self, fullname, path, target=None):
            if any(fullname == n or fullname.startswith(n + ".") for n in names):
                return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
            return None

        def create_module(self, spec):
            m = _Any(spec.name)
            m.__path__ = []
            return m

        def exec_module(self, module):
            pass

    sys.meta_path.insert(0, _Finder())


def main():
    os.chdir(REF)
    _stub_missing({"diffusers", "compressai", "matplotlib", "seaborn", "wmar.augmentations.diffpure",
                   "wmar.augmentations.neuralcompression", "deps.saberi_wmr", "xformers", "wmar.models.chameleon_wrapper",
                   "deps.chameleon", "omegaconf", "wmar.models.rar_wrapper", "wmar.models.taming_wrapper", "wmar.utils.analyzer",
                   "wmar.sync", "syncseal", "wmar.watermarking.synchronization", "deps.watermark_anything", "skimage"})
    import generate as ref_generate
    from helpers import ToyTokenizerModel
    from wmar.augmentations.geometric import HorizontalFlip, UpperLeftCropWithResizeBack
    from wmar.augmentations.valuemetric import Brightness, GaussianBlur
    from wmar.utils.metrics import compute_metric
    from wmar.utils.utils import chw_to_pillow
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy

    model = ToyTokenizerModel("cpu")
    V = 64
    alive = torch.tensor([i for i in range(V) if i % 9 != 4], dtype=torch.long)
    dead = torch.tensor([i for i in range(V) if i % 9 == 4], dtype=torch.long)
    vq = {"alive_ids": alive, "dead_ids": dead, "embedding": torch.zeros(V, 4)}
    wm = GentimeWatermark(vq, V, SeedStrategy.LINEAR, SplitStrategy.RANDOM_STRATIFIED, 1, 2.0, 0.25, device="cpu")
    g = torch.Generator().manual_seed(5)
    codes = torch.randint(0, V, (3, 16), generator=g)
    augs = [("gaussian-blur", lambda x, k: GaussianBlur()(x, k), [0, 3, 7]),
            ("brightness", lambda x, b: Brightness()(x, b), [1, 1.5, 2.5]),
            ("flip-h", lambda x, do: HorizontalFlip()(x) if do else x, [0, 1]),
            ("upperleft-crop", lambda x, f: UpperLeftCropWithResizeBack()(x, f), [1.0, 0.8, 0.5])]
    log = {}
    ref_generate.fill_batch_log(log, "wm", model, codes, {"max_roundtrips": 2, "augmentations": augs})
    out = {"codes": codes.numpy(), "alive": alive.numpy(), "dead": dead.numpy()}
    orig_codes, orig_imgs = log["wm"]["roundtrips"][0][1], log["wm"]["roundtrips"][0][2]
    for transform, entries in log["wm"].items():
        for j, (param, c, imgs, _) in enumerate(entries):
            out[f"{transform}/{j}/param"] = np.float64(param)
            out[f"{transform}/{j}/codes"] = c
            for name in ("l0", "psnr", "pvalue"):
                vals = []
                for b in range(c.shape[0]):
                    img, oimg = chw_to_pillow(torch.from_numpy(imgs[b])), chw_to_pillow(torch.from_numpy(orig_imgs[b]))
                    with np.errstate(divide="ignore"):
                        vals.append(compute_metric(name, c[b], orig_codes[b], img, oimg, wm, transform, param))
                out[f"{transform}/{j}/{name}"] = np.asarray(vals, dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, "evallog.npz"), **out)
    print("evallog.npz:", len(out), "arrays;", {k: v.tolist() for k, v in out.items() if k.endswith("1/l0")})


if __name__ == "__main__":
    main()
