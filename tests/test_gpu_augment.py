"""GPU parity: the augmentation kernels (csrc/augment.cu through wmar_b200.augmentations) vs goldens produced by the
reference's own classes (tests/golden/augment.npz, oracle/gen_golden_augment.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import G

pytestmark = pytest.mark.gpu


def test_augmentations_match_reference_classes():
    from wmar_b200 import augmentations as A
    g = np.load(os.path.join(G, "augment.npz"))
    img = torch.from_numpy(g["image"]).cuda()
    # element-wise ops: bit-exact
    for f in (1.25, 2.5):
        np.testing.assert_array_equal(A.Brightness()(img, f).cpu().numpy(), g[f"brightness/{f}"])
    torch.manual_seed(5)                                   # the CUDA generator draws other numbers than the CPU one:
    got = A.GaussianNoise()(img, 0.1)                      # check the fused op with the golden's own draw instead
    from wmar_b200.augmentations import _run
    got = _run("noise", img, (0.1,), torch.from_numpy(g["noise/draw"]).cuda())
    np.testing.assert_array_equal(got.cpu().numpy(), g["noise/0.1"])
    np.testing.assert_array_equal(A.HorizontalFlip()(img).cpu().numpy(), g["hflip"])
    for f in (0.95, 0.75, 0.5):
        np.testing.assert_array_equal(A.UpperLeftCropWithPadBack()(img, f).cpu().numpy(), g[f"crop_pad/{f}"])
        # bilinear up-scaling: fp32 interpolation weights computed in a different order than torch's antialias kernel
        np.testing.assert_allclose(A.UpperLeftCropWithResizeBack()(img, f).cpu().numpy(), g[f"crop_resize/{f}"], atol=3e-6, rtol=0)
    # Gaussian blur: same weights, different summation order than the library convolution
    for k in (3, 9, 19):
        np.testing.assert_allclose(A.GaussianBlur()(img, k).cpu().numpy(), g[f"blur/{k}"], atol=3e-6, rtol=0)
    assert A.GaussianBlur()(img, 0) is img
    # nearest-neighbour rotation: identical except where a source coordinate lands within fp32 rounding of x.5
    for a in (-20, -5, 10, 20):
        got = A.Rotate()(img, a).cpu().numpy()
        want = g[f"rotate/{a}"]
        assert got.shape == want.shape
        frac = float((got != want).mean())
        assert frac <= 0.005, (a, frac)
    assert torch.equal(A.Rotate()(img, 0), img)


def test_default_augmentation_list_runs_on_device():
    from wmar_b200 import augmentations as A
    torch.manual_seed(0)
    img = torch.rand(2, 3, 64, 64, device="cuda")
    for name, fn, params in A.default_augmentations():
        for p in params:
            out = fn(img, p).clamp(0, 1)
            assert out.shape == img.shape and out.is_cuda and torch.isfinite(out).all(), (name, p)


def test_round_trip_evaluation_on_device():
    """fill_batch_log + compute_metrics (generate.py:111-164, metrics.py:19-45) on a small Taming wrapper: layout of the
    log, metric definitions, and the watermark surviving the identity transforms."""
    from oracle import gpt as ogpt
    from oracle import vqgan as ov
    from wmar_b200.evaluate import compute_metrics, fill_batch_log
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    V, steps = 16384, 64
    gpt_cfg = dict(vocab_size=V, block_size=steps, n_layer=2, n_head=4, n_embd=256)
    dd = dict(ov.TAMING_CFG, ch=128, ch_mult=(1, 2), resolution=16, attn_resolutions=(8,))
    gw = ogpt.synthetic_gpt_weights(V, steps, 2, 4, 256, seed=7)
    vw = ov.synthetic_taming_vqgan_weights(dd, seed=8)
    state = {"transformer." + k: v for k, v in gw.items()}
    state.update({"first_stage_model." + k: v for k, v in vw.items()})
    model = TamingARMMWrapper(state_dict=state, gpt_cfg=gpt_cfg, dd_cfg=dd, device="cuda", max_batch=4)
    wm = create_watermarker_from_string(model.get_vq(), V, "linear-stratifiedrand-h=1-d=4.0-g=0.25", "cuda")
    model.set_watermarker(wm)
    torch.manual_seed(0)
    codes = model.sample([1, 9, 232, 975], {"temperature": 1.0, "top_k": 250, "top_p": 0.92}, apply_watermark=True)
    from wmar_b200 import augmentations as A
    augs = [("brightness", lambda x, b: A.Brightness()(x, b), [1, 1.5]), ("flip-h", lambda x, do: A.HorizontalFlip()(x) if do else x, [0, 1])]
    log = fill_batch_log({}, "wm", model, codes, {"max_roundtrips": 1, "augmentations": augs})
    assert set(log["wm"]) == {"roundtrips", "brightness", "flip-h"} and len(log["wm"]["roundtrips"]) == 2
    m = compute_metrics(log, "wm", wm)
    p0, m0 = m["roundtrips"][0]
    assert p0 == 0 and float(m0["l0"].max()) == 0.0 and torch.isinf(m0["psnr"]).all() and float(m0["pvalue"].max()) < 1e-6
    # brightness 1 and "no flip" re-encode the untouched image: same codes as round trip 1
    rt1 = log["wm"]["roundtrips"][1][1]
    assert torch.equal(log["wm"]["brightness"][0][1], rt1) and torch.equal(log["wm"]["flip-h"][0][1], rt1)
    # reference definition of l0 on one image (metrics.py:34)
    c, o = log["wm"]["flip-h"][1][1][0], codes[0]
    assert abs(float(m["flip-h"][1][1]["l0"][0]) - (o != c).sum().item() / o.shape[0]) < 1e-12


def test_evaluation_log_values_match_the_reference():
    """Value-level check of row f1: wmar_b200.evaluate.fill_batch_log + compute_metrics on the GPU (CUDA augmentations,
    device detector) against tests/golden/evallog.npz = the REFERENCE's fill_batch_log (generate.py:111-164) +
    compute_metric (metrics.py:25-45) + its own augmentation classes + GentimeWatermark.detect, run on the CPU over the
    same toy tokenizer (helpers.ToyTokenizerModel).  Codes and l0 bit-exact; PSNR to 1e-9 wherever the 8-bit images of
    the two sides are identical, else within 0.05 dB (blur / resize taps differ in the last float bit, which can move
    one of 3072 8-bit pixels); p-values to 1e-9 relative."""
    import os
    from helpers import G, ToyTokenizerModel
    from wmar_b200 import augmentations as A
    from wmar_b200.evaluate import compute_metrics, fill_batch_log
    from wmar_b200.watermarking import GentimeWatermark, SeedStrategy, SplitStrategy
    g = np.load(os.path.join(G, "evallog.npz"))
    model = ToyTokenizerModel("cuda")
    V = 64
    vq = {"alive_ids": torch.from_numpy(g["alive"]), "dead_ids": torch.from_numpy(g["dead"])}
    wm = GentimeWatermark(vq, V, SeedStrategy.LINEAR, SplitStrategy.RANDOM_STRATIFIED, 1, 2.0, 0.25, device="cuda")
    codes = torch.from_numpy(g["codes"]).cuda()
    augs = [("gaussian-blur", lambda x, k: A.GaussianBlur()(x, k), [0, 3, 7]),
            ("brightness", lambda x, b: A.Brightness()(x, b), [1, 1.5, 2.5]),
            ("flip-h", lambda x, do: A.HorizontalFlip()(x) if do else x, [0, 1]),
            ("upperleft-crop", lambda x, f: A.UpperLeftCropWithResizeBack()(x, f), [1.0, 0.8, 0.5])]
    log = fill_batch_log({}, "wm", model, codes, {"max_roundtrips": 2, "augmentations": augs})
    m = compute_metrics(log, "wm", wm)
    n_checked = 0
    for transform, entries in log["wm"].items():
        for j, (param, c, imgs, _) in enumerate(entries):
            assert float(param) == float(g[f"{transform}/{j}/param"])
            want_codes = g[f"{transform}/{j}/codes"]
            same_codes = np.array_equal(c.cpu().numpy(), want_codes)
            got = m[transform][j][1]
            if same_codes:
                np.testing.assert_array_equal(got["l0"].cpu().numpy(), g[f"{transform}/{j}/l0"])
                np.testing.assert_allclose(got["pvalue"].cpu().numpy(), g[f"{transform}/{j}/pvalue"], rtol=1e-9)
                n_checked += 1
            else:   # a block mean within float rounding of two colours: at most one code of the batch may differ
                assert (c.cpu().numpy() != want_codes).sum() <= 1, (transform, j)
            want_psnr, got_psnr = g[f"{transform}/{j}/psnr"], got["psnr"].cpu().numpy()
            for b in range(len(want_psnr)):
                if np.isinf(want_psnr[b]):
                    assert np.isinf(got_psnr[b]), (transform, j, b)
                else:
                    assert abs(got_psnr[b] - want_psnr[b]) <= 0.05, (transform, j, b, got_psnr[b], want_psnr[b])
    assert n_checked >= 12


def test_detect_host_batches_equals_plain_loop():
    """Bulk detection over pinned host batches (copy stream + two staging buffers) == the plain per-batch loop."""
    from oracle import gpt as ogpt
    from oracle import vqgan as ov
    from wmar_b200.evaluate import detect_host_batches
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    V = 16384
    gpt_cfg = dict(vocab_size=V, block_size=64, n_layer=1, n_head=4, n_embd=256)
    dd = dict(ov.TAMING_CFG, ch=128, ch_mult=(1, 2), resolution=32, attn_resolutions=(16,))
    state = {"transformer." + k: v for k, v in ogpt.synthetic_gpt_weights(V, 64, 1, 4, 256, seed=7).items()}
    state.update({"first_stage_model." + k: v for k, v in ov.synthetic_taming_vqgan_weights(dd, seed=8).items()})
    m = TamingARMMWrapper(state_dict=state, gpt_cfg=gpt_cfg, dd_cfg=dd, device="cuda", max_batch=4)
    wm = create_watermarker_from_string(m.get_vq(), V, "linear-stratifiedrand-h=1-d=2.0-g=0.25", "cuda")
    g = torch.Generator().manual_seed(3)
    batches = [(torch.rand(4, 3, 32, 32, generator=g) * 2 - 1).pin_memory() for _ in range(5)]
    batches.append((torch.rand(2, 3, 32, 32, generator=g) * 2 - 1).pin_memory())        # ragged tail
    got = detect_host_batches(m, wm, batches)
    assert got.shape == (6, 4, 4)
    for i, hb in enumerate(batches):
        st = wm.detect_stats(m.images_to_codes(hb.cuda()))
        for j, k in enumerate(("n_green", "n_scored", "z", "pvalue")):
            assert torch.equal(got[i, : hb.shape[0], j], st[k].double().cpu()), (i, k)
    assert detect_host_batches(m, wm, []).shape[0] == 0
