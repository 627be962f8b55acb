"""GPU parity: the RAR decode engine (CFG + watermark + sampler fused) vs the oracle and the reference-generated
goldens (tests/golden/rar.npz, produced by importing deps/rar/modeling/rar.py -- see oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import G, make_wm

pytestmark = pytest.mark.gpu


def _engine(name):
    from oracle import rar as orar
    from wmar_b200.models.rar_engine import RAREngine
    g = np.load(os.path.join(G, "rar.npz"))
    d, depth, heads, mlp, steps, B, seed = [int(x) for x in g[f"{name}/cfg"]]
    w = orar.synthetic_rar_weights(d, depth, heads, mlp, seed=seed)
    return g, w, RAREngine(w, depth, heads), (d, depth, heads, mlp, steps, B)


def _noise(seed, steps, B, V, skip_uniform):
    torch.manual_seed(seed)
    torch.rand(skip_uniform)  # RAR.preprocess_condition burns B uniforms first (rar.py:305)
    return torch.empty(steps, B, V).exponential_(1)


@pytest.mark.parametrize("name", ["tiny", "narrow"])
def test_rar_engine_matches_reference_golden(name):
    from oracle import rar as orar
    from wmar_b200 import _lib
    g, w, eng, (d, depth, heads, mlp, steps, B) = _engine(name)
    wm = make_wm("rar")
    cond = torch.from_numpy(g[f"{name}/cond"]).long()
    ids, logits = eng.sample(cond, steps, 4.0, 1.0, wm, greedy=True, return_logits=True)
    # guided logits of every step vs the oracle fed with the engine's own tokens
    o = orar.RAROracle(w, depth, heads)
    rows = torch.cat([cond + 1025, torch.full_like(cond, 2025)])
    ids_c = ids.cpu()
    for s in range(0, steps, 17):
        o.reset()
        for t in range(s + 1):
            last = torch.cat([ids_c[:, t - 1], ids_c[:, t - 1]]) if t > 0 else None
            lg = o.step(t, rows, last)
        want = lg[B:] + (lg[:B] - lg[B:]) * 4.0
        np.testing.assert_allclose(logits[s].cpu().numpy(), want.numpy(), rtol=2e-3, atol=5e-4, err_msg=f"step {s}")
        if s > 40:
            break
    np.testing.assert_array_equal(ids_c.numpy(), g[f"{name}/greedy_wm"])
    ids = eng.sample(cond, steps, 4.0, 1.0, wm, noise=_noise(3, steps, B, 1024, B).cuda())
    np.testing.assert_array_equal(ids.cpu().numpy(), g[f"{name}/sample_wm_seed3"])
    _lib.check(_lib.lib().wmar_check_device_flag(_lib.current_stream()))


def test_rar_engine_rows_independent_and_watermark_detectable():
    g, w, eng, (d, depth, heads, mlp, steps, B) = _engine("narrow")
    wm = make_wm("rar")
    cond = torch.from_numpy(g["narrow/cond"]).long()
    full = eng.sample(cond, steps, 4.0, 1.0, wm, greedy=True)
    part = eng.sample(cond[:3], steps, 4.0, 1.0, wm, greedy=True)
    assert torch.equal(full[:3], part)
    assert torch.equal(full, eng.sample(cond, steps, 4.0, 1.0, wm, greedy=True))
    st = wm.detect_stats(eng.sample(cond, steps, 4.0, 1.0, wm, seed=5))
    st0 = wm.detect_stats(eng.sample(cond, steps, 4.0, 1.0, None, seed=5))
    assert st["n_green"].float().mean() > st0["n_green"].float().mean() + 3


def test_rar_wrapper_surface():
    """RarARMMWrapper.sample / codes_to_images / images_to_codes shapes and ranges at reduced width (full depth of the
    API, not of the model)."""
    from wmar_b200.models import RarARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    m = RarARMMWrapper(rar_cfg=dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512),
                       max_batch=4)
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), "linear-stratifiedrand-h=1-d=2.0-g=0.25",
                                        "cuda")
    m.set_watermarker(wm)
    torch.manual_seed(0)
    codes = m.sample([1, 9, 232, 340, 975], None, apply_watermark=True)
    assert codes.shape == (5, 256) and codes.dtype == torch.int64 and int(codes.max()) < 1024
    imgs = m.codes_to_images(codes)
    assert imgs.shape == (5, 3, 256, 256) and float(imgs.min()) >= -1 and float(imgs.max()) <= 1
    back = m.images_to_codes(imgs)
    assert back.shape == (5, 256)
    p = wm.detect(codes)
    assert p.shape == (5,) and p.dtype == torch.float64
    with pytest.raises(AssertionError):
        m.codes_to_images(codes[:, :100])
