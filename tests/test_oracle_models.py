"""CPU: the torch-fp32 restatements (oracle/gpt.py, rar.py, vqgan.py) against fixtures from the real reference."""
import os

import numpy as np
import pytest
import torch

from oracle import gpt as ogpt
from oracle import rar as orar
from oracle import vqgan as ov
from oracle import wm

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(os.path.dirname(HERE), "wmar_b200", "assets")
G = os.path.join(HERE, "golden")


def predraw_noise(seed, steps, B, V, skip_uniform=0):
    """torch.multinomial draws q = empty(B,V).exponential_(1) once per step from the default CPU generator;
    one [steps,B,V] draw from the same seed is the same stream (re-checked by these tests).
    RAR.generate first burns B uniforms in preprocess_condition (rar.py:305, torch.rand_like)."""
    torch.manual_seed(seed)
    if skip_uniform:
        torch.rand(skip_uniform)
    return torch.empty(steps, B, V).exponential_(1)


@pytest.mark.parametrize("name", ["tiny", "narrow"])
def test_gpt_oracle_matches_reference(name):
    g = np.load(os.path.join(G, "gpt.npz"))
    V, block, L, H, d, steps, B, seed = [int(x) for x in g[f"{name}/cfg"]]
    w = ogpt.synthetic_gpt_weights(V, block, L, H, d, seed=seed)
    o = ogpt.GPTOracle(w, L, H)
    cond = torch.from_numpy(g[f"{name}/cond"]).long()
    np.testing.assert_allclose(o.step(cond, 0)[:, :64].numpy(), g[f"{name}/logits0_head"], rtol=2e-4, atol=2e-5)
    alive, dead = wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "vqgan_alive_ids.txt")), V)
    rows = wm.GreenRows(V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    codes = ogpt.sample_with_past(o, cond, steps, 1.0, 250, 0.92, rows, 2.0, greedy=True)
    np.testing.assert_array_equal(codes.numpy(), g[f"{name}/greedy_wm"])
    codes = ogpt.sample_with_past(o, cond, steps, 1.0, 250, 0.92, rows, 2.0, noise=predraw_noise(1, steps, B, V))
    np.testing.assert_array_equal(codes.numpy(), g[f"{name}/sample_wm_seed1"])
    codes = ogpt.sample_with_past(o, cond, steps, 0.8, 600, 0.5, None, 0.0, noise=predraw_noise(2, steps, B, V))
    np.testing.assert_array_equal(codes.numpy(), g[f"{name}/sample_nowm_seed2"])


@pytest.mark.parametrize("name", ["tiny", "narrow"])
def test_rar_oracle_matches_reference(name):
    g = np.load(os.path.join(G, "rar.npz"))
    d, depth, heads, mlp, steps, B, seed = [int(x) for x in g[f"{name}/cfg"]]
    w = orar.synthetic_rar_weights(d, depth, heads, mlp, seed=seed)
    o = orar.RAROracle(w, depth, heads)
    cond = torch.from_numpy(g[f"{name}/cond"]).long()
    alive, dead = wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "rar_all_ids.txt")), 1024)
    rows = wm.GreenRows(1024, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    ids = orar.generate(o, cond, steps, 4.0, 1.0, rows, 2.0, greedy=True)
    np.testing.assert_array_equal(ids.numpy(), g[f"{name}/greedy_wm"])
    ids = orar.generate(o, cond, steps, 4.0, 1.0, rows, 2.0, noise=predraw_noise(3, steps, B, 1024, skip_uniform=B))
    np.testing.assert_array_equal(ids.numpy(), g[f"{name}/sample_wm_seed3"])


@pytest.mark.parametrize("name", ["taming_small", "maskgit_small"])
def test_vqgan_oracle_matches_reference_small(name):
    g = np.load(os.path.join(G, "vqgan.npz"))
    if name.startswith("taming"):
        cfg = dict(ov.TAMING_CFG, ch=32, ch_mult=(1, 2, 2), resolution=64, z_channels=64, n_embed=512, embed_dim=64)
        w = ov.synthetic_taming_vqgan_weights(cfg, seed=3)
        gen = torch.Generator().manual_seed(11)
        enc, dec = ov.taming_images_to_codes, ov.taming_codes_to_images
        n_codes = cfg["n_embed"]
    else:
        cfg = dict(ov.MASKGIT_CFG, hidden_channels=32, channel_mult=(1, 2, 2), resolution=64, z_channels=64,
                   num_embeddings=256)
        w = ov.synthetic_maskgit_weights(cfg, seed=5)
        gen = torch.Generator().manual_seed(13)
        enc, dec = ov.rar_images_to_codes, ov.rar_codes_to_images
        n_codes = cfg["num_embeddings"]
    img = torch.rand(2, 3, 64, 64, generator=gen) * 2 - 1
    codes_in = torch.randint(0, n_codes, (2, 16 * 16), generator=gen)
    np.testing.assert_array_equal(codes_in.numpy(), g[f"{name}/codes_in"])
    np.testing.assert_array_equal(enc(img, w, cfg).numpy(), g[f"{name}/codes"])
    rec = dec(codes_in, w, cfg)
    np.testing.assert_allclose(rec[:, :, ::2, ::2].numpy(), g[f"{name}/rec_sub"], atol=2e-5)


def test_config1_roundtrip_plumbing():
    """BASELINE config 1: Taming VQGAN 256x256 encode -> decode round trip, batch 1, CPU, through the oracle."""
    g = np.load(os.path.join(G, "vqgan.npz"))
    w = ov.synthetic_taming_vqgan_weights(ov.TAMING_CFG, seed=3)
    gen = torch.Generator().manual_seed(11)
    img = torch.rand(1, 3, 256, 256, generator=gen) * 2 - 1
    codes = ov.taming_images_to_codes(img, w)
    assert codes.shape == (1, 256) and codes.dtype == torch.int64
    np.testing.assert_array_equal(codes.numpy(), g["taming_full/codes"])
    rec = ov.taming_codes_to_images(codes, w)
    assert rec.shape == (1, 3, 256, 256) and float(rec.abs().max()) <= 1.0
    codes2 = ov.taming_images_to_codes(rec, w)
    assert codes2.shape == (1, 256)
