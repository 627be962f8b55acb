"""CPU: the plain-C oracle against fixtures produced by the real reference (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest

from oracle import wm

HERE = os.path.dirname(os.path.abspath(__file__))
ASSETS = os.path.join(os.path.dirname(HERE), "wmar_b200", "assets")
G = os.path.join(HERE, "golden")


def _assets(name):
    if name == "taming":
        return wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "vqgan_alive_ids.txt")), 16384) + (16384,)
    if name == "rar":
        return wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "rar_all_ids.txt")), 1024) + (1024,)
    if name == "chameleon":
        # chameleon_all_ids.txt = BPE ids 4..8195 and 16384..65535
        alive = list(range(4, 8196)) + list(range(16384, 65536))
        return wm.alive_dead(alive, 8192) + (65536,)
    raise KeyError(name)


def test_randperm_kat():
    kat = np.load(os.path.join(G, "randperm_kat.npz"))
    for key in kat.files:
        _, seed, n = key.split("_")
        np.testing.assert_array_equal(wm.randperm(int(seed), int(n)), kat[key])


def test_context_seed_bigint():
    for s in (0, 1, 16383, 65535 * 3, 2 ** 40, 2 ** 63 + 12345):
        assert wm.context_seed(s) == (wm.SALT * s) % (2 ** 64 - 1)


@pytest.mark.parametrize("model", ["taming", "rar", "chameleon"])
def test_greenlist_matches_reference(model):
    gl = np.load(os.path.join(G, "greenlist.npz"))
    alive, dead, V = _assets(model)
    for split in ("stratifiedrand", "rand"):
        for c in (0, 1, 5, 975, min(16383, V - 1)):
            seed = wm.context_seed(c)
            ids = wm.greenlist_ids(V, 0.25, split, alive, dead, seed)
            assert len(ids) == int(gl[f"{model}/{split}/ctx{c}/n"])
            np.testing.assert_array_equal(ids[:16], gl[f"{model}/{split}/ctx{c}/head"])
            bits = wm.greenlist_bitmask(V, 0.25, split, alive, dead, seed)
            np.testing.assert_array_equal(bits.view(np.uint8), gl[f"{model}/{split}/ctx{c}/bits"])
    bits = wm.greenlist_bitmask(V, 0.5, "stratifiedrand", alive, dead, 0)
    np.testing.assert_array_equal(bits.view(np.uint8), gl[f"{model}/fixed_g0.5/bits"])


def test_greenlist_sizes_survey_facts():
    alive, dead, V = _assets("taming")
    assert (len(alive), len(dead)) == (971, 15413)
    assert len(wm.greenlist_ids(V, 0.25, "stratifiedrand", alive, dead, 0)) == 242 + 3854
    alive, dead, V = _assets("rar")
    assert len(wm.greenlist_ids(V, 0.25, "stratifiedrand", alive, dead, 0)) == 256
    alive, dead, V = _assets("chameleon")
    assert len(dead) == 4 and len(wm.greenlist_ids(V, 0.25, "stratifiedrand", alive, dead, 0)) == 14336 + 4


CASES = {
    "taming_linear_h1": ("taming", "linear", "stratifiedrand", 1, 0.25),
    "taming_linear_h2": ("taming", "linear", "stratifiedrand", 2, 0.25),
    "taming_rand_h1": ("taming", "linear", "rand", 1, 0.5),
    "taming_spatial_h1": ("taming", "spatial", "stratifiedrand", 1, 0.25),
    "taming_spatial_h3": ("taming", "spatial", "stratifiedrand", 3, 0.25),
    "rar_linear_h1": ("rar", "linear", "stratifiedrand", 1, 0.25),
    "rar_fixed_h0": ("rar", "fixed", "stratifiedrand", 0, 0.25),
    "cham_fixed_h0": ("chameleon", "fixed", "stratifiedrand", 0, 0.25),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_detect_matches_reference(case):
    ops = np.load(os.path.join(G, "watermark_ops.npz"))
    model, ss, sp, h, gamma = CASES[case]
    alive, dead, V = _assets(model)
    codes = ops[f"{case}/codes"].astype(np.int64)
    ng, ns, masks = wm.detect_counts(codes, V, gamma, sp, ss, h, alive, dead, return_mask=True)
    p = wm.pvalue(ng, ns, gamma)
    np.testing.assert_allclose(p, ops[f"{case}/pvalues"], rtol=1e-12, atol=0)
    for b in range(codes.shape[0]):
        np.testing.assert_array_equal(np.asarray(masks[b], dtype=np.int8), ops[f"{case}/mask{b}"])


def test_detect_survey_pin():
    """torch.manual_seed(0); randint(0,16384,(2,256)) -> p = [0.93299..., 0.56710...]."""
    import torch
    torch.manual_seed(0)
    codes = torch.randint(0, 16384, (2, 256)).numpy()
    alive, dead, V = _assets("taming")
    ng, ns = wm.detect_counts(codes, V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    np.testing.assert_allclose(wm.pvalue(ng, ns, 0.25), [0.9329939855645604, 0.5671041022942398], rtol=1e-12)


def test_detect_too_short_raises():
    alive, dead, V = _assets("rar")
    with pytest.raises(ValueError):
        wm.detect_counts(np.zeros((1, 1), dtype=np.int64), V, 0.25, "stratifiedrand", "linear", 1, alive, dead)


@pytest.mark.parametrize("case", [c for c in sorted(CASES) if not c.startswith("cham")])
def test_process_logits_rows_match_reference(case):
    ops = np.load(os.path.join(G, "watermark_ops.npz"))
    model, ss, sp, h, gamma = CASES[case]
    alive, dead, V = _assets(model)
    rows = wm.GreenRows(V, gamma, sp, ss, h, alive, dead)
    for key in [k for k in ops.files if k.startswith(case + "/proc_t") and k.endswith("/past")]:
        past = ops[key].astype(np.int64)
        ref_bits = ops[key.replace("/past", "/bits")]
        got = rows(past)
        for b in range(past.shape[0]):
            want = ref_bits[b]
            if got[b] is None:
                assert not want.any(), f"{key} row {b}: reference watermarked a row the oracle skipped"
            else:
                np.testing.assert_array_equal(got[b].view(np.uint8), want)
