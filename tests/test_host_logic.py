"""CPU: host-side logic that mirrors the reference (batching / chunking / output tree, generate.py:79-108,179-207),
the C-ABI exports, and the product path's refusal to run without CUDA."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _reference_plan(all_inputs, batch_size, chunk_id, num_chunks):
    """literal restatement of the loop in generate.py:179-207"""
    batches = []
    for i in range(len(all_inputs) // batch_size):
        batches.append(all_inputs[i * batch_size:(i + 1) * batch_size])
    if len(all_inputs) % batch_size != 0:
        batches.append(all_inputs[(len(all_inputs) // batch_size) * batch_size:])
    base, out = {}, []
    for batch_idx, batch in enumerate(batches):
        ci = []
        for c in batch:
            if isinstance(c, tuple):
                c = c[0]
            if c not in base:
                base[c] = 0
            base[c] += 1
            ci.append(base[c])
        if batch_idx % num_chunks != chunk_id:
            continue
        out.append((batch_idx, batch, ci))
    return out


@pytest.mark.parametrize("n_per,bs,chunks", [(3, 4, 1), (5, 16, 2), (7, 3, 8), (1, 10, 4)])
def test_plan_batches_matches_reference_loop(n_per, bs, chunks):
    from wmar_b200.generate import expand_conditionings, plan_batches
    inputs = expand_conditionings("1,9,232,340,568", n_per)
    assert len(inputs) == 5 * n_per and inputs[:n_per] == [1] * n_per
    seen = []
    for cid in range(chunks):
        got = plan_batches(inputs, bs, cid, chunks)
        assert got == _reference_plan(inputs, bs, cid, chunks)
        seen += [(c, k) for _, b, ci in got for c, k in zip(b, ci)]
    # every (conditioning, running index) is produced exactly once across the chunks
    assert sorted(seen) == sorted((c, k + 1) for c in [1, 9, 232, 340, 568] for k in range(n_per))


def test_output_tree_names():
    from wmar_b200.generate import chw_to_uint8, output_paths
    png, npy, js = output_paths("o", 975, 3, "linear-stratifiedrand-h=1-d=2.0-g=0.25", orig_only=False)
    assert png == "o/c=975,idx=3/0003_linear-stratifiedrand-h=1-d=2.0-g=0.25_roundtrips_0.png"
    assert npy.endswith("_roundtrips_0.npy") and js.endswith("_roundtrips_0.json")
    png, npy, js = output_paths("o", (12, "a cat"), 1, "None", orig_only=True)
    assert png == "o/images/12:0001.png" and npy == "o/codes/12:0001.npy" and js is None
    img = np.stack([np.full((2, 2), -1.0), np.zeros((2, 2)), np.full((2, 2), 1.0)])
    u8 = chw_to_uint8(img)
    assert u8.shape == (2, 2, 3) and u8[0, 0].tolist() == [0, 128, 255]


def test_abi_exports_every_declared_symbol():
    """include/wmar_b200.h <-> libwmar_b200.so <-> wmar_b200/_lib.py agree (no compute calls: no GPU here)."""
    from wmar_b200 import _lib, build
    build.build()
    header = open(os.path.join(ROOT, "include", "wmar_b200.h")).read()
    declared = set(re.findall(r"\b(wmar_[a-z0-9_]+)\s*\(", header))
    declared -= {"wmar_status"}
    L = ctypes.CDLL(_lib.SO_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    _lib.lib()
    assert _lib.MISSING == []
    assert _lib.lib().wmar_version() >= 1


def test_no_cpu_fallback():
    from wmar_b200 import _lib
    from wmar_b200.models import RarARMMWrapper, TamingARMMWrapper
    from wmar_b200.watermarking import GentimeWatermark, SeedStrategy, SplitStrategy
    vq = {"alive_ids": torch.arange(8), "dead_ids": torch.arange(8, 16)}
    with pytest.raises(_lib.WmarError):
        GentimeWatermark(vq, 16, SeedStrategy.LINEAR, SplitStrategy.RANDOM_STRATIFIED, 1, 2.0, 0.25, device="cpu")
    with pytest.raises(_lib.WmarError):
        TamingARMMWrapper(device="cpu")
    with pytest.raises(_lib.WmarError):
        RarARMMWrapper(device="cpu")


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under wmar_b200/ may import or execute it"""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wmar_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), os.path.join(dirpath, f)
                assert "libwm_oracle" not in src


def test_state_module_keeps_reference_keys_and_applies_deltas(tmp_path):
    from wmar_b200.models.state import StateModule, update_weights
    flat = {"encoder.conv_in.weight": torch.randn(4, 3, 3, 3), "encoder.down.0.block.0.norm1.bias": torch.randn(4),
            "quantize.embedding.weight": torch.randn(8, 4)}
    m = StateModule(flat)
    assert set(m.state_dict()) == set(flat)
    assert set(m.encoder.state_dict()) == {"conv_in.weight", "down.0.block.0.norm1.bias"}
    delta = {"conv_in.weight": torch.ones(4, 3, 3, 3)}
    p = tmp_path / "enc_delta.pth"
    torch.save(delta, p)
    before = m.encoder.conv_in.weight.clone()
    update_weights(m.encoder, str(p))
    assert torch.allclose(m.encoder.conv_in.weight, before + 1.0)


def test_split_k_handoff_flags_are_unique_per_generation():
    """csrc/gemm.cuh: flag = (epoch + 1) * 1024 + salt with salt = 1-based launch index within the step.  The engines
    clear the workspace at the start of a generation, so within one generation no two launches may share a flag, no flag
    may be 0 (cleared memory) and all must fit 32 bits -- for the launch counts the three engines produce."""
    cases = {"taming": (256, 4 * 48 + 1), "rar_xl": (257, 5 * 32 + 2), "anole_7b": (2048, 4 * 32 + 1)}
    for name, (steps, launches) in cases.items():
        assert launches < 1024, name
        flags = {(epoch + 1) * 1024 + salt for epoch in range(steps) for salt in range(1, launches + 1)}
        assert len(flags) == steps * launches, name
        assert min(flags) > 0 and max(flags) < 2 ** 32, name


def test_psnr_and_l0_follow_the_reference_definitions():
    """wmar/utils/metrics.py:19-21,34 + utils.py:69-80 restated in numpy (images go through chw_to_pillow = clip + ROUND to
    uint8 before the PSNR) against wmar_b200.evaluate on CPU tensors, including values that sit on .5 boundaries."""
    import numpy as np
    import torch
    from wmar_b200.evaluate import psnr_uint8, to_uint8

    def ref_u8(x):                       # chw_to_pillow without the PIL container
        y = (255 * ((x.transpose(1, 2, 0) + 1.0) / 2.0)).clip(0, 255)
        return np.round(y).astype(np.uint8)

    g = torch.Generator().manual_seed(3)
    a = torch.rand(3, 3, 16, 16, generator=g) * 2.4 - 1.2          # some values outside [-1, 1]
    b = (a + 0.05 * torch.randn(3, 3, 16, 16, generator=g))
    a[0, 0, 0, :8] = torch.tensor([(2 * k + 1) / 255.0 - 1.0 for k in range(8)])   # exactly k + 0.5 after rescaling
    for i in range(3):
        np.testing.assert_array_equal(to_uint8(a[i:i + 1])[0].permute(1, 2, 0).numpy(), ref_u8(a[i].numpy()))
        ra, rb = ref_u8(a[i].numpy()), ref_u8(b[i].numpy())
        mse = np.mean((ra * 1.0 - rb * 1.0) ** 2)
        assert abs(float(psnr_uint8(a[i:i + 1], b[i:i + 1])[0]) - 10 * np.log10(255.0 ** 2 / mse)) < 1e-9


def test_full_mode_writes_the_reference_evaluation_tree(tmp_path):
    """generate.py:37-108,111-164 with stand-in model / watermarker objects on the CPU: one png + npy + json per
    (image, transform, parameter), the reference's names, metric keys and metric definitions; --orig_only tree."""
    import torch
    from wmar_b200.evaluate import fill_batch_log
    from wmar_b200.generate import save_batch_log

    class Model:                      # codes <-> images: a fixed invertible toy mapping (4 codes -> 2x2 image)
        def codes_to_images(self, codes):
            x = (codes.float() / 7.0 * 2.0 - 1.0).reshape(-1, 1, 2, 2).repeat(1, 3, 1, 1)
            return x.clamp(-1, 1)

        def images_to_codes(self, imgs):
            return torch.round((imgs[:, 0].reshape(-1, 4) + 1.0) / 2.0 * 7.0).long().clamp(0, 7)

    class WM:
        device = "cpu"

        def detect(self, codes):
            return (codes.sum(dim=1) % 5).double() / 10.0

        def __str__(self):
            return "linear-stratifiedrand-h=1-d=2.0-g=0.25"

    model, wm = Model(), WM()
    codes = torch.tensor([[0, 1, 2, 3], [7, 6, 5, 4], [3, 3, 3, 3]])
    augs = [("brightness", lambda x, b: x * b, [1, 2.0])]
    ev = {"metric_names": ["pvalue", "l0", "psnr", "bpp"], "augmentations": augs, "max_roundtrips": 1, "orig_only": False}
    batch = [1, (9, "a prompt"), 1]
    log = {"batch": batch}
    fill_batch_log(log, str(wm), model, codes, ev)
    save_batch_log(log, str(tmp_path), wm, ev, cond_indices=[1, 1, 2])
    stems = {(1, 1): "c=1,idx=1/0001", (9, 1): "c=9,idx=1/0001", (1, 2): "c=1,idx=2/0002"}
    expected = set()
    for stem in stems.values():
        for tp in ("roundtrips_0", "roundtrips_1", "brightness_1", "brightness_2.0"):
            for ext in ("png", "npy", "json"):
                expected.add(f"{stem}_{wm}_{tp}.{ext}")
    found = {os.path.relpath(os.path.join(d, f), tmp_path) for d, _, fs in os.walk(tmp_path) for f in fs}
    found = {f.replace(".png.npy", ".png") for f in found}     # save_png falls back to .npy without PIL
    assert found == expected
    j0 = json.load(open(tmp_path / f"c=9,idx=1/0001_{wm}_roundtrips_0.json"))
    assert list(j0) == ["pvalue", "l0", "psnr", "bpp"]
    assert j0["l0"] == 0.0 and j0["psnr"] == float("inf") and j0["bpp"] is None
    assert j0["pvalue"] == float(wm.detect(codes)[1])             # detector on the ORIGINAL codes (metrics.py:43)
    jb = json.load(open(tmp_path / f"c=1,idx=1/0001_{wm}_brightness_2.0.json"))
    re_codes = model.images_to_codes(((model.codes_to_images(codes) / 2 + 0.5) * 2.0).clamp(0, 1) * 2 - 1)
    assert jb["l0"] == float((re_codes[0] != codes[0]).sum()) / 4 and jb["pvalue"] == float(wm.detect(re_codes)[0])
    np.testing.assert_array_equal(np.load(tmp_path / f"c=1,idx=2/0002_{wm}_roundtrips_0.npy"), codes[2].numpy())
    # --orig_only: images/ + codes/ with "<conditioning>:<idx>" names, no metrics
    ev0 = {"metric_names": [], "augmentations": [], "max_roundtrips": 0, "orig_only": True}
    log0 = {"batch": batch}
    fill_batch_log(log0, str(wm), model, codes, ev0)
    out0 = tmp_path / "orig"
    save_batch_log(log0, str(out0), wm, ev0, cond_indices=[1, 1, 2])
    assert sorted(os.listdir(out0 / "codes")) == ["1:0001.npy", "1:0002.npy", "9:0001.npy"]
    assert len(os.listdir(out0 / "images")) == 3


def test_chameleon_alive_ids_asset_is_the_reference_key():
    """The stratified split permutes len(alive) ids, so the alive list is part of the watermark key.  The reference reads
    assets/chameleon_all_ids.txt (chameleon_wrapper.py:32) with vq.n_e = 8192 (armm_wrapper.py:42-55): 57344 alive BPE
    ids = [4, 8196) and [16384, 65536), dead = set(range(8192)) - alive = {0, 1, 2, 3}."""
    from wmar_b200.models.armm_wrapper import load_ids
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    alive = load_ids(os.path.join(here, "wmar_b200", "assets", "chameleon_all_ids.txt"))
    assert alive == list(range(4, 8196)) + list(range(16384, 65536))
    assert sorted(set(range(8192)) - set(alive)) == [0, 1, 2, 3]
    src = open(os.path.join(here, "wmar_b200", "models", "chameleon_wrapper.py")).read()
    assert "init_alivecodes" in src and "torch.arange(IMAGE_TOKEN_LO" not in src


def test_clustering_split_matches_the_reference_golden():
    """CLUSTERING split (gentime_watermark.py:175-216): the host routine reproduces the greenlist the reference's own
    GentimeWatermark built from the same codebook (tests/golden/clustering.npz, oracle/gen_golden_clustering.py), id for
    id and in the same order; the bitmask row it becomes on the device has exactly those bits."""
    from wmar_b200.watermarking.clustering import clustering_greenlist_ids, ids_to_bitmask_row
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "clustering.npz"))
    green = clustering_greenlist_ids(torch.from_numpy(g["emb"]), g["alive"], g["dead"])
    np.testing.assert_array_equal(np.asarray(green, dtype=np.int64), g["green"])
    V = g["emb"].shape[0]
    row = ids_to_bitmask_row(green, V)
    bits = np.unpackbits(row.view(np.uint8), bitorder="little")[:V]
    assert sorted(np.nonzero(bits)[0].tolist()) == sorted(g["green"].tolist())


def test_precompute_imagenet_codes_host_logic(tmp_path):
    """The bulk tokeniser (precompute_imagenet_codes.py equivalent): selection with the reference's numpy calls under
    seed 1, the reference's file names, batching through images_to_codes -- with a stand-in model on the CPU."""
    from PIL import Image
    from wmar_b200 import precompute_imagenet_codes as pc
    root = tmp_path / "imagenet"
    wnids = ["n01440764", "n01443537", "n15075141"]
    (root / "train").mkdir(parents=True)
    (root / "labels.txt").write_text("".join(f"{w},name{i}\n" for i, w in enumerate(wnids)))
    rng = np.random.default_rng(0)
    for w in wnids:
        (root / "train" / w).mkdir()
        for k in range(5):
            arr = rng.integers(0, 255, size=(40 + 3 * k, 48, 3), dtype=np.uint8)
            Image.fromarray(arr).save(root / "train" / w / f"{w}_{k}.JPEG")
    (tmp_path / "idx.json").write_text(json.dumps({"0": [wnids[0], "a"], "7": [wnids[1], "b"], "999": [wnids[2], "c"]}))

    class StandIn:
        device = "cpu"
        calls = []

        def images_to_codes(self, x):
            assert x.shape[1:] == (3, 32, 32) and float(x.min()) >= -1 and float(x.max()) <= 1
            self.calls.append(x.shape[0])
            return (x.flatten(1)[:, :16] * 100).long()

    labels = pc.load_labels(str(root))
    assert labels == wnids
    np.random.seed(1)
    want = {}
    for w in wnids:                                   # literal restatement of precompute_imagenet_codes.py:74-82
        cls_paths = [os.path.join(str(root), "train", w, p) for p in os.listdir(os.path.join(str(root), "train", w))]
        want[w] = np.random.choice(cls_paths, size=3, replace=False)
        np.random.shuffle(want[w])
    np.random.seed(1)
    got = pc.select_paths(str(root), labels, {w: 3 for w in wnids})
    assert all(list(got[w]) == list(want[w]) for w in wnids)

    m = StandIn()
    np.random.seed(1)
    torch.manual_seed(1)
    out = tmp_path / "out"
    # monkeypatched counts: 3 per class instead of 50
    orig = pc.counts_per_label
    pc.counts_per_label = lambda labels, size, split=None: {w: 3 for w in labels}
    try:
        n = pc.run(m, str(root), str(out), 32, batch_size=2, classes={0, 999}, max_per_class=None,
                   class_index_path=str(tmp_path / "idx.json"), log=lambda s: None)
    finally:
        pc.counts_per_label = orig
    assert n == 6 and m.calls == [2, 1, 2, 1]
    assert sorted(os.listdir(out / "codes")) == [f"{c}:{k:04}.npy" for c in (0, 999) for k in range(3)]
    assert sorted(os.listdir(out / "images")) == [f"{c}:{k:04}.png" for c in (0, 999) for k in range(3)]
    assert np.load(out / "codes" / "0:0000.npy").shape == (16,)


def test_group_batches_preserves_order_and_content():
    """generate --lanes: consecutive planned batches are grouped for one wrapper call; flattening the groups gives back the
    planned batches in order (so the wrapper's chunking by max_batch == batch_size reproduces them)."""
    from wmar_b200.generate import expand_conditionings, group_batches, plan_batches
    inputs = expand_conditionings(",".join(str(i) for i in range(23)), 2)
    for bs, chunks in ((16, 1), (5, 2), (7, 3)):
        for chunk_id in range(chunks):
            planned = list(plan_batches(inputs, bs, chunk_id, chunks))
            for lanes in (1, 2, 3):
                groups = list(group_batches(iter(planned), lanes))
                assert [b for g in groups for b in g] == planned
                assert all(len(g) == lanes for g in groups[:-1]) and 1 <= len(groups[-1]) <= lanes if groups else True
                flat = [c for g in groups for _, batch, _ in g for c in batch]
                assert flat == [c for _, batch, _ in planned for c in batch]
