"""GPU parity: greenlist engine, logit processor, fused sampler and detector vs the oracle and reference goldens.
Everything goes through the C-ABI (wmar_b200/_lib.py -> libwmar_b200.so)."""
import os

import numpy as np
import pytest
import torch

from helpers import G, assets, make_wm, unpack_rows

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["taming", "rar", "chameleon"])
def test_greenlist_table_matches_reference(model):
    from oracle import wm as owm
    gl = np.load(os.path.join(G, "greenlist.npz"))
    alive, dead, V = assets(model)
    for split in ("stratifiedrand", "rand"):
        if model == "chameleon":
            n_rows = 1024  # a prefix of the table is enough (rows are independent)
            w = make_wm(model, split=split)
            tab = w.table[:n_rows].cpu().numpy()
        else:
            w = make_wm(model, split=split)
            tab = w.table.cpu().numpy()
        for c in (0, 1, 5, 975, min(16383, V - 1)):
            if c >= tab.shape[0]:
                continue
            np.testing.assert_array_equal(tab[c].view(np.uint8), gl[f"{model}/{split}/ctx{c}/bits"])
        # device build == host build of the product library == oracle, on a sample of rows
        host = make_wm(model, split=split, build_on="host")
        rows = [0, 1, 2, 77, tab.shape[0] - 1]
        np.testing.assert_array_equal(host.table[rows].cpu().numpy(), tab[rows])
        o = owm.greenlist_bitmask(V, 0.25, split, alive, dead, owm.context_seed(77))
        np.testing.assert_array_equal(tab[77].view(np.uint32), o)
    w = make_wm(model, seed_strategy="fixed", h=0, gamma=0.5)
    np.testing.assert_array_equal(w.table[0].cpu().numpy().view(np.uint8), gl[f"{model}/fixed_g0.5/bits"])


def test_greenlist_full_table_device_equals_host():
    dev = make_wm("taming")
    host = make_wm("taming", build_on="host")
    assert torch.equal(dev.table, host.table)
    pop = unpack_rows(dev.table[:64].cpu().numpy(), 16384).sum(axis=1)
    assert (pop == 4096).all()


CASES = {
    "taming_linear_h1": ("taming", "linear", "stratifiedrand", 1, 2.0, 0.25),
    "taming_linear_h2": ("taming", "linear", "stratifiedrand", 2, 2.0, 0.25),
    "taming_rand_h1": ("taming", "linear", "rand", 1, 4.0, 0.5),
    "taming_spatial_h1": ("taming", "spatial", "stratifiedrand", 1, 2.0, 0.25),
    "taming_spatial_h3": ("taming", "spatial", "stratifiedrand", 3, 2.0, 0.25),
    "rar_linear_h1": ("rar", "linear", "stratifiedrand", 1, 2.0, 0.25),
    "rar_fixed_h0": ("rar", "fixed", "stratifiedrand", 0, 2.0, 0.25),
    "cham_fixed_h0": ("chameleon", "fixed", "stratifiedrand", 0, 2.0, 0.25),
}


@pytest.mark.parametrize("case", [c for c in sorted(CASES) if not c.startswith("cham")])
def test_logit_processor_matches_reference(case):
    ops = np.load(os.path.join(G, "watermark_ops.npz"))
    model, ss, sp, h, delta, gamma = CASES[case]
    w = make_wm(model, ss, sp, h, delta, gamma)
    proc = w.spawn_logit_processor()
    V = w.vocab_size
    for key in [k for k in ops.files if k.startswith(case + "/proc_t") and k.endswith("/past")]:
        past = torch.from_numpy(ops[key].astype(np.int64)).cuda()
        want = np.unpackbits(ops[key.replace("/past", "/bits")], axis=-1, bitorder="little")[:, :V].astype(bool)
        logits = torch.randn(past.shape[0], V, device="cuda")
        before = logits.clone()
        out = proc(past_ids=past, logits=logits)
        assert out.data_ptr() == logits.data_ptr()  # in place, like the reference
        diff = (out - before).cpu().numpy()
        assert ((diff != 0) == want).all(), key
        np.testing.assert_allclose(diff[want], delta, atol=1e-5)


@pytest.mark.parametrize("case", sorted(CASES))
def test_detect_matches_reference(case):
    from oracle import wm as owm
    ops = np.load(os.path.join(G, "watermark_ops.npz"))
    model, ss, sp, h, delta, gamma = CASES[case]
    if model == "chameleon":
        pytest.skip("65536-row table not needed: FIXED seeding has one row")
    w = make_wm(model, ss, sp, h, delta, gamma)
    codes = torch.from_numpy(ops[f"{case}/codes"].astype(np.int64)).cuda()
    st = w.detect_stats(codes, return_masks=True)
    alive, dead, V = assets(model)
    ng, ns, masks = owm.detect_counts(codes.cpu().numpy(), V, gamma, sp, ss, h, alive, dead, return_mask=True)
    np.testing.assert_array_equal(st["n_green"].cpu().numpy(), ng)
    np.testing.assert_array_equal(st["n_scored"].cpu().numpy(), ns)
    np.testing.assert_allclose(st["pvalue"].cpu().numpy(), ops[f"{case}/pvalues"], rtol=1e-9, atol=1e-300)
    np.testing.assert_allclose(st["z"].cpu().numpy(), owm.zscore(ng, ns, gamma), rtol=1e-12)
    for b in range(codes.shape[0]):
        assert st["masks"][b] == ops[f"{case}/mask{b}"].tolist() == masks[b]
    p2 = w.detect(codes)
    assert p2.dtype == torch.float64 and p2.shape == (codes.shape[0],)


def test_detect_chameleon_fixed():
    ops = np.load(os.path.join(G, "watermark_ops.npz"))
    w = make_wm("chameleon", "fixed", "stratifiedrand", 0, 2.0, 0.25)
    codes = torch.from_numpy(ops["cham_fixed_h0/codes"].astype(np.int64)).cuda()
    np.testing.assert_allclose(w.detect(codes).cpu().numpy(), ops["cham_fixed_h0/pvalues"], rtol=1e-9)


def test_detect_errors_and_extremes():
    w = make_wm("rar")
    with pytest.raises(ValueError):
        w.detect(torch.zeros((2, 1), dtype=torch.long, device="cuda"))
    # all-green passage: tiny p-value must match scipy's betainc in log space
    from oracle import wm as owm
    row5 = w.greenlist_ids_for_sum(5)
    codes = torch.full((1, 256), 5, dtype=torch.long, device="cuda")
    codes[0, 1::2] = row5[:128]
    st = w.detect_stats(codes)
    alive, dead, V = assets("rar")
    ng, ns = owm.detect_counts(codes.cpu().numpy(), V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    assert int(st["n_green"][0]) == int(ng[0]) and int(st["n_scored"][0]) == int(ns[0])
    np.testing.assert_allclose(st["pvalue"].cpu().numpy(), owm.pvalue(ng, ns, 0.25), rtol=1e-9, atol=1e-300)


@pytest.mark.parametrize("cfg", [
    dict(model="taming", T=1.0, top_k=250, top_p=0.92, wm=True),
    dict(model="taming", T=0.8, top_k=600, top_p=0.5, wm=False),
    dict(model="taming", T=1.3, top_k=None, top_p=0.9, wm=True),
    dict(model="taming", T=1.0, top_k=250, top_p=None, wm=True),
    dict(model="rar", T=1.0, top_k=None, top_p=None, wm=True),
    dict(model="rar", T=0.7, top_k=40, top_p=0.95, wm=True),
])
@pytest.mark.parametrize("greedy", [False, True])
def test_fused_sampler_matches_oracle(cfg, greedy):
    import ctypes
    from oracle import sampling, wm as owm
    from wmar_b200 import _lib
    alive, dead, V = assets(cfg["model"])
    w = make_wm(cfg["model"])
    rows_fn = owm.GreenRows(V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    g = torch.Generator().manual_seed(5)
    B = 16
    for t in (0, 1, 7):
        logits = torch.randn(B, V, generator=g) * 2.0
        past = torch.randint(0, V, (B, t), generator=g)
        noise = torch.empty(B, V).exponential_(1, generator=g)
        want = sampling.sample_step(logits, rows_fn(past.numpy()) if cfg["wm"] else None, 2.0, cfg["T"], cfg["top_k"],
                                    cfg["top_p"], noise, greedy=greedy)
        sp = _lib.SampleParams(cfg["T"], cfg["top_k"] or 0, cfg["top_p"] or 0.0, 1 if greedy else 0, 0)
        wmp = w.c_params() if cfg["wm"] else _lib.WmParams(None, 0, V, 0, 0, 16, 0.0, 0.0)
        out = torch.empty(B, dtype=torch.long, device="cuda")
        lg, ps, nz = logits.cuda(), past.cuda().contiguous(), noise.cuda()
        _lib.check(_lib.lib().wmar_wm_sample(ctypes.byref(wmp), ctypes.byref(sp), _lib.ptr(ps) if t else None, B, t,
                                             t, _lib.ptr(lg), _lib.ptr(nz), _lib.ptr(out), _lib.current_stream()))
        _lib.check(_lib.lib().wmar_check_device_flag(_lib.current_stream()))
        np.testing.assert_array_equal(out.cpu().numpy(), want.numpy(), err_msg=f"t={t}")


def test_sampler_philox_distribution():
    """Without pre-drawn noise the kernel draws q itself: the empirical distribution must follow softmax(logits)."""
    import ctypes
    from wmar_b200 import _lib
    V = 1024
    logits = torch.zeros(1, V)
    logits[0, :4] = torch.tensor([3.0, 2.0, 1.0, 0.0])
    logits[0, 4:] = -30.0
    B = 4096
    lg = logits.expand(B, V).contiguous().cuda()
    out = torch.empty(B, dtype=torch.long, device="cuda")
    sp = _lib.SampleParams(1.0, 0, 0.0, 0, 1234)
    wmp = _lib.WmParams(None, 0, V, 0, 0, 16, 0.0, 0.0)
    _lib.check(_lib.lib().wmar_wm_sample(ctypes.byref(wmp), ctypes.byref(sp), None, B, 0, 0, _lib.ptr(lg), None,
                                         _lib.ptr(out), _lib.current_stream()))
    freq = torch.bincount(out.cpu(), minlength=V)[:4].double() / B
    want = torch.softmax(logits[0, :4].double(), 0)
    assert torch.allclose(freq, want, atol=0.03), (freq, want)


def test_clustering_split_on_device_matches_reference_golden():
    """GentimeWatermark(FIXED, CLUSTERING): the one-row device table carries exactly the reference's greenlist
    (tests/golden/clustering.npz), the logit processor adds delta on those ids only, and linear seeding is refused like
    in the reference (gentime_watermark.py:177)."""
    from wmar_b200.watermarking import GentimeWatermark, SeedStrategy, SplitStrategy
    g = np.load(os.path.join(G, "clustering.npz"))
    V = g["emb"].shape[0]
    vq = {"alive_ids": torch.from_numpy(g["alive"]), "dead_ids": torch.from_numpy(g["dead"]),
          "embedding": torch.from_numpy(g["emb"])}
    wm = GentimeWatermark(vq, V, SeedStrategy.FIXED, SplitStrategy.CLUSTERING, 0, 2.0, 0.5, device="cuda")
    assert str(wm) == "fixed-clustering-h=0-d=2.0-g=0.50"
    assert sorted(wm.greenlist_ids_for_sum(0).cpu().tolist()) == sorted(g["green"].tolist())
    logits = torch.zeros(2, V, device="cuda")
    out = wm._process_logits(torch.zeros(2, 3, dtype=torch.long, device="cuda"), logits)
    want = torch.zeros(V)
    want[torch.from_numpy(g["green"])] = 2.0
    assert torch.equal(out[0].cpu(), want) and torch.equal(out[1].cpu(), want)
    with pytest.raises(AssertionError):
        GentimeWatermark(vq, V, SeedStrategy.LINEAR, SplitStrategy.CLUSTERING, 1, 2.0, 0.5, device="cuda")


@pytest.mark.parametrize("rows,rowlen", [(16, 16384), (8, 1024), (5, 65536), (3, 1000), (16, 100000)])
def test_in_kernel_torch_philox_stream_is_bit_identical(rows, rowlen):
    """rng_mode 1 of the sampler replicates torch's CUDA generator: the Exp(1) tensor that `torch.multinomial` divides by
    (`empty_like(probs).exponential_(1)`, mingpt.py:363 / rar.py:454) is reproduced element by element, bit for bit, from
    (seed, offset) -- including the generator's offset bookkeeping over consecutive calls and tensors large enough for
    the grid-stride loop to use all four values of a Philox call."""
    from wmar_b200 import _lib
    from wmar_b200.models.armm_wrapper import AutoregressiveMultimodalModelWrapper

    class W(AutoregressiveMultimodalModelWrapper):
        device = torch.device("cuda", torch.cuda.current_device())

    torch.manual_seed(1234 + rows)
    torch.rand(7, device="cuda")                       # some earlier use of the generator
    gen = torch.cuda.default_generators[torch.cuda.current_device()]
    st = gen.get_state()
    steps = 3
    want = [torch.empty(rows, rowlen, device="cuda").exponential_(1) for _ in range(steps)]
    end_offset = gen.get_offset()
    gen.set_state(st)
    p = W()._torch_stream(steps, rows, rowlen)
    assert gen.get_offset() == end_offset              # the generator is advanced exactly like torch would have
    iters = (p["torch_numel"] - 1) // (p["torch_threads"] * 4) + 1
    for s in range(steps):
        got = torch.empty(rows, rowlen, device="cuda")
        _lib.check(_lib.lib().wmar_debug_torch_exponential(p["seed"], p["torch_offset"] + 4 * iters * s, rows, rowlen,
                                                           p["torch_threads"], _lib.ptr(got), _lib.current_stream()))
        assert torch.equal(got, want[s]), (s, (got != want[s]).sum().item())


def test_wrapper_sampling_equals_torch_buffer_mode():
    """TamingARMMWrapper / RarARMMWrapper: rng="torch" (stream drawn inside the sampler) and rng="torch_buffer" (the
    same stream pre-drawn by torch, the round-1 path pinned to the reference goldens) give identical codes under the same
    seed, and leave torch's generator in the same state."""
    from oracle import gpt as ogpt
    from oracle import vqgan as ov
    from wmar_b200.models import RarARMMWrapper, TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    V, steps = 16384, 64
    gpt_cfg = dict(vocab_size=V, block_size=steps, n_layer=2, n_head=4, n_embd=256)
    dd = dict(ov.TAMING_CFG, ch=128, ch_mult=(1, 2), resolution=16, attn_resolutions=(8,))
    gw = ogpt.synthetic_gpt_weights(V, steps, 2, 4, 256, seed=7)
    vw = ov.synthetic_taming_vqgan_weights(dd, seed=8)
    state = {"transformer." + k: v for k, v in gw.items()}
    state.update({"first_stage_model." + k: v for k, v in vw.items()})
    gp = {"temperature": 1.0, "top_k": 250, "top_p": 0.92}
    out, tail = {}, {}
    for rng in ("torch", "torch_buffer"):
        m = TamingARMMWrapper(state_dict=state, gpt_cfg=gpt_cfg, dd_cfg=dd, device="cuda", max_batch=4, rng=rng)
        wm = create_watermarker_from_string(m.get_vq(), V, "linear-stratifiedrand-h=1-d=2.0-g=0.25", "cuda")
        m.set_watermarker(wm)
        torch.manual_seed(11)
        out[rng] = m.sample([1, 9, 232, 975, 3, 4], gp, apply_watermark=True).cpu()      # two engine calls (4 + 2 rows)
        tail[rng] = torch.rand(4, device="cuda").cpu()
    assert torch.equal(out["torch"], out["torch_buffer"]) and torch.equal(tail["torch"], tail["torch_buffer"])
    out = {}
    for rng in ("torch", "torch_buffer"):
        m = RarARMMWrapper(rar_cfg=dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512),
                           max_batch=4, rng=rng)
        torch.manual_seed(12)
        out[rng] = m.sample([1, 9, 232, 340, 975], None, apply_watermark=False).cpu()
    assert torch.equal(out["torch"], out["torch_buffer"])


def test_wrapper_lanes_equal_sequential_chunk_loop():
    """TamingARMMWrapper.sample with more conditionings than max_batch: chunk i runs on engine lane i % lanes (own KV
    cache / scratch / step graph, shared weights) on that lane's CUDA stream.  The codes must equal the sequential chunk
    loop (lanes=1) bit for bit -- sampled (torch's Philox stream, offsets assigned per chunk in order) and greedy -- on
    repeated calls (lanes are reused), and torch's generator must end in the same state."""
    from oracle import gpt as ogpt
    from oracle import vqgan as ov
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    V, steps = 16384, 64
    gpt_cfg = dict(vocab_size=V, block_size=steps, n_layer=2, n_head=4, n_embd=256)
    dd = dict(ov.TAMING_CFG, ch=128, ch_mult=(1, 2), resolution=16, attn_resolutions=(8,))
    gw = ogpt.synthetic_gpt_weights(V, steps, 2, 4, 256, seed=7)
    vw = ov.synthetic_taming_vqgan_weights(dd, seed=8)
    state = {"transformer." + k: v for k, v in gw.items()}
    state.update({"first_stage_model." + k: v for k, v in vw.items()})
    gp = {"temperature": 1.0, "top_k": 250, "top_p": 0.92}
    conds = [1, 9, 232, 975, 3, 4, 77, 500, 640, 12]                    # 3 chunks of max_batch = 4 (4 + 4 + 2 rows)
    res = {}
    for lanes in (1, 2, 3):
        m = TamingARMMWrapper(state_dict=state, gpt_cfg=gpt_cfg, dd_cfg=dd, device="cuda", max_batch=4, lanes=lanes)
        wm = create_watermarker_from_string(m.get_vq(), V, "linear-stratifiedrand-h=1-d=2.0-g=0.25", "cuda")
        m.set_watermarker(wm)
        torch.manual_seed(21)
        a = m.sample(conds, gp, apply_watermark=True)
        b = m.sample(conds, gp, apply_watermark=True)                   # second call: lanes reused, generator advanced
        g = m.sample(conds, gp, apply_watermark=False, greedy=True)
        img = m.codes_to_images(a)                                      # 10 rows > max_batch: chunked tokenizer calls
        back = m.images_to_codes(img)
        res[lanes] = (a.cpu(), b.cpu(), g.cpu(), torch.rand(4, device="cuda").cpu(), back.cpu())
        assert img.shape == (10, 3, 16, 16) and back.shape == a.shape
    for lanes in (2, 3):
        for x, y in zip(res[1], res[lanes]):
            assert torch.equal(x, y), lanes
    assert not torch.equal(res[1][0], res[1][1])


def test_rar_wrapper_lanes_equal_sequential_chunk_loop():
    """RarARMMWrapper: same property as the Taming wrapper's lanes (chunks of max_batch on concurrent engine lanes ==
    the sequential chunk loop, bit for bit, including the torch.rand_like draw RAR makes per chunk)."""
    from wmar_b200.models import RarARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    res = {}
    conds = [1, 9, 232, 340, 975, 17, 600, 3, 44]                       # 3 chunks of max_batch = 4 (4 + 4 + 1 rows)
    for lanes in (1, 2):
        m = RarARMMWrapper(rar_cfg=dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512),
                           max_batch=4, lanes=lanes)
        wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), "linear-stratifiedrand-h=1-d=2.0-g=0.25", "cuda")
        m.set_watermarker(wm)
        torch.manual_seed(12)
        a = m.sample(conds, None, apply_watermark=True)
        b = m.sample(conds, None, apply_watermark=False, greedy=True)
        res[lanes] = (a.cpu(), b.cpu(), torch.rand(4, device="cuda").cpu())
    for x, y in zip(res[1], res[2]):
        assert torch.equal(x, y)
