"""GPU parity: bf16 skinny GEMM, the Chameleon logits-processor/token-selector operator (vs goldens made by the
reference's own classes) and the Chameleon decode engine (vs the CPU restatement, teacher-forced)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from helpers import G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,K,split", [(256, 256, 1), (768, 256, 2), (4096, 4096, 0), (4096, 11008, 4), (512, 352, 1)])
def test_skinny_gemm_bf16(N, K, split):
    from wmar_b200 import _lib
    g = torch.Generator().manual_seed(N + K)
    x = torch.randn(16, K, generator=g).bfloat16().float().cuda()
    w = (torch.randn(N, K, generator=g) * 0.05).bfloat16().cuda()
    y = torch.empty(16, N, device="cuda")
    _lib.check(_lib.lib().wmar_skinny_gemm_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y), N, K, split, _lib.current_stream()))
    ref = (x.double() @ w.double().t())
    # fp32 accumulation of exact bf16 products, rounded once to bf16: within one bf16 ulp (2^-8 relative) of the exact value
    err = (y.double() - ref).abs()
    tol = ref.abs() * 2.0 ** -8 + 1e-3 * ref.abs().max() * 2.0 ** -8
    assert bool((err <= tol).all()), float((err - tol).max())
    assert torch.equal(y, y.bfloat16().float())            # bf16 values
    y2 = torch.empty_like(y)
    _lib.check(_lib.lib().wmar_skinny_gemm_bf16(_lib.ptr(x), _lib.ptr(w), _lib.ptr(y2), N, K, split, _lib.current_stream()))
    assert torch.equal(y, y2)                              # deterministic split-K


def _green_params(green, V, delta):
    from wmar_b200 import _lib
    bits = np.zeros(((V + 31) // 32) * 32, dtype=np.uint8)
    bits[green] = 1
    table = torch.from_numpy(np.packbits(bits, bitorder="little").view(np.int32).copy()).cuda().reshape(1, -1)
    return _lib.WmParams(table.data_ptr(), 1, V, 0, 0, 16, float(delta), 0.25), table


def test_select_operator_matches_reference_classes():
    from wmar_b200 import _lib
    g = np.load(os.path.join(G, "chameleon_sampling.npz"))
    V, lo, hi, B = [int(x) for x in g["meta"]]
    logits3 = torch.from_numpy(g["logits"]).cuda()
    wm, keep = _green_params(g["green"], V, 2.0)
    for name in "abc":
        use_wm, temp, top_p, greedy = g[f"{name}/cfg"]
        sp = _lib.SampleParams(float(temp), 0, float(top_p), int(greedy), 0)
        noise = torch.from_numpy(g[f"{name}/noise"]).cuda()
        out = torch.empty(B, dtype=torch.long, device="cuda")
        mixed = torch.empty(B, hi - lo, device="cuda")
        _lib.check(_lib.lib().wmar_cham_select(ctypes.byref(wm) if use_wm else None, ctypes.byref(sp), _lib.ptr(logits3), B, V,
                                               lo, hi, 3.0, 1.2, None, 0, 0, _lib.ptr(noise), _lib.ptr(out), _lib.ptr(mixed),
                                               _lib.current_stream()))
        np.testing.assert_array_equal(out.cpu().numpy(), g[f"{name}/ids"][:B])      # token ids: bit-exact
    _lib.check(_lib.lib().wmar_check_device_flag(_lib.current_stream()))


def _tiny():
    from oracle import chameleon as oc
    from wmar_b200.models.cham_engine import ChameleonEngine
    V, d, L, H, Fh = 1024, 256, 2, 2, 384
    lo, hi = 4, 516
    w = oc.synthetic_chameleon_weights(V, d, L, H, H, Fh, seed=5)
    eng = ChameleonEngine(w, L, H, image_tokens=(lo, hi), max_seq=64, max_batch=2)
    orc = oc.ChameleonOracle(w, L, H, H)
    boi, bos = 700, 0
    full = [[bos, 901, 902, 903, 904, boi], [bos, 950, 951, boi]]
    img = [[bos, boi], [bos, boi]]
    unc = [[bos, boi], [bos, boi]]
    return eng, orc, full + img + unc, (V, lo, hi)


def test_chameleon_engine_vs_oracle_teacher_forced():
    from oracle import chameleon as oc
    from wmar_b200 import _lib
    eng, orc, prompts3, (V, lo, hi) = _tiny()
    steps, B = 12, 2
    ids, mixed = eng.sample(prompts3, steps, 3.0, 1.2, temperature=0.9, top_p=0.9, greedy=True, return_logits=True)
    ids_c = ids.cpu()
    assert int(ids_c.min()) >= lo and int(ids_c.max()) < hi               # only image tokens are ever selected
    _, want = oc.generate(orc, prompts3, B, steps, 3.0, 1.2, lo, hi, 0.9, 0.9, greedy=True, forced_ids=ids_c)
    want = want[:, :, lo:hi]
    got = mixed.cpu()
    # bf16 model: every Linear / norm output is rounded to bf16 (2^-8 relative); different summation orders flip the last
    # bit here and there, which the 3-way guidance (scales 3.0 / 1.2) amplifies -> tolerance 4 % of the logit range
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    assert err <= 0.04 * scale, (err, scale)
    # greedy choice: the engine's token must be (near-)optimal under the oracle's logits at every step
    for s in range(steps):
        for b in range(B):
            assert want[s, b, ids_c[b, s] - lo] >= want[s, b].max() - 0.08 * scale
    again = eng.sample(prompts3, steps, 3.0, 1.2, temperature=0.9, top_p=0.9, greedy=True)
    assert torch.equal(ids, again)                                        # deterministic, cache re-initialised per call
    _lib.check(_lib.lib().wmar_check_device_flag(_lib.current_stream()))


@pytest.mark.parametrize("name", ["tiny", "gqa"])
def test_chameleon_engine_vs_reference_module_logits(name):
    """The CUDA engine against logits of the reference's OWN Transformer module (tests/golden/chameleon_transformer.npz,
    oracle/gen_golden_chameleon_transformer.py): ragged prompts, three guidance groups, teacher-forced tokens.  Teacher
    forcing goes through the public noise argument: q = 1e-30 at the forced id makes argmax(p / q) select it."""
    from oracle import chameleon as oc
    from wmar_b200 import _lib
    from wmar_b200.models.cham_engine import ChameleonEngine
    g = np.load(os.path.join(G, "chameleon_transformer.npz"))
    V, d, L, H, Hkv, Fh, steps, seed = [int(x) for x in g[f"{name}/meta"]]
    lens = [int(x) for x in g[f"{name}/prompt_lens"]]
    flat = [int(x) for x in g[f"{name}/prompts_flat"]]
    prompts3, o = [], 0
    for n in lens:
        prompts3.append(flat[o:o + n])
        o += n
    B = len(prompts3) // 3
    forced = torch.from_numpy(g[f"{name}/forced"])[:B]                     # [B, steps]
    ref = torch.from_numpy(g[f"{name}/logits"])                            # [steps + 1, 3B, V] fp32 (bf16 values)
    lo, hi = 4, 516
    w = oc.synthetic_chameleon_weights(V, d, L, H, Hkv, Fh, seed=seed)
    eng = ChameleonEngine(w, L, H, n_kv_head=Hkv, image_tokens=(lo, hi), max_seq=64, max_batch=2)
    noise = torch.ones(steps + 1, B, V)
    for s in range(steps):
        noise[s, torch.arange(B), forced[:, s]] = 1e-30
    ids, mixed = eng.sample(prompts3, steps + 1, 3.0, 1.2, temperature=1.0, top_p=None, noise=noise.cuda(), return_logits=True)
    np.testing.assert_array_equal(ids.cpu()[:, :steps].numpy(), forced.numpy())
    f, im, u = ref[:, :B], ref[:, B:2 * B], ref[:, 2 * B:]
    want = (u + 1.2 * (im - u) + 3.0 * (f - im))[:, :, lo:hi]              # logits_processor.py:312-335
    got = mixed.cpu()
    scale = want.abs().max().item()
    err = (got - want).abs().max().item()
    rms = (got - want).pow(2).mean().sqrt().item()
    print(f"chameleon {name}: range {scale:.2f}, max |diff| {err:.4f}, rms {rms:.5f}")
    # bf16 model, guidance scales 3.0 / 1.2 amplify single-ulp flips of the three rows (see the teacher-forced test above)
    assert err <= 0.04 * scale and rms <= 0.004 * scale, (err, rms, scale)
    _lib.check(_lib.lib().wmar_check_device_flag(_lib.current_stream()))


def test_chameleon_engine_sampling_with_reference_noise_and_watermark():
    """Sampling path: with the Exp(1) draws of torch.multinomial and a fixed greenlist the engine's token at every step
    equals what the restated reference pipeline selects from the ENGINE's own mixed logits (isolates the fused
    processors + selector from bf16 noise in the transformer)."""
    import ctypes
    from transformers import TopPLogitsWarper
    from oracle import chameleon as oc
    from wmar_b200 import _lib
    eng, orc, prompts3, (V, lo, hi) = _tiny()
    steps, B = 10, 2
    torch.manual_seed(3)
    noise = torch.empty(steps, B, V).exponential_(1)
    green = torch.randperm(V)[: V // 4].numpy()
    wm, keep = _green_params(green, V, 2.0)

    class _WM:                                   # minimal stand-in for GentimeWatermark.c_params()
        def c_params(self):
            return wm

    ids, mixed = eng.sample(prompts3, steps, 3.0, 1.2, temperature=0.9, top_p=0.9, watermarker=_WM(), noise=noise.cuda(),
                            return_logits=True)
    ids_c, mixed_c = ids.cpu(), mixed.cpu()
    for s in range(steps):
        l = torch.full((B, V), -float("inf"))
        l[:, lo:hi] = mixed_c[s]
        gmask = torch.zeros(V, dtype=torch.bool)
        gmask[green] = True
        l[:, gmask] += 2.0
        l = TopPLogitsWarper(0.9)(None, l / 0.9)
        want = (l.softmax(dim=1) / noise[s]).argmax(dim=1)
        np.testing.assert_array_equal(ids_c[:, s].numpy(), want.numpy())
    frac = np.isin(ids_c.numpy(), green).mean()
    assert frac > 0.25                            # the bias is visible


def test_chameleon_wrapper_roundtrip_small():
    """ChameleonARMMWrapper at reduced shapes: sample -> codes_to_images -> images_to_codes -> detect keep the reference's
    shapes / ranges / vocab conventions (chameleon_wrapper.py:139-186)."""
    from wmar_b200.models.chameleon_wrapper import (BEGIN_IMAGE, IMAGE_TOKEN_HI, IMAGE_TOKEN_LO, ChameleonARMMWrapper)
    from wmar_b200.watermarking import create_watermarker_from_string
    cfg = dict(vocab_size=16384, dim=256, n_layers=2, n_heads=2, n_kv_heads=2, ffn_hidden=384, norm_eps=1e-5, rope_theta=10000.0,
               qk_normalization=True)
    vq = dict(ch=128, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(), in_channels=3, resolution=32,
              z_channels=256, n_embed=8192, embed_dim=256)
    m = ChameleonARMMWrapper(model_cfg=cfg, vq_cfg=vq, max_batch=2, image_tokens_per_image=256,
                             tokenize=lambda p: [16384 - 1 - (len(w) % 7) for w in p.split()])
    assert m.get_total_vocab_size() == 16384 and m.codes_size == 16 and m.image_size == 32
    rows = m.prompt_rows(["a red bus", "two cats on a couch"])
    assert len(rows) == 6 and all(r[-1] == BEGIN_IMAGE for r in rows) and rows[2] == [0, BEGIN_IMAGE] and rows[4] == [0, BEGIN_IMAGE]
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), "fixed-stratifiedrand-h=0-d=4.0-g=0.25", m.device)
    m.set_watermarker(wm)
    torch.manual_seed(0)
    cond = [(0, "a red bus"), (1, "two cats on a couch"), (2, "a dog")]          # 3 prompts, max_batch 2 -> two engine calls
    codes = m.sample(cond, {"temperature": 0.9, "top_p": 0.9}, apply_watermark=True)
    assert codes.shape == (3, 256) and int(codes.min()) >= IMAGE_TOKEN_LO and int(codes.max()) < IMAGE_TOKEN_HI
    imgs = m.codes_to_images(codes)
    assert imgs.shape == (3, 3, 32, 32) and float(imgs.min()) >= -1.0 and float(imgs.max()) <= 1.0
    back = m.images_to_codes(imgs)
    assert back.shape == codes.shape and int(back.min()) >= IMAGE_TOKEN_LO and int(back.max()) < IMAGE_TOKEN_HI
    p_wm = wm.detect(codes)
    codes0 = m.sample(cond, {"temperature": 0.9, "top_p": 0.9}, apply_watermark=False)
    p_0 = wm.detect(codes0)
    assert float(p_wm.max()) < 1e-6 < float(p_0.min())                          # delta 4: the watermark is unmistakable


def test_chameleon_wrapper_lanes_equal_sequential_chunk_loop():
    """ChameleonARMMWrapper.sample with more prompts than max_batch: chunks on concurrent engine lanes (shared weights, own
    KV cache / stream) == the sequential chunk loop, bit for bit, sampled and greedy, including the first call (lanes are
    created before any work is enqueued)."""
    from wmar_b200.models.chameleon_wrapper import ChameleonARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    cfg = dict(vocab_size=16384, dim=256, n_layers=2, n_heads=2, n_kv_heads=2, ffn_hidden=384, norm_eps=1e-5, rope_theta=10000.0,
               qk_normalization=True)
    vq = dict(ch=128, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, attn_resolutions=(), in_channels=3, resolution=32,
              z_channels=256, n_embed=8192, embed_dim=256)
    cond = [(i, p) for i, p in enumerate(["a red bus", "two cats on a couch", "a dog", "a kitchen with a stove", "pizza"])]
    res = {}
    for lanes in (1, 2):
        m = ChameleonARMMWrapper(model_cfg=cfg, vq_cfg=vq, max_batch=2, image_tokens_per_image=256, lanes=lanes,
                                 tokenize=lambda p: [16384 - 1 - (len(w) % 7) for w in p.split()])
        wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), "fixed-stratifiedrand-h=0-d=4.0-g=0.25", m.device)
        m.set_watermarker(wm)
        torch.manual_seed(3)
        a = m.sample(cond, {"temperature": 0.9, "top_p": 0.9}, apply_watermark=True)       # 3 chunks: 2 + 2 + 1 prompts
        g = m.sample(cond, {"temperature": 0.9, "top_p": 0.9}, apply_watermark=True, greedy=True)
        res[lanes] = (a.cpu(), g.cpu(), torch.rand(4, device="cuda").cpu())
    for x, y in zip(res[1], res[2]):
        assert torch.equal(x, y)


def test_two_group_mode_is_bit_identical_to_three_groups():
    """Text-only prompts: image-conditioned rows == unconditioned rows; computing them once (n_groups = 2) must give the
    same ids and the same mixed logits as the reference's three row groups."""
    eng, orc, prompts3, (V, lo, hi) = _tiny()
    B = 2
    assert prompts3[B:2 * B] == prompts3[2 * B:]
    ids2, mixed2 = eng.sample(prompts3, 12, 3.0, 1.2, temperature=0.9, top_p=0.9, greedy=True, return_logits=True)
    assert eng.last_n_groups == 2
    # force three groups by making one image-conditioned row formally different (same tokens, list vs tuple is equal, so
    # perturb and restore through the C-ABI: call with an explicit 3-group layout)
    import ctypes
    from wmar_b200 import _lib
    R = 3 * B
    p_max = max(len(p) for p in prompts3)
    pr = torch.zeros((R, p_max), dtype=torch.long)
    for r, p in enumerate(prompts3):
        pr[r, :len(p)] = torch.tensor(p)
    pr = pr.cuda()
    plen = torch.tensor([len(p) for p in prompts3], dtype=torch.int32, device="cuda")
    out = torch.empty((B, 12), dtype=torch.long, device="cuda")
    logits = torch.empty((12, B, hi - lo), device="cuda")
    sp = _lib.SampleParams(0.9, 0, 0.9, 1, 0)
    _lib.check(_lib.lib().wmar_cham_sample(eng.handle, None, ctypes.byref(sp), _lib.ptr(pr), _lib.ptr(plen), p_max, p_max, B, 3,
                                           3.0, 1.2, 12, None, _lib.ptr(out), _lib.ptr(logits), _lib.current_stream()))
    assert torch.equal(out, ids2) and torch.equal(logits, mixed2)
