"""Generate tests/golden/*.npz by running the REAL reference code from /root/reference (TEST INFRASTRUCTURE ONLY).

Run in the build container only (the GPU box has no /root/reference):   python -m oracle.gen_golden
The reference has no tests or golden vectors for this path, so these fixtures -- outputs of the
unmodified reference classes on seeded inputs -- are what pins the oracle, and through it the CUDA path.
Stubs: pytorch_lightning / matplotlib / omegaconf / timm are not installed here; they are replaced by the minimal
stand-ins below (timm.layers.Mlp = fc1 -> GELU -> fc2, as in timm).  Every fixture stores the seeds/configs needed
to rebuild its inputs with oracle/*.synthetic_* so nothing large is committed.
"""
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def install_stubs():
    import transformers  # noqa: F401  (must be imported before the timm stub exists)
    from transformers.generation.utils import TopKLogitsWarper  # noqa: F401
    import torch.nn as nn

    pl = types.ModuleType("pytorch_lightning")
    pl.LightningModule = nn.Module
    sys.modules["pytorch_lightning"] = pl
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules[name] = types.ModuleType(name)
    timm = types.ModuleType("timm")
    layers = types.ModuleType("timm.layers")

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0, **kw):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, out_features or in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    layers.Mlp = Mlp
    timm.layers = layers
    sys.modules["timm"] = timm
    sys.modules["timm.layers"] = layers
    sys.path.insert(0, REF)


class AttrDict(dict):
    __getattr__ = dict.__getitem__

    def get(self, k, d=None):
        return dict.get(self, k, d)


def ad(d):
    return AttrDict({k: ad(v) if isinstance(v, dict) else v for k, v in d.items()})


def pack(mask_bool):
    return np.packbits(mask_bool.astype(np.uint8), bitorder="little")


# --------------------------------------------------------------------------------------------------------------
def gen_greenlist():
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy
    from oracle import wm

    out = {}
    kat = {}
    for seed, n in [(0, 10), (1, 971), (15485863, 15413), (2 ** 40 + 7, 1024), (15485863 * 16383, 4)]:
        g = torch.Generator(device="cpu")
        g.manual_seed(seed)
        kat[f"randperm_{seed}_{n}"] = torch.randperm(n, generator=g).numpy()
    np.savez_compressed(os.path.join(OUT, "randperm_kat.npz"), **kat)

    assets = {
        "taming": (os.path.join(REF, "assets/vqgan_alive_ids.txt"), 16384, 16384),
        "rar": (os.path.join(REF, "assets/rar_all_ids.txt"), 1024, 1024),
        "chameleon": (os.path.join(REF, "assets/chameleon_all_ids.txt"), 8192, 65536),
    }
    for name, (path, n_e, V) in assets.items():
        alive, dead = wm.alive_dead(wm.load_ids(path), n_e)
        vq = {"alive_ids": torch.from_numpy(alive), "dead_ids": torch.from_numpy(dead),
              "embedding": torch.zeros(4, 4)}
        for split in ("stratifiedrand", "rand"):
            w = GentimeWatermark(vq, V, SeedStrategy.LINEAR, SplitStrategy(split), 1, 2.0, 0.25, "cpu")
            for c in (0, 1, 5, 975, min(16383, V - 1)):
                seed = (w.salt_key * c) % (2 ** 64 - 1)
                ids = w._split_with_seed(seed).numpy()
                m = np.zeros(V, dtype=bool)
                m[ids] = True
                out[f"{name}/{split}/ctx{c}/bits"] = pack(m)
                out[f"{name}/{split}/ctx{c}/n"] = np.int64(len(ids))
                out[f"{name}/{split}/ctx{c}/head"] = ids[:16]
        # gamma variant exercising int() truncation
        w = GentimeWatermark(vq, V, SeedStrategy.FIXED, SplitStrategy.RANDOM_STRATIFIED, 0, 2.0, 0.5, "cpu")
        ids = w.fixed_greenlist.numpy()
        m = np.zeros(V, dtype=bool)
        m[ids] = True
        out[f"{name}/fixed_g0.5/bits"] = pack(m)
        out[f"{name}/fixed_g0.5/n"] = np.int64(len(ids))
    np.savez_compressed(os.path.join(OUT, "greenlist.npz"), **out)
    print("greenlist.npz:", len(out), "arrays")


def gen_process_logits_and_detect():
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy
    from oracle import wm

    out = {}
    cases = [
        ("taming_linear_h1", "assets/vqgan_alive_ids.txt", 16384, 16384, "linear", "stratifiedrand", 1, 2.0, 0.25, 256),
        ("taming_linear_h2", "assets/vqgan_alive_ids.txt", 16384, 16384, "linear", "stratifiedrand", 2, 2.0, 0.25, 64),
        ("taming_rand_h1", "assets/vqgan_alive_ids.txt", 16384, 16384, "linear", "rand", 1, 4.0, 0.5, 64),
        ("taming_spatial_h1", "assets/vqgan_alive_ids.txt", 16384, 16384, "spatial", "stratifiedrand", 1, 2.0, 0.25, 256),
        ("taming_spatial_h3", "assets/vqgan_alive_ids.txt", 16384, 16384, "spatial", "stratifiedrand", 3, 2.0, 0.25, 256),
        ("rar_linear_h1", "assets/rar_all_ids.txt", 1024, 1024, "linear", "stratifiedrand", 1, 2.0, 0.25, 256),
        ("rar_fixed_h0", "assets/rar_all_ids.txt", 1024, 1024, "fixed", "stratifiedrand", 0, 2.0, 0.25, 256),
        ("cham_fixed_h0", "assets/chameleon_all_ids.txt", 8192, 65536, "fixed", "stratifiedrand", 0, 2.0, 0.25, 1024),
    ]
    for name, path, n_e, V, ss, sp, h, delta, gamma, L in cases:
        alive, dead = wm.alive_dead(wm.load_ids(os.path.join(REF, path)), n_e)
        vq = {"alive_ids": torch.from_numpy(alive), "dead_ids": torch.from_numpy(dead),
              "embedding": torch.zeros(4, 4)}
        w = GentimeWatermark(vq, V, SeedStrategy(ss), SplitStrategy(sp), h, delta, gamma, "cpu")
        g = torch.Generator().manual_seed(1234)
        hi = V if "cham" not in name else 8196
        lo = 0 if "cham" not in name else 4
        # detection on random codes, some rows made repetitive so that dedup matters
        B = 3
        codes = torch.randint(lo, hi, (B, L), generator=g)
        codes[1, L // 2:] = codes[1, : L - L // 2]
        codes[2, :] = codes[2, :8].repeat(L // 8)
        import io
        from contextlib import redirect_stdout
        with redirect_stdout(io.StringIO()):
            pv, masks = w.detect(codes, return_masks=True)
        out[f"{name}/codes"] = codes.numpy().astype(np.int32)
        out[f"{name}/pvalues"] = pv.numpy()
        for b in range(B):
            out[f"{name}/mask{b}"] = np.asarray(masks[b], dtype=np.int8)
        # the logit processor on a short history (incl. histories too short for the context)
        if "cham" in name:
            continue
        for t in ([0, 1, 2, 5] if ss != "spatial" else [1, 15, 16, 17, 33]):
            past = torch.randint(lo, hi, (4, t), generator=g)
            logits = torch.randn(4, V, generator=g)
            before = logits.clone()
            with redirect_stdout(io.StringIO()):
                after = w._process_logits(past, logits)
            changed = (after != before).numpy()
            assert np.allclose((after - before).numpy()[changed], delta, atol=1e-5)
            out[f"{name}/proc_t{t}/past"] = past.numpy().astype(np.int32)
            out[f"{name}/proc_t{t}/bits"] = np.stack([pack(changed[b]) for b in range(4)])
    np.savez_compressed(os.path.join(OUT, "watermark_ops.npz"), **out)
    print("watermark_ops.npz:", len(out), "arrays")


def gen_gpt():
    """Tiny + narrow-but-real-shaped minGPT through the reference sample_with_past with the reference watermark."""
    from deps.taming.modules.transformer.mingpt import GPT, sample_with_past
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy
    from oracle import gpt as ogpt
    from oracle import wm

    out = {}
    alive, dead = wm.alive_dead(wm.load_ids(os.path.join(REF, "assets/vqgan_alive_ids.txt")), 16384)
    vq = {"alive_ids": torch.from_numpy(alive), "dead_ids": torch.from_numpy(dead), "embedding": torch.zeros(4, 4)}
    cfgs = {"tiny": dict(vocab_size=16384, block_size=256, n_layer=2, n_head=2, n_embd=128, steps=24, B=4),
            "narrow": dict(vocab_size=16384, block_size=256, n_layer=3, n_head=6, n_embd=384, steps=48, B=16)}
    for name, c in cfgs.items():
        steps, B = c.pop("steps"), c.pop("B")
        weights = ogpt.synthetic_gpt_weights(seed=7, **c)
        model = GPT(**c)
        missing, unexpected = model.load_state_dict(weights, strict=False)
        assert not unexpected and all(k.endswith("attn.mask") for k in missing), (missing, unexpected)
        model.eval()
        w = GentimeWatermark(vq, 16384, SeedStrategy.LINEAR, SplitStrategy.RANDOM_STRATIFIED, 1, 2.0, 0.25, "cpu")
        cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:B].view(-1, 1)
        # greedy, watermarked
        codes_g = sample_with_past(cond, model, steps, temperature=1.0, sample_logits=False, top_k=250, top_p=0.92,
                                   logit_processor=w.spawn_logit_processor())
        # sampled, watermarked: CPU generator seeded; multinomial draws exponential_ over [B,V] each step
        torch.manual_seed(1)
        codes_s = sample_with_past(cond, model, steps, temperature=1.0, sample_logits=True, top_k=250, top_p=0.92,
                                   logit_processor=w.spawn_logit_processor())
        # sampled, no watermark, other params
        torch.manual_seed(2)
        codes_n = sample_with_past(cond, model, steps, temperature=0.8, sample_logits=True, top_k=600, top_p=0.5,
                                   logit_processor=None)
        # first-step logits for a numerics pin
        lg, _, _ = model.forward_with_past(cond, past=None, past_length=0)
        out[f"{name}/cond"] = cond.numpy()[:, 0]
        out[f"{name}/greedy_wm"] = codes_g.numpy().astype(np.int32)
        out[f"{name}/sample_wm_seed1"] = codes_s.numpy().astype(np.int32)
        out[f"{name}/sample_nowm_seed2"] = codes_n.numpy().astype(np.int32)
        out[f"{name}/logits0_head"] = lg[:, 0, :64].detach().numpy()
        out[f"{name}/cfg"] = np.asarray([c["vocab_size"], c["block_size"], c["n_layer"], c["n_head"], c["n_embd"],
                                         steps, B, 7])
    np.savez_compressed(os.path.join(OUT, "gpt.npz"), **out)
    print("gpt.npz:", len(out), "arrays")


def gen_vqgan():
    from deps.taming.modules.diffusionmodules.model import Decoder, Encoder
    from deps.taming.modules.vqvae.quantize import VectorQuantizer2
    from deps.rar.modeling.modules import maskgit_vqgan as mg
    from oracle import vqgan as ov

    out = {}
    # ---- Taming VQModel glue (vqgan.py:64-73, cond_transformer.py:169-192) around the real Encoder/Decoder/VQ2
    for name, cfg, B in (("taming_small", dict(ov.TAMING_CFG, ch=32, ch_mult=(1, 2, 2), resolution=64, z_channels=64,
                                                n_embed=512, embed_dim=64), 2),
                         ("taming_full", ov.TAMING_CFG, 1)):
        w = ov.synthetic_taming_vqgan_weights(cfg, seed=3)
        dd = dict(ch=cfg["ch"], out_ch=cfg["out_ch"], ch_mult=cfg["ch_mult"], num_res_blocks=cfg["num_res_blocks"],
                  attn_resolutions=list(cfg["attn_resolutions"]), dropout=0.0, in_channels=cfg["in_channels"],
                  resolution=cfg["resolution"], z_channels=cfg["z_channels"], double_z=False)
        enc, dec = Encoder(**dd).eval(), Decoder(**dd).eval()
        vq = VectorQuantizer2(cfg["n_embed"], cfg["embed_dim"], beta=0.25)
        qc = torch.nn.Conv2d(cfg["z_channels"], cfg["embed_dim"], 1)
        pqc = torch.nn.Conv2d(cfg["embed_dim"], cfg["z_channels"], 1)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")})
        dec.load_state_dict({k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")})
        vq.load_state_dict({"embedding.weight": w["quantize.embedding.weight"]})
        qc.load_state_dict({"weight": w["quant_conv.weight"], "bias": w["quant_conv.bias"]})
        pqc.load_state_dict({"weight": w["post_quant_conv.weight"], "bias": w["post_quant_conv.bias"]})
        g = torch.Generator().manual_seed(11)
        res = cfg["resolution"]
        img = torch.rand(B, 3, res, res, generator=g) * 2 - 1
        s = res // 2 ** (len(cfg["ch_mult"]) - 1)
        codes_in = torch.randint(0, cfg["n_embed"], (B, s * s), generator=g)
        with torch.no_grad():
            _, _, info = vq(qc(enc(img)))
            codes = info[2].view(B, -1)
            zq = vq.get_codebook_entry(codes_in.reshape(-1), shape=(B, s, s, cfg["embed_dim"]))
            rec = dec(pqc(zq)).clamp(-1, 1)
            o_codes = ov.taming_images_to_codes(img, w, cfg)
            o_rec = ov.taming_codes_to_images(codes_in, w, cfg)
        assert torch.equal(o_codes, codes), "oracle VQGAN encode != reference"
        assert torch.allclose(o_rec, rec, atol=1e-5), float((o_rec - rec).abs().max())
        out[f"{name}/codes"] = codes.numpy().astype(np.int32)
        out[f"{name}/codes_in"] = codes_in.numpy().astype(np.int32)
        stride = max(1, res // 32)
        out[f"{name}/rec_sub"] = rec[:, :, ::stride, ::stride].numpy()
        out[f"{name}/rec_mean_abs"] = np.float64(rec.abs().double().mean().item())
        print(name, "ok; rec mean|x| =", out[f"{name}/rec_mean_abs"], "unique codes", codes.unique().numel())
    # ---- MaskGIT VQGAN (RAR tokenizer)
    for name, cfg, B in (("maskgit_small", dict(ov.MASKGIT_CFG, hidden_channels=32, channel_mult=(1, 2, 2),
                                                 resolution=64, z_channels=64, num_embeddings=256), 2),
                         ("maskgit_full", ov.MASKGIT_CFG, 1)):
        w = ov.synthetic_maskgit_weights(cfg, seed=5)
        conf = ad(dict(channel_mult=list(cfg["channel_mult"]), num_resolutions=len(cfg["channel_mult"]), dropout=0.0,
                       hidden_channels=cfg["hidden_channels"], num_channels=3,
                       num_res_blocks=cfg["num_res_blocks"], resolution=cfg["resolution"],
                       z_channels=cfg["z_channels"]))
        enc, dec = mg.Encoder(conf).eval(), mg.Decoder(conf).eval()
        vq = mg.VectorQuantizer(cfg["num_embeddings"], cfg["z_channels"], 0.25)
        enc.load_state_dict({k[len("encoder."):]: v for k, v in w.items() if k.startswith("encoder.")})
        dec.load_state_dict({k[len("decoder."):]: v for k, v in w.items() if k.startswith("decoder.")})
        vq.load_state_dict({"embedding.weight": w["quantize.embedding.weight"]})
        g = torch.Generator().manual_seed(13)
        res = cfg["resolution"]
        img = torch.rand(B, 3, res, res, generator=g) * 2 - 1
        s = res // 2 ** (len(cfg["channel_mult"]) - 1)
        codes_in = torch.randint(0, cfg["num_embeddings"], (B, s * s), generator=g)
        with torch.no_grad():
            codes = vq(enc((img + 1.0) / 2.0))[1]                       # rar_wrapper.py:124-125, titok.py:75-78
            rec = torch.clamp(dec(vq.get_codebook_entry(codes_in)), 0.0, 1.0)   # titok.py:81-85
            rec = torch.clamp(rec * 2.0 - 1.0, -1.0, 1.0)               # rar_wrapper.py:112-114
            o_codes = ov.rar_images_to_codes(img, w, cfg)
            o_rec = ov.rar_codes_to_images(codes_in, w, cfg)
        assert torch.equal(o_codes, codes), "oracle MaskGIT encode != reference"
        assert torch.allclose(o_rec, rec, atol=1e-5), float((o_rec - rec).abs().max())
        out[f"{name}/codes"] = codes.numpy().astype(np.int32)
        out[f"{name}/codes_in"] = codes_in.numpy().astype(np.int32)
        stride = max(1, res // 32)
        out[f"{name}/rec_sub"] = rec[:, :, ::stride, ::stride].numpy()
        print(name, "ok; unique codes", codes.unique().numel())
    np.savez_compressed(os.path.join(OUT, "vqgan.npz"), **out)
    print("vqgan.npz:", len(out), "arrays")


def gen_rar():
    from deps.rar.modeling.rar import RAR
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy
    from oracle import rar as orar
    from oracle import wm

    out = {}
    alive, dead = wm.alive_dead(wm.load_ids(os.path.join(REF, "assets/rar_all_ids.txt")), 1024)
    vq = {"alive_ids": torch.from_numpy(alive), "dead_ids": torch.from_numpy(dead), "embedding": torch.zeros(4, 4)}
    for name, (d, depth, heads, mlp, steps, B) in {"tiny": (128, 2, 4, 256, 24, 3),
                                                    "narrow": (320, 3, 4, 640, 40, 8)}.items():
        conf = ad({"model": {"generator": {"hidden_size": d, "num_hidden_layers": depth, "num_attention_heads": heads,
                                           "intermediate_size": mlp, "image_seq_len": 256,
                                           "condition_num_classes": 1000, "dropout": 0.0, "attn_drop": 0.0},
                             "vq_model": {"codebook_size": 1024}}})
        model = RAR(conf)
        weights = orar.synthetic_rar_weights(d, depth, heads, mlp, seed=9)
        missing, unexpected = model.load_state_dict(weights, strict=False)
        assert not missing and not unexpected, (missing, unexpected)
        model.eval()
        model.set_random_ratio(0)
        w = GentimeWatermark(vq, 1024, SeedStrategy.LINEAR, SplitStrategy.RANDOM_STRATIFIED, 1, 2.0, 0.25, "cpu")
        cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814])[:B].view(-1, 1)
        # the reference generate() hard-codes image_seq_len steps; run the full 256 only for 'tiny'
        model.image_seq_len_backup = model.image_seq_len
        kw = dict(guidance_scale=4.0, guidance_decay="constant", guidance_scale_pow=0.0, randomize_temperature=1.0,
                  softmax_temperature_annealing=False, num_sample_steps=8)
        torch.manual_seed(3)
        ids = model.generate(condition=cond.clone(), logit_processor=w.spawn_logit_processor(), **kw)
        out[f"{name}/cond"] = cond.numpy()[:, 0]
        out[f"{name}/sample_wm_seed3"] = ids.numpy().astype(np.int32)
        out[f"{name}/cfg"] = np.asarray([d, depth, heads, mlp, 256, B, 9])
        # greedy variant: RAR has no greedy switch; patch torch.multinomial for the call (both sides do argmax)
        orig = torch.multinomial
        torch.multinomial = lambda p, num_samples=1: torch.argmax(p, dim=-1, keepdim=True)
        try:
            ids_g = model.generate(condition=cond.clone(), logit_processor=w.spawn_logit_processor(), **kw)
        finally:
            torch.multinomial = orig
        out[f"{name}/greedy_wm"] = ids_g.numpy().astype(np.int32)
        print("rar", name, "done")
    np.savez_compressed(os.path.join(OUT, "rar.npz"), **out)
    print("rar.npz:", len(out), "arrays")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.dirname(HERE))
    install_stubs()
    which = sys.argv[1:] or ["greenlist", "ops", "gpt", "vqgan", "rar"]
    torch.set_grad_enabled(False)
    if "greenlist" in which:
        gen_greenlist()
    if "ops" in which:
        gen_process_logits_and_detect()
    if "gpt" in which:
        gen_gpt()
    if "vqgan" in which:
        gen_vqgan()
    if "rar" in which:
        gen_rar()
