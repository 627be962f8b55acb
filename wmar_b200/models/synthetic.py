"""Seeded random-init state dicts at the reference's shapes (there are no checkpoints offline).

Key names equal the reference modules' ``state_dict()`` keys so that the same packers serve real checkpoints:
  Taming Net2NetTransformer  (deps/taming/models/cond_transformer.py): ``transformer.*`` (minGPT) and
      ``first_stage_model.*`` (VQModel: encoder, decoder, quantize.embedding, quant_conv, post_quant_conv)
  RAR  (deps/rar/modeling/rar.py) and its MaskGIT-VQGAN tokenizer (deps/rar/modeling/modules/maskgit_vqgan.py)
Tensors are created directly on ``device`` (1.4 G parameters take seconds on the GPU, minutes on the host).
"""
import torch

TAMING_GPT_CFG = dict(vocab_size=16384, block_size=256, n_layer=48, n_head=24, n_embd=1536)
TAMING_VQGAN_DDCONFIG = dict(ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,),
                             in_channels=3, resolution=256, z_channels=256, n_embed=16384, embed_dim=256)
RAR_XL_CFG = dict(hidden_size=1280, num_hidden_layers=32, num_attention_heads=16, intermediate_size=5120,
                  codebook_size=1024, image_seq_len=256, condition_num_classes=1000)
MASKGIT_VQGAN_CFG = dict(hidden_channels=128, channel_mult=(1, 1, 2, 2, 4), num_res_blocks=2, num_channels=3,
                         resolution=256, z_channels=256, num_embeddings=1024)


class _Gen:
    def __init__(self, seed, device):
        self.device = torch.device(device)
        self.g = torch.Generator(device=self.device).manual_seed(seed)

    def rn(self, *shape, std=1.0, mean=0.0):
        t = torch.randn(*shape, generator=self.g, device=self.device, dtype=torch.float32)
        return t.mul_(std).add_(mean) if (std != 1.0 or mean != 0.0) else t


def gpt_state(cfg=TAMING_GPT_CFG, seed=0, device="cpu", prefix=""):
    """GPT._init_weights (mingpt.py:155-163): N(0, 0.02) Linear / Embedding; LN near identity; small biases."""
    G = _Gen(seed, device)
    V, T, L, d = cfg["vocab_size"], cfg["block_size"], cfg["n_layer"], cfg["n_embd"]
    w = {"tok_emb.weight": G.rn(V, d, std=0.02), "pos_emb": G.rn(1, T, d, std=0.02)}
    for i in range(L):
        p = f"blocks.{i}."
        for ln in ("ln1", "ln2"):
            w[p + ln + ".weight"] = G.rn(d, std=0.1, mean=1.0)
            w[p + ln + ".bias"] = G.rn(d, std=0.05)
        for nm in ("key", "query", "value", "proj"):
            w[p + f"attn.{nm}.weight"] = G.rn(d, d, std=0.02)
            w[p + f"attn.{nm}.bias"] = G.rn(d, std=0.01)
        w[p + "mlp.0.weight"] = G.rn(4 * d, d, std=0.02)
        w[p + "mlp.0.bias"] = G.rn(4 * d, std=0.01)
        w[p + "mlp.2.weight"] = G.rn(d, 4 * d, std=0.02)
        w[p + "mlp.2.bias"] = G.rn(d, std=0.01)
    w["ln_f.weight"] = G.rn(d, std=0.1, mean=1.0)
    w["ln_f.bias"] = G.rn(d, std=0.05)
    w["head.weight"] = G.rn(V, d, std=0.02)
    return {prefix + k: v for k, v in w.items()}


def _conv(w, G, p, cout, cin, k, bias=True):
    w[p + ".weight"] = G.rn(cout, cin, k, k, std=1.0 / (cin * k * k) ** 0.5)
    if bias:
        w[p + ".bias"] = G.rn(cout, std=0.05)


def _norm(w, G, p, c):
    w[p + ".weight"] = G.rn(c, std=0.1, mean=1.0)
    w[p + ".bias"] = G.rn(c, std=0.05)


def taming_vqgan_state(cfg=TAMING_VQGAN_DDCONFIG, seed=0, device="cpu", prefix=""):
    """Taming / Chameleon VQModel (diffusionmodules/model.py:343-538, models/vqgan.py:27-45)."""
    G = _Gen(seed, device)
    w = {}

    def res(p, cin, cout):
        _norm(w, G, p + ".norm1", cin)
        _conv(w, G, p + ".conv1", cout, cin, 3)
        _norm(w, G, p + ".norm2", cout)
        _conv(w, G, p + ".conv2", cout, cout, 3)
        if cin != cout:
            _conv(w, G, p + ".nin_shortcut", cout, cin, 1)

    def attn(p, c):
        _norm(w, G, p + ".norm", c)
        for nm in ("q", "k", "v", "proj_out"):
            _conv(w, G, p + "." + nm, c, c, 1)

    ch, mult, nrb = cfg["ch"], tuple(cfg["ch_mult"]), cfg["num_res_blocks"]
    nres = len(mult)
    _conv(w, G, "encoder.conv_in", ch, cfg["in_channels"], 3)
    res_now = cfg["resolution"]
    in_mult = (1,) + mult
    bin_ = ch
    for lvl in range(nres):
        bin_, bout = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
            if res_now in cfg["attn_resolutions"]:
                attn(f"encoder.down.{lvl}.attn.{b}", bin_)
        if lvl != nres - 1:
            _conv(w, G, f"encoder.down.{lvl}.downsample.conv", bin_, bin_, 3)
            res_now //= 2
    res("encoder.mid.block_1", bin_, bin_)
    attn("encoder.mid.attn_1", bin_)
    res("encoder.mid.block_2", bin_, bin_)
    _norm(w, G, "encoder.norm_out", bin_)
    _conv(w, G, "encoder.conv_out", cfg["z_channels"], bin_, 3)
    bin_ = ch * mult[-1]
    _conv(w, G, "decoder.conv_in", bin_, cfg["z_channels"], 3)
    res("decoder.mid.block_1", bin_, bin_)
    attn("decoder.mid.attn_1", bin_)
    res("decoder.mid.block_2", bin_, bin_)
    for lvl in reversed(range(nres)):
        bout = ch * mult[lvl]
        for b in range(nrb + 1):
            res(f"decoder.up.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
            if res_now in cfg["attn_resolutions"]:
                attn(f"decoder.up.{lvl}.attn.{b}", bin_)
        if lvl != 0:
            _conv(w, G, f"decoder.up.{lvl}.upsample.conv", bin_, bin_, 3)
            res_now *= 2
    _norm(w, G, "decoder.norm_out", bin_)
    _conv(w, G, "decoder.conv_out", cfg["out_ch"], bin_, 3)
    w["quantize.embedding.weight"] = G.rn(cfg["n_embed"], cfg["embed_dim"])
    _conv(w, G, "quant_conv", cfg["embed_dim"], cfg["z_channels"], 1)
    _conv(w, G, "post_quant_conv", cfg["z_channels"], cfg["embed_dim"], 1)
    return {prefix + k: v for k, v in w.items()}


def maskgit_vqgan_state(cfg=MASKGIT_VQGAN_CFG, seed=0, device="cpu", prefix=""):
    """MaskGIT-VQGAN tokenizer of RAR (maskgit_vqgan.py:160-321): bias-free convs except conv_out / decoder.conv_in /
    upsample_conv; nin_shortcut takes the block's OUTPUT channel count (it is applied to h, :87-88)."""
    G = _Gen(seed, device)
    w = {}

    def res(p, cin, cout):
        _norm(w, G, p + ".norm1", cin)
        _conv(w, G, p + ".conv1", cout, cin, 3, False)
        _norm(w, G, p + ".norm2", cout)
        _conv(w, G, p + ".conv2", cout, cout, 3, False)
        if cin != cout:
            _conv(w, G, p + ".nin_shortcut", cout, cout, 1, False)

    hc, mult, nrb = cfg["hidden_channels"], tuple(cfg["channel_mult"]), cfg["num_res_blocks"]
    nres = len(mult)
    in_mult = (1,) + mult
    _conv(w, G, "encoder.conv_in", hc, cfg["num_channels"], 3, False)
    for lvl in range(nres):
        bin_, bout = hc * in_mult[lvl], hc * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
    mid = hc * mult[-1]
    for b in range(nrb):
        res(f"encoder.mid.{b}", mid, mid)
    _norm(w, G, "encoder.norm_out", mid)
    _conv(w, G, "encoder.conv_out", cfg["z_channels"], mid, 1, True)
    _conv(w, G, "decoder.conv_in", mid, cfg["z_channels"], 3, True)
    for b in range(nrb):
        res(f"decoder.mid.{b}", mid, mid)
    for lvl in reversed(range(nres)):
        bin_ = hc * mult[-1] if lvl == nres - 1 else hc * mult[lvl + 1]
        bout = hc * mult[lvl]
        for b in range(nrb):
            res(f"decoder.up.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
        if lvl != 0:
            _conv(w, G, f"decoder.up.{lvl}.upsample_conv", bout, bout, 3, True)
    _norm(w, G, "decoder.norm_out", hc * mult[0])
    _conv(w, G, "decoder.conv_out", cfg["num_channels"], hc * mult[0], 3, True)
    w["quantize.embedding.weight"] = G.rn(cfg["num_embeddings"], cfg["z_channels"])
    return {prefix + k: v for k, v in w.items()}


def taming_net2net_state(gpt_cfg=TAMING_GPT_CFG, dd_cfg=TAMING_VQGAN_DDCONFIG, seed=0, device="cpu"):
    """State dict of a Net2NetTransformer checkpoint (cond_transformer.py:27-60; cond stage = Labelator, no weights)."""
    s = gpt_state(gpt_cfg, seed, device, prefix="transformer.")
    s.update(taming_vqgan_state(dd_cfg, seed + 1, device, prefix="first_stage_model."))
    return s


def rar_state(cfg=RAR_XL_CFG, seed=0, device="cpu", prefix=""):
    """RAR generator (deps/rar/modeling/rar.py:186-260).  The reference zero-inits the adaLN layers; random values are
    used instead so that every term of the step is exercised."""
    G = _Gen(seed, device)
    d, depth, heads, mlp = cfg["hidden_size"], cfg["num_hidden_layers"], cfg["num_attention_heads"], cfg["intermediate_size"]
    codebook, n_cls, seq = cfg["codebook_size"], cfg["condition_num_classes"], cfg["image_seq_len"]
    hd = d // heads
    w = {"cls_token": G.rn(1, 1, d, std=0.02), "embeddings.weight": G.rn(codebook + 1 + n_cls + 1, d, std=0.02),
         "pos_embed": G.rn(1, seq + 1024, d, std=0.02), "target_aware_pos_embed": G.rn(1, seq + 1024, d, std=0.02),
         "timesteps_embeddings": G.rn(1, seq + 100, d, std=0.02)}
    for i in range(depth):
        p = f"blocks.{i}."
        for nm in ("norm1", "norm2"):
            w[p + nm + ".weight"] = G.rn(d, std=0.1, mean=1.0)
            w[p + nm + ".bias"] = G.rn(d, std=0.05)
        w[p + "attn.qkv.weight"] = G.rn(3 * d, d, std=0.02)
        w[p + "attn.qkv.bias"] = G.rn(3 * d, std=0.01)
        for nm in ("q_norm", "k_norm"):
            w[p + f"attn.{nm}.weight"] = G.rn(hd, std=0.1, mean=1.0)
            w[p + f"attn.{nm}.bias"] = G.rn(hd, std=0.05)
        w[p + "attn.proj.weight"] = G.rn(d, d, std=0.02)
        w[p + "attn.proj.bias"] = G.rn(d, std=0.01)
        w[p + "mlp.fc1.weight"] = G.rn(mlp, d, std=0.02)
        w[p + "mlp.fc1.bias"] = G.rn(mlp, std=0.01)
        w[p + "mlp.fc2.weight"] = G.rn(d, mlp, std=0.02)
        w[p + "mlp.fc2.bias"] = G.rn(d, std=0.01)
        w[p + "adaLN_modulation.1.weight"] = G.rn(6 * d, d, std=0.02)
        w[p + "adaLN_modulation.1.bias"] = G.rn(6 * d, std=0.05)
    w["adaln_before_head.adaLN_modulation.1.weight"] = G.rn(2 * d, d, std=0.02)
    w["adaln_before_head.adaLN_modulation.1.bias"] = G.rn(2 * d, std=0.05)
    w["lm_head.weight"] = G.rn(codebook, d, std=0.02)
    w["lm_head.bias"] = G.rn(codebook, std=0.01)
    return {prefix + k: v for k, v in w.items()}


# Anole-7B / Chameleon-7B (deps/chameleon/inference/transformer.py ModelArgs as loaded from params.json of the 7B
# checkpoint: dim 4096, 32 layers, 32 heads, ffn hidden 11008, vocab 65536, qk_normalization, no swin norm) and its
# 512-pixel VQGAN (deps/chameleon/inference/vqgan.py ddconfig: 6 levels, attention only in `mid`, 8192 codes)
ANOLE_7B_CFG = dict(vocab_size=65536, dim=4096, n_layers=32, n_heads=32, n_kv_heads=32, ffn_hidden=11008,
                    norm_eps=1e-5, rope_theta=10000.0, qk_normalization=True)
CHAMELEON_VQGAN_DDCONFIG = dict(ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(),
                                in_channels=3, resolution=512, z_channels=256, n_embed=8192, embed_dim=256)


def chameleon_state(cfg=ANOLE_7B_CFG, seed=0, device="cpu"):
    """bf16 state dict with the reference Transformer's parameter names; created layer by layer on `device`."""
    G = _Gen(seed, device)
    bf = torch.bfloat16
    V, d, L, H, Hkv, F = cfg["vocab_size"], cfg["dim"], cfg["n_layers"], cfg["n_heads"], cfg["n_kv_heads"], cfg["ffn_hidden"]
    hd = d // H
    w = {"tok_embeddings.weight": G.rn(V, d, std=1.0).to(bf), "norm.weight": G.rn(d, std=0.1, mean=1.0).to(bf),
         "output.weight": G.rn(V, d, std=0.02).to(bf)}
    for i in range(L):
        p = f"layers.{i}."
        w[p + "attention_norm.weight"] = G.rn(d, std=0.1, mean=1.0).to(bf)
        w[p + "ffn_norm.weight"] = G.rn(d, std=0.1, mean=1.0).to(bf)
        w[p + "attention.wqkv.weight"] = G.rn((H + 2 * Hkv) * hd, d, std=0.02).to(bf)
        w[p + "attention.wo.weight"] = G.rn(d, H * hd, std=0.02).to(bf)
        if cfg.get("qk_normalization", True):
            for nm in ("q_normalization", "k_normalization"):
                w[p + f"attention.{nm}.weight"] = G.rn(hd, std=0.1, mean=1.0).to(bf)
                w[p + f"attention.{nm}.bias"] = G.rn(hd, std=0.05).to(bf)
        w[p + "feed_forward.w13.weight"] = G.rn(2 * F, d, std=0.02).to(bf)
        w[p + "feed_forward.w2.weight"] = G.rn(d, F, std=0.02).to(bf)
    return w
