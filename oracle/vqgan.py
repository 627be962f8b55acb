"""VQGAN tokenizers restated in functional torch fp32 on CPU (TEST INFRASTRUCTURE ONLY).

Taming / Chameleon VQGAN: deps/taming/modules/diffusionmodules/model.py (ResnetBlock :79-138, AttnBlock :141-193,
  Up/Downsample :39-76, Encoder :343-434, Decoder :437-538), deps/taming/models/vqgan.py:64-73 (quant convs),
  deps/taming/modules/vqvae/quantize.py:272-331 (VectorQuantizer2 argmin / get_codebook_entry),
  deps/taming/models/cond_transformer.py:169-192 (encode_to_z / decode_to_img), wmar/models/taming_wrapper.py:79-92.
MaskGIT-VQGAN (RAR): deps/rar/modeling/modules/maskgit_vqgan.py (ResnetBlock :54-90 incl. the shortcut quirk :87-88,
  Encoder :160-197, Decoder :200-245, VectorQuantizer :248-321), deps/rar/modeling/titok.py:75-85,
  wmar/models/rar_wrapper.py:109-128.
Weights: dicts keyed like the reference modules' state_dict.
"""
import torch
import torch.nn.functional as F

TAMING_CFG = dict(ch=128, out_ch=3, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolutions=(16,),
                  in_channels=3, resolution=256, z_channels=256, n_embed=16384, embed_dim=256)
MASKGIT_CFG = dict(hidden_channels=128, channel_mult=(1, 1, 2, 2, 4), num_res_blocks=2, num_channels=3,
                   resolution=256, z_channels=256, num_embeddings=1024)


def _gn(x, w, p):
    return F.group_norm(x, 32, w[p + ".weight"], w[p + ".bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def _conv(x, w, p, stride=1, padding=1):
    return F.conv2d(x, w[p + ".weight"], w.get(p + ".bias"), stride=stride, padding=padding)


# ----------------------------------------------------------------------------------------------- Taming
def _t_res(x, w, p):
    h = _conv(_swish(_gn(x, w, p + ".norm1")), w, p + ".conv1")
    h = _conv(_swish(_gn(h, w, p + ".norm2")), w, p + ".conv2")
    if (p + ".nin_shortcut.weight") in w:
        x = _conv(x, w, p + ".nin_shortcut", padding=0)
    return x + h


def _t_attn(x, w, p):
    h = _gn(x, w, p + ".norm")
    q = _conv(h, w, p + ".q", padding=0)
    k = _conv(h, w, p + ".k", padding=0)
    v = _conv(h, w, p + ".v", padding=0)
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    a = torch.bmm(q, k) * (int(c) ** (-0.5))
    a = F.softmax(a, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, a.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(h, w, p + ".proj_out", padding=0)


def taming_encoder(x, w, cfg, prefix="encoder"):
    nres = len(cfg["ch_mult"])
    res = cfg["resolution"]
    h = _conv(x, w, prefix + ".conv_in")
    for lvl in range(nres):
        for b in range(cfg["num_res_blocks"]):
            h = _t_res(h, w, f"{prefix}.down.{lvl}.block.{b}")
            if res in cfg["attn_resolutions"]:
                h = _t_attn(h, w, f"{prefix}.down.{lvl}.attn.{b}")
        if lvl != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = _conv(h, w, f"{prefix}.down.{lvl}.downsample.conv", stride=2, padding=0)
            res //= 2
    h = _t_res(h, w, prefix + ".mid.block_1")
    h = _t_attn(h, w, prefix + ".mid.attn_1")
    h = _t_res(h, w, prefix + ".mid.block_2")
    return _conv(_swish(_gn(h, w, prefix + ".norm_out")), w, prefix + ".conv_out")


def taming_decoder(z, w, cfg, prefix="decoder"):
    nres = len(cfg["ch_mult"])
    res = cfg["resolution"] // 2 ** (nres - 1)
    h = _conv(z, w, prefix + ".conv_in")
    h = _t_res(h, w, prefix + ".mid.block_1")
    h = _t_attn(h, w, prefix + ".mid.attn_1")
    h = _t_res(h, w, prefix + ".mid.block_2")
    for lvl in reversed(range(nres)):
        for b in range(cfg["num_res_blocks"] + 1):
            h = _t_res(h, w, f"{prefix}.up.{lvl}.block.{b}")
            if res in cfg["attn_resolutions"]:
                h = _t_attn(h, w, f"{prefix}.up.{lvl}.attn.{b}")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(h, w, f"{prefix}.up.{lvl}.upsample.conv")
            res *= 2
    return _conv(_swish(_gn(h, w, prefix + ".norm_out")), w, prefix + ".conv_out")


def vq2_argmin(z, emb):
    """quantize.py:277-285; z fp32[B,C,H,W] -> indices int64[B*H*W] (first index on ties)."""
    zf = z.permute(0, 2, 3, 1).contiguous().view(-1, emb.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1) - 2 * torch.einsum(
        "bd,dn->bn", zf, emb.t())
    return torch.argmin(d, dim=1)


@torch.no_grad()
def taming_images_to_codes(images, w, cfg=TAMING_CFG):
    h = taming_encoder(images, w, cfg)
    h = _conv(h, w, "quant_conv", padding=0)
    return vq2_argmin(h, w["quantize.embedding.weight"]).view(images.shape[0], -1)


@torch.no_grad()
def taming_codes_to_images(codes, w, cfg=TAMING_CFG):
    B, L = codes.shape
    s = int(round(L ** 0.5))
    zq = w["quantize.embedding.weight"][codes.reshape(-1)].view(B, s, s, -1).permute(0, 3, 1, 2).contiguous()
    zq = _conv(zq, w, "post_quant_conv", padding=0)
    return taming_decoder(zq, w, cfg).clamp(-1, 1)


def synthetic_taming_vqgan_weights(cfg=TAMING_CFG, seed=0):
    """Seeded synthetic VQModel state_dict at the reference's shapes.  Conv weights ~ N(0, 1/sqrt(fan_in)) so that
    activations stay O(1); GroupNorm affine near identity; codebook ~ N(0,1) (the reference's default
    uniform(+-1/n_e) init makes the fp32 argmin pure rounding noise, which no implementation can match)."""
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(p, cout, cin, k, bias=True):
        w[p + ".weight"] = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        if bias:
            w[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def norm(p, c):
        w[p + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        w[p + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cin, 1)

    def attn(p, c):
        norm(p + ".norm", c)
        for nm in ("q", "k", "v", "proj_out"):
            conv(p + "." + nm, c, c, 1)

    ch, mult, nrb = cfg["ch"], cfg["ch_mult"], cfg["num_res_blocks"]
    nres = len(mult)
    # encoder
    conv("encoder.conv_in", ch, cfg["in_channels"], 3)
    res_now = cfg["resolution"]
    in_mult = (1,) + tuple(mult)
    for lvl in range(nres):
        bin_, bout = ch * in_mult[lvl], ch * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
            if res_now in cfg["attn_resolutions"]:
                attn(f"encoder.down.{lvl}.attn.{b}", bin_)
        if lvl != nres - 1:
            conv(f"encoder.down.{lvl}.downsample.conv", bin_, bin_, 3)
            res_now //= 2
    res("encoder.mid.block_1", bin_, bin_)
    attn("encoder.mid.attn_1", bin_)
    res("encoder.mid.block_2", bin_, bin_)
    norm("encoder.norm_out", bin_)
    conv("encoder.conv_out", cfg["z_channels"], bin_, 3)
    # decoder
    bin_ = ch * mult[-1]
    res_now = cfg["resolution"] // 2 ** (nres - 1)
    conv("decoder.conv_in", bin_, cfg["z_channels"], 3)
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
            conv(f"decoder.up.{lvl}.upsample.conv", bin_, bin_, 3)
            res_now *= 2
    norm("decoder.norm_out", bin_)
    conv("decoder.conv_out", cfg["out_ch"], bin_, 3)
    w["quantize.embedding.weight"] = torch.randn(cfg["n_embed"], cfg["embed_dim"], generator=g)
    conv("quant_conv", cfg["embed_dim"], cfg["z_channels"], 1)
    conv("post_quant_conv", cfg["z_channels"], cfg["embed_dim"], 1)
    return w


# ----------------------------------------------------------------------------------------------- MaskGIT (RAR)
def _same_conv(x, w, p):
    k = w[p + ".weight"].shape[-1]
    return F.conv2d(x, w[p + ".weight"], w.get(p + ".bias"), padding=k // 2)  # stride 1: 'same' == k//2 each side


def _m_res(x, w, p):
    h = _same_conv(F.silu(_gn(x, w, p + ".norm1")), w, p + ".conv1")
    h = _same_conv(F.silu(_gn(h, w, p + ".norm2")), w, p + ".conv2")
    if (p + ".nin_shortcut.weight") in w:
        x = _same_conv(h, w, p + ".nin_shortcut")  # quirk: shortcut applied to the post-conv activations (:87-88)
    return h + x


def maskgit_encoder(x, w, cfg, prefix="encoder"):
    nres = len(cfg["channel_mult"])
    h = _same_conv(x, w, prefix + ".conv_in")
    for lvl in range(nres):
        for b in range(cfg["num_res_blocks"]):
            h = _m_res(h, w, f"{prefix}.down.{lvl}.block.{b}")
        if lvl != nres - 1:
            h = F.avg_pool2d(h, kernel_size=2, stride=2)
    for b in range(cfg["num_res_blocks"]):
        h = _m_res(h, w, f"{prefix}.mid.{b}")
    return _same_conv(F.silu(_gn(h, w, prefix + ".norm_out")), w, prefix + ".conv_out")


def maskgit_decoder(z, w, cfg, prefix="decoder"):
    nres = len(cfg["channel_mult"])
    h = _same_conv(z, w, prefix + ".conv_in")
    for b in range(cfg["num_res_blocks"]):
        h = _m_res(h, w, f"{prefix}.mid.{b}")
    for lvl in reversed(range(nres)):
        for b in range(cfg["num_res_blocks"]):
            h = _m_res(h, w, f"{prefix}.up.{lvl}.block.{b}")
        if lvl != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _same_conv(h, w, f"{prefix}.up.{lvl}.upsample_conv")
    return _same_conv(F.silu(_gn(h, w, prefix + ".norm_out")), w, prefix + ".conv_out")


def maskgit_argmin(z, emb):
    """maskgit_vqgan.py:308-321 + :283-284: addmm(|z|^2 + |e|^2, z, e^T, alpha=-2) -> argmin."""
    zf = z.permute(0, 2, 3, 1).contiguous().reshape(-1, emb.shape[1])
    et = emb.t()
    d = torch.addmm(zf.pow(2.0).sum(dim=1, keepdim=True) + et.pow(2.0).sum(dim=0, keepdim=True), zf, et, alpha=-2.0)
    return torch.argmin(d, dim=1)


@torch.no_grad()
def rar_images_to_codes(images, w, cfg=MASKGIT_CFG):
    x = (images + 1.0) / 2.0
    h = maskgit_encoder(x, w, cfg)
    return maskgit_argmin(h, w["quantize.embedding.weight"]).view(images.shape[0], -1)


@torch.no_grad()
def rar_codes_to_images(codes, w, cfg=MASKGIT_CFG):
    B, L = codes.shape
    s = int(round(L ** 0.5))
    zq = w["quantize.embedding.weight"][codes].reshape(B, s, s, -1).permute(0, 3, 1, 2)
    img = torch.clamp(maskgit_decoder(zq, w, cfg), 0.0, 1.0)
    return torch.clamp(img * 2.0 - 1.0, -1.0, 1.0)


def synthetic_maskgit_weights(cfg=MASKGIT_CFG, seed=0):
    g = torch.Generator().manual_seed(seed)
    w = {}

    def conv(p, cout, cin, k, bias):
        w[p + ".weight"] = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
        if bias:
            w[p + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def norm(p, c):
        w[p + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=g)
        w[p + ".bias"] = 0.05 * torch.randn(c, generator=g)

    def res(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3, False)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3, False)
        if cin != cout:
            conv(p + ".nin_shortcut", cout, cout, 1, False)

    hc, mult, nrb = cfg["hidden_channels"], cfg["channel_mult"], cfg["num_res_blocks"]
    nres = len(mult)
    in_mult = (1,) + tuple(mult)
    conv("encoder.conv_in", hc, cfg["num_channels"], 3, False)
    for lvl in range(nres):
        bin_, bout = hc * in_mult[lvl], hc * mult[lvl]
        for b in range(nrb):
            res(f"encoder.down.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
    mid = hc * mult[-1]
    for b in range(nrb):
        res(f"encoder.mid.{b}", mid, mid)
    norm("encoder.norm_out", mid)
    conv("encoder.conv_out", cfg["z_channels"], mid, 1, True)
    conv("decoder.conv_in", mid, cfg["z_channels"], 3, True)
    for b in range(nrb):
        res(f"decoder.mid.{b}", mid, mid)
    for lvl in reversed(range(nres)):
        bin_ = hc * mult[-1] if lvl == nres - 1 else hc * mult[lvl + 1]
        bout = hc * mult[lvl]
        for b in range(nrb):
            res(f"decoder.up.{lvl}.block.{b}", bin_, bout)
            bin_ = bout
        if lvl != 0:
            conv(f"decoder.up.{lvl}.upsample_conv", bout, bout, 3, True)
    norm("decoder.norm_out", hc * mult[0])
    conv("decoder.conv_out", cfg["num_channels"], hc * mult[0], 3, True)
    w["quantize.embedding.weight"] = torch.randn(cfg["num_embeddings"], cfg["z_channels"], generator=g)
    return w
