"""Host handle of the VQGAN tokenizer engine (wmar_vqgan_* in include/wmar_b200.h) and its weight packer.

The packer walks the reference architecture in exactly the order documented at the top of csrc/vqgan.cu and turns
the reference state dict (Taming VQModel: encoder.*, decoder.*, quantize.embedding.weight, quant_conv.*,
post_quant_conv.*; MaskGIT PretrainedTokenizer: encoder.*, decoder.*, quantize.embedding.weight) into a table of
device pointers: conv weights transposed to [Cout_pad64][ky][kx][Cin_pad32] (NHWC implicit-GEMM layout), biases padded.
"""
import ctypes

import torch

from .. import _lib

TAMING_CFG = dict(family=0, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolution=16, resolution=256,
                  z_channels=256, embed_dim=256, n_embed=16384)
MASKGIT_CFG = dict(family=1, ch=128, ch_mult=(1, 1, 2, 2, 4), num_res_blocks=2, attn_resolution=0, resolution=256,
                   z_channels=256, embed_dim=256, n_embed=1024)


def _rup(v, m):
    return (v + m - 1) // m * m


class _Packer:
    def __init__(self, state, device):
        self.s, self.dev, self.out = state, device, []

    def raw(self, key):
        self.out.append(self.s[key].detach().to(self.dev, torch.float32).contiguous())

    def conv(self, prefix):
        w = self.s[prefix + ".weight"].detach().to(self.dev, torch.float32)
        cout, cin, kh, kw = w.shape
        wp = torch.zeros((_rup(cout, 64), kh, kw, _rup(cin, 32)), device=self.dev, dtype=torch.float32)
        wp[:cout, :, :, :cin] = w.permute(0, 2, 3, 1)
        bp = torch.zeros(_rup(cout, 64), device=self.dev, dtype=torch.float32)
        if (prefix + ".bias") in self.s:
            bp[:cout] = self.s[prefix + ".bias"].detach().to(self.dev, torch.float32)
        self.out += [wp.contiguous(), bp]

    def norm(self, prefix):
        self.raw(prefix + ".weight")
        self.raw(prefix + ".bias")

    def res(self, prefix):
        self.norm(prefix + ".norm1")
        self.conv(prefix + ".conv1")
        self.norm(prefix + ".norm2")
        self.conv(prefix + ".conv2")
        if (prefix + ".nin_shortcut.weight") in self.s:
            self.conv(prefix + ".nin_shortcut")

    def attn(self, prefix):
        self.norm(prefix + ".norm")
        for n in ("q", "k", "v", "proj_out"):
            self.conv(prefix + "." + n)


def pack_vqgan(state, cfg, device):
    p = _Packer(state, device)
    nl, nrb = len(cfg["ch_mult"]), cfg["num_res_blocks"]
    p.raw("quantize.embedding.weight")
    if cfg["family"] == 0:
        p.conv("encoder.conv_in")
        res = cfg["resolution"]
        for l in range(nl):
            for k in range(nrb):
                p.res(f"encoder.down.{l}.block.{k}")
                if res == cfg["attn_resolution"]:
                    p.attn(f"encoder.down.{l}.attn.{k}")
            if l != nl - 1:
                p.conv(f"encoder.down.{l}.downsample.conv")
                res //= 2
        p.res("encoder.mid.block_1")
        p.attn("encoder.mid.attn_1")
        p.res("encoder.mid.block_2")
        p.norm("encoder.norm_out")
        p.conv("encoder.conv_out")
        p.conv("quant_conv")
        p.conv("post_quant_conv")
        p.conv("decoder.conv_in")
        p.res("decoder.mid.block_1")
        p.attn("decoder.mid.attn_1")
        p.res("decoder.mid.block_2")
        for l in reversed(range(nl)):
            for k in range(nrb + 1):
                p.res(f"decoder.up.{l}.block.{k}")
                if res == cfg["attn_resolution"]:
                    p.attn(f"decoder.up.{l}.attn.{k}")
            if l != 0:
                p.conv(f"decoder.up.{l}.upsample.conv")
                res *= 2
        p.norm("decoder.norm_out")
        p.conv("decoder.conv_out")
    else:
        p.conv("encoder.conv_in")
        for l in range(nl):
            for k in range(nrb):
                p.res(f"encoder.down.{l}.block.{k}")
        for k in range(nrb):
            p.res(f"encoder.mid.{k}")
        p.norm("encoder.norm_out")
        p.conv("encoder.conv_out")
        p.conv("decoder.conv_in")
        for k in range(nrb):
            p.res(f"decoder.mid.{k}")
        for l in reversed(range(nl)):
            for k in range(nrb):
                p.res(f"decoder.up.{l}.block.{k}")
            if l != 0:
                p.conv(f"decoder.up.{l}.upsample_conv")
        p.norm("decoder.norm_out")
        p.conv("decoder.conv_out")
    return p.out


class VQGANEngine:
    def __init__(self, state, cfg, device="cuda", max_batch=16, precision="3xtf32"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WmarError("VQGANEngine is CUDA only (no CPU fallback)")
        self.cfg = dict(cfg)
        self.max_batch = max_batch
        self.precision = {"3xtf32": 0, "tf32": 1, "bf16x3": 2, "bf16x3-dec": 3}[precision]
        self.handle = None
        self.latent = cfg["resolution"] // 2 ** (len(cfg["ch_mult"]) - 1)
        self.sync_weights(state)

    def sync_weights(self, state):
        self._tensors = pack_vqgan(state, self.cfg, self.device)
        L = _lib.lib()
        if self.handle is not None:
            L.wmar_vqgan_destroy(self.handle)
            self.handle = None
        c = self.cfg
        mult = (ctypes.c_int * 8)(*(list(c["ch_mult"]) + [0] * (8 - len(c["ch_mult"]))))
        cc = _lib.VqganConfig(c["family"], c["ch"], len(c["ch_mult"]), mult, c["num_res_blocks"], c["attn_resolution"],
                              c["resolution"], c["z_channels"], c["embed_dim"], c["n_embed"], self.max_batch,
                              self.precision)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.wmar_vqgan_create(ctypes.byref(cc), _lib.pointer_table(self._tensors), len(self._tensors),
                                           ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if self.handle is not None:
                _lib.lib().wmar_vqgan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @torch.no_grad()
    def decode(self, codes):
        """codes int64 [B, s*s] -> images fp32 [B, 3, S, S] in [-1, 1]"""
        codes = codes.to(self.device, torch.long).contiguous()
        B = codes.shape[0]
        R = self.cfg["resolution"]
        out = torch.empty((B, 3, R, R), dtype=torch.float32, device=self.device)
        for i in range(0, B, self.max_batch):
            n = min(self.max_batch, B - i)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().wmar_vqgan_decode(self.handle, _lib.ptr(codes[i:i + n]), n,
                                                        _lib.ptr(out[i:i + n]), _lib.current_stream()))
        return out

    @torch.no_grad()
    def encode(self, images):
        """images fp32 [B, 3, S, S] in [-1, 1] -> codes int64 [B, s*s]"""
        images = images.to(self.device, torch.float32).contiguous()
        B = images.shape[0]
        out = torch.empty((B, self.latent * self.latent), dtype=torch.long, device=self.device)
        for i in range(0, B, self.max_batch):
            n = min(self.max_batch, B - i)
            with torch.cuda.device(self.device):
                _lib.check(_lib.lib().wmar_vqgan_encode(self.handle, _lib.ptr(images[i:i + n]), n,
                                                        _lib.ptr(out[i:i + n]), _lib.current_stream()))
        return out

    def flops(self, decode=True):
        return float(_lib.lib().wmar_vqgan_flops(self.handle, 1 if decode else 0))
