"""Host-side mirror of ``wmar.models.taming_wrapper.TamingARMMWrapper`` (taming_wrapper.py:22-92).

``sample`` replaces ``taming_sample_with_past`` (mingpt.py:326-368) by the decode engine (one C-ABI call for all 256
tokens, watermark and sampler fused behind the lm_head); ``codes_to_images`` / ``images_to_codes`` replace
``Net2NetTransformer.decode_to_img`` / ``encode_to_z`` (cond_transformer.py:169-192) by the VQGAN engine.
"""
import os

import torch

from .. import _lib
from .armm_wrapper import AutoregressiveMultimodalModelWrapper
from .gpt_engine import TamingGPTEngine
from .state import StateModule, split_prefix
from .synthetic import TAMING_GPT_CFG, TAMING_VQGAN_DDCONFIG, taming_net2net_state
from .vqgan_engine import VQGANEngine

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def _load_net2net(modelpath):
    """configs/net2net.yaml + checkpoints/net2net.ckpt (taming_wrapper.py:26-34), without omegaconf/lightning."""
    import yaml
    with open(os.path.join(modelpath, "configs/net2net.yaml")) as f:
        cfg = yaml.safe_load(f)["model"]["params"]
    t = cfg["transformer_config"]["params"]
    gpt_cfg = dict(vocab_size=t["vocab_size"], block_size=t["block_size"], n_layer=t["n_layer"], n_head=t["n_head"],
                   n_embd=t["n_embd"])
    fs = cfg["first_stage_config"]["params"]
    dd = dict(fs["ddconfig"], n_embed=fs["n_embed"], embed_dim=fs["embed_dim"])
    dd["ch_mult"] = tuple(dd["ch_mult"])
    dd["attn_resolutions"] = tuple(dd["attn_resolutions"])
    state = torch.load(os.path.join(modelpath, "checkpoints/net2net.ckpt"), map_location="cpu", weights_only=False)
    if "state_dict" in state:
        state = state["state_dict"]
    return state, gpt_cfg, dd


class TamingARMMWrapper(AutoregressiveMultimodalModelWrapper):
    def __init__(self, modelpath=None, *, state_dict=None, gpt_cfg=None, dd_cfg=None, device="cuda", max_batch=16,
                 vqgan_precision="bf16x3", seed=0, alive_ids_path=None, rng="torch"):
        """modelpath: directory of the reference's Taming download (README.md); None -> seeded random-init weights at
        ``gpt_cfg`` / ``dd_cfg`` shapes (default: the reference's cin_transformer shapes), or an explicit
        ``state_dict`` with Net2NetTransformer keys.  rng: "torch" draws torch.multinomial's own CUDA Philox stream inside
        the sampler kernel (same seeds -> same tokens as the reference, no noise buffer), "torch_buffer" lets torch pre-draw
        that stream into a [steps, B, V] buffer (round 1), "philox" uses an independent in-kernel stream."""
        super().__init__()
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.WmarError("TamingARMMWrapper runs on CUDA only (no CPU fallback)")
        if modelpath is not None:
            state_dict, gpt_cfg, dd_cfg = _load_net2net(modelpath)
        gpt_cfg = dict(gpt_cfg or TAMING_GPT_CFG)
        dd_cfg = dict(dd_cfg or TAMING_VQGAN_DDCONFIG)
        if state_dict is None:
            state_dict = taming_net2net_state(gpt_cfg, dd_cfg, seed=seed, device=self._device)
        self.gpt_cfg, self.dd_cfg = gpt_cfg, dd_cfg
        self.rng = rng
        self.max_batch = max_batch
        self.vqgan_precision = vqgan_precision
        self.model = StateModule({k: v for k, v in state_dict.items()
                                  if k.startswith(("transformer.", "first_stage_model."))
                                  and not k.startswith("first_stage_model.loss")
                                  and not k.endswith(".attn.mask")}).to(self._device)
        self.model.eval()
        vq = self.get_vq()
        vq.n_e = dd_cfg["n_embed"]
        self.init_alivecodes(alive_ids_path or os.path.join(ASSETS, "vqgan_alive_ids.txt"))
        nres = len(dd_cfg["ch_mult"])
        self.codes_size = dd_cfg["resolution"] // 2 ** (nres - 1)
        self.image_size = dd_cfg["resolution"]
        self.dim_z = dd_cfg["embed_dim"]
        self._gpt = None
        self._vqgan = None
        self._step_seed = seed
        self.sync_weights()

    def __repr__(self):
        return "TamingARMMWrapper"

    def get_image_tokenizer(self):
        return self.model.first_stage_model

    def get_total_vocab_size(self):
        return self.get_vq().n_e

    def sync_weights(self):
        """(Re-)pack the weights of the nn.Module tree into the engines -- call after update_weights()."""
        flat = dict(self.model.state_dict())
        gstate = split_prefix(flat, "transformer.")
        vstate = split_prefix(flat, "first_stage_model.")
        if self._gpt is None:
            self._gpt = TamingGPTEngine(gstate, self.gpt_cfg["n_layer"], self.gpt_cfg["n_head"], device=self._device,
                                        max_batch=self.max_batch)
        else:
            self._gpt.sync_weights(gstate)
        dd = self.dd_cfg
        ecfg = dict(family=0, ch=dd["ch"], ch_mult=tuple(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"],
                    attn_resolution=(dd["attn_resolutions"][0] if dd["attn_resolutions"] else 0),
                    resolution=dd["resolution"], z_channels=dd["z_channels"], embed_dim=dd["embed_dim"],
                    n_embed=dd["n_embed"])
        if self._vqgan is None:
            self._vqgan = VQGANEngine(vstate, ecfg, device=self._device, max_batch=self.max_batch,
                                      precision=self.vqgan_precision)
        else:
            self._vqgan.sync_weights(vstate)

    # conditioning: list of size [b] (class ids).  Returns detached codes [b, codes_size**2]  (taming_wrapper.py:61-77)
    def sample(self, conditioning, gen_params, apply_watermark=False, greedy=False):
        cond = torch.as_tensor(conditioning, device=self.device).view(-1).long()
        steps = self.codes_size * self.codes_size
        wm = self.watermarker if apply_watermark else None
        out = []
        for i in range(0, cond.numel(), self.max_batch):
            c = cond[i:i + self.max_batch]
            noise, stream = None, None
            if not greedy and self.rng == "torch":            # torch.multinomial's draws, generated inside the sampler
                stream = self._torch_stream(steps, c.numel(), self.gpt_cfg["vocab_size"])
            elif not greedy and self.rng == "torch_buffer":   # the same draws, pre-drawn by torch into a buffer
                noise = self._draw_noise(steps, c.numel(), self.gpt_cfg["vocab_size"])
            self._step_seed += 1
            out.append(self._gpt.sample(c, steps, temperature=gen_params["temperature"], top_k=gen_params["top_k"],
                                        top_p=gen_params["top_p"], watermarker=wm, noise=noise, greedy=greedy,
                                        seed=self._step_seed, torch_stream=stream))
        codes = out[0] if len(out) == 1 else torch.cat(out, dim=0)
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes

    def codes_to_images(self, codes):
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        images = self._vqgan.decode(codes)
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        return images

    def images_to_codes(self, images):
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        codes = self._vqgan.encode(images)
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes
