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
                 vqgan_precision="bf16x3", seed=0, alive_ids_path=None, rng="torch", lanes=3):
        """modelpath: directory of the reference's Taming download (README.md); None -> seeded random-init weights at
        ``gpt_cfg`` / ``dd_cfg`` shapes (default: the reference's cin_transformer shapes), or an explicit
        ``state_dict`` with Net2NetTransformer keys.  rng: "torch" draws torch.multinomial's own CUDA Philox stream inside
        the sampler kernel (same seeds -> same tokens as the reference, no noise buffer), "torch_buffer" lets torch pre-draw
        that stream into a [steps, B, V] buffer (round 1), "philox" uses an independent in-kernel stream.
        lanes: how many engine lanes (KV cache + scratch + step graph each, weights shared) ``sample`` may run concurrently
        when it is given more than ``max_batch`` conditionings: chunk i goes to lane i % lanes on that lane's CUDA stream.
        Results are identical to the sequential chunk loop (same per-chunk Philox offsets / seeds); lanes=1 is that loop."""
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
        self.lanes = max(1, int(lanes))
        self._lane_engines, self._lane_streams = [], []
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
        self._lane_engines, self._lane_streams = [], []   # lanes borrow the engine's tensors: rebuilt on demand
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

    def _lane(self, k):
        """Engine + stream of lane k (lane 0 = the wrapper's own engine on the caller's stream)."""
        if k == 0:
            return self._gpt, None
        while len(self._lane_engines) < k:
            self._lane_engines.append(self._gpt.clone_lane())
            self._lane_streams.append(torch.cuda.Stream(device=self.device))
        return self._lane_engines[k - 1], self._lane_streams[k - 1]

    # conditioning: list of size [b] (class ids).  Returns detached codes [b, codes_size**2]  (taming_wrapper.py:61-77)
    def sample(self, conditioning, gen_params, apply_watermark=False, greedy=False):
        cond = torch.as_tensor(conditioning, device=self.device).view(-1).long()
        steps = self.codes_size * self.codes_size
        wm = self.watermarker if apply_watermark else None
        n_chunks = (cond.numel() + self.max_batch - 1) // self.max_batch
        n_lanes = min(self.lanes, n_chunks)
        if n_lanes - 1 > len(self._lane_engines):
            # create the missing lanes BEFORE any work of this call is enqueued: engine creation zero-fills its buffers on
            # the legacy default stream, which would queue behind lane 0's generation and land in the middle of lane 1's
            self._lane(n_lanes - 1)
            torch.cuda.synchronize(self.device)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)                                # everything the caller enqueued before this call (cond, weights)
        out = []
        for ci, i in enumerate(range(0, cond.numel(), self.max_batch)):
            c = cond[i:i + self.max_batch]
            noise, stream = None, None
            if not greedy and self.rng == "torch":            # torch.multinomial's draws, generated inside the sampler
                stream = self._torch_stream(steps, c.numel(), self.gpt_cfg["vocab_size"])
            elif not greedy and self.rng == "torch_buffer":   # the same draws, pre-drawn by torch into a buffer
                noise = self._draw_noise(steps, c.numel(), self.gpt_cfg["vocab_size"])
            self._step_seed += 1
            eng, lane_stream = self._lane(ci % n_lanes)
            kw = dict(temperature=gen_params["temperature"], top_k=gen_params["top_k"], top_p=gen_params["top_p"],
                      watermarker=wm, noise=noise, greedy=greedy, seed=self._step_seed, torch_stream=stream)
            if lane_stream is None:
                out.append(eng.sample(c, steps, defer_check=n_lanes > 1, **kw))
            else:
                # NOT wait_stream(main): lane 0's generation was just enqueued there and the lanes must overlap it
                if noise is not None:
                    lane_stream.wait_stream(main)         # (legacy torch_buffer mode: the noise was drawn on `main`)
                else:
                    lane_stream.wait_event(ready)
                with torch.cuda.stream(lane_stream):
                    o = eng.sample(c, steps, defer_check=True, **kw)
                o.record_stream(main)
                out.append(o)
        if n_lanes > 1:
            for k in range(1, n_lanes):
                main.wait_stream(self._lane_streams[k - 1])
            _lib.check_device_flag()
        codes = out[0] if len(out) == 1 else torch.cat(out, dim=0)
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes

    def codes_to_images(self, codes):
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        mb = self.max_batch
        images = self._vqgan.decode(codes) if codes.shape[0] <= mb else torch.cat(
            [self._vqgan.decode(codes[i:i + mb]) for i in range(0, codes.shape[0], mb)], dim=0)
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        return images

    def images_to_codes(self, images):
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        mb = self.max_batch
        codes = self._vqgan.encode(images) if images.shape[0] <= mb else torch.cat(
            [self._vqgan.encode(images[i:i + mb]) for i in range(0, images.shape[0], mb)], dim=0)
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes
