"""Host-side mirror of ``wmar.models.rar_wrapper.RarARMMWrapper`` (rar_wrapper.py:17-128).

``sample`` replaces ``RAR.generate`` (guidance_scale 4.0, constant decay, temperature 1.0, rar_wrapper.py:93-102) by the
RAR decode engine; ``codes_to_images`` / ``images_to_codes`` replace the MaskGIT-VQGAN ``PretrainedTokenizer``
(titok.py:41-89) by the VQGAN engine (family 1).  Images cross the boundary in [-1, 1] like the reference's.
"""
import os

import torch

from .. import _lib
from .armm_wrapper import AutoregressiveMultimodalModelWrapper
from .rar_engine import RAR_SIZES, RAREngine
from .state import StateModule
from .synthetic import MASKGIT_VQGAN_CFG, maskgit_vqgan_state, rar_state
from .vqgan_engine import VQGANEngine

ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


class RarARMMWrapper(AutoregressiveMultimodalModelWrapper):
    def __init__(self, modelpath=None, rar_size="rar_xl", *, state_dict=None, tokenizer_state_dict=None, rar_cfg=None,
                 vq_cfg=None, device="cuda", max_batch=8, vqgan_precision="bf16x3", seed=0, alive_ids_path=None, lanes=3,
                 rng="torch"):
        """modelpath: directory holding ``{rar_size}.bin`` and ``maskgit-vqgan-imagenet-f16-256.bin`` (the files the
        reference downloads, rar_wrapper.py:27-34); None -> seeded random-init weights at the ``rar_size`` shapes."""
        super().__init__()
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.WmarError("RarARMMWrapper runs on CUDA only (no CPU fallback)")
        self.rar_size = rar_size
        cfg = dict(codebook_size=1024, image_seq_len=256, condition_num_classes=1000)
        cfg.update(RAR_SIZES[rar_size])
        cfg.update(rar_cfg or {})
        vq_cfg = dict(vq_cfg or MASKGIT_VQGAN_CFG)
        if modelpath is not None:
            state_dict = torch.load(os.path.join(modelpath, f"{rar_size}.bin"), map_location="cpu")
            tokenizer_state_dict = torch.load(os.path.join(modelpath, "maskgit-vqgan-imagenet-f16-256.bin"),
                                              map_location="cpu")
        if state_dict is None:
            state_dict = rar_state(cfg, seed=seed, device=self._device)
        if tokenizer_state_dict is None:
            tokenizer_state_dict = maskgit_vqgan_state(vq_cfg, seed=seed + 1, device=self._device)
        self.rar_cfg, self.vq_cfg = cfg, vq_cfg
        self.rng = rng
        self.max_batch = max_batch
        self.vqgan_precision = vqgan_precision
        self.model = StateModule({k: v for k, v in state_dict.items() if k != "attn_mask"}).to(self._device)
        self.model.eval()
        self.tokenizer = StateModule(dict(tokenizer_state_dict)).to(self._device)
        self.tokenizer.eval()
        self.tokenizer.quantize.num_embeddings = vq_cfg["num_embeddings"]
        self.init_alivecodes(alive_ids_path or os.path.join(ASSETS, "rar_all_ids.txt"))
        self.codes_size = int(round(cfg["image_seq_len"] ** 0.5))
        self.image_size = self.codes_size * 16
        self.dim_z = vq_cfg["z_channels"]
        self._rar = None
        self._vqgan = None
        self.lanes = max(1, int(lanes))
        self._lane_engines, self._lane_streams = [], []
        self._step_seed = seed
        self.sync_weights()

    def __repr__(self):
        return "RarARMMWrapper"

    def get_image_tokenizer(self):
        return self.tokenizer

    def get_total_vocab_size(self):
        return self.get_vq().num_embeddings

    def sync_weights(self):
        rstate = dict(self.model.state_dict())
        c = self.rar_cfg
        if self._rar is None:
            self._rar = RAREngine(rstate, c["num_hidden_layers"], c["num_attention_heads"], c["codebook_size"],
                                  c["condition_num_classes"], c["image_seq_len"], device=self._device,
                                  max_batch=min(self.max_batch, 8))
        else:
            self._rar.sync_weights(rstate)
        self._lane_engines, self._lane_streams = [], []   # lanes borrow the engine's tensors: rebuilt on demand
        v = self.vq_cfg
        ecfg = dict(family=1, ch=v["hidden_channels"], ch_mult=tuple(v["channel_mult"]),
                    num_res_blocks=v["num_res_blocks"], attn_resolution=0, resolution=v["resolution"],
                    z_channels=v["z_channels"], embed_dim=v["z_channels"], n_embed=v["num_embeddings"])
        tstate = dict(self.tokenizer.state_dict())
        if self._vqgan is None:
            self._vqgan = VQGANEngine(tstate, ecfg, device=self._device, max_batch=max(self.max_batch, 1),
                                      precision=self.vqgan_precision)
        else:
            self._vqgan.sync_weights(tstate)

    def _lane(self, k):
        """Engine + stream of lane k (lane 0 = the wrapper's own engine on the caller's stream); see TamingARMMWrapper."""
        if k == 0:
            return self._rar, None
        while len(self._lane_engines) < k:
            self._lane_engines.append(self._rar.clone_lane())
            self._lane_streams.append(torch.cuda.Stream(device=self.device))
        return self._lane_engines[k - 1], self._lane_streams[k - 1]

    # conditioning: list of size [b] of class indices.  Returns detached codes [b, 256]  (rar_wrapper.py:89-107)
    def sample(self, conditioning, gen_params=None, apply_watermark=False, greedy=False):
        cond = torch.as_tensor(conditioning, device=self.device).view(-1).long()
        steps = self.codes_size * self.codes_size
        wm = self.watermarker if apply_watermark else None
        V = self.rar_cfg["codebook_size"]
        out = []
        mb = self._rar.max_batch
        n_chunks = (cond.numel() + mb - 1) // mb
        n_lanes = min(self.lanes, n_chunks)      # chunk i -> lane i % n_lanes, concurrently (weights shared)
        if n_lanes - 1 > len(self._lane_engines):
            # create the missing lanes BEFORE any work of this call is enqueued: engine creation zero-fills its buffers on
            # the legacy default stream, which would queue behind lane 0's generation and land in the middle of lane 1's
            self._lane(n_lanes - 1)
            torch.cuda.synchronize(self.device)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)                                # everything the caller enqueued before this call (cond, weights)
        for ci, i in enumerate(range(0, cond.numel(), mb)):
            c = cond[i:i + mb]
            noise, stream = None, None
            if not greedy and self.rng in ("torch", "torch_buffer"):
                # RAR.preprocess_condition draws torch.rand_like(condition) first (rar.py:305)
                torch.rand(c.shape, device=self.device)
                if self.rng == "torch":
                    stream = self._torch_stream(steps, c.numel(), V)
                else:
                    noise = self._draw_noise(steps, c.numel(), V)
            self._step_seed += 1
            eng, lane_stream = self._lane(ci % n_lanes)
            kw = dict(guidance_scale=4.0, temperature=1.0, watermarker=wm, noise=noise, greedy=greedy, seed=self._step_seed,
                      torch_stream=stream)
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
        mb = self._vqgan.max_batch           # clamp(0,1) * 2 - 1 inside the kernel (rar_wrapper.py:114-116)
        images = self._vqgan.decode(codes) if codes.shape[0] <= mb else torch.cat(
            [self._vqgan.decode(codes[i:i + mb]) for i in range(0, codes.shape[0], mb)], dim=0)
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        return images

    def images_to_codes(self, images):
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        mb = self._vqgan.max_batch           # (x + 1) / 2 inside the kernel (rar_wrapper.py:124)
        codes = self._vqgan.encode(images) if images.shape[0] <= mb else torch.cat(
            [self._vqgan.encode(images[i:i + mb]) for i in range(0, images.shape[0], mb)], dim=0)
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes
