"""Host-side mirror of ``wmar.models.chameleon_wrapper.ChameleonARMMWrapper`` (chameleon_wrapper.py:139-186) for
text -> image generation with Anole-7B / Chameleon-7B.

``sample`` replaces ``ChameleonInferenceModel.generate`` with ``Options(txt=False)`` -- the worker thread, request /
response queues, TokenManager and ImageDecoder of deps/chameleon/inference/chameleon.py -- by ONE call of the
Chameleon decode engine (csrc/chameleon.cu); ``codes_to_images`` / ``images_to_codes`` replace
``ImageTokenizer`` + the 512-pixel VQGAN (deps/chameleon/inference/image_tokenizer.py, vqgan.py) by the VQGAN engine
(family 0, attention only in ``mid``).

Codes cross the boundary as BPE ids of the image tokens, like the reference's (the watermark works on the 65536-wide
vocabulary).  Offline there is neither ``text_tokenizer.json`` nor a checkpoint, so
  * the text tokenizer is a parameter (``tokenize(str) -> list[int]``); the default is a deterministic stand-in that
    maps whitespace-separated words to ids in the text range (prompt LENGTHS match, contents do not),
  * the BPE <-> codebook translation of VocabTranslation (vocab.py:76-123, a permutation read from the tokenizer file)
    is a parameter (``bpe2img``); the default is the affine map bpe = code + 4 (image tokens occupy ids 4..8195).
"""
import os
import zlib

import torch

from .. import _lib
from .armm_wrapper import AutoregressiveMultimodalModelWrapper
from .cham_engine import ChameleonEngine
from .state import StateModule
from .synthetic import ANOLE_7B_CFG, CHAMELEON_VQGAN_DDCONFIG, chameleon_state, taming_vqgan_state
from .vqgan_engine import VQGANEngine

# ids of the released Chameleon tokenizer (vocab.py:15-44 looks them up by name; stated here because the file is absent)
BOS_ID, PAD_ID, EOS_ID = 0, 1, 2
IMAGE_TOKEN_LO, IMAGE_TOKEN_HI = 4, 8196
END_IMAGE, BEGIN_IMAGE = 8196, 8197
EOT_ID = 8710            # "<reserved08706>", the <END-OF-TURN> sentinel
TEXT_LO = 16384
ASSETS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "assets")


def _stand_in_tokenizer(prompt):
    return [TEXT_LO + (zlib.crc32(w.encode("utf-8")) % (65536 - TEXT_LO)) for w in prompt.split()]


class ChameleonARMMWrapper(AutoregressiveMultimodalModelWrapper):
    def __init__(self, modelpath=None, *, state_dict=None, tokenizer_state_dict=None, model_cfg=None, vq_cfg=None,
                 tokenize=None, bpe2img=None, device="cuda", max_batch=8, vqgan_precision="bf16x3", seed=0, rng="torch", lanes=3,
                 guidance_text=3.0, guidance_image=1.2, image_tokens_per_image=1024, alive_ids_path=None):
        super().__init__()
        self._device = torch.device(device)
        if self._device.type != "cuda":
            raise _lib.WmarError("ChameleonARMMWrapper runs on CUDA only (no CPU fallback)")
        if modelpath is not None:
            raise _lib.WmarError("loading Anole checkpoints needs the files the reference downloads (consolidated.pth, "
                                 "text_tokenizer.json, vqgan.ckpt); pass state_dict / tokenizer_state_dict / tokenize / bpe2img")
        self.cfg = dict(ANOLE_7B_CFG)
        self.cfg.update(model_cfg or {})
        self.vq_cfg = dict(vq_cfg or CHAMELEON_VQGAN_DDCONFIG)
        self.tokenize = tokenize or _stand_in_tokenizer
        self.guidance_text, self.guidance_image = guidance_text, guidance_image
        self.seed, self.rng, self.max_batch, self.vqgan_precision = seed, rng, max_batch, vqgan_precision
        if state_dict is None:
            state_dict = chameleon_state(self.cfg, seed=seed, device=self._device)
        if tokenizer_state_dict is None:
            tokenizer_state_dict = taming_vqgan_state(self.vq_cfg, seed=seed + 1, device=self._device)
        self._state = state_dict
        self.tokenizer = StateModule(dict(tokenizer_state_dict)).to(self._device)
        self.tokenizer.eval()
        self.tokenizer.quantize.n_e = self.vq_cfg["n_embed"]
        n_img = IMAGE_TOKEN_HI - IMAGE_TOKEN_LO
        self.bpe2img = (torch.as_tensor(bpe2img, dtype=torch.long) if bpe2img is not None
                        else torch.arange(n_img) ).to(self._device)          # index: bpe - 4 -> codebook id
        self.img2bpe = torch.empty_like(self.bpe2img)
        self.img2bpe[self.bpe2img] = torch.arange(n_img, device=self._device)
        # alive / dead ids of the watermark exactly as the reference derives them (chameleon_wrapper.py:32 ->
        # armm_wrapper.py:42-55): alive = the 57344 BPE ids of assets/chameleon_all_ids.txt ([4, 8196) and
        # [16384, 65536)), dead = set(range(vq.n_e = 8192)) - alive = {0, 1, 2, 3}.  The stratified split permutes
        # len(alive) ids, so the id list IS part of the watermark key.
        self.init_alivecodes(alive_ids_path or os.path.join(ASSETS, "chameleon_all_ids.txt"))
        vq = self.tokenizer.quantize
        if self.cfg["vocab_size"] < 65536:     # reduced test vocabularies only: ids the model cannot emit are dropped
            vq.alive_ids = vq.alive_ids[vq.alive_ids < self.cfg["vocab_size"]]
        self.codes_size = int(round(image_tokens_per_image ** 0.5))
        self.image_size = self.vq_cfg["resolution"]
        self.dim_z = self.vq_cfg["z_channels"]
        self._eng = None
        self._vqgan = None
        self.lanes = max(1, int(lanes))
        self._lane_engines, self._lane_streams = [], []
        self._step_seed = seed
        self.sync_weights()

    def __repr__(self):
        return "ChameleonARMMWrapper"

    def get_image_tokenizer(self):
        return self.tokenizer

    def get_total_vocab_size(self):
        return self.cfg["vocab_size"]

    def sync_weights(self):
        c = self.cfg
        if self._eng is None:
            self._eng = ChameleonEngine(self._state, c["n_layers"], c["n_heads"], c["n_kv_heads"],
                                        image_tokens=(IMAGE_TOKEN_LO, IMAGE_TOKEN_HI),
                                        max_seq=self.codes_size * self.codes_size + 96, max_batch=min(self.max_batch, 8),
                                        norm_eps=c["norm_eps"], rope_theta=c["rope_theta"],
                                        qk_norm=c.get("qk_normalization", True), device=self._device)
        else:
            self._eng.sync_weights(self._state)
        self._lane_engines, self._lane_streams = [], []   # lanes borrow the engine's tensors: rebuilt on demand
        v = self.vq_cfg
        ecfg = dict(family=0, ch=v["ch"], ch_mult=tuple(v["ch_mult"]), num_res_blocks=v["num_res_blocks"],
                    attn_resolution=0, resolution=v["resolution"], z_channels=v["z_channels"], embed_dim=v["embed_dim"],
                    n_embed=v["n_embed"])
        tstate = dict(self.tokenizer.state_dict())
        if self._vqgan is None:
            self._vqgan = VQGANEngine(tstate, ecfg, device=self._device, max_batch=max(self.max_batch, 1),
                                      precision=self.vqgan_precision)
        else:
            self._vqgan.sync_weights(tstate)

    def prompt_rows(self, prompts):
        """The three CFG row groups of ImageDecoder._split_inputs_for_cfg (chameleon.py:351-372) for text prompts:
        full = <s> text <END-OF-TURN> <boi>; image-conditioned keeps only bos / image tokens / boi / eoi; unconditioned =
        <s> <boi>."""
        full = [[BOS_ID] + list(self.tokenize(p)) + [EOT_ID, BEGIN_IMAGE] for p in prompts]
        keep = lambda t: t == BOS_ID or t == BEGIN_IMAGE or t == END_IMAGE or IMAGE_TOKEN_LO <= t < IMAGE_TOKEN_HI
        img = [[t for t in row if keep(t)] for row in full]
        unc = [[BOS_ID, BEGIN_IMAGE] for _ in full]
        return full + img + unc

    def _lane(self, k):
        """Engine + stream of lane k (lane 0 = the wrapper's own engine on the caller's stream); see TamingARMMWrapper."""
        if k == 0:
            return self._eng, None
        while len(self._lane_engines) < k:
            self._lane_engines.append(self._eng.clone_lane())
            self._lane_streams.append(torch.cuda.Stream(device=self.device))
        return self._lane_engines[k - 1], self._lane_streams[k - 1]

    # conditioning: list of size [b] of (index, prompt) coco tuples.  Returns detached BPE codes [b, 1024]
    def sample(self, conditioning, gen_params=None, apply_watermark=False, greedy=False):
        gen_params = gen_params or {}
        prompts = [p for _, p in conditioning]
        steps = self.codes_size * self.codes_size
        wm = self.watermarker if apply_watermark else None
        V = self.cfg["vocab_size"]
        out = []
        mb = self._eng.max_batch
        n_chunks = (len(prompts) + mb - 1) // mb
        n_lanes = min(self.lanes, n_chunks)      # chunk i -> lane i % n_lanes, concurrently (weights shared)
        if n_lanes - 1 > len(self._lane_engines):
            self._lane(n_lanes - 1)              # before any work of this call is enqueued (see TamingARMMWrapper.sample)
            torch.cuda.synchronize(self.device)
        main = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(main)
        for ci, i in enumerate(range(0, len(prompts), mb)):
            rows = self.prompt_rows(prompts[i:i + mb])
            b = len(rows) // 3
            noise, stream = None, None
            if not greedy and self.rng == "torch":            # multinomial over probs[b, V] (token_selector.py:26-47)
                stream = self._torch_stream(steps, b, V)
            elif not greedy and self.rng == "torch_buffer":
                noise = self._draw_noise(steps, b, V)
            self._step_seed += 1
            eng, lane_stream = self._lane(ci % n_lanes)
            kw = dict(temperature=gen_params.get("temperature", 1.0), top_p=gen_params.get("top_p"), watermarker=wm,
                      noise=noise, greedy=greedy, seed=self._step_seed, torch_stream=stream)
            if lane_stream is None:
                out.append(eng.sample(rows, steps, self.guidance_text, self.guidance_image, defer_check=n_lanes > 1, **kw))
            else:
                if noise is not None:
                    lane_stream.wait_stream(main)
                else:
                    lane_stream.wait_event(ready)
                with torch.cuda.stream(lane_stream):
                    o = eng.sample(rows, steps, self.guidance_text, self.guidance_image, defer_check=True, **kw)
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
        img_ids = self.bpe2img[(codes - IMAGE_TOKEN_LO).clamp_(0, self.bpe2img.numel() - 1)]
        images = self._vqgan.decode(img_ids)
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        return images

    def images_to_codes(self, images):
        assert self.is_images_shaped(images), f"Images shape: {images.shape}"
        # the reference re-tokenises through an 8-bit PIL image (chameleon_wrapper.py:176-183 -> image_tokenizer.py:100-114
        # `_pil_from_chw_tensor`: clamp, (x + 1) / 2 * 255 TRUNCATED to uint8; :72-91 `_vqgan_input_from`: u8 / 255 * 2 - 1;
        # resize and centre crop are the identity at the model's own resolution): same quantisation here, on the device
        q = torch.floor((images.float().clamp(-1.0, 1.0) + 1.0) / 2.0 * 255.0)
        images = (q / 255.0 * 2.0 - 1.0).to(images.dtype)
        codes = self.img2bpe[self._vqgan.encode(images)] + IMAGE_TOKEN_LO
        assert self.is_codes_shaped(codes), f"Codes shape: {codes.shape}"
        return codes
