"""Host handle of the RAR decode engine (wmar_rar_* in include/wmar_b200.h)."""
import ctypes

import torch

from .. import _lib

RAR_SIZES = {  # rar_wrapper.py:44-53
    "rar_b": dict(hidden_size=768, num_hidden_layers=24, num_attention_heads=16, intermediate_size=3072),
    "rar_l": dict(hidden_size=1024, num_hidden_layers=24, num_attention_heads=16, intermediate_size=4096),
    "rar_xl": dict(hidden_size=1280, num_hidden_layers=32, num_attention_heads=16, intermediate_size=5120),
    "rar_xxl": dict(hidden_size=1408, num_hidden_layers=40, num_attention_heads=16, intermediate_size=6144),
}


class RAREngine:
    """Packs a RAR state dict (keys as deps/rar/modeling/rar.py RAR.state_dict()) into the borrowed weight table and
    runs RAR.generate (rar.py:408-459) on the device: CFG, watermark, sampler fused, no host work per token."""

    def __init__(self, state, n_layer, n_head, codebook_size=1024, n_classes=1000, image_seq_len=256, device="cuda",
                 max_batch=8):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WmarError("RAREngine is CUDA only (no CPU fallback)")
        self.n_layer, self.n_head = n_layer, n_head
        self.codebook_size, self.n_classes, self.image_seq_len = codebook_size, n_classes, image_seq_len
        self.max_batch = max_batch
        self.handle = None
        self.sync_weights(state)

    def sync_weights(self, state):
        def dev(k):
            return state[k].detach().to(device=self.device, dtype=torch.float32).contiguous()

        d = state["cls_token"].shape[-1]
        self.hidden = d
        self.mlp = state["blocks.0.mlp.fc1.weight"].shape[0]
        t = [dev("cls_token").reshape(d), dev("embeddings.weight"), dev("pos_embed").reshape(-1, d),
             dev("target_aware_pos_embed").reshape(-1, d), dev("timesteps_embeddings").reshape(-1, d)]
        assert t[2].shape[0] >= self.image_seq_len + 2 and t[3].shape[0] >= self.image_seq_len + 2
        assert t[4].shape[0] >= self.image_seq_len + 1
        assert t[1].shape[0] >= self.codebook_size + 1 + self.n_classes + 1
        for i in range(self.n_layer):
            p = f"blocks.{i}."
            for k in ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.q_norm.weight",
                      "attn.q_norm.bias", "attn.k_norm.weight", "attn.k_norm.bias", "attn.proj.weight", "attn.proj.bias",
                      "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias",
                      "adaLN_modulation.1.weight", "adaLN_modulation.1.bias"):
                t.append(dev(p + k))
        t += [dev("adaln_before_head.adaLN_modulation.1.weight"), dev("adaln_before_head.adaLN_modulation.1.bias"),
              dev("lm_head.weight"), dev("lm_head.bias")]
        self._tensors = t
        self._create()

    def clone_lane(self):
        """A second engine over the SAME weight tensors with its own KV cache / scratch / graphs: an independent lane
        whose generations run concurrently with this one's on another CUDA stream (gpt_engine.TamingGPTEngine.clone_lane)."""
        lane = object.__new__(type(self))
        for k in ("device", "n_layer", "n_head", "codebook_size", "n_classes", "image_seq_len", "max_batch", "hidden", "mlp"):
            setattr(lane, k, getattr(self, k))
        lane.handle = None
        lane._tensors = self._tensors
        lane._create()
        return lane

    def _create(self):
        t = self._tensors
        L = _lib.lib()
        if self.handle is not None:
            L.wmar_rar_destroy(self.handle)
            self.handle = None
        cfg = _lib.RarConfig(self.codebook_size, self.n_classes, self.image_seq_len, self.n_layer, self.n_head,
                             self.hidden, self.mlp, self.max_batch)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.wmar_rar_create(ctypes.byref(cfg), _lib.pointer_table(t), len(t), ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if self.handle is not None:
                _lib.lib().wmar_rar_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @torch.no_grad()
    def sample(self, cond, steps=None, guidance_scale=4.0, temperature=1.0, watermarker=None, noise=None, greedy=False,
               seed=0, return_logits=False, torch_stream=None, defer_check=False):
        """cond int64[B] class ids -> ids int64[B, steps]; noise fp32[steps,B,V] ~ Exp(1) or None (in-kernel Philox)."""
        steps = steps or self.image_seq_len
        cond = torch.as_tensor(cond, dtype=torch.long, device=self.device).reshape(-1).contiguous()
        B = cond.numel()
        V = self.codebook_size
        out = torch.empty((B, steps), dtype=torch.long, device=self.device)
        logits = torch.empty((steps, B, V), dtype=torch.float32, device=self.device) if return_logits else None
        sp = _lib.SampleParams(float(temperature), 0, 0.0, 1 if greedy else 0, int(seed))
        if torch_stream is not None and noise is None and not greedy:   # torch's own CUDA Philox stream, drawn in the kernel
            sp.seed, sp.rng_mode = int(torch_stream["seed"]), 1
            sp.torch_offset, sp.torch_threads = int(torch_stream["torch_offset"]), int(torch_stream["torch_threads"])
            sp.torch_numel, sp.torch_rowlen = int(torch_stream["torch_numel"]), int(torch_stream["torch_rowlen"])
        wm = watermarker.c_params() if watermarker is not None else None
        if noise is not None:
            assert noise.shape == (steps, B, V) and noise.dtype == torch.float32 and noise.is_cuda
            noise = noise.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().wmar_rar_sample(self.handle, ctypes.byref(wm) if wm is not None else None,
                                                  ctypes.byref(sp), _lib.ptr(cond), B, steps, float(guidance_scale),
                                                  _lib.ptr(noise), _lib.ptr(out), _lib.ptr(logits),
                                                  _lib.current_stream()))
            if not defer_check:        # (lanes: the caller checks once after joining the streams)
                _lib.check_device_flag()   # raises on an out-of-range context sum / top-p overflow (device-side checks)
        self._keepalive = (cond, noise)
        return (out, logits) if return_logits else out

    def algorithmic_bytes(self, B, steps):
        return float(_lib.lib().wmar_rar_algorithmic_bytes(self.handle, B, steps))
