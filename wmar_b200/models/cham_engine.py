"""Host handle of the Chameleon / Anole image-token decode engine (wmar_cham_* in include/wmar_b200.h)."""
import ctypes

import torch

from .. import _lib


class ChameleonEngine:
    """Packs a Chameleon Transformer state dict (keys as deps/chameleon/inference/transformer.py Transformer.state_dict())
    into the weight table the C side borrows and runs ImageDecoder's whole generation loop (chameleon.py:299-389) on the
    device: per-row prompt prefill, 3-way classifier-free guidance, watermark, allowed-token mask, temperature, top-p,
    multinomial, token replication -- no Python per token."""

    def __init__(self, state, n_layer, n_head, n_kv_head=None, image_tokens=(4, 8196), max_seq=1100, max_batch=8,
                 norm_eps=1e-5, rope_theta=10000.0, qk_norm=True, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WmarError("ChameleonEngine is CUDA only (no CPU fallback)")
        self.n_layer, self.n_head, self.n_kv_head = n_layer, n_head, n_kv_head or n_head
        self.image_tokens, self.max_seq, self.max_batch = tuple(image_tokens), max_seq, max_batch
        self.norm_eps, self.rope_theta, self.qk_norm = norm_eps, rope_theta, qk_norm
        self.handle = None
        self.sync_weights(state)

    def sync_weights(self, state):
        bf = lambda t: t.detach().to(device=self.device, dtype=torch.bfloat16).contiguous()
        f32 = lambda t: t.detach().to(device=self.device, dtype=torch.bfloat16).float().contiguous()  # bf16 VALUES in fp32
        tok = bf(state["tok_embeddings.weight"])
        self.vocab_size, self.dim = tok.shape
        hd = self.dim // self.n_head
        tensors = [tok]
        for i in range(self.n_layer):
            p = f"layers.{i}."
            if self.qk_norm:
                qn = [f32(state[p + f"attention.{n}.{k}"]) for n in ("q_normalization", "k_normalization") for k in ("weight", "bias")]
            else:
                one, zero = torch.ones(hd, device=self.device), torch.zeros(hd, device=self.device)
                qn = [one, zero, one.clone(), zero.clone()]
            w13 = bf(state[p + "feed_forward.w13.weight"])
            self.ffn_hidden = F = w13.shape[0] // 2
            # interleave per 64-row tile: 32 rows of w1 (x1), then the matching 32 rows of w3 (x3) -- see wmar_b200.h
            w13 = torch.stack((w13[:F].view(F // 32, 32, -1), w13[F:].view(F // 32, 32, -1)), dim=1).reshape(2 * F, -1).contiguous()
            tensors += [f32(state[p + "attention_norm.weight"]), bf(state[p + "attention.wqkv.weight"]), *qn,
                        bf(state[p + "attention.wo.weight"]), f32(state[p + "ffn_norm.weight"]), w13,
                        bf(state[p + "feed_forward.w2.weight"])]
        tensors += [f32(state["norm.weight"]), bf(state["output.weight"])]
        self._tensors = tensors
        self._create()

    def clone_lane(self):
        """A second engine over the SAME weight tensors with its own KV cache / scratch / graphs: an independent lane whose
        generations run concurrently with this one's on another CUDA stream (gpt_engine.TamingGPTEngine.clone_lane)."""
        lane = object.__new__(type(self))
        for k in ("device", "n_layer", "n_head", "n_kv_head", "image_tokens", "max_seq", "max_batch", "norm_eps", "rope_theta",
                  "qk_norm", "vocab_size", "dim", "ffn_hidden"):
            setattr(lane, k, getattr(self, k))
        lane.handle = None
        lane._tensors = self._tensors
        lane._create()
        return lane

    def _create(self):
        L = _lib.lib()
        if self.handle is not None:
            L.wmar_cham_destroy(self.handle)
            self.handle = None
        cfg = _lib.ChamConfig(self.vocab_size, self.dim, self.n_layer, self.n_head, self.n_kv_head, self.ffn_hidden,
                              self.max_seq, self.max_batch, self.image_tokens[0], self.image_tokens[1],
                              float(self.norm_eps), float(self.rope_theta), 1 if self.qk_norm else 0)
        table = _lib.pointer_table(self._tensors)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.wmar_cham_create(ctypes.byref(cfg), table, len(self._tensors), ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if self.handle is not None:
                _lib.lib().wmar_cham_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @torch.no_grad()
    def sample(self, prompts3, steps=1024, guidance_text=3.0, guidance_image=1.2, temperature=1.0, top_p=None,
               watermarker=None, noise=None, greedy=False, seed=0, return_logits=False, torch_stream=None, defer_check=False):
        """prompts3: 3B token-id lists (B full-conditioned, B image-conditioned, B unconditioned rows, each ending in
        <boi>; chameleon.py:351-372).  Returns ids int64[B, steps] (+ the mixed logits of the image-token window)."""
        R = len(prompts3)
        assert R % 3 == 0 and R > 0
        B = R // 3
        n_groups = 3
        if all(list(prompts3[B + b]) == list(prompts3[2 * B + b]) for b in range(B)):
            # text-only prompts: the image-conditioned rows are the unconditioned rows -> compute them once
            prompts3 = list(prompts3[:B]) + list(prompts3[2 * B:])
            n_groups, R = 2, 2 * B
        assert n_groups * B <= 16 and B <= self.max_batch, "too many images for one call"
        p_max = max(len(p) for p in prompts3)
        assert min(len(p) for p in prompts3) >= 1
        pr = torch.zeros((R, p_max), dtype=torch.long)
        for r, p in enumerate(prompts3):
            pr[r, :len(p)] = torch.as_tensor(p, dtype=torch.long)
        pr = pr.to(self.device)
        plen = torch.tensor([len(p) for p in prompts3], dtype=torch.int32, device=self.device)
        W = self.image_tokens[1] - self.image_tokens[0]
        out = torch.empty((B, steps), dtype=torch.long, device=self.device)
        logits = torch.empty((steps, B, W), dtype=torch.float32, device=self.device) if return_logits else None
        sp = _lib.SampleParams(float(temperature), 0, float(top_p) if top_p else 0.0, 1 if greedy else 0, int(seed))
        if torch_stream is not None and noise is None and not greedy:   # torch's own CUDA Philox stream, drawn in the kernel
            sp.seed, sp.rng_mode = int(torch_stream["seed"]), 1
            sp.torch_offset, sp.torch_threads = int(torch_stream["torch_offset"]), int(torch_stream["torch_threads"])
            sp.torch_numel, sp.torch_rowlen = int(torch_stream["torch_numel"]), int(torch_stream["torch_rowlen"])
        wm = watermarker.c_params() if watermarker is not None else None
        if noise is not None:
            assert noise.shape == (steps, B, self.vocab_size) and noise.dtype == torch.float32 and noise.is_cuda
            noise = noise.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().wmar_cham_sample(self.handle, ctypes.byref(wm) if wm is not None else None, ctypes.byref(sp),
                                                   _lib.ptr(pr), _lib.ptr(plen), p_max, p_max, B, n_groups, float(guidance_text),
                                                   float(guidance_image), steps, _lib.ptr(noise), _lib.ptr(out),
                                                   _lib.ptr(logits), _lib.current_stream()))
            if not defer_check:        # (lanes: the caller checks once after joining the streams)
                _lib.check_device_flag()   # raises on an out-of-range context sum / top-p overflow (device-side checks)
        self._keepalive = (pr, plen, noise)
        self.last_n_groups = n_groups
        return (out, logits) if return_logits else out

    def algorithmic_bytes(self, B, p_max, steps, n_groups=None):
        n_groups = n_groups or getattr(self, "last_n_groups", 3)
        return float(_lib.lib().wmar_cham_algorithmic_bytes(self.handle, B, n_groups, p_max, steps))
