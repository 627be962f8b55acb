"""Host handle of the Taming minGPT decode engine (wmar_gpt_* in include/wmar_b200.h)."""
import ctypes

import torch

from .. import _lib


class TamingGPTEngine:
    """Packs a minGPT state dict (keys as deps/taming/modules/transformer/mingpt.py GPT.state_dict()) into the
    weight table the C side borrows, and runs the whole sample_with_past loop (mingpt.py:326-368) on the device."""

    def __init__(self, state, n_layer, n_head, device="cuda", max_batch=16):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WmarError("TamingGPTEngine is CUDA only (no CPU fallback)")
        self.n_layer, self.n_head = n_layer, n_head
        self.handle = None
        self.max_batch = max_batch
        self.sync_weights(state)

    def sync_weights(self, state):
        def dev(t):
            return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

        tok = dev(state["tok_emb.weight"])
        pos = dev(state["pos_emb"]).reshape(-1, tok.shape[1]).contiguous()
        self.vocab_size, self.n_embd = tok.shape
        self.block_size = pos.shape[0]
        tensors = [tok, pos]
        for i in range(self.n_layer):
            p = f"blocks.{i}."
            wqkv = torch.cat([dev(state[p + f"attn.{n}.weight"]) for n in ("query", "key", "value")], dim=0).contiguous()
            bqkv = torch.cat([dev(state[p + f"attn.{n}.bias"]) for n in ("query", "key", "value")], dim=0).contiguous()
            tensors += [dev(state[p + "ln1.weight"]), dev(state[p + "ln1.bias"]), wqkv, bqkv,
                        dev(state[p + "attn.proj.weight"]), dev(state[p + "attn.proj.bias"]),
                        dev(state[p + "ln2.weight"]), dev(state[p + "ln2.bias"]),
                        dev(state[p + "mlp.0.weight"]), dev(state[p + "mlp.0.bias"]),
                        dev(state[p + "mlp.2.weight"]), dev(state[p + "mlp.2.bias"])]
        tensors += [dev(state["ln_f.weight"]), dev(state["ln_f.bias"]), dev(state["head.weight"])]
        self._tensors = tensors  # keep the borrowed storage alive
        self._create()

    def clone_lane(self):
        """A second engine over the SAME weight tensors (device pointers shared, nothing copied) with its own KV cache,
        scratch buffers and step graph: an independent 'lane' whose generations can run concurrently with this one's
        on another CUDA stream.  The decode step is a latency chain that leaves most of the HBM bandwidth idle at
        16 rows; two lanes interleave their chains (measured 1.44x images/s on Taming C2, profiles/r02_summary.md)."""
        lane = object.__new__(type(self))
        lane.device, lane.n_layer, lane.n_head, lane.max_batch = self.device, self.n_layer, self.n_head, self.max_batch
        lane.vocab_size, lane.n_embd, lane.block_size = self.vocab_size, self.n_embd, self.block_size
        lane.handle = None
        lane._tensors = self._tensors
        lane._create()
        return lane

    def _create(self):
        L = _lib.lib()
        if self.handle is not None:
            L.wmar_gpt_destroy(self.handle)
            self.handle = None
        cfg = _lib.GptConfig(self.vocab_size, self.block_size, self.n_layer, self.n_head, self.n_embd, self.max_batch)
        table = _lib.pointer_table(self._tensors)
        h = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.wmar_gpt_create(ctypes.byref(cfg), table, len(self._tensors), ctypes.byref(h)))
        self.handle = h

    def __del__(self):
        try:
            if self.handle is not None:
                _lib.lib().wmar_gpt_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @torch.no_grad()
    def sample(self, cond, steps, temperature=1.0, top_k=None, top_p=None, watermarker=None, noise=None,
               greedy=False, seed=0, return_logits=False, torch_stream=None, defer_check=False):
        """cond int64[B] -> codes int64[B, steps].  noise fp32[steps,B,V] ~ Exp(1) reproduces torch.multinomial's
        draws (see tests); None draws from an in-kernel Philox stream keyed by `seed`."""
        cond = torch.as_tensor(cond, dtype=torch.long, device=self.device).reshape(-1).contiguous()
        B = cond.numel()
        out = torch.empty((B, steps), dtype=torch.long, device=self.device)
        logits = torch.empty((steps, B, self.vocab_size), dtype=torch.float32, device=self.device) if return_logits else None
        sp = _lib.SampleParams(float(temperature), int(top_k) if top_k else 0, float(top_p) if top_p else 0.0,
                               1 if greedy else 0, int(seed))
        if torch_stream is not None and noise is None and not greedy:   # torch's own CUDA Philox stream, drawn in the kernel
            sp.seed, sp.rng_mode = int(torch_stream["seed"]), 1
            sp.torch_offset, sp.torch_threads = int(torch_stream["torch_offset"]), int(torch_stream["torch_threads"])
            sp.torch_numel, sp.torch_rowlen = int(torch_stream["torch_numel"]), int(torch_stream["torch_rowlen"])
        wm = watermarker.c_params() if watermarker is not None else None
        if noise is not None:
            assert noise.shape == (steps, B, self.vocab_size) and noise.dtype == torch.float32 and noise.is_cuda
            noise = noise.contiguous()
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().wmar_gpt_sample(self.handle, ctypes.byref(wm) if wm is not None else None,
                                                  ctypes.byref(sp), _lib.ptr(cond), B, steps, _lib.ptr(noise),
                                                  _lib.ptr(out), _lib.ptr(logits), _lib.current_stream()))
            if not defer_check:        # (lanes: the caller checks once after joining the streams)
                _lib.check_device_flag()   # raises on an out-of-range context sum / top-p overflow (device-side checks)
        self._keepalive = (cond, noise)
        return (out, logits) if return_logits else out

    def algorithmic_bytes(self, B, steps):
        return float(_lib.lib().wmar_gpt_algorithmic_bytes(self.handle, B, steps))

    def launches_per_step(self):
        return int(_lib.lib().wmar_gpt_launches_per_step(self.handle))
