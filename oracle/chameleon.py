"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (torch, bf16 tensors like the reference model) of the Chameleon / Anole image-token generation path:

  * Transformer.forward_with_attn_bias                 deps/chameleon/inference/transformer.py:97-337
  * ChameleonModelAdapter (per-row key ranges)          deps/chameleon/inference/model_adapter.py:51-118
  * ChameleonGenerator.__next__                         deps/chameleon/inference/generation.py:68-103
  * ImageDecoder (CFG row groups, processor order)      deps/chameleon/inference/chameleon.py:299-389
  * InBatchInstructCFGLogitsProcessor, AllowOnlyTokens  deps/chameleon/inference/logits_processor.py:312-335,135-151
  * ReplicatedInputTokenSelector                        deps/chameleon/inference/token_selector.py:26-47

Pinning of the transformer: tests/golden/chameleon_transformer.npz holds logits of the reference's OWN `Transformer` module
(imported unmodified, built like loader.py:16-33) on ragged prefills + teacher-forced single-token passes, multi-head and
grouped-query (oracle/gen_golden_chameleon_transformer.py); this restatement reproduces them to within single bf16
rounding flips (tests/test_oracle_chameleon.py).  What that run could NOT execute are the three xformers operators the
module imports (RMSNorm, rope_padded, fmha.memory_efficient_attention_forward with
BlockDiagonalCausalWithOffsetPaddedKeysMask; xformers is unpinned in the reference's README.md:37, not installable here
and has no CPU kernels): they came from oracle/xformers_stub, written from the documented semantics --
RMSNorm = x * rsqrt(mean(x^2) + eps) * weight in fp32, stored in the input dtype; rope_padded rotates ADJACENT pairs
(x[2j], x[2j+1]) by position * theta^(-2j/hd) in fp32, stores bf16 and appends k, v to the padded cache; attention is
softmax(q k^T / sqrt(hd)) v over the row's own keys 0..p with fp32 softmax.  So: module graph, fused-weight layouts, norm
placement, GQA expansion, residual order and every bf16 rounding point of the Linear layers are PINNED to the reference
run; the inside of those three operators is pinned to their documentation only ("parity unpinned" for them).
The logits processors / token selector ARE pinned: tests/golden/chameleon_sampling.npz is produced by the imported
reference classes (oracle/gen_golden_chameleon.py).
"""
import math

import torch
import torch.nn.functional as F

BF = torch.bfloat16


def synthetic_chameleon_weights(V, d, L, H, Hkv, Fh, seed=0, qk_norm=True):
    """State dict with the reference's parameter names (transformer.py), bf16, N(0, 0.02)-style random init."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: (torch.randn(*s, generator=g) * std).to(BF)
    hd = d // H
    w = {"tok_embeddings.weight": rn(V, d, std=1.0), "norm.weight": (1 + torch.randn(d, generator=g) * 0.1).to(BF),
         "output.weight": rn(V, d, std=0.05)}
    for i in range(L):
        p = f"layers.{i}."
        w[p + "attention_norm.weight"] = (1 + torch.randn(d, generator=g) * 0.1).to(BF)
        w[p + "ffn_norm.weight"] = (1 + torch.randn(d, generator=g) * 0.1).to(BF)
        w[p + "attention.wqkv.weight"] = rn((H + 2 * Hkv) * hd, d, std=0.05)
        w[p + "attention.wo.weight"] = rn(d, H * hd, std=0.05)
        if qk_norm:
            for nm in ("q_normalization", "k_normalization"):
                w[p + f"attention.{nm}.weight"] = (1 + torch.randn(hd, generator=g) * 0.1).to(BF)
                w[p + f"attention.{nm}.bias"] = (torch.randn(hd, generator=g) * 0.1).to(BF)
        w[p + "feed_forward.w13.weight"] = rn(2 * Fh, d, std=0.05)
        w[p + "feed_forward.w2.weight"] = rn(d, Fh, std=0.05)
    return w


def rms_norm(x, weight, eps):
    xf = x.float()
    return (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps) * weight.float()).to(x.dtype)


def rope_pairs(x, pos, theta):
    """x [..., hd] (bf16), adjacent pairs rotated by pos * theta^(-2j/hd); fp32 math, bf16 result."""
    hd = x.shape[-1]
    j = torch.arange(hd // 2, dtype=torch.float32)
    freq = torch.pow(torch.tensor(theta, dtype=torch.float32), -2.0 * j / hd)
    ang = float(pos) * freq
    cs, sn = torch.cos(ang), torch.sin(ang)
    xf = x.float().reshape(*x.shape[:-1], hd // 2, 2)
    x0, x1 = xf[..., 0], xf[..., 1]
    out = torch.stack([x0 * cs - x1 * sn, x0 * sn + x1 * cs], dim=-1)
    return out.reshape(x.shape).to(x.dtype)


class ChameleonOracle:
    """One row at a time, one position at a time (a causal prefill with a cache == feeding the prompt token by token)."""

    def __init__(self, w, n_layer, n_head, n_kv_head, norm_eps=1e-5, rope_theta=10000.0, qk_norm=True):
        self.w, self.L, self.H, self.Hkv = w, n_layer, n_head, n_kv_head
        self.eps, self.theta, self.qk_norm = norm_eps, rope_theta, qk_norm
        self.d = w["tok_embeddings.weight"].shape[1]
        self.hd = self.d // n_head
        self.cache = {}

    def reset(self):
        self.cache = {}

    @torch.no_grad()
    def step_row(self, row, token, pos):
        """Feeds `token` at position `pos` of row `row`; returns the fp32 logits [V] (transformer.py:297-319)."""
        w, H, Hkv, hd = self.w, self.H, self.Hkv, self.hd
        h = w["tok_embeddings.weight"][token].clone()                                   # bf16 [d]
        for i in range(self.L):
            p = f"layers.{i}."
            xn = rms_norm(h, w[p + "attention_norm.weight"], self.eps)
            qkv = F.linear(xn, w[p + "attention.wqkv.weight"])                          # transformer.py:111
            q = qkv[: H * hd].view(H, hd)
            k = qkv[H * hd: (H + Hkv) * hd].view(Hkv, hd)
            v = qkv[(H + Hkv) * hd:].view(Hkv, hd)
            if self.qk_norm:                                                            # :116-123
                q = F.layer_norm(q, (hd,), w[p + "attention.q_normalization.weight"], w[p + "attention.q_normalization.bias"])
                k = F.layer_norm(k, (hd,), w[p + "attention.k_normalization.weight"], w[p + "attention.k_normalization.bias"])
            q = rope_pairs(q, pos, self.theta)                                          # rope_padded, :130-138
            k = rope_pairs(k, pos, self.theta)
            ck, cv = self.cache.setdefault((i, row), ([], []))
            assert len(ck) == pos, "positions must be fed in order"
            ck.append(k)
            cv.append(v)
            Kc = torch.stack(ck, 0).float()                                             # [t, Hkv, hd]
            Vc = torch.stack(cv, 0).float()
            grp = H // Hkv
            Kc = Kc.repeat_interleave(grp, dim=1)                                       # GQA expand, :141-147
            Vc = Vc.repeat_interleave(grp, dim=1)
            s = torch.einsum("hd,thd->ht", q.float(), Kc) / math.sqrt(hd)
            pr = torch.softmax(s, dim=-1)
            y = torch.einsum("ht,thd->hd", pr, Vc).to(BF).reshape(H * hd)               # :149-156
            h = h + F.linear(y, w[p + "attention.wo.weight"])                            # :158,238-244
            hn = rms_norm(h, w[p + "ffn_norm.weight"], self.eps)
            x13 = F.linear(hn, w[p + "feed_forward.w13.weight"])                         # :215-219
            x1, x3 = x13.chunk(2, -1)
            h = h + F.linear(F.silu(x1) * x3, w[p + "feed_forward.w2.weight"])
        return F.linear(rms_norm(h, w["norm.weight"], self.eps), w["output.weight"]).float()   # :314-319


def instruct_cfg(logits3, s_txt, s_img):
    """InBatchInstructCFGLogitsProcessor (logits_processor.py:312-335) for logits [3B, V]; returns the mixed [B, V]."""
    f, im, u = logits3.chunk(3)
    return u + s_img * (im - u) + s_txt * (f - im)


def allow_only(logits, lo, hi):
    """AllowOnlyTokensLogitsProcessor (logits_processor.py:135-151) for a contiguous id range."""
    out = torch.full_like(logits, -math.inf)
    out[:, lo:hi] = logits[:, lo:hi]
    return out


@torch.no_grad()
def generate(oracle, prompts3, B, steps, s_txt, s_img, lo, hi, temperature, top_p, wm_process=None, noise=None,
             greedy=False, forced_ids=None):
    """ImageDecoder + ChameleonGenerator (chameleon.py:299-389, generation.py:68-103).  prompts3: 3B token lists (full,
    image-conditioned, unconditioned), each ending in <boi>.  wm_process(past_ids [B,t], logits [B,V]) -> logits.
    noise [steps,B,V] ~ Exp(1) reproduces torch.multinomial (argmax p / q).  forced_ids [B,steps]: teacher forcing.
    Returns (ids [B,steps], mixed logits [steps,B,V])."""
    from transformers import TopPLogitsWarper
    oracle.reset()
    R = 3 * B
    logits = [None] * R
    for r in range(R):                                      # prefill: every row over its own prompt
        for pos, tok in enumerate(prompts3[r]):
            logits[r] = oracle.step_row(r, tok, pos)
    seqs = [list(p) for p in prompts3]
    ids, mixed_all = [], []
    warp = TopPLogitsWarper(top_p) if top_p is not None and 0 < top_p < 1 else None
    for s in range(steps):
        l3 = torch.stack(logits, 0)
        mixed = instruct_cfg(l3, s_txt, s_img)
        mixed_all.append(mixed.clone())
        l = mixed.clone()
        past = torch.tensor([seqs[b] for b in range(B)], dtype=torch.long) if len({len(seqs[b]) for b in range(B)}) == 1 else None
        if wm_process is not None:
            l = wm_process(past, l)
        l = allow_only(l, lo, hi)
        l = l / temperature
        if warp is not None:
            l = warp(None, l)
        probs = l.softmax(dim=1)
        if greedy:
            nxt = probs.argmax(dim=1)
        elif noise is not None:
            nxt = (probs / noise[s]).argmax(dim=1)
        else:
            nxt = probs.multinomial(1).squeeze(1)
        if forced_ids is not None:
            nxt = forced_ids[:, s]
        ids.append(nxt.clone())
        if s + 1 < steps:
            for r in range(R):
                tok = int(nxt[r % B])
                logits[r] = oracle.step_row(r, tok, len(seqs[r]))
                seqs[r].append(tok)
    return torch.stack(ids, 1), torch.stack(mixed_all, 0)
