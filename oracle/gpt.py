"""Taming minGPT decode step and sampling loop, restated in plain torch fp32 (TEST INFRASTRUCTURE ONLY).

Follows deps/taming/modules/transformer/mingpt.py: Block :98-122, CausalSelfAttention :42-95,
GPT.forward_with_past :183-214, sample_with_past :326-368.  Weights are a dict keyed like the reference GPT's
state_dict (tok_emb.weight, pos_emb, blocks.{i}.ln1.weight, blocks.{i}.attn.{key,query,value,proj}.{weight,bias},
blocks.{i}.mlp.{0,2}.{weight,bias}, ln_f.{weight,bias}, head.weight).
"""
import math

import torch
import torch.nn.functional as F

from . import sampling


def synthetic_gpt_weights(vocab_size, block_size, n_layer, n_head, n_embd, seed=0, dtype=torch.float32):
    """Seeded N(0, 0.02) Linear/Embedding weights like GPT._init_weights (:155-163); non-trivial biases, LN
    parameters and pos_emb so that every term of the step is exercised."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02):
        return (torch.randn(*shape, generator=g, dtype=torch.float32) * std).to(dtype)

    w = {"tok_emb.weight": rn(vocab_size, n_embd), "pos_emb": rn(1, block_size, n_embd)}
    for i in range(n_layer):
        p = f"blocks.{i}."
        for ln in ("ln1", "ln2"):
            w[p + ln + ".weight"] = 1.0 + rn(n_embd, std=0.1)
            w[p + ln + ".bias"] = rn(n_embd, std=0.05)
        for nm in ("key", "query", "value", "proj"):
            w[p + f"attn.{nm}.weight"] = rn(n_embd, n_embd)
            w[p + f"attn.{nm}.bias"] = rn(n_embd, std=0.01)
        w[p + "mlp.0.weight"] = rn(4 * n_embd, n_embd)
        w[p + "mlp.0.bias"] = rn(4 * n_embd, std=0.01)
        w[p + "mlp.2.weight"] = rn(n_embd, 4 * n_embd)
        w[p + "mlp.2.bias"] = rn(n_embd, std=0.01)
    w["ln_f.weight"] = 1.0 + rn(n_embd, std=0.1)
    w["ln_f.bias"] = rn(n_embd, std=0.05)
    w["head.weight"] = rn(vocab_size, n_embd)
    return w


class GPTOracle:
    def __init__(self, weights, n_layer, n_head):
        self.w = weights
        self.n_layer = n_layer
        self.n_head = n_head
        self.reset()

    def reset(self):
        self.k = [None] * self.n_layer
        self.v = [None] * self.n_layer

    def step(self, idx, pos):
        """idx int64[B] (one token per row), pos = past_length -> logits fp32[B, V]  (:183-214)."""
        w = self.w
        x = w["tok_emb.weight"][idx] + w["pos_emb"][0, pos]
        B, C = x.shape
        H = self.n_head
        hd = C // H
        for i in range(self.n_layer):
            p = f"blocks.{i}."
            a = F.layer_norm(x, (C,), w[p + "ln1.weight"], w[p + "ln1.bias"], 1e-5)
            q = F.linear(a, w[p + "attn.query.weight"], w[p + "attn.query.bias"]).view(B, H, 1, hd)
            k = F.linear(a, w[p + "attn.key.weight"], w[p + "attn.key.bias"]).view(B, H, 1, hd)
            v = F.linear(a, w[p + "attn.value.weight"], w[p + "attn.value.bias"]).view(B, H, 1, hd)
            self.k[i] = k if self.k[i] is None else torch.cat((self.k[i], k), dim=-2)
            self.v[i] = v if self.v[i] is None else torch.cat((self.v[i], v), dim=-2)
            att = (q @ self.k[i].transpose(-2, -1)) * (1.0 / math.sqrt(hd))
            att = F.softmax(att, dim=-1)
            y = (att @ self.v[i]).transpose(1, 2).contiguous().view(B, C)
            x = x + F.linear(y, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
            m = F.layer_norm(x, (C,), w[p + "ln2.weight"], w[p + "ln2.bias"], 1e-5)
            m = F.gelu(F.linear(m, w[p + "mlp.0.weight"], w[p + "mlp.0.bias"]))
            x = x + F.linear(m, w[p + "mlp.2.weight"], w[p + "mlp.2.bias"])
        x = F.layer_norm(x, (C,), w["ln_f.weight"], w["ln_f.bias"], 1e-5)
        return F.linear(x, w["head.weight"])


@torch.no_grad()
def sample_with_past(oracle, cond, steps, temperature, top_k, top_p, green_row_fn=None, delta=0.0, noise=None,
                     greedy=False, return_logits=False):
    """cond int64[B] (class ids); green_row_fn(past_ids int64[B,t]) -> list of u32 bitmask rows (or None) per row.

    noise fp32[steps, B, V] ~ Exp(1) (pre-drawn q of torch.multinomial).  Returns codes int64[B, steps]."""
    oracle.reset()
    x = cond.clone()
    sample = cond.view(-1, 1).clone()
    all_logits = []
    for n in range(steps):
        logits = oracle.step(x, n)
        if return_logits:
            all_logits.append(logits.clone())
        rows = green_row_fn(sample) if green_row_fn is not None else None
        x = sampling.sample_step(logits, rows, delta, temperature, top_k, top_p,
                                 None if noise is None else noise[n], greedy=greedy)
        sample = torch.cat((sample, x.view(-1, 1)), dim=1)
    codes = sample[:, 1:]
    if return_logits:
        return codes, torch.stack(all_logits)
    return codes
