"""RAR decode step / generate loop restated in plain torch fp32 (TEST INFRASTRUCTURE ONLY).

Follows deps/rar/modeling/rar.py: Attention :56-118, FinalLayer :123-134, Block :138-183, forward_fn :319-405,
generate :408-459; timm.layers.Mlp (third-party, unpinned) = fc1 -> GELU(erf) -> fc2.
Weights: dict keyed like the reference RAR state_dict.
"""
import torch
import torch.nn.functional as F

from . import sampling


def synthetic_rar_weights(d, depth, heads, mlp, codebook=1024, n_cls=1000, seq=256, seed=0):
    """Seeded trunc-normal-like weights; adaLN layers get non-zero weights (the reference zero-inits them, which
    would make the synthetic model degenerate)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*s, std=0.02):
        return torch.randn(*s, generator=g) * std

    hd = d // heads
    w = {"cls_token": rn(1, 1, d), "embeddings.weight": rn(codebook + 1 + n_cls + 1, d),
         "pos_embed": rn(1, seq + 1024, d), "target_aware_pos_embed": rn(1, seq + 1024, d),
         "timesteps_embeddings": rn(1, seq + 100, d)}
    for i in range(depth):
        p = f"blocks.{i}."
        for nm in ("norm1", "norm2"):
            w[p + nm + ".weight"] = 1.0 + rn(d, std=0.1)
            w[p + nm + ".bias"] = rn(d, std=0.05)
        w[p + "attn.qkv.weight"] = rn(3 * d, d)
        w[p + "attn.qkv.bias"] = rn(3 * d, std=0.01)
        for nm in ("q_norm", "k_norm"):
            w[p + f"attn.{nm}.weight"] = 1.0 + rn(hd, std=0.1)
            w[p + f"attn.{nm}.bias"] = rn(hd, std=0.05)
        w[p + "attn.proj.weight"] = rn(d, d)
        w[p + "attn.proj.bias"] = rn(d, std=0.01)
        w[p + "mlp.fc1.weight"] = rn(mlp, d)
        w[p + "mlp.fc1.bias"] = rn(mlp, std=0.01)
        w[p + "mlp.fc2.weight"] = rn(d, mlp)
        w[p + "mlp.fc2.bias"] = rn(d, std=0.01)
        w[p + "adaLN_modulation.1.weight"] = rn(6 * d, d)
        w[p + "adaLN_modulation.1.bias"] = rn(6 * d, std=0.05)
    w["adaln_before_head.adaLN_modulation.1.weight"] = rn(2 * d, d)
    w["adaln_before_head.adaLN_modulation.1.bias"] = rn(2 * d, std=0.05)
    w["lm_head.weight"] = rn(codebook, d)
    w["lm_head.bias"] = rn(codebook, std=0.01)
    return w


class RAROracle:
    def __init__(self, weights, depth, heads, codebook=1024, n_cls=1000):
        self.w = weights
        self.depth = depth
        self.heads = heads
        self.codebook = codebook
        self.none_id = n_cls + codebook + 1
        self.reset()

    def reset(self):
        self.k = [None] * self.depth
        self.v = [None] * self.depth

    def _embed(self, tok_ids, positions, cond_ids):
        """tok_ids: int64[R, n] embedding rows (-1 = cls token); positions: list[int]; cond_ids int64[R]."""
        w = self.w
        x = []
        for j, i in enumerate(positions):
            t = w["cls_token"][0, 0].expand(tok_ids.shape[0], -1) if i == 0 else w["embeddings.weight"][tok_ids[:, j]]
            t = t + w["pos_embed"][0, i]
            if i >= 1:
                t = t + w["target_aware_pos_embed"][0, i + 1]
            x.append(t)
        x = torch.stack(x, dim=1)
        c = torch.stack([w["embeddings.weight"][cond_ids] + w["timesteps_embeddings"][0, i] for i in positions], dim=1)
        return x, c

    def forward_tokens(self, x, c, causal):
        """x, c fp32[R, n, d] (n = 2 at step 0, else 1) -> logits fp32[R, n, codebook]."""
        w = self.w
        R, n, d = x.shape
        H = self.heads
        hd = d // H
        for i in range(self.depth):
            p = f"blocks.{i}."
            mod = F.linear(F.silu(c), w[p + "adaLN_modulation.1.weight"], w[p + "adaLN_modulation.1.bias"])
            sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
            a = F.layer_norm(x, (d,), w[p + "norm1.weight"], w[p + "norm1.bias"], 1e-6) * (1 + sc1) + sh1
            qkv = F.linear(a, w[p + "attn.qkv.weight"], w[p + "attn.qkv.bias"]).reshape(R, n, 3, H, hd)
            q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
            q = F.layer_norm(q, (hd,), w[p + "attn.q_norm.weight"], w[p + "attn.q_norm.bias"], 1e-6)
            k = F.layer_norm(k, (hd,), w[p + "attn.k_norm.weight"], w[p + "attn.k_norm.bias"], 1e-6)
            self.k[i] = k if self.k[i] is None else torch.cat([self.k[i], k], dim=-2)
            self.v[i] = v if self.v[i] is None else torch.cat([self.v[i], v], dim=-2)
            att = (q @ self.k[i].transpose(-2, -1)) * (hd ** -0.5)
            if causal and n > 1:
                m = torch.full((n, n), float("-inf")).triu_(1)
                att = att + m
            y = (F.softmax(att, dim=-1) @ self.v[i]).transpose(1, 2).reshape(R, n, d)
            x = x + g1 * F.linear(y, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
            m2 = F.layer_norm(x, (d,), w[p + "norm2.weight"], w[p + "norm2.bias"], 1e-6) * (1 + sc2) + sh2
            m2 = F.linear(F.gelu(F.linear(m2, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])),
                          w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
            x = x + g2 * m2
        mod = F.linear(F.silu(c), w["adaln_before_head.adaLN_modulation.1.weight"],
                       w["adaln_before_head.adaLN_modulation.1.bias"])
        scale, shift = mod.chunk(2, dim=-1)  # scale FIRST (:131)
        x = F.layer_norm(x, (d,), None, None, 1e-6) * (1 + scale) + shift
        return F.linear(x, w["lm_head.weight"], w["lm_head.bias"])

    def step(self, step_idx, cond_rows, last_ids):
        """cond_rows int64[R] embedding ids of the (cond | none-cond) rows; last_ids int64[R] previous image token."""
        if step_idx == 0:
            tok = torch.stack([torch.full_like(cond_rows, -1), cond_rows], dim=1)
            x, c = self._embed(tok, [0, 1], cond_rows)
            return self.forward_tokens(x, c, causal=True)[:, -1]
        x, c = self._embed(last_ids.view(-1, 1), [step_idx + 1], cond_rows)
        return self.forward_tokens(x, c, causal=False)[:, -1]


@torch.no_grad()
def generate(oracle, condition, steps=256, guidance_scale=4.0, temperature=1.0, green_row_fn=None, delta=0.0,
             noise=None, greedy=False, return_logits=False):
    """condition int64[B] class ids -> ids int64[B, steps]  (rar.py:408-459 with guidance_scale_pow = 0)."""
    oracle.reset()
    B = condition.shape[0]
    cond = condition + oracle.codebook + 1
    rows = torch.cat([cond, torch.full_like(cond, oracle.none_id)])
    ids = torch.zeros((B, 0), dtype=torch.long)
    all_logits = []
    for s in range(steps):
        scale_pow = torch.ones(1) * 0.0
        scale_step = (1 - torch.cos(((s / steps) ** scale_pow) * torch.pi)) * 1 / 2
        cfg = ((guidance_scale - 1) * scale_step + 1).item()
        last = torch.cat([ids[:, -1], ids[:, -1]]) if s > 0 else None
        lg = oracle.step(s, rows, last)
        logits = lg[B:] + (lg[:B] - lg[B:]) * cfg
        if return_logits:
            all_logits.append(logits.clone())
        mrows = None
        if green_row_fn is not None:
            mrows = green_row_fn(ids)
        nxt = sampling.sample_step(logits, mrows, delta, temperature, None, None,
                                   None if noise is None else noise[s], greedy=greedy)
        ids = torch.cat([ids, nxt.view(-1, 1)], dim=-1)
    if return_logits:
        return ids, torch.stack(all_logits)
    return ids
