"""xformers.ops.fmha stand-in: memory_efficient_attention_forward for the padded-keys block-diagonal causal bias."""
import math

import torch

from . import attn_bias  # noqa: F401


def memory_efficient_attention_forward(query, key, value, attn_bias=None, p=0.0, scale=None, *, op=None):
    """query [1, Mq, Hkv, G, K], key / value [1, Mk, Hkv, G, K] (grouped-query layout).  Query i of sequence b (n_b new
    tokens, L_b keys) attends to keys 0 .. L_b - n_b + i of its own cache block; softmax in fp32, output in q's dtype."""
    assert query.dim() == 5 and query.shape[0] == 1 and p == 0.0
    K = query.shape[-1]
    sc = 1.0 / math.sqrt(K) if scale is None else scale
    q_start = attn_bias.q_seqinfo.seqstart_py
    k_start = attn_bias.k_seqinfo.seqstart_py
    k_len = attn_bias.k_seqinfo.seqlen.tolist()
    out = torch.zeros_like(query)
    for b in range(len(k_len)):
        q0, q1 = q_start[b], q_start[b + 1]
        n, L = q1 - q0, k_len[b]
        if n == 0:
            continue
        q = query[0, q0:q1].float()                                 # [n, Hkv, G, K]
        k = key[0, k_start[b]:k_start[b] + L].float()               # [L, Hkv, G, K]
        v = value[0, k_start[b]:k_start[b] + L].float()
        s = torch.einsum("nhgk,lhgk->hgnl", q, k) * sc
        i = torch.arange(n)[:, None]
        l = torch.arange(L)[None, :]
        s = s.masked_fill(l > (L - n + i), float("-inf"))
        pr = torch.softmax(s, dim=-1)
        out[0, q0:q1] = torch.einsum("hgnl,lhgk->nhgk", pr, v).to(query.dtype)
    return out
