"""xformers.ops stand-in: RMSNorm, rope_padded, fmha (see ../__init__.py)."""
import torch
from torch import nn

from . import fmha  # noqa: F401


class RMSNorm(nn.Module):
    """x * rsqrt(mean(x^2) + eps) * weight, evaluated in fp32, returned in x's dtype (xformers.ops.RMSNorm)."""

    def __init__(self, dim, include_weight=True, eps=1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim)) if include_weight else None

    def forward(self, x):
        xf = x.float()
        y = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + self.eps)
        if self.weight is not None:
            y = y * self.weight.float()
        return y.to(x.dtype)


def rope_padded(xq, xk, xv, cache_k, cache_v, attn_bias, *, theta=10000.0, out_q=None, adjacents=True, internal_dtype=""):
    """Rotary embedding of the new queries / keys of every sequence at their absolute positions, keys and values
    appended to the padded cache (sequence b owns cache rows k_seqstart[b] .. +kv_padding).  adjacents=True rotates the
    pairs (x[2j], x[2j+1]) by pos * theta^(-2j/hd); math in fp32, results stored in the tensors' dtype."""
    assert adjacents and xq.dim() == 4 and xq.shape[0] == 1
    hd = xq.shape[-1]
    j = torch.arange(hd // 2, dtype=torch.float32)
    freq = torch.pow(torch.tensor(float(theta), dtype=torch.float32), -2.0 * j / hd)
    q_start = attn_bias.q_seqinfo.seqstart_py
    k_start = attn_bias.k_seqinfo.seqstart_py
    k_len = attn_bias.k_seqinfo.seqlen.tolist()
    out = torch.empty_like(xq) if out_q is None else out_q

    def rot(x, pos):            # x [n, H, hd], pos [n]
        ang = pos.to(torch.float32)[:, None, None] * freq
        cs, sn = torch.cos(ang), torch.sin(ang)
        xf = x.float().reshape(*x.shape[:-1], hd // 2, 2)
        x0, x1 = xf[..., 0], xf[..., 1]
        return torch.stack([x0 * cs - x1 * sn, x0 * sn + x1 * cs], dim=-1).reshape(x.shape).to(x.dtype)

    for b in range(len(k_len)):
        q0, q1 = q_start[b], q_start[b + 1]
        n = q1 - q0
        if n == 0:
            continue
        pos = torch.arange(k_len[b] - n, k_len[b])
        out[0, q0:q1] = rot(xq[0, q0:q1], pos)
        c0 = k_start[b] + k_len[b] - n
        cache_k[0, c0:c0 + n] = rot(xk[0, q0:q1], pos)
        cache_v[0, c0:c0 + n] = xv[0, q0:q1]
    return out
