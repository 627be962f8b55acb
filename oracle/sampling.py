"""One sampling step, restated in plain torch fp32 on CPU (TEST INFRASTRUCTURE ONLY).

Operator order:
  Taming  mingpt.py:349-363        +delta on green(ctx) -> /T -> top-k (keep ties) -> top-p -> softmax -> multinomial
  RAR     rar.py:441-454           u + (c-u)*s -> +delta on green (skipped when history is empty) -> /T -> softmax -> multinomial
Third-party bodies restated (not under /root/reference):
  transformers TopKLogitsWarper / TopPLogitsWarper (logits_process.py, 5.x): see top_k_filter / top_p_filter
  torch.multinomial(p, 1) == argmax(p / q), q ~ Exp(1) drawn with exponential_ over the whole [B, V] tensor
"""
import numpy as np
import torch


def apply_green_bias(logits, mask_rows, delta):
    """logits fp32[B,V] (modified in place, like gentime_watermark.py:267); mask_rows u32[B, V/32] or None rows.

    mask_rows[b] is None when the reference skips the row (history too short, :268-270)."""
    V = logits.shape[1]
    for b, row in enumerate(mask_rows):
        if row is None:
            continue
        bits = np.unpackbits(np.ascontiguousarray(row).view(np.uint8), bitorder="little")[:V].astype(bool)
        logits[b, torch.from_numpy(bits)] += delta
    return logits


def top_k_filter(scores, top_k):
    top_k = min(top_k, scores.size(-1))
    kth = torch.topk(scores, top_k)[0][..., -1, None]
    return scores.masked_fill(scores < kth, -float("inf"))


def top_p_filter(scores, top_p):
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    remove = cumulative_probs <= (1 - top_p)
    remove[..., -1:] = 0
    indices_to_remove = remove.scatter(1, sorted_indices, remove)
    return scores.masked_fill(indices_to_remove, -float("inf"))


def sample_step(logits, mask_rows, delta, temperature, top_k, top_p, noise, greedy=False):
    """logits fp32[B,V] -> ids int64[B].  noise fp32[B,V] ~ Exp(1) (the q of multinomial), ignored when greedy."""
    logits = logits.clone().float()
    if mask_rows is not None:
        logits = apply_green_bias(logits, mask_rows, delta)
    logits = logits / temperature
    if top_k is not None:
        logits = top_k_filter(logits, top_k)
    if top_p is not None:
        logits = top_p_filter(logits, top_p)
    probs = torch.softmax(logits, dim=-1)
    if greedy:
        return torch.topk(probs, k=1, dim=-1)[1][:, 0]  # mingpt.py:360-361 (first max index)
    return torch.argmax(probs / noise, dim=-1)
