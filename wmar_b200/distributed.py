"""Multi-GPU plumbing: one process per GPU, models replicated, independent images sharded.

The reference shards with `--chunk_id/--num_chunks` and no communication (generate.py:204-207).  Here rank r is chunk r;
torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests) is used only to broadcast the weights from rank 0
before generation and to gather the per-rank outputs afterwards -- nothing per token.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def broadcast_state(state, src=0):
    """In-place broadcast of every tensor of a flat state dict (sorted key order on every rank)."""
    _, n = world()
    if n == 1:
        return state
    for k in sorted(state):
        dist.broadcast(state[k], src=src)
    return state


def gather_rows(local, pad_value=-1):
    """All-gather tensors whose dim 0 differs per rank (e.g. codes int64[n_r, 256]); returns the list per rank."""
    rank, n = world()
    if n == 1:
        return [local]
    cnt = torch.tensor([local.shape[0]], dtype=torch.long, device=local.device)
    counts = [torch.zeros_like(cnt) for _ in range(n)]
    dist.all_gather(counts, cnt)
    m = int(max(c.item() for c in counts))
    padded = torch.full((m,) + tuple(local.shape[1:]), pad_value, dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(n)]
    dist.all_gather(out, padded)
    return [o[: int(c.item())] for o, c in zip(out, counts)]
