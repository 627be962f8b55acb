"""xformers.ops.fmha.attn_bias stand-in: BlockDiagonalCausalWithOffsetPaddedKeysMask (fields the reference touches:
q_seqinfo.{seqstart, seqstart_py, max_seqlen}, k_seqinfo.{seqstart_py, seqlen}; model_adapter.py:77-118)."""
from dataclasses import dataclass

import torch


@dataclass
class _SeqLenInfo:
    seqstart: torch.Tensor
    seqstart_py: list
    max_seqlen: int


@dataclass
class _PaddedSeqLenInfo(_SeqLenInfo):
    seqlen: torch.Tensor = None
    padding: int = 0


@dataclass
class BlockDiagonalCausalWithOffsetPaddedKeysMask:
    q_seqinfo: _SeqLenInfo
    k_seqinfo: _PaddedSeqLenInfo

    @classmethod
    def from_seqlens(cls, q_seqlen, kv_padding, kv_seqlen, causal_diagonal=None):
        assert len(q_seqlen) == len(kv_seqlen)
        qs = [0]
        for n in q_seqlen:
            qs.append(qs[-1] + int(n))
        ks = [kv_padding * i for i in range(len(kv_seqlen) + 1)]
        q = _SeqLenInfo(torch.tensor(qs, dtype=torch.int32), qs, max(q_seqlen) if len(q_seqlen) else 0)
        k = _PaddedSeqLenInfo(torch.tensor(ks, dtype=torch.int32), ks, max(kv_seqlen) if len(kv_seqlen) else 0,
                              seqlen=torch.tensor([int(x) for x in kv_seqlen], dtype=torch.int32), padding=kv_padding)
        return cls(q, k)
