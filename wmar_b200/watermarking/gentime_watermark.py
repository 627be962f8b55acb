"""Host-side mirror of the reference operator ``wmar.watermarking.gentime_watermark.GentimeWatermark``.

Same constructor, ``spawn_logit_processor()``, ``detect()``, ``__str__`` and ``create_watermarker_from_string`` as
wmar/watermarking/gentime_watermark.py:110-366, but every call lands in one CUDA kernel of libwmar_b200.so:

  * construction builds the greenlist BITMASK TABLE on the device once (one row per context sum) instead of
    re-seeding torch's CPU generator and running two ``randperm`` per row per token (:161-226);
  * the logit processor ``f(past_ids, logits)`` is a single launch with no ``.item()`` sync (:229-271);
  * ``detect`` de-duplicates n-grams, looks the bits up and evaluates the p-value on the device (:285-344).
"""
from enum import Enum
from functools import partial
from typing import Union

import torch

from .. import _lib


class SeedStrategy(Enum):
    FIXED = "fixed"
    LINEAR = "linear"
    SPATIAL = "spatial"


class SplitStrategy(Enum):
    RANDOM = "rand"
    RANDOM_STRATIFIED = "stratifiedrand"
    CLUSTERING = "clustering"


_SEED_CODE = {SeedStrategy.FIXED: 0, SeedStrategy.LINEAR: 1, SeedStrategy.SPATIAL: 2}
_SPLIT_CODE = {SplitStrategy.RANDOM: 0, SplitStrategy.RANDOM_STRATIFIED: 1}


class GentimeWatermark:
    def __init__(
        self,
        vq: Union[object, dict],
        vocab_size: int,
        seed_strategy: SeedStrategy,
        split_strategy: SplitStrategy,
        context_size: int,
        delta: float,
        gamma: float,
        device="cuda",
        spatial_dim=16,
        salt_key=15485863,
        build_on="device",
    ) -> None:
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.WmarError("wmar_b200.GentimeWatermark runs on CUDA only (no CPU fallback); got device="
                                 f"{device!r}")
        self.vocab_size = vocab_size
        if isinstance(vq, dict):
            alive, dead = vq["alive_ids"], vq["dead_ids"]
        else:
            alive, dead = vq.alive_ids, vq.dead_ids
        self.alive_ids = torch.as_tensor(alive, dtype=torch.long).to(self.device).contiguous()
        self.dead_ids = torch.as_tensor(dead, dtype=torch.long).to(self.device).contiguous()

        self.salt_key = salt_key
        self.seed_strategy = SeedStrategy(seed_strategy)
        self.split_strategy = SplitStrategy(split_strategy)
        self.context_size = int(context_size)
        self.delta = float(delta)
        self.gamma = float(gamma)
        self.greenlist_size = int(self.vocab_size * self.gamma)
        self.spatial_dim = spatial_dim
        if self.split_strategy is SplitStrategy.CLUSTERING:
            # gentime_watermark.py:175-177: one fixed greenlist from t-SNE + KMeans of the alive codebook vectors (host,
            # once per watermarker) -> the single row of the bitmask table
            assert self.seed_strategy is SeedStrategy.FIXED, "Clustering only with fixed seeding"
            from .clustering import clustering_greenlist_ids, ids_to_bitmask_row
            emb = vq["embedding"] if isinstance(vq, dict) else vq.embedding.weight
            ids = clustering_greenlist_ids(emb, self.alive_ids.cpu(), self.dead_ids.cpu())
            self.n_rows = 1
            self.table = torch.from_numpy(ids_to_bitmask_row(ids, self.vocab_size)).view(1, -1).to(self.device)
            self._params = _lib.WmParams(self.table.data_ptr(), self.n_rows, self.vocab_size,
                                         _SEED_CODE[self.seed_strategy], self.context_size, self.spatial_dim, self.delta,
                                         self.gamma)
            return
        if self.seed_strategy is SeedStrategy.SPATIAL and self.context_size not in (1, 3):
            raise AssertionError("Spatial seeding only implemented for context size in [1,3]")

        # one table row per possible context SUM (only the sum enters the seed, gentime_watermark.py:225)
        if self.seed_strategy is SeedStrategy.FIXED or self.context_size == 0:
            self.n_rows = 1          # context_size 0: the context sum is always 0, only row 0 is reachable
        else:
            self.n_rows = max(self.context_size, 1) * (self.vocab_size - 1) + 1
        words = (self.vocab_size + 31) // 32
        L = _lib.lib()
        if build_on == "device":
            self.table = torch.empty((self.n_rows, words), dtype=torch.int32, device=self.device)
            with torch.cuda.device(self.device):
                _lib.check(L.wmar_greenlist_build_device(
                    self.vocab_size, self.gamma, _SPLIT_CODE[self.split_strategy], _SEED_CODE[self.seed_strategy],
                    self.salt_key, _lib.ptr(self.alive_ids), self.alive_ids.numel(), _lib.ptr(self.dead_ids),
                    self.dead_ids.numel(), self.n_rows, _lib.ptr(self.table), _lib.current_stream()))
        else:  # host build (bit-identical), then one H2D copy
            host = torch.empty((self.n_rows, words), dtype=torch.int32)
            a, d = self.alive_ids.cpu().contiguous(), self.dead_ids.cpu().contiguous()
            _lib.check(L.wmar_greenlist_build_host(
                self.vocab_size, self.gamma, _SPLIT_CODE[self.split_strategy], _SEED_CODE[self.seed_strategy],
                self.salt_key, _lib.ptr(a), a.numel(), _lib.ptr(d), d.numel(), self.n_rows, _lib.ptr(host), 0))
            self.table = host.to(self.device)
        self._params = _lib.WmParams(self.table.data_ptr(), self.n_rows, self.vocab_size,
                                     _SEED_CODE[self.seed_strategy], self.context_size, self.spatial_dim, self.delta,
                                     self.gamma)

    def __str__(self):
        ret = f"{self.seed_strategy.value}-{self.split_strategy.value}-"
        ret += f"h={self.context_size}-d={self.delta:.1f}-g={self.gamma:.2f}"
        return ret

    # -- C-ABI view used by the engines ---------------------------------------------------------------------
    def c_params(self):
        return self._params

    def greenlist_ids_for_sum(self, ctx_sum: int) -> torch.LongTensor:
        """ids whose bit is set in table row `ctx_sum` (sorted; the reference returns them in shuffle order)."""
        row = self.table[0 if self.seed_strategy is SeedStrategy.FIXED else ctx_sum]
        bits = (row.view(-1, 1) >> torch.arange(32, device=row.device, dtype=torch.int32)) & 1
        return torch.nonzero(bits.reshape(-1)[: self.vocab_size]).flatten()

    # past_ids: [B, len], logits: [B, vocab_size]  -- in place, and returned (gentime_watermark.py:229-271)
    def _process_logits(self, past_ids: torch.LongTensor, logits: torch.Tensor) -> torch.Tensor:
        assert logits.shape[-1] == self.vocab_size, f"Logits shape mismatch: {logits.shape} vs {self.vocab_size}"
        assert logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous(), \
            "logits must be a contiguous fp32 CUDA tensor"
        B = logits.shape[0]
        if past_ids.dtype != torch.long or not past_ids.is_cuda:
            past_ids = past_ids.to(device=logits.device, dtype=torch.long)
        t = past_ids.shape[1]
        stride = past_ids.stride(0) if t > 0 else 0
        if t > 0 and past_ids.stride(1) != 1:
            past_ids = past_ids.contiguous()
            stride = past_ids.stride(0)
        with torch.cuda.device(logits.device):
            _lib.check(_lib.lib().wmar_wm_process_logits(
                self._params, ctypes_ptr_or_none(past_ids, t), B, t, stride, ctypes_ptr(logits),
                _lib.current_stream()))
        return logits

    def spawn_logit_processor(self):
        return partial(self._process_logits)

    def detect_stats(self, codes: torch.LongTensor, return_masks: bool = False):
        """codes [B, L] -> dict(n_green int32[B], n_scored int32[B], z f64[B], pvalue f64[B][, masks list])."""
        codes = codes.to(device=self.device, dtype=torch.long).contiguous()
        B, L = codes.shape
        ng = torch.empty(B, dtype=torch.int32, device=self.device)
        ns = torch.empty(B, dtype=torch.int32, device=self.device)
        z = torch.empty(B, dtype=torch.float64, device=self.device)
        p = torch.empty(B, dtype=torch.float64, device=self.device)
        mask = mlen = None
        stride = 0
        if return_masks:
            stride = L + self.context_size + 1
            mask = torch.full((B, stride), -2, dtype=torch.int8, device=self.device)
            mlen = torch.empty(B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.lib().wmar_detect(self._params, _lib.ptr(codes), B, L, _lib.ptr(ng), _lib.ptr(ns),
                                              _lib.ptr(z), _lib.ptr(p), _lib.ptr(mask), stride, _lib.ptr(mlen),
                                              _lib.current_stream()))
            _lib.check_device_flag()   # codes outside the greenlist table raise (the reference's indexing would)
        out = {"n_green": ng, "n_scored": ns, "z": z, "pvalue": p}
        if return_masks:
            lens = mlen.cpu().tolist()
            m = mask.cpu()
            out["masks"] = [m[b, : lens[b]].tolist() for b in range(B)]
        return out

    # codes: [B, len] -> p-values (float64 tensor on self.device)  (gentime_watermark.py:322-344)
    def detect(self, codes: torch.LongTensor, return_masks: bool = False):
        st = self.detect_stats(codes, return_masks)
        if return_masks:
            return st["pvalue"], st["masks"]
        return st["pvalue"]


def ctypes_ptr(t):
    return _lib.ptr(t)


def ctypes_ptr_or_none(t, n):
    return _lib.ptr(t) if n > 0 else None


# For example: fixed-stratifiedrand-h=0-d=8.0-g=0.50
def create_watermarker_from_string(vq, vocab_size: int, method: str, device: str) -> GentimeWatermark:
    parts = method.split("-")
    seed_strategy = parts[0]
    split_strategy = parts[1]
    context_size = int(parts[2].split("=")[1])
    delta = float(parts[3].split("=")[1])
    gamma = float(parts[4].split("=")[1])
    return GentimeWatermark(vq, vocab_size, SeedStrategy(seed_strategy), SplitStrategy(split_strategy), context_size,
                            delta, gamma, device=device)
