"""Generates tests/golden/chameleon_sampling.npz by running the reference's own Chameleon logits processors and token
selector (imported from /root/reference; cannot travel to the GPU box) on seeded logits:

    InBatchInstructCFGLogitsProcessor -> [watermark] -> AllowOnlyTokensLogitsProcessor -> TemperatureLogitsWarper ->
    TopPLogitsWarper -> softmax -> ReplicatedInputTokenSelector(Multinomial | Argmax, n=3)
    (deps/chameleon/inference/chameleon.py:312-327,338-346; generation.py:86-97)

    python oracle/gen_golden_chameleon.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, "/root/reference/deps")

from transformers import LogitsProcessorList, TemperatureLogitsWarper, TopPLogitsWarper  # noqa: E402

from chameleon.inference.logits_processor import AllowOnlyTokensLogitsProcessor, InBatchInstructCFGLogitsProcessor  # noqa: E402
from chameleon.inference.token_selector import (ArgmaxTokenSelector, MultinomialTokenSelector,  # noqa: E402
                                                ReplicatedInputTokenSelector)


def main():
    out = {}
    V, lo, hi, B = 2048, 4, 516, 3
    torch.manual_seed(11)
    logits = torch.randn(3 * B, V) * 3.0
    input_ids = torch.zeros(3 * B, 5, dtype=torch.long)
    green = torch.randperm(V)[: V // 4]
    delta = 2.0

    def wm(ids, lg):  # a fixed greenlist processor with the reference's callback protocol (chameleon.py:320-321)
        lg[:, green] += delta
        return lg

    for name, use_wm, temp, top_p, greedy in (("a", True, 0.9, 0.9, False), ("b", False, 1.0, 0.5, False), ("c", True, 0.7, 0.9, True)):
        procs = [InBatchInstructCFGLogitsProcessor(3.0, 1.2)]
        if use_wm:
            procs.append(wm)
        procs += [AllowOnlyTokensLogitsProcessor(list(range(lo, hi))), TemperatureLogitsWarper(temp), TopPLogitsWarper(top_p)]
        lp = LogitsProcessorList(procs)
        l = lp(input_ids, logits.clone())
        probs = l.softmax(dim=1)
        sel = ReplicatedInputTokenSelector(ArgmaxTokenSelector() if greedy else MultinomialTokenSelector(), n=3)
        torch.manual_seed(5)
        ids = sel(input_ids, probs)
        # the Exp(1) draws torch.multinomial consumed for the primary rows (probs[:B], full vocabulary)
        torch.manual_seed(5)
        noise = torch.empty(B, V).exponential_(1)
        out[f"{name}/ids"] = ids.numpy()
        out[f"{name}/noise"] = noise.numpy()
        out[f"{name}/processed"] = l[:B].numpy()
        out[f"{name}/cfg"] = np.array([int(use_wm), temp, top_p, int(greedy)], dtype=np.float64)
    out["logits"] = logits.numpy()
    out["green"] = green.numpy()
    out["meta"] = np.array([V, lo, hi, B], dtype=np.int64)
    path = os.path.join(ROOT, "tests", "golden", "chameleon_sampling.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
