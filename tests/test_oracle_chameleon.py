"""CPU: the Chameleon sampling restatement (oracle/chameleon.py) vs goldens produced by the reference's own logits
processors and token selector (oracle/gen_golden_chameleon.py); the transformer restatement vs logits of the reference's
own `Transformer` module run in the build container (oracle/gen_golden_chameleon_transformer.py), plus self-consistency."""
import os

import numpy as np
import torch

from helpers import G


def _pipeline(logits3, green, use_wm, temp, top_p, greedy, noise, lo, hi):
    from transformers import TopPLogitsWarper
    from oracle import chameleon as oc
    l = oc.instruct_cfg(logits3, 3.0, 1.2)
    if use_wm:
        l[:, green] += 2.0
    l = oc.allow_only(l, lo, hi) / temp
    l = TopPLogitsWarper(top_p)(None, l)
    probs = l.softmax(dim=1)
    ids = probs.argmax(dim=1) if greedy else (probs / noise).argmax(dim=1)
    return l, ids


def test_sampling_pipeline_matches_reference_classes():
    g = np.load(os.path.join(G, "chameleon_sampling.npz"))
    V, lo, hi, B = [int(x) for x in g["meta"]]
    logits3 = torch.from_numpy(g["logits"])
    green = torch.from_numpy(g["green"])
    for name in "abc":
        use_wm, temp, top_p, greedy = g[f"{name}/cfg"]
        l, ids = _pipeline(logits3.clone(), green, bool(use_wm), float(temp), float(top_p), bool(greedy),
                           torch.from_numpy(g[f"{name}/noise"]), lo, hi)
        np.testing.assert_array_equal(l.numpy(), g[f"{name}/processed"])
        np.testing.assert_array_equal(ids.repeat(3).numpy(), g[f"{name}/ids"])   # ReplicatedInputTokenSelector(n=3)


def test_transformer_restatement_is_causal_and_row_local():
    """Feeding a prompt token by token with the cache == the per-row key ranges of the reference's attention bias:
    a row's logits depend only on its own tokens, and the cache makes the order of rows irrelevant."""
    from oracle import chameleon as oc
    V, d, L, H, Fh = 320, 256, 2, 2, 128
    w = oc.synthetic_chameleon_weights(V, d, L, H, H, Fh, seed=3)
    o = oc.ChameleonOracle(w, L, H, H)
    toks = [5, 17, 200, 31]
    a = [o.step_row(0, t, i) for i, t in enumerate(toks)]
    o.reset()
    for i, t in enumerate([9, 8, 7]):
        o.step_row(1, t, i)                       # another row in between must not matter
    b = [o.step_row(0, t, i) for i, t in enumerate(toks)]
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert a[-1].dtype == torch.float32 and a[-1].shape == (V,)
    # bf16 logits: every value is exactly representable in bf16 (output.float(), transformer.py:319)
    assert torch.equal(a[-1], a[-1].to(torch.bfloat16).float())


def _golden_case(g, name):
    V, d, L, H, Hkv, Fh, steps, seed = [int(x) for x in g[f"{name}/meta"]]
    lens = [int(x) for x in g[f"{name}/prompt_lens"]]
    flat = [int(x) for x in g[f"{name}/prompts_flat"]]
    prompts, o = [], 0
    for n in lens:
        prompts.append(flat[o:o + n])
        o += n
    return (V, d, L, H, Hkv, Fh, steps, seed), prompts, torch.from_numpy(g[f"{name}/forced"]), torch.from_numpy(g[f"{name}/logits"])


def test_transformer_restatement_matches_reference_module():
    """Logits of the reference's Transformer (transformer.py:97-337, run with the xformers stand-in) on a ragged prefill +
    teacher-forced single-token passes, multi-head and grouped-query.  The model is bf16: the reference evaluates the
    prefill as one batched matmul, the restatement row by row, so individual bf16 roundings may flip (one bf16 ulp at
    the logit magnitude 2..4 is 2^-6); everything else is bit-equal."""
    from oracle import chameleon as oc
    g = np.load(os.path.join(G, "chameleon_transformer.npz"))
    for name in ("tiny", "gqa"):
        (V, d, L, H, Hkv, Fh, steps, seed), prompts, forced, want = _golden_case(g, name)
        o = oc.ChameleonOracle(oc.synthetic_chameleon_weights(V, d, L, H, Hkv, Fh, seed=seed), L, H, Hkv)
        n_exact = n_all = 0
        for r, p in enumerate(prompts):
            for pos, t in enumerate(p):
                lg = o.step_row(r, t, pos)
            rows = [lg] + [o.step_row(r, int(forced[r, s]), len(p) + s) for s in range(steps)]
            got = torch.stack(rows)
            assert (got - want[:, r]).abs().max().item() <= 2 ** -5
            n_exact += int((got == want[:, r]).sum())
            n_all += got.numel()
        assert n_exact >= 0.98 * n_all, (name, n_exact, n_all)
