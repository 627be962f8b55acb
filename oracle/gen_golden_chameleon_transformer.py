"""Generates tests/golden/chameleon_transformer.npz by running the reference's OWN Chameleon `Transformer`
(deps/chameleon/inference/transformer.py, imported unmodified from /root/reference; built the way loader.py:16-33 does,
default dtype bf16) on the CPU of the build container.  The three xformers operators it imports (RMSNorm, rope_padded,
fmha.memory_efficient_attention_forward + the padded-keys causal bias) are not installable here; they come from
oracle/xformers_stub (documented semantics in plain torch).  Everything else -- the module graph, the fused wqkv / w13
layouts, q/k LayerNorm placement, GQA expansion, residual order, every bf16 rounding point of the Linear layers, the
final `.float()` -- is the reference's code running.

Protocol = ChameleonModelAdapter (model_adapter.py:72-118): one ragged prefill over all rows, then one token per row per
call with the same bias object and k_seqinfo.seqlen += 1.  The next tokens are a fixed seeded sequence (teacher forcing),
so the golden holds logits only; rows are the three guidance groups of B images
(chameleon.py:351-372) and the three rows of an image are fed the same token, as ImageDecoder does.

    python oracle/gen_golden_chameleon_transformer.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "xformers_stub"))
sys.path.insert(0, "/root/reference")

from deps.chameleon.inference.transformer import ModelArgs, Transformer, make_cache  # noqa: E402
from deps.chameleon.inference import transformer as ref_tf  # noqa: E402

from oracle import chameleon as oc  # noqa: E402

CASES = {
    # name: (V, d, L, H, Hkv, ffn_dim_multiplier, multiple_of -> Fh, prompts, steps, seed)
    "tiny": (1024, 256, 2, 2, 2, 0.5, 128, 384, [[0, 901, 902, 903, 904, 700], [0, 950, 951, 700], [0, 700], [0, 700], [0, 700], [0, 700]], 12, 5),
    "gqa": (768, 512, 3, 4, 2, 0.5, 128, 768, [[0, 11, 12, 13, 700], [0, 21, 700], [0, 700]], 8, 9),
}


def build(V, d, L, H, Hkv, mult, mo, Fh, seed):
    w = oc.synthetic_chameleon_weights(V, d, L, H, Hkv, Fh, seed=seed)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)              # loader.py:17-20
    model = Transformer(ModelArgs(dim=d, n_layers=L, n_heads=H, n_kv_heads=Hkv, vocab_size=V, ffn_dim_multiplier=mult,
                                  multiple_of=mo, norm_eps=1e-5, rope_theta=10000.0, qk_normalization=True, swin_norm=False))
    torch.set_default_dtype(old)
    res = model.load_state_dict(w, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    return w, model.eval()


def run(model, prompts, forced, max_seq):
    """ChameleonModelAdapter.__call__ restated around the reference model's forward_with_attn_bias."""
    R = len(prompts)
    cache = make_cache(model.args, R * max_seq, dtype=torch.bfloat16)
    lens = [len(p) for p in prompts]
    bias = ref_tf.AttnBias.from_seqlens(q_seqlen=lens, kv_seqlen=lens, kv_padding=max_seq)
    flat = model.forward_with_attn_bias(torch.tensor([t for p in prompts for t in p]), bias, cache)
    last = torch.stack([flat[sum(lens[: r + 1]) - 1] for r in range(R)])
    out = [last]
    bias.q_seqinfo.seqstart.copy_(torch.arange(R + 1, dtype=torch.int))
    bias.q_seqinfo.max_seqlen = 1
    bias.q_seqinfo.seqstart_py = bias.q_seqinfo.seqstart.tolist()
    for s in range(forced.shape[1]):
        bias.k_seqinfo.seqlen.add_(1)
        out.append(model.forward_with_attn_bias(forced[:, s].clone(), bias, cache))
    return torch.stack(out)                                # [steps + 1, R, V] fp32


def main():
    out = {}
    for name, (V, d, L, H, Hkv, mult, mo, Fh, prompts, steps, seed) in CASES.items():
        w, model = build(V, d, L, H, Hkv, mult, mo, Fh, seed)
        assert model.layers[0].feed_forward.w2.weight.shape[1] == Fh
        g = torch.Generator().manual_seed(100 + seed)
        B = len(prompts) // 3                              # rows = [full | image-conditioned | unconditioned] x B images;
        forced = torch.randint(4, min(V, 516), (B, steps), generator=g).repeat(3, 1)   # one token per image, fed to its 3 rows
        with torch.no_grad():
            logits = run(model, prompts, forced, 64)
        out[f"{name}/logits"] = logits.numpy()
        out[f"{name}/forced"] = forced.numpy()
        out[f"{name}/meta"] = np.array([V, d, L, H, Hkv, Fh, steps, seed], dtype=np.int64)
        out[f"{name}/prompt_lens"] = np.array([len(p) for p in prompts], dtype=np.int64)
        out[f"{name}/prompts_flat"] = np.array([t for p in prompts for t in p], dtype=np.int64)
        # the restatement on the same inputs, for the record printed below
        o = oc.ChameleonOracle(w, L, H, Hkv)
        worst = 0.0
        for r, p in enumerate(prompts):
            lg = None
            for pos, t in enumerate(p):
                lg = o.step_row(r, t, pos)
            worst = max(worst, (lg - logits[0, r]).abs().max().item())
            for s in range(steps):
                lg = o.step_row(r, int(forced[r, s]), len(p) + s)
                worst = max(worst, (lg - logits[s + 1, r]).abs().max().item())
        print(f"{name}: logits {tuple(logits.shape)}, range {logits.abs().max().item():.3f}, restatement max |diff| {worst:.3e}")
    path = os.path.join(ROOT, "tests", "golden", "chameleon_transformer.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
