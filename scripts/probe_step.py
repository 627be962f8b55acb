"""On-box probe: full-size Taming decode (B=16, 256 tokens) with the persistent step kernel vs the per-GEMM graph path;
times both and compares the sampled ids.  Not the bench."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from helpers import make_wm  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402


def weights(V, block, L, H, d):
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, std=0.02: torch.randn(*s, device="cuda", generator=g) * std
    w = {"tok_emb.weight": rn(V, d), "pos_emb": rn(1, block, d), "ln_f.weight": 1 + rn(d, std=0.1),
         "ln_f.bias": rn(d, std=0.1), "head.weight": rn(V, d)}
    for i in range(L):
        p = f"blocks.{i}."
        for ln in ("ln1", "ln2"):
            w[p + ln + ".weight"] = 1 + rn(d, std=0.1)
            w[p + ln + ".bias"] = rn(d, std=0.1)
        for nm in ("key", "query", "value", "proj"):
            w[p + f"attn.{nm}.weight"] = rn(d, d)
            w[p + f"attn.{nm}.bias"] = rn(d, std=0.01)
        w[p + "mlp.0.weight"] = rn(4 * d, d)
        w[p + "mlp.0.bias"] = rn(4 * d, std=0.01)
        w[p + "mlp.2.weight"] = rn(d, 4 * d)
        w[p + "mlp.2.bias"] = rn(d, std=0.01)
    return w


def main():
    small = "--small" in sys.argv
    V, block, L, H, d = (16384, 64, 4, 4, 256) if small else (16384, 256, 48, 24, 1536)
    steps = int(sys.argv[sys.argv.index("--steps") + 1]) if "--steps" in sys.argv else block
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 3
    B = 16
    w = weights(V, block, L, H, d)
    wm = make_wm("taming")
    cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:B]
    out = {}
    codes = {}
    modes = sys.argv[sys.argv.index("--modes") + 1].split(",") if "--modes" in sys.argv else ["fused", "graph", "graph_v0"]
    for mode in modes:
        os.environ["WMAR_STEP"] = "graph" if mode.startswith("graph") else "fused"
        import ctypes
        from wmar_b200 import _lib
        _lib.lib().wmar_debug_set_gemm_engine.argtypes = [ctypes.c_int]
        if mode in ("graph_v0", "graph_tc"):   # "graph" / "fused" keep the library default (mma.sync engine)
            _lib.lib().wmar_debug_set_gemm_engine(1 if mode == "graph_v0" else 0)
        eng = TamingGPTEngine(w, L, H)
        for rep in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            c = eng.sample(cond, steps, 1.0, 250, 0.92, wm, seed=1, greedy=(rep == 0))
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            by = eng.algorithmic_bytes(B, steps)
            print(f"{mode}: B={B} x{steps}: {ms:.1f} ms  {B/ms*1e3:.2f} img/s  {by/ms/1e6:.0f} GB/s algorithmic "
                  f"({ms/steps*1e3:.0f} us/token)", flush=True)
            out.setdefault(mode, []).append({"ms": ms, "GBps": by / ms / 1e6})
            if rep == 0:
                codes[mode] = c.cpu()
        st = wm.detect_stats(c)
        print(mode, "n_green", st["n_green"].tolist()[:4], flush=True)
        del eng
        torch.cuda.empty_cache()
    ks = list(codes)
    for k in ks[1:]:
        same = (codes[ks[0]] == codes[k]).float().mean().item()
        first_diff = (codes[ks[0]] != codes[k]).any(0).float().argmax().item() if same < 1 else -1
        print(f"greedy ids {ks[0]} vs {k}: {same*100:.2f}% equal, first differing step {first_diff}", flush=True)
        out[f"same_{k}"] = same
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_step.json", "w"), indent=1)


if __name__ == "__main__":
    main()
