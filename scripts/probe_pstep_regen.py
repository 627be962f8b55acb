"""Debug: several generations back to back on ONE engine (persistent step kernel vs per-GEMM graph), small or full shapes.

    python scripts/probe_pstep_regen.py [narrow|full] [steps] [generations]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpt as ogpt  # noqa: E402
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402
from wmar_b200.models.synthetic import TAMING_GPT_CFG, gpt_state  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "narrow"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 24
gens = int(sys.argv[3]) if len(sys.argv) > 3 else 4
if which == "full":
    c = TAMING_GPT_CFG
    w = gpt_state(c, seed=0, device="cuda")
    L, H = c["n_layer"], c["n_head"]
else:
    g = np.load("tests/golden/gpt.npz")
    V, block, L, H, d, _, _, seed = [int(x) for x in g["narrow/cfg"]]
    w = ogpt.synthetic_gpt_weights(V, block, L, H, d, seed=seed)
    steps = min(steps, block)
res = {}
for mode in ("graph", "pstep"):
    os.environ["WMAR_STEP"] = mode
    eng = TamingGPTEngine(w, L, H)
    out = []
    for i in range(gens):
        B = 16 if i % 2 == 0 else 5
        cond = (torch.arange(B) * 7 + i) % 1000
        try:
            ids = eng.sample(cond, steps, 1.0, 250, 0.92, None, greedy=True)
            torch.cuda.synchronize()
        except Exception as e:
            print(mode, "generation", i, "FAILED:", str(e)[:200], flush=True)
            sys.exit(1)
        out.append(ids.cpu())
        print(mode, "generation", i, "B", B, "ok", flush=True)
    res[mode] = out
    del eng
for i in range(gens):
    same = (res["graph"][i] == res["pstep"][i]).float().mean().item()
    print(f"generation {i}: ids equal {same * 100:.2f}%")
