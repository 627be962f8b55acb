"""Debug: persistent step kernel vs per-GEMM graph on the small golden configs, logits of the first steps."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import gpt as ogpt  # noqa: E402
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402

g = np.load("tests/golden/gpt.npz")
for name in ("tiny", "narrow"):
    V, block, L, H, d, steps, B, seed = [int(x) for x in g[f"{name}/cfg"]]
    w = ogpt.synthetic_gpt_weights(V, block, L, H, d, seed=seed)
    for Bt in (4, 16, 9, 1):
        cond = torch.arange(Bt) * 7 % 1000
        res = {}
        for mode in ("graph", "pstep"):
            os.environ["WMAR_STEP"] = mode
            eng = TamingGPTEngine(w, L, H)
            codes, logits = eng.sample(cond, 6, 1.0, None, None, None, greedy=True, return_logits=True)
            rc = _lib.lib().wmar_check_device_flag(_lib.current_stream())
            if rc:
                print(mode, Bt, _lib.lib().wmar_last_error().decode())
            res[mode] = (codes.cpu(), logits.cpu(), rc)
        for n in range(2):
            dl = (res["graph"][1][n] - res["pstep"][1][n]).abs()
            print(f"{name} d={d} L={L} H={H} B={Bt} step {n}: max|dlogit| {dl.max():.3e} per row {[f'{x:.1e}' for x in dl.max(dim=1).values.tolist()]} "
                  f"flag rc graph {res['graph'][2]} pstep {res['pstep'][2]}", flush=True)
