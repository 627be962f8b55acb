"""On-box probe: time-stamps (globaltimer) of the fused attention / MLP block kernels of one layer at one token step."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
STEP = int(sys.argv[1]) if len(sys.argv) > 1 else 200
os.environ["WMAR_STEP_TRACE"] = str(STEP)
from helpers import make_wm  # noqa: E402
from probe_step import weights  # noqa: E402
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402

NAMES = {0: "start", 1: "P.first", 2: "P.last", 3: "C.a_full0", 4: "C.done0", 5: "C.last", 6: "M.b1", 7: "M.commit1",
         8: "M.b2", 9: "M.last", 10: "W.dep", 11: "W.b1", 12: "W.xbar1", 13: "W.qkv", 14: "W.att", 15: "W.xbar2",
         16: "W.b2", 17: "E.acc0", 18: "E.part", 19: "E.last", 20: "end", 21: "W.scores", 22: "W.softmax", 24: "kv0", 25: "kv1", 26: "kv2", 27: "kv3", 28: "kv4", 29: "kv5", 30: "kv6", 31: "kv7", 32: "B1.combined", 33: "B1.bar", 34: "B1.stored", 35: "B1.stats_in", 36: "B1.x0_in", 37: "B1.xN_in", 40: "ALL.start.min", 41: "ALL.start.max", 42: "ALL.end.min", 43: "ALL.end.max"}


def main():
    V, block, L, H, d = 16384, 256, 48, 24, 1536
    w = weights(V, block, L, H, d)
    eng = TamingGPTEngine(w, L, H)
    cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:16]
    eng.sample(cond, block, 1.0, 250, 0.92, make_wm("taming"), seed=1)
    out = np.zeros(144, dtype=np.uint64)
    fn = _lib.lib().wmar_gpt_debug_trace
    fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    _lib.check(fn(eng.handle, out.ctypes.data))
    t0 = int(out[0])
    for k, nm in ((0, "ATT"), (1, "MLP"), (2, "ATT next layer")):
        print(f"=== {nm} block, CTA (0,0), step {STEP}: us since the attention kernel's start")
        for e in range(48):
            v = int(out[48 * k + e])
            if v and e in NAMES:
                print(f"  {NAMES[e]:10s} {(v - t0) / 1e3:9.2f}")


if __name__ == "__main__":
    sys.path.insert(0, "scripts")
    main()
