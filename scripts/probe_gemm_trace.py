"""On-box probe: globaltimer stamps inside one kind of skinny GEMM (last launch of the run) on the default decode path."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "scripts")
kind = sys.argv[1] if len(sys.argv) > 1 else "11"
os.environ["WMAR_GEMM_TRACE"] = kind
from helpers import make_wm  # noqa: E402
from probe_step import weights  # noqa: E402
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402

NAMES = {0: "start", 1: "dep", 2: "stats", 3: "it0 w", 4: "it1 w", 5: "it2 w", 6: "it3 w", 7: "loop end", 8: "not last", 9: "is last",
         10: "reduced", 11: "stored"}
w = weights(16384, 256, 48, 24, 1536)
eng = TamingGPTEngine(w, 48, 24)
cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:16]
eng.sample(cond, 200, 1.0, 250, 0.92, make_wm("taming"), seed=1)
out = np.zeros(16, dtype=np.uint64)
fn = _lib.lib().wmar_debug_gemm_trace
fn.argtypes = [ctypes.c_void_p]
assert fn(out.ctypes.data) == 0
t0 = int(out[0])
print(f"kind {kind}:", "  ".join(f"{NAMES[e]} {(int(out[e]) - t0) / 1e3:.2f}" for e in range(12) if out[e]))
