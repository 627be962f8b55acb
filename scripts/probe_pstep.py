"""Probe of the persistent step kernel at BASELINE configs[1] shapes: decode-loop time per path, ids equality, trace.

    python scripts/probe_pstep.py [steps] [modes...]     (modes: pstep pstep2 graph)
"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402
from wmar_b200.models.synthetic import TAMING_GPT_CFG, gpt_state  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
modes = sys.argv[2:] or ["pstep", "graph"]
c = TAMING_GPT_CFG
w = gpt_state(c, seed=0, device="cuda")
cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:16]
out = {}
for mode in modes:
    os.environ["WMAR_STEP"] = mode[:5]
    if mode == "pstep2":
        os.environ["WMAR_PSTEP_NG"] = "2"
    t0 = time.time()
    eng = TamingGPTEngine(w, c["n_layer"], c["n_head"])
    torch.cuda.synchronize()
    os.environ.pop("WMAR_PSTEP_NG", None)
    print(f"[{mode}] create {time.time() - t0:.2f} s", flush=True)
    ids = eng.sample(cond, steps, 1.0, 250, 0.92, None, greedy=True)
    torch.cuda.synchronize()
    rc = _lib.lib().wmar_check_device_flag(_lib.current_stream())
    print(f"[{mode}] device flag rc={rc} {_lib.lib().wmar_last_error().decode() if rc else ''}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        ids = eng.sample(cond, steps, 1.0, 250, 0.92, None, greedy=True)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    by = eng.algorithmic_bytes(16, steps)
    print(f"[{mode}] {steps} steps: {best:.1f} ms = {best * 1e3 / steps:.0f} us/token, {by / best / 1e6:.0f} GB/s algorithmic "
          f"= {by / best / 1e6 / 6557.8:.3f} of 6557.8", flush=True)
    out[mode] = ids.cpu()
    if mode.startswith("pstep") and os.environ.get("WMAR_PSTEP_TRACE"):
        L = _lib.lib()
        G = 148
        buf = (ctypes.c_ulonglong * (G * 512))()
        L.wmar_gpt_debug_pstep_trace.restype = ctypes.c_int
        L.wmar_gpt_debug_pstep_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        n = L.wmar_gpt_debug_pstep_trace(eng.handle, buf, G * 512)
        if n > 0:
            tr = np.frombuffer(buf, dtype=np.uint64).reshape(G, 512)[:n].astype(np.int64)
            t0s = tr[:, 0].min()
            ev = tr[:, 1:1 + 48 * 5].reshape(n, 48, 5) - t0s
            end = tr[:, 241] - t0s
            names = ["qkv+att", "proj", "fc1", "fc2", "-"]
            print(f"[{mode}] last step: kernel span {end.max() / 1e3:.1f} us; start skew {(tr[:, 0].max() - t0s) / 1e3:.1f} us")
            prev = np.concatenate([np.zeros((n, 1), dtype=np.int64) + 0, ev[:, :-1, 3]], axis=1)   # end of previous layer
            for l in (0, 1, 24, 47):
                seg = []
                last = prev[:, l] if l > 0 else (tr[:, 0] - t0s)
                for k in range(4):
                    seg.append(f"{names[k]} {np.median(ev[:, l, k] - last) / 1e3:5.2f} (max {np.max(ev[:, l, k] - last) / 1e3:5.2f})")
                    last = ev[:, l, k]
                print(f"   layer {l:2d}: " + " | ".join(seg) + f" | layer {np.median(ev[:, l, 3] - (prev[:, l] if l else 0)) / 1e3:.1f} us")
            per_layer = np.diff(np.median(ev[:, :, 3], axis=0))
            print(f"   median layer period {np.median(per_layer) / 1e3:.2f} us (min {per_layer.min() / 1e3:.2f}, max {per_layer.max() / 1e3:.2f})")
            it = tr[:, 300:316].reshape(n, 4, 4) - t0s          # [cta][phase][x staged, loop done, reduced, (qkv: attention start)]
            L2 = c["n_layer"] // 2
            base = ev[:, L2 - 1, 3]
            for k, nm in enumerate(["qkv", "proj", "fc1", "fc2"]):
                st = prev_end = (base if k == 0 else ev[:, L2, k - 1])
                print(f"   layer {L2} {nm}: phase start -> X staged {np.median(it[:, k, 0] - st) / 1e3:5.2f} (min {np.min(it[:, k, 0] - st) / 1e3:5.2f} max {np.max(it[:, k, 0] - st) / 1e3:5.2f})"
                      f" | main loop {np.median(it[:, k, 1] - it[:, k, 0]) / 1e3:5.2f} (max {np.max(it[:, k, 1] - it[:, k, 0]) / 1e3:5.2f})"
                      f" | reduce {np.median(it[:, k, 2] - it[:, k, 1]) / 1e3:5.2f} | hand-off+rest {np.median(ev[:, L2, k] - it[:, k, 2]) / 1e3:5.2f} (max {np.max(ev[:, L2, k] - it[:, k, 2]) / 1e3:5.2f})")
            print(f"   layer {L2} attention (after last qkv item): {np.median(ev[:, L2, 0] - it[:, 0, 3]) / 1e3:5.2f} us (max {np.max(ev[:, L2, 0] - it[:, 0, 3]) / 1e3:5.2f})")
            np.save(os.path.join("gpurun_out", f"pstep_trace_{mode}.npy"), tr)
    del eng
    torch.cuda.empty_cache()
ms = list(out)
for m in ms[1:]:
    same = (out[ms[0]] == out[m]).float().mean().item()
    print(f"ids {ms[0]} vs {m}: {same * 100:.2f}% equal", flush=True)
