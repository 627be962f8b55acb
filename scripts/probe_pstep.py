"""Probe of the persistent step kernel at BASELINE configs[1] shapes: decode-loop time per path, ids equality, trace.

    python scripts/probe_pstep.py [steps] [modes...]     (modes: pstep graph; WMAR_PSTEP_TRACE=1 prints the phase timeline of the last step)
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
    t0 = time.time()
    eng = TamingGPTEngine(w, c["n_layer"], c["n_head"])
    torch.cuda.synchronize()
    print(f"[{mode}] create {time.time() - t0:.2f} s", flush=True)
    try:
        ids = eng.sample(cond, steps, 1.0, 250, 0.92, None, greedy=True)
    except Exception as ex:
        if not os.environ.get("WMAR_PSTEP_DBG"):
            raise
        ids = torch.zeros(16, steps, dtype=torch.long)
    torch.cuda.synchronize()
    rc = _lib.lib().wmar_check_device_flag(_lib.current_stream())
    print(f"[{mode}] device flag rc={rc} {_lib.lib().wmar_last_error().decode() if rc else ''}", flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        try:
            ids = eng.sample(cond, steps, 1.0, 250, 0.92, None, greedy=True)
        except Exception as ex:            # knock-out runs (WMAR_PSTEP_DBG) produce garbage logits: time them anyway
            if not os.environ.get("WMAR_PSTEP_DBG"):
                raise
            print("   (ignored:", str(ex)[:80], ")")
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    by = eng.algorithmic_bytes(16, steps)
    print(f"[{mode}] {steps} steps: {best:.1f} ms = {best * 1e3 / steps:.0f} us/token, {by / best / 1e6:.0f} GB/s algorithmic "
          f"= {by / best / 1e6 / 6557.8:.3f} of 6557.8", flush=True)
    out[mode] = ids.cpu()
    if mode.startswith("pstep") and os.environ.get("WMAR_PSTEP_TRACE"):
        L = _lib.lib()
        G, EV, PER = 148, 1024, 20
        buf = (ctypes.c_ulonglong * (G * EV))()
        L.wmar_gpt_debug_pstep_trace.restype = ctypes.c_int
        L.wmar_gpt_debug_pstep_trace.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        n = L.wmar_gpt_debug_pstep_trace(eng.handle, buf, G * EV)
        if n > 0:
            tr = np.frombuffer(buf, dtype=np.uint64).reshape(G, EV)[:n].astype(np.int64)
            t0s = tr[:, 0].min()
            nl = c["n_layer"]
            raw = tr[:, 1:1 + nl * PER].reshape(n, nl, PER)
            ev = (raw - t0s).astype(np.float64) / 1e3
            ev[raw == 0] = np.nan
            end = (tr[:, 2 + 48 * PER] - t0s) / 1e3
            names = ["x flags", "x loaded", "qkv loop", "qkv stored", "qkv signal", "qkv flags", "attention", "att signal",
                     "y flags", "y loaded", "proj loop", "proj stored", "xb signal", "xb flags", "xb loaded", "fc1", "fc2",
                     "fc2 signal", "part flags", "reduced+signal"]
            print(f"[{mode}] last step: kernel span {np.nanmax(end):.1f} us; start skew {(tr[:, 0].max() - t0s) / 1e3:.1f} us; "
                  f"head starts at {np.nanmedian((tr[:, 1 + 48 * PER] - t0s) / 1e3):.1f} us")
            for l in (1, nl // 2):
                prev = ev[:, l - 1, PER - 1]
                last = prev
                print(f"   layer {l}: (median over CTAs of the time since the CTA's previous event; max; CTAs that ran it)")
                for k in range(PER):
                    cur = ev[:, l, k]
                    dt = cur - last
                    if np.isfinite(dt).any():
                        print(f"      {names[k]:>15s} +{np.nanmedian(dt):6.2f}  max {np.nanmax(dt):6.2f}  n={int(np.isfinite(dt).sum()):3d}   "
                              f"(abs median {np.nanmedian(cur - np.nanmin(prev)):6.2f})")
                    last = np.where(np.isnan(cur), last, cur)
                print(f"      layer time {np.nanmedian(ev[:, l, PER - 1] - prev):.2f} us")
            per_layer = np.diff(np.nanmedian(ev[:, :, PER - 1], axis=0))
            print(f"   median layer period {np.median(per_layer):.2f} us (min {per_layer.min():.2f}, max {per_layer.max():.2f})")
            np.save(os.path.join("gpurun_out", f"pstep_trace_{mode}.npy"), tr)
    del eng
    torch.cuda.empty_cache()
ms = list(out)
for m in ms[1:]:
    same = (out[ms[0]] == out[m]).float().mean().item()
    print(f"ids {ms[0]} vs {m}: {same * 100:.2f}% equal", flush=True)
