"""Quick on-box probe: skinny-GEMM bandwidth per shape and full-size Taming sampling time (not the bench)."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from wmar_b200 import _lib  # noqa: E402


def time_gemm(N, K, split, iters=50):
    # rotate over enough weight copies that nothing is served from the 126 MB L2
    copies = max(2, int(400e6 // (N * K * 4)) + 1)
    ws = [(torch.randn(N, K, device="cuda") * 0.02) for _ in range(copies)]
    x = torch.randn(16, K, device="cuda")
    b = torch.randn(N, device="cuda")
    y = torch.empty(16, N, device="cuda")
    L = _lib.lib()
    for i in range(5):
        _lib.check(L.wmar_skinny_gemm(_lib.ptr(x), _lib.ptr(ws[i % copies]), _lib.ptr(b), _lib.ptr(y), N, K, split,
                                      _lib.current_stream()))
    torch.cuda.synchronize()
    # capture the launches in a CUDA graph so that the host launch rate (ctypes) is not what is measured
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            L.wmar_skinny_gemm(_lib.ptr(x), _lib.ptr(ws[i % copies]), _lib.ptr(b), _lib.ptr(y), N, K, split,
                               _lib.current_stream())
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, N * K * 4 / us / 1e3  # GB/s


def main():
    out = {"gemm": []}
    import ctypes
    dbg = _lib.lib().wmar_debug_set_gemm_mode
    dbg.argtypes = [ctypes.c_int]
    for mode in (0, 1, 2):
        dbg(mode)
        for (N, K) in [(4608, 1536), (1536, 1536), (6144, 1536), (1536, 6144), (16384, 1536)]:
            for split in (0, 1):
                us, gbs = time_gemm(N, K, split)
                out["gemm"].append({"mode": mode, "N": N, "K": K, "split": split, "us": round(us, 2), "GBps": round(gbs, 1)})
                print(out["gemm"][-1], flush=True)
    dbg(0)
    # full-size Taming engine with synthetic weights generated on the device
    from wmar_b200.models.gpt_engine import TamingGPTEngine
    from helpers import make_wm
    V, block, L, H, d = 16384, 256, 48, 24, 1536
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, std=0.02: torch.randn(*s, device="cuda", generator=g) * std
    w = {"tok_emb.weight": rn(V, d), "pos_emb": rn(1, block, d), "ln_f.weight": torch.ones(d, device="cuda"),
         "ln_f.bias": torch.zeros(d, device="cuda"), "head.weight": rn(V, d)}
    for i in range(L):
        p = f"blocks.{i}."
        for ln in ("ln1", "ln2"):
            w[p + ln + ".weight"] = torch.ones(d, device="cuda")
            w[p + ln + ".bias"] = torch.zeros(d, device="cuda")
        for nm in ("key", "query", "value", "proj"):
            w[p + f"attn.{nm}.weight"] = rn(d, d)
            w[p + f"attn.{nm}.bias"] = rn(d, std=0.01)
        w[p + "mlp.0.weight"] = rn(4 * d, d)
        w[p + "mlp.0.bias"] = rn(4 * d, std=0.01)
        w[p + "mlp.2.weight"] = rn(d, 4 * d)
        w[p + "mlp.2.bias"] = rn(d, std=0.01)
    eng = TamingGPTEngine(w, L, H)
    del w
    wm = make_wm("taming")
    cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:16]
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        codes = eng.sample(cond, 256, 1.0, 250, 0.92, wm, seed=rep)
        torch.cuda.synchronize()
        dt = time.time() - t0
        by = eng.algorithmic_bytes(16, 256)
        print(f"taming B=16 x256: {dt*1e3:.1f} ms  {16/dt:.2f} img/s  {by/dt/1e9:.0f} GB/s algorithmic", flush=True)
        out.setdefault("taming", []).append({"ms": dt * 1e3, "img_s": 16 / dt, "GBps": by / dt / 1e9})
    st = wm.detect_stats(codes)
    print("n_green", st["n_green"].tolist(), "z", [round(v, 2) for v in st["z"].tolist()])
    json.dump(out, open("gpurun_out/probe_gpt.json", "w"), indent=1)


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    main()
