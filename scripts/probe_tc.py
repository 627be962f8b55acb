"""On-box probe of the tcgen05 skinny GEMM: correctness vs fp64 per shape, then per-shape bandwidth of both engines
(tcgen05 vs mma.sync) replayed from a CUDA graph, with and without programmatic dependent launch.  Not the bench."""
import ctypes
import json
import sys

import torch

sys.path.insert(0, ".")
from wmar_b200 import _lib  # noqa: E402

L = _lib.lib()
for name in ("wmar_debug_set_gemm_engine", "wmar_debug_set_pdl", "wmar_debug_set_gemm_mode", "wmar_debug_set_tc_dbg"):
    getattr(L, name).argtypes = [ctypes.c_int]
    getattr(L, name).restype = None

SHAPES = [(256, 64), (128, 256), (1536, 1536), (4608, 1536), (6144, 1536), (1536, 6144), (16384, 1536), (1024, 1280),
          (3840, 1280), (7680, 1280)]


def check(N, K):
    g = torch.Generator().manual_seed(N + K)
    x = torch.randn(16, K, generator=g).cuda()
    w = (torch.randn(N, K, generator=g) * 0.02).cuda()
    b = torch.randn(N, generator=g).cuda()
    ref64 = x.double() @ w.double().t() + b.double()
    ref32 = torch.nn.functional.linear(x, w, b)
    out = {}
    for eng in (0, 1):
        L.wmar_debug_set_gemm_engine(eng)
        y = torch.full((16, N), float("nan"), device="cuda")
        _lib.check(L.wmar_skinny_gemm(_lib.ptr(x), _lib.ptr(w), _lib.ptr(b), _lib.ptr(y), N, K, 0, _lib.current_stream()))
        torch.cuda.synchronize()
        out["tc" if eng == 0 else "v0"] = (y.double() - ref64).abs().max().item()
    out["fp32"] = (ref32.double() - ref64).abs().max().item()
    out["scale"] = ref64.abs().max().item()
    L.wmar_debug_set_gemm_engine(0)
    return out


def time_gemm(N, K, iters=60):
    copies = max(2, int(600e6 // (N * K * 4)) + 1)
    ws = [(torch.randn(N, K, device="cuda") * 0.02) for _ in range(copies)]
    x = torch.randn(16, K, device="cuda")
    b = torch.randn(N, device="cuda")
    y = torch.empty(16, N, device="cuda")
    for i in range(copies):
        _lib.check(L.wmar_skinny_gemm(_lib.ptr(x), _lib.ptr(ws[i]), _lib.ptr(b), _lib.ptr(y), N, K, 0, _lib.current_stream()))
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            _lib.check(L.wmar_skinny_gemm(_lib.ptr(x), _lib.ptr(ws[i % copies]), _lib.ptr(b), _lib.ptr(y), N, K, 0,
                                          _lib.current_stream()))
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    return us, N * K * 4 / us / 1e3


def main():
    res = {"check": [], "time": []}
    for (N, K) in SHAPES:
        r = check(N, K)
        r.update(N=N, K=K)
        res["check"].append(r)
        print(r, flush=True)
    if "--dbg" in sys.argv:
        L.wmar_debug_set_pdl(0)
        L.wmar_debug_set_gemm_engine(0)
        for bits in (0, 1, 2, 3, 7):
            L.wmar_debug_set_tc_dbg(bits)
            for (N, K) in [(6144, 1536), (16384, 1536), (16384, 6144)]:
                us, gbs = time_gemm(N, K)
                r = {"dbg": bits, "N": N, "K": K, "us": round(us, 2), "GBps": round(gbs, 1)}
                res["time"].append(r)
                print(r, flush=True)
        L.wmar_debug_set_tc_dbg(0)
    elif "--check-only" not in sys.argv:
        for pdl in (1, 0):
            L.wmar_debug_set_pdl(pdl)
            for eng in (0, 1):
                L.wmar_debug_set_gemm_engine(eng)
                for (N, K) in [(4608, 1536), (1536, 1536), (6144, 1536), (1536, 6144), (16384, 1536)]:
                    try:
                        us, gbs = time_gemm(N, K)
                    except Exception as e:  # graph capture of PDL launches may be refused
                        print("time failed", N, K, eng, pdl, repr(e), flush=True)
                        continue
                    r = {"engine": "tc" if eng == 0 else "v0", "pdl": pdl, "N": N, "K": K, "us": round(us, 2),
                         "GBps": round(gbs, 1)}
                    res["time"].append(r)
                    print(r, flush=True)
        L.wmar_debug_set_gemm_engine(0)
        L.wmar_debug_set_pdl(1)
    with open("gpurun_out/probe_tc.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
