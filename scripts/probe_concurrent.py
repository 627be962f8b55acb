"""On-box probe: N independent Taming generations (B=16 each, separate engines sharing ONE set of weights, separate KV
caches / workspaces) enqueued on N CUDA streams, against the same N generations back to back on one stream.  The decode
step is a latency chain that leaves HBM ~60 % idle; does a second chain fill the gaps?"""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
sys.path.insert(0, "scripts")
from helpers import make_wm  # noqa: E402
from probe_step import weights  # noqa: E402
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.gpt_engine import TamingGPTEngine  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = 256
w = weights(16384, 256, 48, 24, 1536)
engs = [TamingGPTEngine(w, 48, 24) for _ in range(N)]
for e in engs[1:]:                      # share the first engine's weight tensors (same device pointers)
    pass
wm = make_wm("taming")
cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814, 937, 975] * 2)[:16].cuda()
streams = [torch.cuda.Stream() for _ in range(N)]
outs = [torch.empty((16, steps), dtype=torch.long, device="cuda") for _ in range(N)]
L = _lib.lib()
wmp = wm.c_params()


def enqueue(i, seed):
    sp = _lib.SampleParams(1.0, 250, 0.92, 0, seed)
    with torch.cuda.stream(streams[i]):
        _lib.check(L.wmar_gpt_sample(engs[i].handle, ctypes.byref(wmp), ctypes.byref(sp), _lib.ptr(cond), 16, steps, None,
                                     _lib.ptr(outs[i]), None, _lib.current_stream()))


torch.cuda.synchronize()
for i in range(N):
    enqueue(i, 1 + i)
torch.cuda.synchronize()
ref = [o.clone() for o in outs]
for mode in ("sequential", "concurrent", "sequential", "concurrent"):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode == "sequential":
        for i in range(N):
            sp = _lib.SampleParams(1.0, 250, 0.92, 0, 1 + i)
            _lib.check(L.wmar_gpt_sample(engs[i].handle, ctypes.byref(wmp), ctypes.byref(sp), _lib.ptr(cond), 16, steps, None,
                                         _lib.ptr(outs[i]), None, _lib.current_stream()))
    else:
        cur = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(cur)
        for i in range(N):
            enqueue(i, 1 + i)
        for s in streams:
            cur.wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    same = all(torch.equal(o, r) for o, r in zip(outs, ref))
    print(f"{mode}: {N} x 16 images in {ms:.1f} ms = {N * 16 / ms * 1e3:.2f} img/s; ids identical to the first run: {same}", flush=True)
_lib.check_device_flag()
