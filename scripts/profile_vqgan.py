"""Workload for the ncu launch list of the VQGAN stacks: one warm decode + encode, then one profiled pair (the profiler
is told to skip the first half through --launch-skip; see scripts/profile_r02_vqgan.sh).  Prints the number of launches
of one (decode, encode) pair so the skip count can be checked."""
import sys

import torch

sys.path.insert(0, ".")
from wmar_b200 import _lib  # noqa: E402
from wmar_b200.models.synthetic import TAMING_VQGAN_DDCONFIG, taming_vqgan_state  # noqa: E402
from wmar_b200.models.vqgan_engine import VQGANEngine  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
dd = dict(TAMING_VQGAN_DDCONFIG)
st = taming_vqgan_state(dd, seed=1, device="cuda")
ecfg = dict(family=0, ch=dd["ch"], ch_mult=tuple(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"], attn_resolution=16,
            resolution=dd["resolution"], z_channels=dd["z_channels"], embed_dim=dd["embed_dim"], n_embed=dd["n_embed"])
eng = VQGANEngine(st, ecfg, max_batch=16, precision=prec)
g = torch.Generator(device="cuda").manual_seed(0)
codes = torch.randint(0, dd["n_embed"], (16, 256), device="cuda", generator=g)
torch.cuda.synchronize()
print("MARK start", flush=True)
for _ in range(2):
    img = eng.decode(codes)
    torch.cuda.synchronize()
    print("MARK decoded", flush=True)
    back = eng.encode(img)
    torch.cuda.synchronize()
    print("MARK encoded", flush=True)
