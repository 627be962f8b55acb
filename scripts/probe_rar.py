"""On-box probe: RAR-XL decode loop (8 images = 16 guided rows), time per pass; not the bench.

    python scripts/probe_rar.py [steps] [reps]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wmar_b200.models.rar_engine import RAR_SIZES, RAREngine  # noqa: E402
from wmar_b200.models.synthetic import rar_state  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
c = dict(codebook_size=1024, image_seq_len=256, condition_num_classes=1000)
c.update(RAR_SIZES["rar_xl"])
w = rar_state(c, seed=0, device="cuda")
eng = RAREngine(w, c["num_hidden_layers"], c["num_attention_heads"], max_batch=8)
cond = torch.tensor([1, 9, 232, 340, 568, 656, 703, 814])
ids = eng.sample(cond, steps, 4.0, 1.0, None, greedy=True)
torch.cuda.synchronize()
best = 1e9
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ids = eng.sample(cond, steps, 4.0, 1.0, None, greedy=True)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
by = eng.algorithmic_bytes(8, steps)
print(f"[rar_xl hoist={os.environ.get('WMAR_RAR_HOIST', '1')}] {steps} steps: {best:.1f} ms = {best * 1e3 / (steps + 1):.0f} us/pass, "
      f"{by / best / 1e6:.0f} GB/s algorithmic = {by / best / 1e6 / 6557.8:.3f} of 6557.8; ids checksum {int(ids.sum())}", flush=True)
