"""On-box probe: Anole-7B shapes (random-init bf16), B images (3B guided rows), time per pass + HBM roofline fraction."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from wmar_b200.models.cham_engine import ChameleonEngine  # noqa: E402
from wmar_b200.models.synthetic import ANOLE_7B_CFG, chameleon_state  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    c = ANOLE_7B_CFG
    w = chameleon_state(c, seed=0, device="cuda")
    eng = ChameleonEngine(w, c["n_layers"], c["n_heads"], c["n_kv_heads"], max_seq=1100, max_batch=8)
    del w
    torch.manual_seed(0)
    full = [[0] + torch.randint(16384, 65536, (12 + 3 * b,)).tolist() + [8710, 8197] for b in range(B)]
    rows = full + [[0, 8197]] * B + [[0, 8197]] * B
    p_max = max(len(r) for r in rows)
    out = {}
    for rep in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ids = eng.sample(rows, steps, 3.0, 1.2, temperature=0.9, top_p=0.9, seed=rep + 1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        passes = p_max + steps - 1
        by = eng.algorithmic_bytes(B, p_max, steps)
        print(f"anole-7b B={B} steps={steps}: {ms:.1f} ms  {ms / passes * 1e3:.0f} us/pass  {by / ms / 1e6:.0f} GB/s algorithmic "
              f"-> {B / (ms / passes * (p_max + 1023)) * 1e3:.3f} img/s at 1024 tokens", flush=True)
        out[f"rep{rep}"] = {"ms": ms, "us_per_pass": ms / passes * 1e3, "GBps": by / ms / 1e6}
    assert int(ids.min()) >= 4 and int(ids.max()) < 8196
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/probe_cham.json", "w"), indent=1)


if __name__ == "__main__":
    main()
