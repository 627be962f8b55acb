"""Golden vector for the CLUSTERING greenlist split, produced by the reference's own GentimeWatermark
(wmar/watermarking/gentime_watermark.py:175-216) imported from /root/reference.  TEST INFRASTRUCTURE ONLY.

    python oracle/gen_golden_clustering.py        -> tests/golden/clustering.npz
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
sys.path.insert(0, REF)
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def inputs(seed=0, vocab=640, dim=8, n_dead=37):
    g = torch.Generator().manual_seed(seed)
    emb = torch.randn(vocab, dim, generator=g)
    emb[: vocab // 2] += 3.0                                # two blobs, so that t-SNE has structure to find
    dead = torch.randperm(vocab, generator=g)[:n_dead].sort().values
    alive = torch.tensor(sorted(set(range(vocab)) - set(dead.tolist())), dtype=torch.long)
    alive = alive[torch.randperm(len(alive), generator=g)]  # file order is not sorted in general
    return emb, alive, dead


def main():
    from wmar.watermarking.gentime_watermark import GentimeWatermark, SeedStrategy, SplitStrategy
    emb, alive, dead = inputs()
    vq = {"alive_ids": alive, "dead_ids": dead, "embedding": emb}
    wm = GentimeWatermark(vq, emb.shape[0], SeedStrategy.FIXED, SplitStrategy.CLUSTERING, 0, 2.0, 0.5, device="cpu")
    green = wm.fixed_greenlist.numpy().astype(np.int64)
    np.savez_compressed(os.path.join(OUT, "clustering.npz"), green=green, alive=alive.numpy(), dead=dead.numpy(),
                        emb=emb.numpy())
    print("clustering.npz: green", len(green), "of", emb.shape[0])


if __name__ == "__main__":
    main()
