import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
ASSETS = os.path.join(ROOT, "wmar_b200", "assets")
G = os.path.join(HERE, "golden")


def assets(name):
    from oracle import wm
    if name == "taming":
        return wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "vqgan_alive_ids.txt")), 16384) + (16384,)
    if name == "rar":
        return wm.alive_dead(wm.load_ids(os.path.join(ASSETS, "rar_all_ids.txt")), 1024) + (1024,)
    if name == "chameleon":
        alive = list(range(4, 8196)) + list(range(16384, 65536))
        return wm.alive_dead(alive, 8192) + (65536,)
    raise KeyError(name)


def make_wm(name, seed_strategy="linear", split="stratifiedrand", h=1, delta=2.0, gamma=0.25, build_on="device"):
    import torch
    from wmar_b200.watermarking import GentimeWatermark, SeedStrategy, SplitStrategy
    alive, dead, V = assets(name)
    vq = {"alive_ids": torch.from_numpy(alive), "dead_ids": torch.from_numpy(dead)}
    return GentimeWatermark(vq, V, SeedStrategy(seed_strategy), SplitStrategy(split), h, delta, gamma, "cuda",
                            build_on=build_on)


def unpack_rows(table_i32, V):
    """int32 [rows, V/32] (numpy) -> bool [rows, V]"""
    u8 = np.ascontiguousarray(table_i32).view(np.uint8)
    return np.unpackbits(u8, axis=-1, bitorder="little")[..., :V].astype(bool)


class ToyTokenizerModel:
    """Device-agnostic stand-in with the wrappers' codes_to_images / images_to_codes surface, used on both sides of the
    evaluation-log golden (oracle/gen_golden_evallog.py runs the REFERENCE's fill_batch_log + compute_metric over it on
    the CPU; tests/test_gpu_augment.py runs wmar_b200.evaluate over it on the GPU).  64 codes, 4 x 4 latent grid, every
    code paints an 8 x 8 block with its RGB colour; encoding = nearest colour of the block mean (first index on ties)."""

    def __init__(self, device="cpu"):
        import torch
        g = torch.Generator().manual_seed(123)
        self.table = (torch.rand(64, 3, generator=g) * 2 - 1).to(device)
        self.device = torch.device(device)

    def codes_to_images(self, codes):
        B = codes.shape[0]
        col = self.table[codes.to(self.device)].view(B, 4, 4, 3).permute(0, 3, 1, 2)
        return col.repeat_interleave(8, dim=2).repeat_interleave(8, dim=3).contiguous()

    def images_to_codes(self, imgs):
        B = imgs.shape[0]
        m = imgs.to(self.device).float().view(B, 3, 4, 8, 4, 8).mean(dim=(3, 5))          # [B, 3, 4, 4]
        d = (m.permute(0, 2, 3, 1).reshape(B, 16, 1, 3) - self.table.view(1, 1, 64, 3)).pow(2).sum(-1)
        return d.argmin(dim=-1)
