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
