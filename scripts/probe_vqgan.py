"""On-box probe: Taming VQGAN decode / encode at 256x256, batch 16: time and accuracy of the tcgen05 conv path vs the
mma.sync path (WMAR_CONV=v0) on the same weights."""
import os
import sys

import torch

sys.path.insert(0, ".")


def run(tag):
    from wmar_b200.models.synthetic import TAMING_VQGAN_DDCONFIG, taming_vqgan_state
    from wmar_b200.models.vqgan_engine import VQGANEngine
    dd = dict(TAMING_VQGAN_DDCONFIG)
    st = taming_vqgan_state(dd, seed=1, device="cuda")
    ecfg = dict(family=0, ch=dd["ch"], ch_mult=tuple(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"], attn_resolution=16,
                resolution=dd["resolution"], z_channels=dd["z_channels"], embed_dim=dd["embed_dim"], n_embed=dd["n_embed"])
    eng = VQGANEngine(st, ecfg, max_batch=16)
    g = torch.Generator(device="cuda").manual_seed(0)
    codes = torch.randint(0, dd["n_embed"], (16, 256), device="cuda", generator=g)
    img = eng.decode(codes)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        img = eng.decode(codes)
    e1.record()
    torch.cuda.synchronize()
    ms_d = e0.elapsed_time(e1) / 3
    back = eng.encode(img)
    e0.record()
    for _ in range(3):
        back = eng.encode(img)
    e1.record()
    torch.cuda.synchronize()
    ms_e = e0.elapsed_time(e1) / 3
    print(f"{tag}: decode {ms_d:.1f} ms ({eng.flops(True) * 16 / ms_d / 1e9:.1f} TFLOP/s useful), encode {ms_e:.1f} ms "
          f"({eng.flops(False) * 16 / ms_e / 1e9:.1f} TFLOP/s useful)", flush=True)
    return img.cpu(), back.cpu()


if __name__ == "__main__":
    mode = sys.argv[1] if len(sys.argv) > 1 else "both"
    if mode == "both":
        import subprocess
        for m in ("tc", "v0"):
            env = dict(os.environ)
            if m == "v0":
                env["WMAR_CONV"] = "v0"
            subprocess.check_call([sys.executable, __file__, m], env=env)
        a = torch.load("gpurun_out/vq_tc.pt")
        b = torch.load("gpurun_out/vq_v0.pt")
        rms = (a["img"] - b["img"]).pow(2).mean().sqrt().item()
        same = (a["codes"] == b["codes"]).float().mean().item()
        print(f"tc vs v0: image RMS {rms:.3e}, max {float((a['img'] - b['img']).abs().max()):.3e}, re-encoded codes equal {same * 100:.2f}%")
    else:
        os.makedirs("gpurun_out", exist_ok=True)
        img, back = run(mode)
        torch.save({"img": img, "codes": back}, f"gpurun_out/vq_{mode}.pt")
