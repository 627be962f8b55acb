"""On-box probe: Taming VQGAN decode / encode at 256x256, batch 16, per conv precision mode (3xtf32 / bf16x3 / tf32):
time, image RMS against the 3xTF32 decode of the same codes, and re-encoded codes against the 3xTF32 encoder's."""
import sys

import torch

sys.path.insert(0, ".")


def main():
    from wmar_b200.models.synthetic import TAMING_VQGAN_DDCONFIG, taming_vqgan_state
    from wmar_b200.models.vqgan_engine import VQGANEngine
    dd = dict(TAMING_VQGAN_DDCONFIG)
    st = taming_vqgan_state(dd, seed=1, device="cuda")
    ecfg = dict(family=0, ch=dd["ch"], ch_mult=tuple(dd["ch_mult"]), num_res_blocks=dd["num_res_blocks"], attn_resolution=16,
                resolution=dd["resolution"], z_channels=dd["z_channels"], embed_dim=dd["embed_dim"], n_embed=dd["n_embed"])
    g = torch.Generator(device="cuda").manual_seed(0)
    codes = torch.randint(0, dd["n_embed"], (16, 256), device="cuda", generator=g)
    ref_img = ref_codes = None
    for mode in sys.argv[1:] or ["3xtf32", "bf16x3", "tf32"]:
        eng = VQGANEngine(st, ecfg, max_batch=16, precision=mode)
        img = eng.decode(codes)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            img = eng.decode(codes)
        e1.record()
        torch.cuda.synchronize()
        ms_d = e0.elapsed_time(e1) / 5
        src = ref_img if ref_img is not None else img
        back = eng.encode(src)
        e0.record()
        for _ in range(5):
            back = eng.encode(src)
        e1.record()
        torch.cuda.synchronize()
        ms_e = e0.elapsed_time(e1) / 5
        line = (f"{mode}: decode {ms_d:.2f} ms ({eng.flops(True) * 16 / ms_d / 1e9:.1f} TFLOP/s useful), encode {ms_e:.2f} ms "
                f"({eng.flops(False) * 16 / ms_e / 1e9:.1f} TFLOP/s useful)")
        if ref_img is None:
            ref_img, ref_codes = img.clone(), back.clone()
            line += f"; round trip codes equal {float((back == codes).float().mean()) * 100:.2f}%"
        else:
            d = (img - ref_img)
            line += (f"; image RMS vs 3xtf32 {d.pow(2).mean().sqrt().item():.3e} max {d.abs().max().item():.3e}; "
                     f"codes of the same images equal to the 3xtf32 encoder's {float((back == ref_codes).float().mean()) * 100:.3f}% "
                     f"({int((back != ref_codes).sum())} of {back.numel()} differ)")
        print(line, flush=True)
        del eng


if __name__ == "__main__":
    main()
