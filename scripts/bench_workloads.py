"""On-box measurements of the BASELINE.json configs other than the one bench.py reports (configs[2], [3], [4]):

    python scripts/bench_workloads.py rar        # RAR-XL 256x256, greenlist watermark, 8 images / GPU (16 guided rows)
    python scripts/bench_workloads.py detect     # detection only: VQGAN encode + z-score on synthetic 256x256 images
    python scripts/bench_workloads.py anole      # Anole-7B text -> image 512x512, watermark on, 8 images / GPU (16 guided rows)

Each prints one JSON line (images/s per GPU, phase times from CUDA events, roofline of the dominant phase against
MEASURED_PEAKS.json).  Synthetic data, seeded random-init weights at the reference's shapes.  Single GPU: the path shards
by independent images with no exchange step, so N GPUs = N independent replicas of this.
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
WM_STRING = "linear-stratifiedrand-h=1-d=2.0-g=0.25"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def timed(fn, reps=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def rar():
    from wmar_b200.models import RarARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    B = 8
    m = RarARMMWrapper(rar_size="rar_xl", max_batch=B)
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), WM_STRING, m.device)
    m.set_watermarker(wm)
    torch.manual_seed(1)
    cond = [1, 9, 232, 340, 568, 656, 703, 814]
    ms_s, codes = timed(lambda: m.sample(cond, None, apply_watermark=True))
    ms_d, imgs = timed(lambda: m.codes_to_images(codes))
    ms_t, st = timed(lambda: wm.detect_stats(codes))
    by = m._rar.algorithmic_bytes(B, 256)
    pk = float(peaks().get("hbm_gbs", 6650.0))
    return {"workload": "rar_xl_256_B8_cfg4_wm_linear_h1_d2_g0.25", "images_per_s_per_gpu": B / ((ms_s + ms_d + ms_t) * 1e-3),
            "phase_ms": {"sample": ms_s, "vqgan_decode": ms_d, "detect": ms_t},
            "roofline": {"bound": "hbm", "kernel": "RAR decode loop (257 passes)", "achieved": by / ms_s / 1e6, "peak": pk,
                         "unit": "GB/s", "frac": by / ms_s / 1e6 / pk, "algorithmic_bytes": by},
            "detector": {"n_green_mean": float(st["n_green"].float().mean()), "z_mean": float(st["z"].mean())}}


def detect():
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    B, n_batches = 16, 8
    m = TamingARMMWrapper(gpt_cfg=dict(vocab_size=16384, block_size=256, n_layer=1, n_head=24, n_embd=1536), max_batch=B)
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), WM_STRING, m.device)
    g = torch.Generator(device="cuda").manual_seed(0)
    imgs = [torch.rand(B, 3, 256, 256, device="cuda", generator=g) * 2 - 1 for _ in range(n_batches)]

    def run():
        out = []
        for x in imgs:
            codes = m.images_to_codes(x)
            out.append(wm.detect_stats(codes))
        return out

    ms, sts = timed(run, reps=2)
    fl = m._vqgan.flops(decode=False) * B * n_batches
    pk = float(peaks().get("bf16_tflops_sustained", 1363.5)) / 2.0    # TF32 dense = half the bf16 rate
    return {"workload": "detect_only_taming_encode_256_B16", "images_per_s_per_gpu": B * n_batches / (ms * 1e-3),
            "ms_per_batch_of_16": ms / n_batches,
            "roofline": {"bound": "tensor", "kernel": "VQGAN encoder conv stack (tcgen05 3xTF32 implicit GEMM) + codebook arg-min",
                         "achieved": fl / ms / 1e9, "peak": pk, "unit": "TFLOP/s (useful fp32-equivalent)", "frac": fl / ms / 1e9 / pk},
            "detector": {"p_mean": float(torch.cat([s["pvalue"] for s in sts]).mean())}}


def anole():
    from wmar_b200.models.chameleon_wrapper import ChameleonARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    B = 8    # text-only prompts: the image-conditioned rows equal the unconditioned ones -> 2 x 8 = 16 guided rows
    m = ChameleonARMMWrapper(max_batch=B)
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), "fixed-stratifiedrand-h=0-d=2.0-g=0.25", m.device)
    m.set_watermarker(wm)
    torch.manual_seed(1)
    prompts = ["a photo of a red bus parked next to a building on a sunny day", "two cats sleeping on a couch",
               "a plate of food with broccoli and rice on a wooden table", "a man riding a wave on a surfboard",
               "a kitchen with a stove a sink and a window", "a group of people flying kites in a park",
               "a close up of a pizza on a table", "a train traveling down tracks next to a forest"]
    cond = list(enumerate(prompts))
    ms_s, codes = timed(lambda: m.sample(cond, {"temperature": 0.9, "top_p": 0.9}, apply_watermark=True), reps=1, warmup=1)
    ms_d, imgs = timed(lambda: m.codes_to_images(codes), reps=2)
    ms_t, st = timed(lambda: wm.detect_stats(codes))
    p_max = max(len(r) for r in m.prompt_rows(prompts))
    by = m._eng.algorithmic_bytes(B, p_max, 1024)
    pk = float(peaks().get("hbm_gbs", 6650.0))
    return {"workload": "anole_7b_512_B8_cfg3.0_1.2_T0.9_p0.9_wm_fixed_h0_d2_g0.25", "images_per_s_per_gpu": B / ((ms_s + ms_d + ms_t) * 1e-3),
            "phase_ms": {"sample": ms_s, "vqgan_decode_512": ms_d, "detect": ms_t},
            "roofline": {"bound": "hbm", "kernel": "Anole-7B decode loop (prompt + 1023 passes, bf16 weights)",
                         "achieved": by / ms_s / 1e6, "peak": pk, "unit": "GB/s", "frac": by / ms_s / 1e6 / pk, "algorithmic_bytes": by},
            "detector": {"n_green_mean": float(st["n_green"].float().mean()), "z_mean": float(st["z"].mean())},
            "image_range": [float(imgs.min()), float(imgs.max())]}


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "rar"
    res = {"rar": rar, "detect": detect, "anole": anole}[which]()
    res["gpu"] = torch.cuda.get_device_name(0)
    print(json.dumps(res), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"workload_{which}.json"), "w") as f:
        json.dump(res, f, indent=1)
