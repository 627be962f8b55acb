#!/usr/bin/env python
"""bench.py -- watermarked 256x256 images/sec end to end (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W                      # this framework (CUDA, C-ABI)
    python bench.py --impl reference --gpus N --steps K --warmup W     # CPU restatement of the reference path

A step = one pass of the hot path over one batch per GPU:
    sample 256 tokens (Taming cin_transformer shapes, greenlist watermark delta=2 gamma=.25, top-k 250, top-p .92)
    -> codes_to_images (VQGAN decode) -> detector (n_green, T, z, p) on the produced codes.
Workload = BASELINE.json configs[1]: batch 16 per GPU, V=16384, L=48, H=24, d=1536, fp32, seeded random-init weights
(no checkpoints offline).  Weak scaling: every rank runs its own batch of independent images, no data-path collective
(SURVEY.md 8e); NCCL only broadcasts the weights before and gathers the counters after the timed region.

`value`  : images/s with the conditioning ids already on the device.
`e2e`    : the same metric through the public wrapper call with HOST buffers: conditioning H2D from pinned memory and
           images + codes + detector statistics D2H inside the timed region every step.
`roofline`: the decode loop (the dominant kernels: the per-token weight-streaming GEMMs + KV attention) timed with
           CUDA events on the launching stream, algorithmic bytes per SURVEY.md 8(d) / DESIGN.md.
`cpu_baseline`: the oracle port of the reference path on this box's host cores, on a bounded sample (stated).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "watermarked 256x256 images/sec end-to-end (sample -> decode -> detect), Taming cin_transformer, batch 16/GPU"
UNIT = "images/s"
WM_STRING = "linear-stratifiedrand-h=1-d=2.0-g=0.25"
CLASSES = [1, 9, 232, 340, 568, 656, 703, 814, 937, 975]
GEN_PARAMS = {"temperature": 1.0, "top_k": 250, "top_p": 0.92}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=16, help="images per GPU per step")
    ap.add_argument("--vqgan-precision", choices=["3xtf32", "tf32"], default="3xtf32")
    ap.add_argument("--rng", choices=["torch", "philox"], default="torch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--small", action="store_true", help="tiny shapes (plumbing check only; number is INVALID)")
    return ap.parse_args()


def shapes(small):
    from wmar_b200.models.synthetic import TAMING_GPT_CFG, TAMING_VQGAN_DDCONFIG
    if not small:
        return dict(TAMING_GPT_CFG), dict(TAMING_VQGAN_DDCONFIG)
    return (dict(vocab_size=16384, block_size=256, n_layer=2, n_head=4, n_embd=256),
            dict(TAMING_VQGAN_DDCONFIG))


# ----------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ----------------------------------------------------------------------------------------------------- CPU arm
def cpu_reference_throughput(gpt_state, vq_state, gpt_cfg, B, budget_s, max_decode_steps=24):
    """images/s of the oracle port (oracle/{gpt,sampling,wm,vqgan}.py: a restatement of sample_with_past +
    GentimeWatermark + VQModel.decode + detect) on the host cores, from a bounded sample extrapolated linearly:
    n of the 256 decode steps at batch B, m of the B image decodes, the detector on all B rows."""
    import numpy as np
    import torch
    from oracle import gpt as ogpt
    from oracle import sampling, vqgan as ov, wm as owm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    V = gpt_cfg["vocab_size"]
    assets = os.path.join(ROOT, "wmar_b200", "assets", "vqgan_alive_ids.txt")
    alive, dead = owm.alive_dead(owm.load_ids(assets), V)
    rows = owm.GreenRows(V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    o = ogpt.GPTOracle(gpt_state, gpt_cfg["n_layer"], gpt_cfg["n_head"])
    cond = torch.tensor([CLASSES[i % len(CLASSES)] for i in range(B)], dtype=torch.long)
    gen = torch.Generator().manual_seed(1)
    x = cond.clone()
    seq = cond.view(-1, 1).clone()
    t_dec = 0.0
    n_dec = 0
    with torch.no_grad():
        while n_dec < max_decode_steps and (n_dec < 3 or t_dec < budget_s * 0.6):
            t0 = time.perf_counter()
            logits = o.step(x, n_dec)
            noise = torch.empty(B, V).exponential_(1, generator=gen)
            x = sampling.sample_step(logits, rows(seq), 2.0, GEN_PARAMS["temperature"], GEN_PARAMS["top_k"],
                                     GEN_PARAMS["top_p"], noise)
            seq = torch.cat((seq, x.view(-1, 1)), dim=1)
            t_dec += time.perf_counter() - t0
            n_dec += 1
        codes = torch.randint(0, V, (B, 256), generator=gen)
        n_img = 0
        t_img = 0.0
        while n_img < B and (n_img < 1 or t_img < budget_s * 0.3):
            t0 = time.perf_counter()
            ov.taming_codes_to_images(codes[n_img:n_img + 1], vq_state)
            t_img += time.perf_counter() - t0
            n_img += 1
        t0 = time.perf_counter()
        ng, ns = owm.detect_counts(codes.numpy(), V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
        owm.pvalue(ng, ns, 0.25)
        t_det = time.perf_counter() - t0
    per_batch = t_dec / n_dec * 256 + t_img / n_img * B + t_det
    return {"value": B / per_batch, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_dec} of 256 decode steps at batch {B} ({t_dec / n_dec * 1e3:.0f} ms/step) + {n_img} of {B} "
                      f"VQGAN decodes ({t_img / n_img * 1e3:.0f} ms/img) + detect on {B} rows ({t_det * 1e3:.0f} ms), "
                      "extrapolated linearly to a full batch; torch fp32 CPU, all host threads"}


def host_states(small, seed=0):
    import torch
    from wmar_b200.models.synthetic import gpt_state, taming_vqgan_state
    gpt_cfg, dd = shapes(small)
    torch.set_num_threads(os.cpu_count() or 1)
    return gpt_state(gpt_cfg, seed, "cpu"), taming_vqgan_state(dd, seed + 1, "cpu"), gpt_cfg


def run_reference(args, rank):
    """The reference arm: the reference is pure Python on torch and cannot travel to the GPU box (no /root/reference
    there), so this times its CPU restatement (oracle/, pinned to the imported reference by tests/golden) on rank 0."""
    if rank != 0:
        return
    gs, vs, gpt_cfg = host_states(args.small)
    vals = []
    per = max(4.0, min(args.cpu_budget_s, 120.0 / max(1, args.steps + args.warmup)))
    res = None
    for i in range(args.warmup + args.steps):
        res = cpu_reference_throughput(gs, vs, gpt_cfg, args.batch, per, max_decode_steps=8)
        if i >= args.warmup:
            vals.append(res["value"])
    v = sum(vals) / len(vals) if vals else res["value"]
    res["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": args.batch / v * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "taming_cin_B16_wm_linear_h1_d2_g0.25", "batch_per_gpu": args.batch,
                       "note": "CPU path runs on rank 0 only; one step = bounded sample of one batch"},
            "cpu_baseline": res,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- ours
def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from wmar_b200 import _lib
    from wmar_b200.distributed import broadcast_state
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.models.synthetic import taming_net2net_state
    from wmar_b200.watermarking import create_watermarker_from_string

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; wmar_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()  # raises if the CUDA library is missing
    gpt_cfg, dd = shapes(args.small)
    B = args.batch

    # weights: rank 0 draws them, NCCL broadcast to the replicas (models are replicated, data is sharded)
    state = taming_net2net_state(gpt_cfg, dd, seed=0, device=dev)
    broadcast_state(state, src=0)
    model = TamingARMMWrapper(state_dict=state, gpt_cfg=gpt_cfg, dd_cfg=dd, device=dev, max_batch=B,
                              vqgan_precision=args.vqgan_precision, rng=args.rng)
    wm = create_watermarker_from_string(model.get_vq(), model.get_total_vocab_size(), WM_STRING, dev)
    model.set_watermarker(wm)
    # reference seeding: args.seed + 1000 * chunk_id (generate.py:304)
    torch.manual_seed(1 + 1000 * rank)
    torch.cuda.manual_seed_all(1 + 1000 * rank)

    steps_tok = model.codes_size ** 2
    cond_list = [CLASSES[(rank * B + i) % len(CLASSES)] for i in range(B)]
    cond_dev = torch.tensor(cond_list, dtype=torch.long, device=dev)
    cond_pin = torch.tensor(cond_list, dtype=torch.long).pin_memory()
    img_pin = torch.empty((B, 3, model.image_size, model.image_size), dtype=torch.float32).pin_memory()
    codes_pin = torch.empty((B, steps_tok), dtype=torch.long).pin_memory()
    stat_pin = torch.empty((B, 4), dtype=torch.float64).pin_memory()

    def hot_path(cond):
        codes = model.sample(cond, GEN_PARAMS, apply_watermark=True)
        ev_s.record()
        imgs = model.codes_to_images(codes)
        ev_d.record()
        st = wm.detect_stats(codes)
        return codes, imgs, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev_s, ev_d = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        hot_path(cond_dev)
    barrier()

    # ---- timed region 1: device-resident inputs ----
    clocks = ClockSampler(local_rank)
    clocks.start()
    launches0 = L.wmar_launch_count()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    evd = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_end = torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.perf_counter()
    last = None
    for i in range(args.steps):
        ev_s, ev_d = evs[i], evd[i]
        ev0[i].record()
        last = hot_path(cond_dev)
    ev_end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.wmar_launch_count() - launches0
    t_total_ms = ev0[0].elapsed_time(ev_end)
    t_sample_ms = sum(ev0[i].elapsed_time(evs[i]) for i in range(args.steps))
    t_decode_ms = sum(evs[i].elapsed_time(evd[i]) for i in range(args.steps))
    _lib.check(L.wmar_check_device_flag(_lib.current_stream()))

    # ---- timed region 2: end to end through the wrapper with host buffers ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        c = cond_pin.to(dev, non_blocking=True)
        codes, imgs, st = hot_path(c)
        img_pin.copy_(imgs, non_blocking=True)
        codes_pin.copy_(codes, non_blocking=True)
        stat_pin[:, 0].copy_(st["n_green"], non_blocking=True)
        stat_pin[:, 1].copy_(st["n_scored"], non_blocking=True)
        stat_pin[:, 2].copy_(st["z"], non_blocking=True)
        stat_pin[:, 3].copy_(st["pvalue"], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the result of every step
    e1.record()
    barrier()
    t_e2e_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    h2d = cond_pin.numel() * 8
    d2h = img_pin.numel() * 4 + codes_pin.numel() * 8 + stat_pin.numel() * 8

    # max over ranks
    tt = torch.tensor([t_total_ms, t_e2e_ms, t_sample_ms, t_decode_ms, t_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total_ms, t_e2e_ms, t_sample_ms, t_decode_ms, t_wall_ms = tt.tolist()
    n_img = world * B * args.steps
    value = n_img / (t_total_ms * 1e-3)
    e2e = n_img / (t_e2e_ms * 1e-3)

    # roofline of the decode loop (HBM bound): algorithmic bytes (weights once per token + KV) / event time
    peaks = {}
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        try:
            peaks = json.load(open(ppath))
        except Exception:
            peaks = {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = model._gpt.algorithmic_bytes(B, steps_tok)
    achieved = alg_bytes * args.steps / (t_sample_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("decode_loop_dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "decode loop (256 token steps: skinny GEMMs + KV attention + fused sampler)",
                "achieved": achieved, "peak": peak, "peak_source": "measured" if peaks else "fallback", "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": t_sample_ms / args.steps,
                "phase_ms_per_step": {"sample": t_sample_ms / args.steps, "vqgan_decode": t_decode_ms / args.steps,
                                      "detect+rest": (t_total_ms - t_sample_ms - t_decode_ms) / args.steps}}

    st = last[2]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": t_total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (3xTF32 tensor-core products, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "taming_cin_B16_wm_linear_h1_d2_g0.25" + ("_SMALL_INVALID" if args.small else ""),
                       "batch_per_gpu": B, "global_batch": B * world, "tokens_per_image": steps_tok,
                       "gpt": gpt_cfg, "watermark": WM_STRING, "gen_params": GEN_PARAMS,
                       "vqgan_precision": args.vqgan_precision, "rng": args.rng, "parallelism": f"replicas x{world}",
                       "l2": "inputs larger than L2 (5.5 GB of weights streamed per token step)"},
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": t_e2e_ms / args.steps},
            "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
            "detector": {"n_green_mean": float(st["n_green"].float().mean()), "z_mean": float(st["z"].mean()),
                         "log10_p_max": float(torch.log10(st["pvalue"].clamp_min(1e-300)).max())},
            "wall_ms_per_step": t_wall_ms / args.steps}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gs = {k[len("transformer."):]: v.cpu() for k, v in state.items() if k.startswith("transformer.")}
        vs = {k[len("first_stage_model."):]: v.cpu() for k, v in state.items() if k.startswith("first_stage_model.")}
        line["cpu_baseline"] = cpu_reference_throughput(gs, vs, gpt_cfg, B, args.cpu_budget_s)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
