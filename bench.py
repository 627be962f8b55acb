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
`cpu_baseline`: the oracle port of the reference path on this box's host cores, on a bounded sample (stated); its
           `gpu_eager` sub-object is the same torch port run eagerly on the GPU (what a wmar user experiences on this
           box; context only).
`extra_workloads`: BASELINE.json configs[2] (RAR-XL, 8 images per GPU = batch 64 over 8 GPUs) and configs[4] (detection
           only) measured in the same run as fully formed blocks (value, e2e, ms_per_step, roofline, gpu_launches), so
           that the driver's bare `bench.py --gpus N` records them at every N.  `--workload taming` skips them.
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
    ap.add_argument("--no-anole", action="store_true", help="skip the Anole-7B extra block (one ~7 s step + warm-up)")
    ap.add_argument("--lanes", type=int, default=3, help="extra block: this many batches of 16 on concurrent engine lanes (1 = skip)")
    ap.add_argument("--vqgan-precision", choices=["3xtf32", "tf32", "bf16x3", "bf16x3-dec"], default="bf16x3")
    ap.add_argument("--rng", choices=["torch", "philox"], default="torch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=20.0)
    ap.add_argument("--small", action="store_true", help="tiny shapes (plumbing check only; number is INVALID)")
    ap.add_argument("--workload", default="all", choices=["all", "taming", "rar_xl", "detect", "anole"],
                    help="all = Taming headline line + extra_workloads {rar_xl, detect_only}; rar_xl / detect = only "
                         "that block as the line")
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
    if args.workload in ("all", "rar_xl"):
        # the same arm for BASELINE configs[2] (RAR-XL): the CPU port of RAR.generate + MaskGIT-VQGAN decode + detect
        try:
            from wmar_b200.models.rar_engine import RAR_SIZES
            from wmar_b200.models.synthetic import MASKGIT_VQGAN_CFG, maskgit_vqgan_state, rar_state
            cfg = dict(codebook_size=1024, image_seq_len=256, condition_num_classes=1000)
            cfg.update(RAR_SIZES["rar_xl"])
            if args.small:
                cfg.update(dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512))
            rs = rar_state(cfg, seed=0, device="cpu")
            ts = maskgit_vqgan_state(dict(MASKGIT_VQGAN_CFG), seed=1, device="cpu")
            r = rar_cpu_reference_throughput(rs, ts, cfg, 8, min(per, 12.0))
            line["extra_workloads"] = {"rar_xl": {
                "impl": "reference", "metric": RAR_METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
                "ms_per_step": 8 / r["value"] * 1e3, "higher_is_better": True, "scaling": "weak", "dtype": "f32",
                "data": "synthetic", "config": {"workload": "rar_xl_256_B8_cfg4_wm_linear_h1_d2_g0.25", "batch_per_gpu": 8},
                "cpu_baseline": r,
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}}
        except Exception as e:
            line["extra_workloads"] = {"rar_xl": {"impl": "reference", "error": repr(e)[:300]}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------- ours
def load_peaks():
    ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(ppath):
        try:
            return json.load(open(ppath))
        except Exception:
            return {}
    return {}


def profile_traffic(key):
    """DRAM bytes per launch from the committed ncu capture (profiles/roofline_traffic.json): a profile-derived
    constant, NOT measured in this run (ncu cannot run inside a timed bench)."""
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        try:
            return json.load(open(tpath)).get(key)
        except Exception:
            return None
    return None


def measure_generation(args, world, dev, model, wm, cond_list, gen_params, steps_tok, L):
    """Times `steps` passes of sample -> codes_to_images -> detect on one batch per GPU: region 1 with the conditioning
    already on the device, region 2 end to end with pinned host buffers both ways.  Returns times as max over ranks."""
    import torch
    import torch.distributed as dist
    from wmar_b200 import _lib
    B = len(cond_list)
    cond_dev = torch.tensor(cond_list, dtype=torch.long, device=dev)
    cond_pin = torch.tensor(cond_list, dtype=torch.long).pin_memory()
    img_pin = torch.empty((B, 3, model.image_size, model.image_size), dtype=torch.float32).pin_memory()
    codes_pin = torch.empty((B, steps_tok), dtype=torch.long).pin_memory()
    stat_pin = torch.empty((B, 4), dtype=torch.float64).pin_memory()
    ev = {}

    def hot_path(cond):
        codes = model.sample(cond, gen_params, apply_watermark=True)
        ev["s"].record()
        imgs = model.codes_to_images(codes)
        ev["d"].record()
        st = wm.detect_stats(codes)
        return codes, imgs, st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ev["s"], ev["d"] = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    last = None
    for _ in range(args.warmup):
        last = hot_path(cond_dev)    # (kept like in the timed loop: the caching allocator then owns BOTH image buffers the
                                     #  loop alternates between; otherwise the second timed step pays a cudaMalloc)
    barrier()

    # ---- timed region 1: device-resident inputs ----
    clocks = ClockSampler(dev.index or 0)
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
        ev["s"], ev["d"] = evs[i], evd[i]
        ev0[i].record()
        last = hot_path(cond_dev)
    ev_end.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = L.wmar_launch_count() - launches0
    t_total_ms = ev0[0].elapsed_time(ev_end)
    t_sample_ms = sum(ev0[i].elapsed_time(evs[i]) for i in range(args.steps))
    t_decode_ms = sum(evs[i].elapsed_time(evd[i]) for i in range(args.steps))
    per_step = {"sample": [round(ev0[i].elapsed_time(evs[i]), 2) for i in range(args.steps)],
                "vqgan_decode": [round(evs[i].elapsed_time(evd[i]), 2) for i in range(args.steps)]}
    _lib.check(L.wmar_check_device_flag(_lib.current_stream()))

    # ---- timed region 2: end to end through the wrapper with host buffers ----
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    def e2e_step():
        c = cond_pin.to(dev, non_blocking=True)
        codes, imgs, st = hot_path(c)
        img_pin.copy_(imgs, non_blocking=True)
        codes_pin.copy_(codes, non_blocking=True)
        stat_pin[:, 0].copy_(st["n_green"], non_blocking=True)
        stat_pin[:, 1].copy_(st["n_scored"], non_blocking=True)
        stat_pin[:, 2].copy_(st["z"], non_blocking=True)
        stat_pin[:, 3].copy_(st["pvalue"], non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the user reads the result of every step

    e2e_step()   # one untimed pass of the end-to-end loop (first use of the pinned buffers / copy paths), then K timed steps
    barrier()
    e0.record()
    e2e_wall, t_e2e0 = [], time.perf_counter()
    for i in range(args.steps):
        e2e_step()
        e2e_wall.append(round((time.perf_counter() - t_e2e0) * 1e3, 1))
    e1.record()
    barrier()
    t_e2e_ms = e0.elapsed_time(e1)
    clk = clocks.stop()
    h2d = cond_pin.numel() * 8
    d2h = img_pin.numel() * 4 + codes_pin.numel() * 8 + stat_pin.numel() * 8

    tt = torch.tensor([t_total_ms, t_e2e_ms, t_sample_ms, t_decode_ms, t_wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_total_ms, t_e2e_ms, t_sample_ms, t_decode_ms, t_wall_ms = tt.tolist()
    st = last[2]
    return {"t_total_ms": t_total_ms, "t_e2e_ms": t_e2e_ms, "t_sample_ms": t_sample_ms, "t_decode_ms": t_decode_ms,
            "t_wall_ms": t_wall_ms, "per_step_ms": dict(per_step, e2e_wall_cumulative=e2e_wall), "launches": int(launches), "h2d": h2d, "d2h": d2h, "clocks": clk,
            "detector": {"n_green_mean": float(st["n_green"].float().mean()), "z_mean": float(st["z"].mean()),
                         "log10_p_max": float(torch.log10(st["pvalue"].clamp_min(1e-300)).max())}}


def generation_block(args, world, B, steps_tok, m, alg_bytes, peaks, kernel, traffic_key):
    n_img = world * B * args.steps
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes * args.steps / (m["t_sample_ms"] * 1e-3) / 1e9
    traffic = profile_traffic(traffic_key)
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak,
                "peak_source": "measured" if peaks else "fallback", "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic,
                "traffic_source": ("profiles/roofline_traffic.json: constant derived from a committed ncu capture, not "
                                   "measured in this run") if traffic is not None else None,
                "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": m["t_sample_ms"] / args.steps,
                "phase_ms_per_step": {"sample": m["t_sample_ms"] / args.steps, "vqgan_decode": m["t_decode_ms"] / args.steps,
                                      "detect+rest": (m["t_total_ms"] - m["t_sample_ms"] - m["t_decode_ms"]) / args.steps},
                "per_step_ms_rank0": m.get("per_step_ms")}
    return {"value": n_img / (m["t_total_ms"] * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": m["t_total_ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "data": "synthetic",
            "e2e": {"value": n_img / (m["t_e2e_ms"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": m["h2d"],
                    "d2h_bytes_per_step": m["d2h"], "ms_per_step": m["t_e2e_ms"] / args.steps},
            "gpu_launches": m["launches"], "clocks": m["clocks"], "roofline": roofline, "detector": m["detector"],
            "wall_ms_per_step": m["t_wall_ms"] / args.steps}


def run_taming(args, rank, world, dev, L, peaks):
    import torch
    from wmar_b200.distributed import broadcast_state
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.models.synthetic import taming_net2net_state
    from wmar_b200.watermarking import create_watermarker_from_string
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
    m = measure_generation(args, world, dev, model, wm, cond_list, GEN_PARAMS, steps_tok, L)
    blk = generation_block(args, world, B, steps_tok, m, model._gpt.algorithmic_bytes(B, steps_tok), peaks,
                           "decode loop (256 token steps: skinny GEMMs + KV attention + fused sampler)",
                           "decode_loop_dram_bytes_per_launch")
    blk.update({"metric": METRIC, "dtype": "f32 (transformer: 3xTF32 tensor-core products, fp32 accumulate; VQGAN 3x3 convs: " + args.vqgan_precision + ")",
                "config": {"workload": "taming_cin_B16_wm_linear_h1_d2_g0.25" + ("_SMALL_INVALID" if args.small else ""),
                           "batch_per_gpu": B, "global_batch": B * world, "tokens_per_image": steps_tok,
                           "gpt": gpt_cfg, "watermark": WM_STRING, "gen_params": GEN_PARAMS,
                           "vqgan_precision": args.vqgan_precision, "rng": args.rng, "parallelism": f"replicas x{world}",
                           "step_path": os.environ.get("WMAR_STEP", "default"),
                           "l2": "inputs larger than L2 (5.5 GB of weights streamed per token step)"}})
    if args.lanes > 1:
        # the same wrapper call with lanes x B conditionings: the batches of B run concurrently on engine lanes (own KV
        # cache / scratch / step graph, shared weights; taming_wrapper.py sample()).  A step here = lanes x B images.
        model.lanes = args.lanes
        cond2 = [CLASSES[(rank * B + i) % len(CLASSES)] for i in range(args.lanes * B)]
        args2 = argparse.Namespace(**vars(args))      # an extra block: at most 4 timed steps after at most 3 warm-up steps
        args2.steps, args2.warmup = min(args.steps, 4), min(args.warmup, 3)
        m2 = measure_generation(args2, world, dev, model, wm, cond2, GEN_PARAMS, steps_tok, L)
        b2 = generation_block(args2, world, args.lanes * B, steps_tok, m2, args.lanes * model._gpt.algorithmic_bytes(B, steps_tok), peaks,
                              f"{args.lanes} concurrent decode loops (each 256 token steps at {B} rows; every loop streams the weights itself)",
                              None)
        b2.update({"metric": METRIC.replace("batch 16/GPU", f"{args.lanes} concurrent batches of 16/GPU"), "dtype": blk["dtype"],
                   "config": dict(blk["config"], workload=f"taming_cin_{args.lanes}xB16_concurrent_lanes_wm_linear_h1_d2_g0.25",
                                  batch_per_gpu=args.lanes * B, global_batch=args.lanes * B * world, lanes=args.lanes,
                                  note=f"NOT the headline configuration: the wrapper's sample() is given {args.lanes} x 16 conditionings and "
                                       f"runs the {args.lanes} batches of 16 on {args.lanes} engine lanes / CUDA streams; ids identical "
                                       "to the sequential chunk loop (tests/test_gpu_watermark.py)")})
        blk["_lanes_block"] = b2
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gs = {k[len("transformer."):]: v.cpu() for k, v in state.items() if k.startswith("transformer.")}
        vs = {k[len("first_stage_model."):]: v.cpu() for k, v in state.items() if k.startswith("first_stage_model.")}
        blk["cpu_baseline"] = cpu_reference_throughput(gs, vs, gpt_cfg, B, args.cpu_budget_s)
        try:
            blk["cpu_baseline"]["gpu_eager"] = gpu_eager_reference(state, gpt_cfg, B, dev)
        except Exception as e:  # context number only
            blk["cpu_baseline"]["gpu_eager"] = {"unavailable": repr(e)[:200]}
    del model, wm, state
    torch.cuda.empty_cache()
    return blk


def gpu_eager_reference(state, gpt_cfg, B, dev, n_steps=12):
    """Context only: the torch restatement of the reference path (oracle/gpt.py, the same code the CPU
    baseline runs) executed EAGERLY ON THE GPU -- transformer step on the device, watermark + warpers on the host like
    the reference's per-row greenlist -- over a bounded sample of decode steps, extrapolated to 256 + VQGAN decode."""
    import torch
    from oracle import gpt as ogpt
    from oracle import sampling, vqgan as ov, wm as owm
    V = gpt_cfg["vocab_size"]
    gs = {k[len("transformer."):]: v for k, v in state.items() if k.startswith("transformer.")}
    vs = {k[len("first_stage_model."):]: v for k, v in state.items() if k.startswith("first_stage_model.")}
    assets = os.path.join(ROOT, "wmar_b200", "assets", "vqgan_alive_ids.txt")
    alive, dead = owm.alive_dead(owm.load_ids(assets), V)
    rows = owm.GreenRows(V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    o = ogpt.GPTOracle(gs, gpt_cfg["n_layer"], gpt_cfg["n_head"])
    cond = torch.tensor([CLASSES[i % len(CLASSES)] for i in range(B)], dtype=torch.long)
    gen = torch.Generator().manual_seed(1)
    x, seq = cond.clone(), cond.view(-1, 1).clone()
    with torch.no_grad():
        t_dec = 0.0
        for n in range(n_steps + 2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            logits = o.step(x.to(dev), n).cpu()
            noise = torch.empty(B, V).exponential_(1, generator=gen)
            x = sampling.sample_step(logits, rows(seq), 2.0, GEN_PARAMS["temperature"], GEN_PARAMS["top_k"],
                                     GEN_PARAMS["top_p"], noise)
            seq = torch.cat((seq, x.view(-1, 1)), dim=1)
            if n >= 2:
                t_dec += time.perf_counter() - t0
        codes = torch.randint(0, V, (B, 256), generator=gen).to(dev)
        ov.taming_codes_to_images(codes[:1], vs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ov.taming_codes_to_images(codes, vs)
        torch.cuda.synchronize()
        t_img = time.perf_counter() - t0
    per_batch = t_dec / n_steps * 256 + t_img
    return {"value": B / per_batch, "unit": UNIT, "kind": "port, torch eager on cuda",
            "sample": f"{n_steps} of 256 decode steps at batch {B} ({t_dec / n_steps * 1e3:.1f} ms/step, transformer on the "
                      f"GPU in fp32, watermark + warpers on the host) + VQGAN decode of {B} images ({t_img * 1e3:.0f} ms), "
                      "extrapolated linearly"}


RAR_METRIC = "watermarked 256x256 images/sec end-to-end (sample -> decode -> detect), RAR-XL, CFG 4.0, batch 8/GPU"
RAR_CLASSES = [1, 9, 232, 340, 568, 656, 703, 814]


def rar_cpu_reference_throughput(state, tstate, cfg, B, budget_s):
    """images/s of the oracle port of RAR.generate + MaskGIT-VQGAN decode + detect on the host cores (bounded sample)."""
    import torch
    from oracle import rar as orar
    from oracle import sampling, vqgan as ov, wm as owm
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    V = cfg["codebook_size"]
    assets = os.path.join(ROOT, "wmar_b200", "assets", "rar_all_ids.txt")
    alive, dead = owm.alive_dead(owm.load_ids(assets), V)
    rows_fn = owm.GreenRows(V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
    o = orar.RAROracle(state, cfg["num_hidden_layers"], cfg["num_attention_heads"])
    cond = torch.tensor([RAR_CLASSES[i % len(RAR_CLASSES)] for i in range(B)], dtype=torch.long)
    rows = torch.cat([cond + V + 1, torch.full_like(cond, o.none_id)])
    gen = torch.Generator().manual_seed(1)
    ids = torch.zeros((B, 0), dtype=torch.long)
    t_dec, n_dec = 0.0, 0
    with torch.no_grad():
        while n_dec < 24 and (n_dec < 3 or t_dec < budget_s * 0.6):
            t0 = time.perf_counter()
            last = torch.cat([ids[:, -1], ids[:, -1]]) if n_dec > 0 else None
            lg = o.step(n_dec, rows, last)
            logits = lg[B:] + (lg[:B] - lg[B:]) * 4.0
            noise = torch.empty(B, V).exponential_(1, generator=gen)
            nxt = sampling.sample_step(logits, rows_fn(ids) if n_dec > 0 else None, 2.0, 1.0, None, None, noise)
            ids = torch.cat([ids, nxt.view(-1, 1)], dim=-1)
            t_dec += time.perf_counter() - t0
            n_dec += 1
        codes = torch.randint(0, V, (B, 256), generator=gen)
        n_img, t_img = 0, 0.0
        while n_img < B and (n_img < 1 or t_img < budget_s * 0.3):
            t0 = time.perf_counter()
            ov.rar_codes_to_images(codes[n_img:n_img + 1], tstate)
            t_img += time.perf_counter() - t0
            n_img += 1
        t0 = time.perf_counter()
        ng, ns = owm.detect_counts(codes.numpy(), V, 0.25, "stratifiedrand", "linear", 1, alive, dead)
        owm.pvalue(ng, ns, 0.25)
        t_det = time.perf_counter() - t0
    per_batch = t_dec / n_dec * 256 + t_img / n_img * B + t_det
    return {"value": B / per_batch, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{n_dec} of 256 guided passes at {2 * B} rows ({t_dec / n_dec * 1e3:.0f} ms/pass) + {n_img} of {B} "
                      f"MaskGIT-VQGAN decodes ({t_img / n_img * 1e3:.0f} ms/img) + detect on {B} rows ({t_det * 1e3:.0f} ms), "
                      "extrapolated linearly to a full batch; torch fp32 CPU, all host threads"}


def run_rar_xl(args, rank, world, dev, L, peaks):
    """BASELINE.json configs[2]: RAR-XL 256x256, greenlist watermark, batch 64 sharded 8 per GPU over 8 GPUs (here: 8 images
    = 16 guided rows per GPU per step at every N, weak scaling)."""
    import torch
    from wmar_b200.distributed import broadcast_state
    from wmar_b200.models import RarARMMWrapper
    from wmar_b200.models.rar_engine import RAR_SIZES
    from wmar_b200.models.synthetic import MASKGIT_VQGAN_CFG, maskgit_vqgan_state, rar_state
    from wmar_b200.watermarking import create_watermarker_from_string
    B = 8
    size = "rar_xl"
    cfg = dict(codebook_size=1024, image_seq_len=256, condition_num_classes=1000)
    cfg.update(RAR_SIZES[size])
    over = None
    if args.small:
        over = dict(hidden_size=256, num_hidden_layers=2, num_attention_heads=4, intermediate_size=512)
        cfg.update(over)
    state = rar_state(cfg, seed=0, device=dev)
    tstate = maskgit_vqgan_state(dict(MASKGIT_VQGAN_CFG), seed=1, device=dev)
    broadcast_state(state, src=0)
    broadcast_state(tstate, src=0)
    model = RarARMMWrapper(rar_size=size, state_dict=state, tokenizer_state_dict=tstate, rar_cfg=over, device=dev,
                           max_batch=B, vqgan_precision=args.vqgan_precision, rng=args.rng)
    wm = create_watermarker_from_string(model.get_vq(), model.get_total_vocab_size(), WM_STRING, dev)
    model.set_watermarker(wm)
    torch.manual_seed(1 + 1000 * rank)
    torch.cuda.manual_seed_all(1 + 1000 * rank)
    steps_tok = model.codes_size ** 2
    cond_list = [RAR_CLASSES[(rank * B + i) % len(RAR_CLASSES)] for i in range(B)]
    m = measure_generation(args, world, dev, model, wm, cond_list, None, steps_tok, L)
    blk = generation_block(args, world, B, steps_tok, m, model._rar.algorithmic_bytes(B, steps_tok), peaks,
                           "RAR decode loop (256 guided passes over 16 rows: skinny GEMMs + KV attention + CFG / watermark / sampler)",
                           "rar_decode_loop_dram_bytes_per_launch")
    blk.update({"metric": RAR_METRIC, "dtype": "f32 (transformer: 3xTF32 tensor-core products, fp32 accumulate; VQGAN 3x3 convs: " + args.vqgan_precision + ")",
                "config": {"workload": "rar_xl_256_B8_cfg4_wm_linear_h1_d2_g0.25" + ("_SMALL_INVALID" if args.small else ""),
                           "batch_per_gpu": B, "global_batch": B * world, "tokens_per_image": steps_tok, "rar": cfg,
                           "guidance_scale": 4.0, "watermark": WM_STRING, "vqgan_precision": args.vqgan_precision,
                           "rng": args.rng, "parallelism": f"replicas x{world}",
                           "l2": "inputs larger than L2 (weights streamed per pass)"}})
    if args.lanes > 1:
        model.lanes = args.lanes     # same wrapper call with lanes x 8 conditionings: concurrent engine lanes (see run_taming)
        cond2 = [RAR_CLASSES[(rank * B + i) % len(RAR_CLASSES)] for i in range(args.lanes * B)]
        args2 = argparse.Namespace(**vars(args))
        args2.steps, args2.warmup = min(args.steps, 4), min(args.warmup, 3)
        m2 = measure_generation(args2, world, dev, model, wm, cond2, None, steps_tok, L)
        b2 = generation_block(args2, world, args.lanes * B, steps_tok, m2, args.lanes * model._rar.algorithmic_bytes(B, steps_tok), peaks,
                              f"{args.lanes} concurrent RAR decode loops (each 256 guided passes over 16 rows; every loop streams the weights itself)",
                              None)
        b2.update({"metric": RAR_METRIC.replace("batch 8/GPU", f"{args.lanes} concurrent batches of 8/GPU"), "dtype": blk["dtype"],
                   "config": dict(blk["config"], workload=f"rar_xl_256_{args.lanes}xB8_concurrent_lanes_cfg4_wm_linear_h1_d2_g0.25",
                                  batch_per_gpu=args.lanes * B, global_batch=args.lanes * B * world, lanes=args.lanes,
                                  note="NOT configuration 3 itself: lanes x 8 images per GPU per step on concurrent engine lanes")})
        blk["lanes_block"] = b2
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        blk["cpu_baseline"] = rar_cpu_reference_throughput({k: v.cpu() for k, v in state.items()},
                                                           {k: v.cpu() for k, v in tstate.items()}, cfg, B,
                                                           min(args.cpu_budget_s, 12.0))
    del model, wm, state, tstate
    torch.cuda.empty_cache()
    return blk


DETECT_METRIC = "detection-only 256x256 images/sec (VQGAN encode + z-score), Taming tokenizer, batch 16/GPU"


def run_detect(args, rank, world, dev, L, peaks):
    """BASELINE.json configs[4]: detection only = VQGAN encode + detector on synthetic 256x256 images (images sharded over
    the GPUs, no exchange).  One step = 8 batches of 16 images per GPU; e2e copies the images H2D and the statistics D2H."""
    import torch
    import torch.distributed as dist
    from wmar_b200 import _lib
    from wmar_b200.models import TamingARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    B, n_batches = 16, 8
    m = TamingARMMWrapper(gpt_cfg=dict(vocab_size=16384, block_size=256, n_layer=1, n_head=24, n_embd=1536), device=dev,
                          max_batch=B, vqgan_precision=args.vqgan_precision)
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), WM_STRING, dev)
    g = torch.Generator().manual_seed(100 + rank)
    imgs_pin = [(torch.rand(B, 3, 256, 256, generator=g) * 2 - 1).pin_memory() for _ in range(n_batches)]
    imgs_dev = [x.to(dev) for x in imgs_pin]
    stat_pin = torch.empty((n_batches, B, 4), dtype=torch.float64).pin_memory()

    def step(src, host):
        out = []
        for i, x in enumerate(src):
            if host:
                x = x.to(dev, non_blocking=True)
            st = wm.detect_stats(m.images_to_codes(x))
            if host:
                for j, k in enumerate(("n_green", "n_scored", "z", "pvalue")):
                    stat_pin[i, :, j].copy_(st[k], non_blocking=True)
            out.append(st)
        if host:
            torch.cuda.current_stream().synchronize()
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(imgs_dev, False)
    barrier()
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    l0 = L.wmar_launch_count()
    a0, a1, e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    a0.record()
    for _ in range(args.steps):
        sts = step(imgs_dev, False)
    a1.record()
    barrier()
    launches = L.wmar_launch_count() - l0
    from wmar_b200.evaluate import detect_host_batches   # the bulk-detection entry: H2D of batch i+1 under the encode of batch i
    detect_host_batches(m, wm, imgs_pin)                 # one untimed pass (allocates the staging / pinned result buffers)
    barrier()
    e0.record()
    for _ in range(args.steps):
        stat_pin = detect_host_batches(m, wm, imgs_pin)
    e1.record()
    barrier()
    clk = clocks.stop()
    tt = torch.tensor([a0.elapsed_time(a1), e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_dev_ms, t_e2e_ms = tt.tolist()
    _lib.check(L.wmar_check_device_flag(_lib.current_stream()))
    n_img = world * B * n_batches * args.steps
    fl = m._vqgan.flops(decode=False) * B * n_batches
    pk = float(peaks.get("bf16_tflops_sustained", 1363.5))
    ach = fl * args.steps / (t_dev_ms * 1e-3) / 1e12
    blk = {"metric": DETECT_METRIC, "value": n_img / (t_dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": t_dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32 (VQGAN convs: " + args.vqgan_precision + " tensor-core products, fp32 accumulate)", "data": "synthetic",
           "config": {"workload": "detect_only_taming_encode_256_B16x8", "batch_per_gpu": B, "batches_per_step": n_batches,
                      "watermark": WM_STRING, "vqgan_precision": args.vqgan_precision, "parallelism": f"replicas x{world}",
                      "l2": "8 x 12.6 MB of images per step + activations larger than L2 between batches"},
           "e2e": {"value": n_img / (t_e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": B * n_batches * 3 * 256 * 256 * 4,
                   "d2h_bytes_per_step": stat_pin.numel() * 8, "ms_per_step": t_e2e_ms / args.steps,
                   "api": "wmar_b200.evaluate.detect_host_batches (pinned host images in, statistics out; copy stream overlaps the encoder)"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "tensor", "kernel": "VQGAN encoder conv stack (persistent tcgen05 implicit GEMM, " + args.vqgan_precision + ") + codebook arg-min",
                        "achieved": ach, "peak": pk, "peak_source": "measured bf16_tflops_sustained" if peaks else "fallback",
                        "unit": "TFLOP/s", "frac": ach / pk, "frac_counting_the_three_products": 3.0 * ach / pk, "traffic": None,
                        "note": ("useful fp32-equivalent FLOPs; every product is 3 bf16 MMAs (two-term split), whose own ceiling is 1/3 of the bf16 peak"
                                 if args.vqgan_precision.startswith("bf16x3") else
                                 "useful fp32-equivalent FLOPs; every product is 3 TF32 MMAs, whose own ceiling is 1/6 of the bf16 peak")},
           "detector": {"p_mean": float(torch.cat([s_["pvalue"] for s_ in sts]).mean())}, "clocks": clk}
    del m, wm
    torch.cuda.empty_cache()
    return blk


ANOLE_METRIC = "watermarked 512x512 images/sec end-to-end (text prompt -> sample -> decode -> detect), Anole-7B shapes, 8 images/GPU"
ANOLE_PROMPTS = ["a photo of a red bus parked next to a building on a sunny day", "two cats sleeping on a couch",
                 "a plate of food with broccoli and rice on a wooden table", "a man riding a wave on a surfboard",
                 "a kitchen with a stove a sink and a window", "a group of people flying kites in a park",
                 "a close up of a pizza on a table", "a train traveling down tracks next to a forest"]


def run_anole(args, rank, world, dev, L, peaks):
    """BASELINE.json configs[3]: Anole-7B text -> image 512x512, watermark on (here: 8 images = 16 guided rows per GPU per
    step at every N, random-init bf16 weights at the 7B shapes, synthetic prompts through the wrapper's stand-in tokenizer).
    A step takes ~7 s, so this block times ONE step after ONE warm-up (stated in `steps` / `warmup`)."""
    import torch
    import torch.distributed as dist
    from wmar_b200 import _lib
    from wmar_b200.models.chameleon_wrapper import ChameleonARMMWrapper
    from wmar_b200.watermarking import create_watermarker_from_string
    B = 8
    torch.manual_seed(0)
    m = ChameleonARMMWrapper(max_batch=B, device=dev, vqgan_precision=args.vqgan_precision)
    wm_string = "fixed-stratifiedrand-h=0-d=2.0-g=0.25"
    wm = create_watermarker_from_string(m.get_vq(), m.get_total_vocab_size(), wm_string, dev)
    m.set_watermarker(wm)
    torch.manual_seed(1 + 1000 * rank)
    cond = list(enumerate(ANOLE_PROMPTS))
    gp = {"temperature": 0.9, "top_p": 0.9}
    img_pin = torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory()

    def step():
        codes = m.sample(cond, gp, apply_watermark=True)
        e_s.record()
        imgs = m.codes_to_images(codes)
        e_d.record()
        st = wm.detect_stats(codes)
        img_pin.copy_(imgs, non_blocking=True)
        return st

    e0, e_s, e_d, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches0 = L.wmar_launch_count()
    e0.record()
    st = step()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    launches = L.wmar_launch_count() - launches0
    tt = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e_s), e_s.elapsed_time(e_d)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, t_s, t_d = tt.tolist()
    _lib.check(L.wmar_check_device_flag(_lib.current_stream()))
    p_max = max(len(r) for r in m.prompt_rows(ANOLE_PROMPTS))
    by = m._eng.algorithmic_bytes(B, p_max, 1024)
    pk = float(peaks.get("hbm_gbs", 6650.0))
    blk = {"metric": ANOLE_METRIC, "value": world * B / (t_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
           "ms_per_step": t_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "bf16 weights / activations as the reference (fp32 accumulate); VQGAN 3x3 convs: " + args.vqgan_precision,
           "data": "synthetic",
           "config": {"workload": "anole_7b_512_B8_cfg3.0_1.2_T0.9_p0.9_wm_fixed_h0_d2_g0.25", "batch_per_gpu": B,
                      "global_batch": B * world, "tokens_per_image": 1024, "watermark": wm_string, "gen_params": gp,
                      "parallelism": f"replicas x{world}", "l2": "inputs larger than L2 (13.5 GB of bf16 weights streamed per pass)"},
           "e2e": {"value": world * B / (t_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": 8 * p_max * 3 * B,
                   "d2h_bytes_per_step": img_pin.numel() * 4, "ms_per_step": t_ms,
                   "note": "the timed step IS the end-to-end call: host prompts in, images copied to pinned host memory"},
           "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "kernel": "Anole-7B decode loop (prompt + 1023 passes over 16 rows, bf16 skinny GEMMs + GQA attention + CFG / watermark / sampler)",
                        "achieved": by / t_s / 1e6, "peak": pk, "peak_source": "measured" if peaks else "fallback", "unit": "GB/s",
                        "frac": by / t_s / 1e6 / pk, "traffic": None, "algorithmic_bytes_per_launch": by, "ms_per_launch": t_s,
                        "phase_ms_per_step": {"sample": t_s, "vqgan_decode_512": t_d, "detect+rest": t_ms - t_s - t_d}},
           "detector": {"n_green_mean": float(st["n_green"].float().mean()), "z_mean": float(st["z"].mean())}}
    if args.lanes > 1:
        # lanes x 8 prompts through the same wrapper call: concurrent engine lanes (see run_taming); one step, one warm-up
        cond2 = [(i, ANOLE_PROMPTS[i % len(ANOLE_PROMPTS)]) for i in range(args.lanes * B)]
        m.lanes = args.lanes
        img2 = torch.empty((args.lanes * B, 3, 512, 512), dtype=torch.float32).pin_memory()

        def step2():
            codes = m.sample(cond2, gp, apply_watermark=True)
            e_s.record()
            imgs = m.codes_to_images(codes)
            st2 = wm.detect_stats(codes)
            img2.copy_(imgs, non_blocking=True)
            return st2

        step2()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        step2()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t2 = torch.tensor([e0.elapsed_time(e1), e0.elapsed_time(e_s)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        t2_ms, t2_s = t2.tolist()
        _lib.check(L.wmar_check_device_flag(_lib.current_stream()))
        blk["lanes_block"] = {"metric": ANOLE_METRIC.replace("8 images/GPU", f"{args.lanes} concurrent batches of 8/GPU"),
                              "value": world * args.lanes * B / (t2_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": 1, "warmup": 1,
                              "ms_per_step": t2_ms, "lanes": args.lanes,
                              "roofline": {"bound": "hbm", "achieved": args.lanes * by / t2_s / 1e6, "peak": pk, "unit": "GB/s",
                                           "frac": args.lanes * by / t2_s / 1e6 / pk,
                                           "kernel": f"{args.lanes} concurrent Anole-7B decode loops (every loop streams the weights itself)"},
                              "note": "NOT configuration 4 itself: lanes x 8 images per GPU per step on concurrent engine lanes"}
    del m, wm
    torch.cuda.empty_cache()
    return blk


def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from wmar_b200 import _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; wmar_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()  # raises if the CUDA library is missing
    peaks = load_peaks()
    runners = {"taming": run_taming, "rar_xl": run_rar_xl, "detect": run_detect, "anole": run_anole}
    if args.workload == "all":
        line = run_taming(args, rank, world, dev, L, peaks)
        extra = {}
        if "_lanes_block" in line:
            extra["taming_concurrent_lanes"] = line.pop("_lanes_block")
        for name, key in (("rar_xl", "rar_xl"), ("detect", "detect_only"), ("anole", "anole_7b")):
            if name == "anole" and (args.no_anole or args.small):
                continue
            try:
                extra[key] = runners[name](args, rank, world, dev, L, peaks)
            except Exception as e:  # an extra block must never take the headline line down
                extra[key] = {"error": repr(e)[:300]}
                if world > 1:
                    raise
        line["extra_workloads"] = extra
    else:
        line = runners[args.workload](args, rank, world, dev, L, peaks)
        if "_lanes_block" in line:
            line["extra_workloads"] = {"taming_concurrent_lanes": line.pop("_lanes_block")}
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
