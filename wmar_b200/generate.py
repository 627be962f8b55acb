"""The reference's generation CLI (generate.py:167-345) wired to the B200 engines.

    python -m wmar_b200.generate --outdir out --model taming [--modelpath DIR] --conditioning 1,9,232 \
        --num_samples_per_conditioning 2 --batch_size 16 --wm_method gentime --wm_seed_strategy linear \
        --wm_split_strategy stratifiedrand --wm_context_size 1 --wm_delta 2 --wm_gamma 0.25 --orig_only true

Same flags, chunking rule (`batch_idx % num_chunks == chunk_id`, seed + 1000 * chunk_id) and output tree as the
reference.  Under ``torchrun`` every rank is one chunk (chunk_id = RANK, num_chunks = WORLD_SIZE): independent
images, no data-path collective.  Without --modelpath the models are seeded random-init at the reference's shapes (no
checkpoints ship with this repo).  Full mode runs the reference's evaluation round trip on the device
(wmar_b200.evaluate: decode -> [round trip | augment] -> re-encode -> detect) and writes one png / npy / json per
(image, transform, parameter) exactly like generate.py:37-108; ``--orig_only true`` writes the images/ + codes/ tree.
The neural-compression and DiffPure transforms (and the ``bpp`` metric they define) are out of scope: ``bpp`` is null.
"""
import argparse
import json
import os
import random
import sys

import numpy as np


def group_batches(batches, lanes):
    """Consecutive runs of `lanes` planned batches (the last group may be shorter); lanes = 1 -> one batch per group."""
    group = []
    for item in batches:
        group.append(item)
        if len(group) == lanes:
            yield group
            group = []
    if group:
        yield group


def plan_batches(all_inputs, batch_size, chunk_id=0, num_chunks=1):
    """generate.py:179-207 -- split into batches (last may be smaller), keep the 1-based running count per
    conditioning even for batches another chunk owns, return this chunk's [(batch_idx, batch, cond_indices)]."""
    batches = [all_inputs[i * batch_size:(i + 1) * batch_size] for i in range(len(all_inputs) // batch_size)]
    if len(all_inputs) % batch_size != 0:
        batches.append(all_inputs[(len(all_inputs) // batch_size) * batch_size:])
    counts, mine = {}, []
    for batch_idx, batch in enumerate(batches):
        cond_indices = []
        for c in batch:
            if isinstance(c, tuple):
                c = c[0]
            counts[c] = counts.get(c, 0) + 1
            cond_indices.append(counts[c])
        if batch_idx % num_chunks != chunk_id:
            continue
        mine.append((batch_idx, batch, cond_indices))
    return mine


def expand_conditionings(conditioning, num_samples_per_conditioning):
    """generate.py:334-347 -- comma separated ImageNet classes or a prompt file, each repeated n times."""
    if ".txt" in conditioning:
        with open(conditioning, "r") as f:
            conds = [(idx, line.strip()) for idx, line in enumerate(f)]
    else:
        conds = [int(c) for c in conditioning.split(",")]
    return [c for c in conds for _ in range(num_samples_per_conditioning)]


def output_paths(outdir, conditioning, cond_index, method, orig_only, transform="roundtrips", param=0):
    """generate.py:79-108 -- file stems of one image."""
    if isinstance(conditioning, tuple):
        conditioning = conditioning[0]
    if orig_only:
        return (os.path.join(outdir, "images", f"{conditioning}:{cond_index:04}.png"),
                os.path.join(outdir, "codes", f"{conditioning}:{cond_index:04}.npy"), None)
    d = os.path.join(outdir, f"c={conditioning},idx={cond_index}")
    stem = os.path.join(d, f"{cond_index:04}_{method}_{transform}_{param}")
    return stem + ".png", stem + ".npy", stem + ".json"


def chw_to_uint8(img):
    """wmar/utils/utils.py chw_to_pillow: [-1,1] CHW float -> HWC uint8"""
    x = np.clip((np.asarray(img, dtype=np.float32) + 1.0) / 2.0, 0.0, 1.0)
    return (x.transpose(1, 2, 0) * 255.0).round().astype(np.uint8)


def save_png(path, hwc_uint8):
    try:
        from PIL import Image
        Image.fromarray(hwc_uint8).save(path)
    except ImportError:  # PIL is optional here: fall back to a raw .npy next to the expected name
        np.save(path + ".npy", hwc_uint8)


def save_batch_log(log, outdir, watermarker, eval_params, cond_indices):
    """generate.py:37-108 compute_metrics_and_save_from_batch_log: metrics per (method, transform, param, image) and the
    reference's file names.  The log holds device (or CPU) tensors (wmar_b200.evaluate.fill_batch_log); the metrics of a
    whole batch are computed at once and only then copied to the host."""
    from .evaluate import compute_metrics, to_uint8
    names = list(eval_params["metric_names"])
    batch = log["batch"]
    for method in [k for k in log.keys() if k != "batch"]:
        metrics = compute_metrics(log, method, watermarker, [n for n in names if n != "bpp"])
        for transform, entries in log[method].items():
            for e_idx, (param, codes, imgs, _) in enumerate(entries):
                codes_h = codes.cpu().numpy() if hasattr(codes, "cpu") else np.asarray(codes)
                u8 = to_uint8(imgs if hasattr(imgs, "cpu") else __import__("torch").as_tensor(imgs)).permute(0, 2, 3, 1).cpu().numpy()
                m = {n: (v.cpu().tolist() if v is not None else None) for n, v in metrics[transform][e_idx][1].items()}
                for i in range(len(codes_h)):
                    conditioning = batch[i]
                    if hasattr(conditioning, "item"):
                        conditioning = conditioning.item()
                    if isinstance(conditioning, tuple):
                        conditioning = conditioning[0]      # only the index if there is a prompt string too
                    cond_index = cond_indices[i]
                    if not eval_params["orig_only"]:
                        d = os.path.join(outdir, f"c={conditioning},idx={cond_index}")
                        os.makedirs(d, exist_ok=True)
                        stem = os.path.join(d, f"{cond_index:04}_{method}_{transform}_{param}")
                        save_png(stem + ".png", u8[i])
                        np.save(stem + ".npy", codes_h[i])
                        row = {n: (None if n == "bpp" or m.get(n) is None else float(m[n][i])) for n in names}
                        with open(stem + ".json", "w") as f:
                            json.dump(row, f)
                    else:
                        assert param == 0 and transform == "roundtrips"
                        os.makedirs(os.path.join(outdir, "images"), exist_ok=True)
                        os.makedirs(os.path.join(outdir, "codes"), exist_ok=True)
                        suffix = f"_{method}" if len(log.keys()) > 2 else ""
                        save_png(os.path.join(outdir, "images", f"{conditioning}:{cond_index:04}{suffix}.png"), u8[i])
                        np.save(os.path.join(outdir, "codes", f"{conditioning}:{cond_index:04}{suffix}.npy"), codes_h[i])


def get_parser():
    def str2bool(v):
        if isinstance(v, bool):
            return v
        if v.lower() in ("yes", "true", "t", "y", "1"):
            return True
        if v.lower() in ("no", "false", "f", "n", "0"):
            return False
        raise argparse.ArgumentTypeError("Boolean value expected.")

    p = argparse.ArgumentParser()
    p.add_argument("--outdir", type=str)
    p.add_argument("--model", type=str, choices=["taming", "chameleon7b", "rar"])
    p.add_argument("--modelpath", type=str, default=None)
    p.add_argument("--encoder_ft_ckpt", type=str)
    p.add_argument("--decoder_ft_ckpt", type=str)
    p.add_argument("--num_samples_per_conditioning", type=int, default=1)
    p.add_argument("--conditioning", type=str)
    p.add_argument("--batch_size", type=int, nargs="?", default=10)
    p.add_argument("--lanes", type=int, default=3,
                   help="consecutive batches sampled concurrently on engine lanes of one GPU (1 = the reference's sequential loop)")
    p.add_argument("--top_k", type=int, nargs="?", default=600)
    p.add_argument("--temperature", type=float, nargs="?", default=1.0)
    p.add_argument("--top_p", type=float, nargs="?", default=0.92)
    p.add_argument("--chunk_id", type=int, nargs="?", default=None)
    p.add_argument("--num_chunks", type=int, nargs="?", default=None)
    p.add_argument("--orig_only", type=str2bool, nargs="?", default=False)
    p.add_argument("--wm_method", type=str, nargs="?", choices=["none", "gentime"], default="none")
    p.add_argument("--wm_seed_strategy", type=str, nargs="?", choices=["fixed", "linear", "spatial"])
    p.add_argument("--wm_split_strategy", type=str, nargs="?", choices=["rand", "stratifiedrand", "clustering"])
    p.add_argument("--wm_context_size", type=int, nargs="?", default=0)
    p.add_argument("--wm_delta", type=float, nargs="?")
    p.add_argument("--wm_gamma", type=float, nargs="?", default=0)
    p.add_argument("--seed", type=int, nargs="?", default=42)
    return p


def main(argv=None):
    import torch
    args, _ = get_parser().parse_known_args(argv)
    assert args.outdir, "Output directory is not set"
    os.makedirs(args.outdir, exist_ok=True)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    chunk_id = args.chunk_id if args.chunk_id is not None else rank
    num_chunks = args.num_chunks if args.num_chunks is not None else world
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"

    seed = args.seed + (1000 * chunk_id)  # generate.py:304
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)

    from wmar_b200.models import ChameleonARMMWrapper, RarARMMWrapper, TamingARMMWrapper
    from wmar_b200.models.state import update_weights
    from wmar_b200.watermarking import GentimeWatermark, SeedStrategy, SplitStrategy
    if args.model == "taming":
        model = TamingARMMWrapper(args.modelpath, device=device, max_batch=min(16, args.batch_size))
    elif args.model == "rar":
        model = RarARMMWrapper(args.modelpath, device=device, max_batch=min(8, args.batch_size))
    elif args.model == "chameleon7b":
        # no checkpoint / text tokenizer offline: Anole-7B shapes with seeded random weights (wrapper docstring)
        model = ChameleonARMMWrapper(args.modelpath, device=device, max_batch=min(5, args.batch_size), seed=args.seed)
    else:
        raise ValueError(f"Model {args.model} not supported by wmar_b200")
    patched = False
    if args.encoder_ft_ckpt not in (None, "none"):
        update_weights(model.get_image_tokenizer().encoder, args.encoder_ft_ckpt)
        patched = True
    if args.decoder_ft_ckpt not in (None, "none"):
        update_weights(model.get_image_tokenizer().decoder, args.decoder_ft_ckpt)
        patched = True
    if patched:
        model.sync_weights()

    all_inputs = expand_conditionings(args.conditioning, args.num_samples_per_conditioning)
    if args.model == "rar":
        assert args.wm_method == "none" or (args.wm_seed_strategy in ["linear", "fixed"]
                                            and args.wm_split_strategy == "stratifiedrand")
    watermarker = None
    if args.wm_method == "gentime":
        watermarker = GentimeWatermark(model.get_vq(), model.get_total_vocab_size(), SeedStrategy(args.wm_seed_strategy),
                                       SplitStrategy(args.wm_split_strategy), args.wm_context_size, args.wm_delta,
                                       args.wm_gamma, model.device)
    model.set_watermarker(watermarker)
    gen_params = {"batch_size": args.batch_size, "temperature": args.temperature, "top_k": args.top_k,
                  "top_p": args.top_p}
    from wmar_b200.augmentations import default_augmentations
    from wmar_b200.evaluate import fill_batch_log
    if args.orig_only:      # generate.py:381-384
        eval_params = {"metric_names": [], "augmentations": [], "max_roundtrips": 0, "orig_only": True}
    else:
        eval_params = {"metric_names": ["pvalue", "l0", "psnr", "bpp"], "augmentations": default_augmentations(),
                       "max_roundtrips": 1, "orig_only": False}
    n_done = 0
    # `--lanes` consecutive batches are sampled by ONE wrapper call: the wrapper cuts the concatenated conditioning back into
    # the same batches (max_batch == batch_size) and runs them on concurrent engine lanes -- same codes as one call per batch
    # (only when a batch is ONE engine call -- otherwise the concatenation would be cut differently than the batches)
    cap = {"taming": 16, "rar": 8}.get(args.model, 5)
    lanes = max(1, args.lanes) if args.batch_size <= cap else 1
    for group in group_batches(plan_batches(all_inputs, args.batch_size, chunk_id, num_chunks), lanes):
        conds = [c for _, batch, _ in group for c in batch]
        codes_all = model.sample(conds, gen_params, apply_watermark=watermarker is not None)
        o = 0
        for batch_idx, batch, cond_indices in group:
            codes = codes_all[o:o + len(batch)]
            o += len(batch)
            batch_log = {"batch": batch}
            fill_batch_log(batch_log, str(watermarker), model, codes, eval_params)
            save_batch_log(batch_log, args.outdir, watermarker, eval_params, cond_indices)
            n_done += len(batch)
    print(f"[chunk {chunk_id}/{num_chunks}] wrote {n_done} images to {args.outdir}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
