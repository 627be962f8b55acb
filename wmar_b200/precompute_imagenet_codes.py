"""The reference's bulk tokeniser (precompute_imagenet_codes.py) on the B200 VQGAN engine.

    python -m wmar_b200.precompute_imagenet_codes --model taming [--modelpath DIR] --imagenet_root data/imagenet/061417/ \
        --outdir out/imagenet_taming [--batch_size 16] [--classes 0,999] [--max_per_class 2]

Same inputs and outputs as the reference script: ``labels.txt`` + ``train/<wnid>/`` under --imagenet_root, 50 images per
class (or the counts of ``assets/imagenet_512_split_50k.txt`` for the 512-pixel Chameleon tokenizer) drawn with
``np.random.choice(..., replace=False)`` + ``np.random.shuffle`` under seed 1 (precompute_imagenet_codes.py:22-25,70-84),
Resize(size) + RandomCrop + ToTensor + 2x-1 per image in the reference's order (so torch's RNG is consumed identically),
``codes/<class idx>:<count:04>.npy`` and ``images/<class idx>:<count:04>.png`` per image (:124-127).
What changes: the images of a class go through ``images_to_codes`` in batches (one encoder launch sequence per
--batch_size images) instead of one ``vqgan.encode`` per image.
The reference carries two hard-coded filters (only class indices 0 and 999, only the first two images of a class,
:112-120); here they are the flags ``--classes`` / ``--max_per_class`` and default to "everything".
"""
import argparse
import json
import os
import random

import numpy as np


def load_labels(imagenet_root):
    """labels.txt lines are "<wnid>,<name>" (precompute_imagenet_codes.py:57-60)."""
    with open(os.path.join(imagenet_root, "labels.txt"), "r") as f:
        return [line.strip().split(",")[0] for line in f.readlines() if line.strip()]


def counts_per_label(labels, size, split_512_path=None):
    """50 per class, or the custom split for 512-pixel images (:62-72)."""
    if size == 512 and split_512_path and os.path.exists(split_512_path):
        with open(split_512_path, "r") as f:
            rows = [line.strip().split(",") for line in f.readlines() if line.strip()]
        return {k: int(v) for k, v in rows}
    return {k: 50 for k in labels}


def select_paths(imagenet_root, labels, cnt_per_label):
    """np.random.choice without replacement, then np.random.shuffle, label by label (:74-82): the caller seeds numpy."""
    paths = {}
    for label in labels:
        cls_dir = os.path.join(imagenet_root, "train", label)
        cls_paths = [os.path.join(cls_dir, p) for p in os.listdir(cls_dir)]
        n = min(cnt_per_label[label], len(cls_paths))
        paths[label] = np.random.choice(cls_paths, size=n, replace=False)
        np.random.shuffle(paths[label])
    return paths


def label_to_index(class_index_path, labels):
    """assets/imagenet_class_index.json maps "idx" -> [wnid, name] (:86-93); without the file: position in labels.txt."""
    if class_index_path and os.path.exists(class_index_path):
        with open(class_index_path, "r") as f:
            idx = json.load(f)
        return {val[0]: k for k, val in idx.items()}
    return {label: str(i) for i, label in enumerate(labels)}


def load_image(path, size):
    """transforms.Compose([Resize(size), RandomCrop((size, size)), ToTensor(), 2x - 1]) (:101-108)."""
    from PIL import Image
    from torchvision import transforms
    img = Image.open(path)
    if not img.mode == "RGB":
        img = img.convert("RGB")
    t = transforms.Compose([transforms.Resize(size), transforms.RandomCrop((size, size)), transforms.ToTensor()])
    return 2.0 * t(img) - 1.0


def run(model, imagenet_root, outdir, size, batch_size=16, classes=None, max_per_class=None, class_index_path=None,
        split_512_path=None, log=print):
    """model: anything with images_to_codes(float[B,3,size,size] in [-1,1] on model.device) -> int64[B, T]."""
    import torch
    from .evaluate import to_uint8
    from PIL import Image
    labels = load_labels(imagenet_root)
    paths = select_paths(imagenet_root, labels, counts_per_label(labels, size, split_512_path))
    l2i = label_to_index(class_index_path, labels)
    os.makedirs(os.path.join(outdir, "codes"), exist_ok=True)
    os.makedirs(os.path.join(outdir, "images"), exist_ok=True)
    device = getattr(model, "device", "cpu")
    n_done = 0
    for label, curr in paths.items():
        conditioning = l2i[label]
        if classes is not None and int(conditioning) not in classes:
            continue
        todo = [(count, p) for count, p in enumerate(curr) if max_per_class is None or count < max_per_class]
        for i in range(0, len(todo), batch_size):
            chunk = todo[i:i + batch_size]
            imgs = torch.stack([load_image(p, size) for _, p in chunk]).to(device)
            codes = model.images_to_codes(imgs).cpu().numpy()
            u8 = to_uint8(imgs).permute(0, 2, 3, 1).cpu().numpy()
            for j, (count, _) in enumerate(chunk):
                Image.fromarray(u8[j]).save(os.path.join(outdir, "images", f"{conditioning}:{count:04}.png"))
                np.save(os.path.join(outdir, "codes", f"{conditioning}:{count:04}.npy"), codes[j].reshape(-1))
                n_done += 1
        log(f"class {conditioning} ({label}): {len(todo)} images")
    return n_done


def main(argv=None):
    import torch
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", type=str, choices=["taming", "chameleon7b", "rar"], required=True)
    ap.add_argument("--modelpath", type=str, default=None)
    ap.add_argument("--imagenet_root", type=str, required=True)
    ap.add_argument("--outdir", type=str, required=True)
    ap.add_argument("--batch_size", type=int, default=16)
    ap.add_argument("--classes", type=str, default=None, help="comma separated class indices (reference: 0,999)")
    ap.add_argument("--max_per_class", type=int, default=None, help="reference: 2")
    ap.add_argument("--class_index", type=str, default=os.path.join("assets", "imagenet_class_index.json"))
    ap.add_argument("--split_512", type=str, default=os.path.join("assets", "imagenet_512_split_50k.txt"))
    args = ap.parse_args(argv)
    random.seed(1)
    np.random.seed(1)
    torch.manual_seed(1)
    torch.cuda.manual_seed_all(1)
    size = 512 if args.model == "chameleon7b" else 256
    from .models import RarARMMWrapper, TamingARMMWrapper
    if args.model == "taming":
        model = TamingARMMWrapper(args.modelpath, max_batch=args.batch_size) if args.modelpath else \
            TamingARMMWrapper(gpt_cfg=dict(vocab_size=16384, block_size=256, n_layer=1, n_head=24, n_embd=1536),
                              max_batch=args.batch_size)     # tokenizer only: a one-layer stand-in transformer
    elif args.model == "rar":
        model = RarARMMWrapper(args.modelpath, max_batch=min(args.batch_size, 8)) if args.modelpath else \
            RarARMMWrapper(rar_cfg=dict(num_hidden_layers=1), max_batch=min(args.batch_size, 8))
    else:
        from .models.chameleon_wrapper import ChameleonARMMWrapper
        model = ChameleonARMMWrapper(model_cfg=dict(n_layers=1), max_batch=min(args.batch_size, 8))
    classes = None if args.classes is None else {int(c) for c in args.classes.split(",")}
    bs = args.batch_size if args.model == "taming" else min(args.batch_size, 8)
    n = run(model, args.imagenet_root, args.outdir, size, bs, classes, args.max_per_class, args.class_index, args.split_512)
    print(f"{n} images tokenised into {args.outdir}")


if __name__ == "__main__":
    main()
