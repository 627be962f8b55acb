"""The evaluation round trip of the reference's generate.py kept on the device:

    fill_batch_log                          generate.py:111-164   decode -> [round trips | augment] -> re-encode
    compute_metric("pvalue" | "l0" | "psnr") wmar/utils/metrics.py:19-45
    compute_metrics_and_save_from_batch_log generate.py:37-108    (the metrics part; files are written by wmar_b200.generate)

The reference copies every intermediate tensor to the host (`.cpu().numpy()`), calls `watermarker.detect` once per
image and converts images to PIL for the PSNR.  Here the log holds device tensors, the detector scores a whole
[B, L] batch per transform / parameter in one launch and the metrics are batched tensor expressions; `to_numpy=True`
reproduces the reference's log layout (tuples of numpy arrays) for callers that need it.
"""
import torch

from .augmentations import default_augmentations


@torch.no_grad()
def fill_batch_log(batch_log, key, model, codes, eval_params, to_numpy=False):
    """batch_log[key][transform] = [(param, codes [B,L], imgs [B,3,S,S] in [-1,1], None), ...]   (generate.py:111-164)"""
    conv = (lambda t: t.cpu().numpy()) if to_numpy else (lambda t: t)
    imgs = model.codes_to_images(codes)
    log = batch_log[key] = {}
    log["roundtrips"] = [(0, conv(codes), conv(imgs), None)]
    curr = imgs
    for T in range(1, eval_params.get("max_roundtrips", 1) + 1):
        curr_codes = model.images_to_codes(curr)
        curr = model.codes_to_images(curr_codes)
        log["roundtrips"].append((T, conv(curr_codes), conv(curr), None))
    for aug_name, aug_fn, aug_params in eval_params.get("augmentations", default_augmentations()):
        log[aug_name] = []
        for p in aug_params:
            x01 = imgs / 2.0 + 0.5                                   # augmentations expect [0, 1]
            aug = aug_fn(x01, p).clamp(0, 1) * 2.0 - 1.0
            aug_codes = model.images_to_codes(aug)
            log[aug_name].append((p, conv(aug_codes), conv(aug), None))
    return batch_log


def to_uint8(x):
    """wmar/utils/utils.py:69-80 chw_to_pillow on a batch: 255 * (x + 1) / 2 in fp32, clip to [0, 255], round half to even."""
    return torch.round((255 * ((x.float() + 1.0) / 2.0)).clamp(0, 255)).to(torch.uint8)


def psnr_uint8(a, b):
    """compute_psnr (metrics.py:19-21) on the 8-bit images the reference converts to (chw_to_pillow); a, b in [-1,1]."""
    qa, qb = to_uint8(a).double(), to_uint8(b).double()
    mse = (qa - qb).pow(2).flatten(1).mean(dim=1)
    return 10.0 * torch.log10(255.0 ** 2 / mse)


@torch.no_grad()
def compute_metrics(batch_log, key, watermarker, metric_names=("pvalue", "l0", "psnr")):
    """{transform: [(param, {metric: tensor[B]})]} for one method of the log (metrics.py:19-45, batched)."""
    log = batch_log[key]
    as_t = lambda x, dev: x if isinstance(x, torch.Tensor) else torch.as_tensor(x, device=dev)
    dev = watermarker.device if watermarker is not None else None
    orig_codes = as_t(log["roundtrips"][0][1], dev)
    orig_imgs = as_t(log["roundtrips"][0][2], dev)
    out = {}
    for transform, entries in log.items():
        out[transform] = []
        for param, codes, imgs, _ in entries:
            codes, imgs = as_t(codes, dev), as_t(imgs, dev)
            m = {}
            for name in metric_names:
                if name == "l0":
                    m[name] = (orig_codes != codes).sum(dim=1).double() / orig_codes.shape[1]
                elif name == "psnr":
                    m[name] = psnr_uint8(imgs, orig_imgs)
                elif name == "pvalue":
                    m[name] = watermarker.detect(codes) if watermarker is not None else None
                else:
                    raise ValueError(f"Metric {name} not found")
            out[transform].append((param, m))
    return out
