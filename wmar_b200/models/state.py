"""nn.Module trees built from flat state dicts.

The reference hands real nn.Modules across its wrapper boundary (``get_image_tokenizer().encoder`` is patched by
``update_weights`` -- generate.py:327-332, wmar/utils/utils.py:47-66).  ``StateModule`` keeps that contract: it is
an nn.Module whose ``state_dict()`` keys equal the reference module's, so delta checkpoints apply unchanged, while
the arithmetic is done by the CUDA engines that borrow the parameter storage after ``sync_weights()``.
"""
import torch
import torch.nn as nn


class StateModule(nn.Module):
    def __init__(self, flat=None):
        super().__init__()
        if flat:
            for key, value in flat.items():
                self._insert(key.split("."), value)

    def _insert(self, parts, value):
        if len(parts) == 1:
            self.register_parameter(parts[0], nn.Parameter(value.detach().float(), requires_grad=False))
            return
        head = parts[0]
        if head not in self._modules:
            self.add_module(head, StateModule())
        self._modules[head]._insert(parts[1:], value)

    def flat(self, prefix=""):
        return {prefix + k: v for k, v in self.state_dict().items()}


def split_prefix(flat, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in flat.items() if k.startswith(prefix)}


def load_ids(path):
    """armm_wrapper.py:46-50"""
    ids = []
    with open(path, "r") as f:
        for line in f:
            ids.extend(int(t) for t in line.split(",") if t.strip())
    return ids


def update_weights(model, ckpt_path, delta=True):
    """wmar/utils/utils.py:47-66 -- additive delta patches (or a plain state dict) applied to a module tree."""
    state_dict = torch.load(ckpt_path, map_location="cpu", weights_only=False)
    if "state_dict" in state_dict:
        state_dict = state_dict["state_dict"]
    if delta:
        to_apply = dict(model.state_dict())
        for key in state_dict:
            if key in to_apply:
                to_apply[key] = to_apply[key] + state_dict[key].to(to_apply[key].device)
            else:
                to_apply[key] = state_dict[key]
    else:
        to_apply = state_dict
    return model.load_state_dict(to_apply, strict=False)
