"""Reference checkpoint files: same dict layout as ``main.py:183-189`` / ``main_utils.py:54-64``.

    {'epoch': next start epoch, 'arch': 'HPLFlowNet' | 'HPLFlowNetShallow', 'state_dict': ..., 'min_loss': ...,
     'optimizer': optimizer.state_dict()}

The reference trains under ``torch.nn.DataParallel`` (``main.py:104``), so its ``state_dict`` keys carry a
``module.`` prefix; the B200 models run one process per GPU without a wrapper.  ``load_checkpoint`` accepts both
and loads ``strict=True`` like ``main.py:122``; ``save_checkpoint`` writes the prefix by default so the file opens
in the reference unchanged.
"""
import os
import shutil

import torch

_PREFIX = "module."


def strip_module_prefix(state_dict):
    """Keys of a DataParallel-wrapped model -> keys of the bare model (no-op if there is no prefix)."""
    if state_dict and all(k.startswith(_PREFIX) for k in state_dict):
        return type(state_dict)((k[len(_PREFIX):], v) for k, v in state_dict.items())
    return state_dict


def add_module_prefix(state_dict):
    if state_dict and all(k.startswith(_PREFIX) for k in state_dict):
        return state_dict
    return type(state_dict)((_PREFIX + k, v) for k, v in state_dict.items())


def make_state(model, optimizer, epoch, min_loss, arch="HPLFlowNet", data_parallel_keys=True):
    sd = model.state_dict()
    return {"epoch": epoch + 1, "arch": arch, "state_dict": add_module_prefix(sd) if data_parallel_keys else sd,
            "min_loss": min_loss, "optimizer": optimizer.state_dict() if optimizer is not None else None}


def save_checkpoint(state, is_best, ckpt_dir, filename="checkpoint.pth.tar"):
    """main_utils.py:54-64: always the rolling file, a numbered copy when epoch % 10 == 1, ``model_best`` when best."""
    path = os.path.join(ckpt_dir, filename)
    torch.save(state, path)
    if state["epoch"] % 10 == 1:
        shutil.copyfile(path, os.path.join(ckpt_dir, "checkpoint_" + str(state["epoch"]) + ".pth.tar"))
    if is_best:
        shutil.copyfile(path, os.path.join(ckpt_dir, "model_best.pth.tar"))
    return path


def load_checkpoint(path, model, optimizer=None, map_location="cpu"):
    """main.py:117-142.  Returns the checkpoint dict (``epoch`` = next start epoch, ``min_loss``)."""
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    model.load_state_dict(strip_module_prefix(ckpt["state_dict"]), strict=True)
    if optimizer is not None and ckpt.get("optimizer") is not None:
        optimizer.load_state_dict(ckpt["optimizer"])
    return ckpt
