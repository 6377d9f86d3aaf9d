"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference modules (when /root/reference is
present, i.e. in the build container, never on the GPU box) through the shims in oracle/shims."""
import importlib
import os
import sys

import torch

REF = os.environ.get("MDC_REFERENCE_DIR", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")


def available():
    return os.path.isfile(os.path.join(REF, "model.py"))


def load(max_len=100, pad_idx=302, bos_idx=300):
    """Returns dict of reference modules with CFG primed the way the working scripts do
    (train_val_epoch.py:26-27, inference_trail_after_good_map.py:130)."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True
    for p in (REF, _SHIMS):
        if p not in sys.path:
            sys.path.insert(0, p)
    af = importlib.import_module("allied_files")
    af.CFG.device = torch.device("cpu")
    af.CFG.max_len = max_len
    af.CFG.pad_idx = pad_idx
    af.CFG.bos_idx = bos_idx
    mods = {"allied_files": af}
    for name in ("utils", "model", "axial_model", "iou_calcualtions", "iou_bbox"):
        mods[name] = importlib.import_module(name)
    return mods
