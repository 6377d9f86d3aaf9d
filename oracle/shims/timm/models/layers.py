"""Shim: reference model.py:4 does `from timm.models.layers import trunc_normal_`."""
from torch.nn.init import trunc_normal_  # noqa: F401
