"""b200dm — B200-native (sm_100a) DDPM/DDIM hot path behind the reference's Python surface.

Public classes mirror models/generative/diffusion/ddpm.py of the reference:
Unet, GaussianDiffusion, DDPM, ModelPrediction.  All arithmetic runs in libb200dm.so.
"""
from . import _lib  # noqa: F401
from ._lib import B200dmError  # noqa: F401

__all__ = ["B200dmError", "Unet", "GaussianDiffusion", "DDPM", "ModelPrediction", "EMA"]


def __getattr__(name):
    # lazy: the model classes import torch.nn machinery; the C-ABI loader alone must stay light
    if name in ("Unet",):
        from .unet import Unet
        return Unet
    if name in ("GaussianDiffusion", "ModelPrediction"):
        from . import diffusion
        return getattr(diffusion, name)
    if name in ("DDPM", "EMA"):
        from . import ddpm
        return getattr(ddpm, name)
    raise AttributeError(name)
