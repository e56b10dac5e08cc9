"""Import the UNMODIFIED reference diffusion module from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
tests/golden/make_golden.py to generate the committed fixtures and by the optional
`test_oracle_vs_live_reference` test, which skips when the reference is absent.

The reference imports three packages at module top that are not installed here
(pytorch_lightning, ema_pytorch, torchinfo — SURVEY.md §8c).  None of them is used by
Unet / GaussianDiffusion, so inert stubs are injected into sys.modules.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("B200DM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models/generative/diffusion/ddpm.py"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference module `models.generative.diffusion.ddpm`."""
    import torch.nn as nn

    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)

    class _LM(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    class _Strategy:  # placeholder types only
        def __init__(self, *a, **k):
            pass

    pl = _stub("pytorch_lightning", LightningModule=_LM)
    st = _stub("pytorch_lightning.strategies", DDPStrategy=_Strategy,
               SingleDeviceStrategy=_Strategy, Strategy=_Strategy)
    pl.strategies = st

    class _EMA(nn.Module):
        def __init__(self, model, **k):
            super().__init__()
            self.model = model

    _stub("ema_pytorch", EMA=_EMA)
    _stub("torchinfo", summary=lambda *a, **k: None)
    try:
        import wandb  # noqa: F401
    except Exception:
        _stub("wandb", Image=lambda *a, **k: None)

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib

    return importlib.import_module("models.generative.diffusion.ddpm")
