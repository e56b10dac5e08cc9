"""Import the UNMODIFIED reference diffusion module (models/generative/diffusion/ddpm.py of
seungjunlee96/lightning-generative-models) for the baseline arms and the golden-fixture generator.

Search order for the reference tree: $B200DM_REFERENCE_ROOT, /root/reference (build container), baseline/_ref
(a git-ignored run-time copy of the handful of reference files the path imports, made by baseline/make_ref.py so
that the reference itself travels to the GPU box with `gpurun`; never committed).

The reference imports three packages at module top that are not installed in this image (pytorch_lightning,
ema_pytorch, torchinfo — SURVEY.md §8c).  None of them is used by Unet / GaussianDiffusion, so inert stubs are
injected into sys.modules.  Baseline / test infrastructure only: nothing under lightning-generative-models_b200/
imports this.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
LOCAL_COPY = os.path.join(HERE, "_ref")
_REL = "models/generative/diffusion/ddpm.py"


def reference_root():
    for root in (os.environ.get("B200DM_REFERENCE_ROOT"), "/root/reference", LOCAL_COPY):
        if root and os.path.isfile(os.path.join(root, _REL)):
            return root
    return None


REF_ROOT = reference_root() or "/root/reference"


def reference_available() -> bool:
    return reference_root() is not None


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

    root = reference_root()
    if root is None:
        raise RuntimeError("reference tree not present (looked at $B200DM_REFERENCE_ROOT, /root/reference, %s)"
                           % LOCAL_COPY)

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

    # the b200 package ships a same-named shim (models.generative.diffusion.ddpm, the loader convention): make sure
    # the reference tree wins for this import and that a previously imported shim is not returned
    for k in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "utils"
              or k.startswith("utils.")]:
        f = getattr(sys.modules[k], "__file__", None) or ""
        if not f.startswith(root):
            del sys.modules[k]
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    return importlib.import_module("models.generative.diffusion.ddpm")
