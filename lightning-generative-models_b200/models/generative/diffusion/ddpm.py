"""Import-path shim: the reference's loader resolves `models.generative.diffusion.ddpm.DDPM`
(reference utils/loader.py:28-36).  Putting `lightning-generative-models_b200/` on sys.path makes the
same lookup land on the B200 implementation."""
from b200dm.ddpm import DDPM, EMA  # noqa: F401
from b200dm.diffusion import GaussianDiffusion, ModelPrediction  # noqa: F401
from b200dm.unet import Unet  # noqa: F401
