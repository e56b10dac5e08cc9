"""`DDPM` LightningModule-style wrapper and `EMA`, mirroring reference
models/generative/diffusion/ddpm.py:949-1094.

* `DDPM(**config["model"]["args"])` takes the reference's constructor arguments unchanged
  (configs/diffusion/*.json) plus keyword-only B200 knobs (precision, objective, ...).
* `training_step(batch)` / `validation_step(batch)` return the scalar loss (with grad) exactly like the
  reference; `configure_optimizers()` returns an Adam with the reference's hyper-parameters (the fused
  flat-arena implementation); `on_train_batch_end` drives the EMA.
* pytorch_lightning is optional: if it is importable the class derives from LightningModule,
  otherwise from nn.Module with the few hooks the reference uses (`log`, `save_hyperparameters`).
* `EMA` restates the behaviour of ema_pytorch.EMA that the reference relies on (ddpm.py:998,1014,1048).
  ema_pytorch is an unpinned, un-vendored third-party dependency (environments/requirements.txt:20):
  PARITY UNPINNED — behaviour restated from the package's documented algorithm:
  deep copy at construction; `update()` every `update_every` steps; plain copy until
  `update_after_step` (100); afterwards ema <- lerp(ema, online, 1 - decay) with
  decay = clamp(1 - (1 + step)^(-2/3), min 0, max beta).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from .diffusion import GaussianDiffusion
from .optim import FusedAdam
from .unet import Unet

try:  # optional
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    _HAS_PL = True
except Exception:  # pragma: no cover - not installed in the build image
    _Base = nn.Module
    _HAS_PL = False


class EMA(nn.Module):
    def __init__(self, model: GaussianDiffusion, beta=0.9999, update_after_step=100, update_every=10,
                 inv_gamma=1.0, power=2 / 3, min_value=0.0, ema_model: Optional[GaussianDiffusion] = None):
        super().__init__()
        self.beta, self.update_after_step, self.update_every = beta, update_after_step, update_every
        self.inv_gamma, self.power, self.min_value = inv_gamma, power, min_value
        self.online_model = model
        if ema_model is None:
            ema_model = _clone_diffusion(model)
        self.ema_model = ema_model
        self.ema_model.requires_grad_(False)
        self.register_buffer("initted", torch.tensor(False))
        self.register_buffer("step", torch.tensor(0))
        # host mirrors of the two buffers (update() must not synchronise with the device every step); they are
        # re-read from the buffers whenever a state dict is loaded (checkpoint resume)
        self._step_host, self._initted_host = 0, False

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)
        self._step_host = int(self.step.item())
        self._initted_host = bool(self.initted.item())

    @property
    def model(self):
        return self.online_model

    def copy_params_from_model_to_ema(self):
        src, _ = self.online_model.model.flat_parameters()
        dst, _ = self.ema_model.model.flat_parameters()
        dst.copy_(src)

    def get_current_decay(self):
        epoch = max(self._step_host - self.update_after_step - 1, 0)
        if epoch <= 0:
            return 0.0
        value = 1 - (1 + epoch / self.inv_gamma) ** -self.power
        return min(max(value, self.min_value), self.beta)

    @torch.no_grad()
    def update(self):
        step = self._step_host
        self._step_host += 1
        self.step += 1
        if step % self.update_every != 0:
            return
        if step <= self.update_after_step:
            self.copy_params_from_model_to_ema()
            return
        if not self._initted_host:
            self.copy_params_from_model_to_ema()
            self._initted_host = True
            self.initted.fill_(True)
        src, _ = self.online_model.model.flat_parameters()
        dst, _ = self.ema_model.model.flat_parameters()
        L.call("b200dm_ema_update", dst.data_ptr(), src.data_ptr(), dst.numel(), self.get_current_decay())
        self.ema_model.model.arena.touch()      # EMA weight pack refresh

    def forward(self, *a, **k):
        return self.ema_model(*a, **k)


def _clone_diffusion(gd: GaussianDiffusion) -> GaussianDiffusion:
    u = gd.model
    unet = Unet(u.dim, channels=u.channels, self_condition=u.self_condition, precision=u.precision,
                device=u._device, use_tc=u._use_tc, cuda_graph=u._cuda_graph)
    unet.arena.flat.copy_(u.arena.flat)
    return GaussianDiffusion(unet, **gd._ctor)          # every constructor argument of the online model


class DDPM(_Base):
    """ddpm.py:949-1094.  Extra keyword-only arguments select B200 execution options; the positional /
    JSON arguments are the reference's."""

    def __init__(self, img_channels: int = 3, img_size: int = 64, dim: int = 64,
                 diffusion_timesteps: int = 1000, sampling_timesteps: Optional[int] = None,
                 lr: float = 2e-5, betas: Tuple[float, float] = (0.9, 0.99), ema_update_every: int = 10,
                 ema_decay: float = 0.995, *, precision: str = "bf16", objective: str = "pred_v",
                 beta_schedule: str = "sigmoid", device=None, rng: str = "philox",
                 overlap_optimizer: bool = False):
        super().__init__()
        if _HAS_PL:
            self.save_hyperparameters()
        self.hparams_ = SimpleNamespace(img_channels=img_channels, img_size=img_size, dim=dim,
                                        diffusion_timesteps=diffusion_timesteps,
                                        sampling_timesteps=sampling_timesteps, lr=lr, betas=tuple(betas),
                                        ema_update_every=ema_update_every, ema_decay=ema_decay)
        model = Unet(dim=dim, channels=img_channels, precision=precision, device=device)
        diffusion_model = GaussianDiffusion(model, img_size=img_size, timesteps=diffusion_timesteps,
                                            sampling_timesteps=sampling_timesteps, objective=objective,
                                            beta_schedule=beta_schedule, rng=rng)
        self.channels = img_channels
        self.img_size = img_size
        self.ema = EMA(diffusion_model, beta=ema_decay, update_every=ema_update_every)
        self._step_count = 0
        self._overlap_optimizer = bool(overlap_optimizer)     # FusedAdam(overlap_with_backward=...)
        self.logged = {}

    # Lightning provides these; minimal stand-ins otherwise
    if not _HAS_PL:
        def log(self, name, value, sync_dist=False, **kw):
            """`sync_dist=True` (ddpm.py:1017-1023) averages the logged scalar over the ranks.  The all-reduce is
            enqueued on the gradient communication stream, i.e. it rides behind the gradient buckets instead of
            stalling the compute stream; `logged[name]` is a device tensor the caller reads when it wants to."""
            import torch.distributed as dist
            if sync_dist and isinstance(value, torch.Tensor) and dist.is_available() and dist.is_initialized() \
                    and dist.get_world_size() > 1:
                v = value.detach().clone()
                sync = self.ema.model.model.grad_sync
                stream = sync.stream if (sync is not None and sync.stream is not None) else None
                if stream is not None:
                    ev = torch.cuda.Event()
                    ev.record()
                    with torch.cuda.stream(stream):
                        stream.wait_event(ev)
                        dist.all_reduce(v)
                        v.div_(dist.get_world_size())
                    v.record_stream(stream)
                else:
                    dist.all_reduce(v)
                    v.div_(dist.get_world_size())
                value = v
            self.logged[name] = value

        @property
        def global_step(self):
            return self._step_count

    def _common_step(self, batch, mode: str) -> torch.Tensor:
        assert mode in ["train", "val", "test"], f"Invalid mode: {mode}"
        data, _ = batch
        model = self.ema.model if self.training else self.ema.ema_model
        loss = model(data)
        self.log(f"{mode}_loss", loss, prog_bar=True, logger=True,
                 sync_dist=torch.cuda.device_count() > 1)
        return loss

    @torch.inference_mode()
    def sample(self, batch_size: int = 64):
        """What the reference's `_log_sample` (ddpm.py:1029-1042) computes before handing the grid to W&B."""
        self.ema.ema_model.eval()
        return self.ema.ema_model.sample(batch_size=batch_size)

    def training_step(self, batch) -> torch.Tensor:
        return self._common_step(batch, "train")

    def on_train_batch_end(self, outputs=None, batch=None, batch_idx=None):
        self._step_count += 1
        self.ema.update()

    def validation_step(self, batch) -> torch.Tensor:
        return self._common_step(batch, "val")

    # ---- checkpoints in the layout Lightning writes for the reference module (train.py:113-117,137-141) -------------
    def checkpoint(self, optimizer=None, epoch: int = 0) -> dict:
        """{"state_dict": ema.online_model.* / ema.ema_model.* / ema.initted / ema.step, "optimizer_states": [Adam in
        torch.optim format], "global_step", "epoch", "hyper_parameters"} — the keys of a Lightning `.ckpt` of the
        reference DDPM (ema_pytorch naming)."""
        h = self.hparams_
        ck = {"state_dict": self.state_dict(), "global_step": self.global_step, "epoch": epoch,
              "hyper_parameters": dict(vars(h)), "pytorch-lightning_version": "b200dm"}
        if optimizer is not None:
            ck["optimizer_states"] = [optimizer.state_dict()]
        return ck

    def save_checkpoint(self, path: str, optimizer=None, epoch: int = 0):
        torch.save(self.checkpoint(optimizer, epoch), path)

    def load_checkpoint(self, path_or_dict, optimizer=None, strict: bool = True):
        """Load a checkpoint written by `save_checkpoint` or by Lightning for the reference module."""
        ck = path_or_dict if isinstance(path_or_dict, dict) else torch.load(path_or_dict, map_location="cpu",
                                                                            weights_only=False)
        sd = ck["state_dict"] if "state_dict" in ck else ck
        out = self.load_state_dict(sd, strict=strict)
        self.ema.online_model.model.arena.touch()
        self.ema.ema_model.model.arena.touch()
        if not _HAS_PL:
            self._step_count = int(ck.get("global_step", self._step_count))
        if optimizer is not None and ck.get("optimizer_states"):
            optimizer.load_state_dict(ck["optimizer_states"][0])
        return out

    def configure_optimizers(self):
        h = self.hparams_
        opt = FusedAdam(self.ema.model.model, lr=h.lr, betas=h.betas, overlap_with_backward=self._overlap_optimizer)
        # the reference relies on Lightning's DDPStrategy when several GPUs are used
        # (utils/lightning_utils.py:41-43): same semantics here when torch.distributed is initialised
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            unet = self.ema.model.model
            sync = unet.grad_sync or unet.enable_data_parallel()
            self.ema.copy_params_from_model_to_ema()
            opt.grad_scale = sync.grad_scale
        return opt
