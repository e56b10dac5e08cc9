"""`GaussianDiffusion` — drop-in for reference models/generative/diffusion/ddpm.py:532-946.

Same constructor keywords, buffers, and methods (`forward`, `p_losses`, `q_sample`,
`model_predictions`, `p_sample`, `p_sample_loop`, `ddim_sample`, `sample`, `predict_*`, `q_posterior`);
the elementwise math runs in the fused kernels of libb200dm (q_sample+normalize+Philox, loss+grad,
DDIM / DDPM step) and the network in `b200dm.Unet`.

Differences that do not change results: the per-step `imgs.append(img.cpu())` device->host copies of
the reference samplers (ddpm.py:775,829) are only made when `return_all_timesteps=True`; noise comes
from an on-device Philox4x32-10 stream keyed by (torch seed, call counter, global element index)
unless `rng="torch"` asks for torch's own generator (then `t`/noise/initial images consume
torch's global RNG exactly in the reference's order: randint, randn_like / randn, randn_like per step).
"""
from __future__ import annotations

from collections import namedtuple
from random import random
from typing import Optional

import torch
from torch import nn

from . import _lib as L
from .schedule import ddim_time_pairs, make_buffers

ModelPrediction = namedtuple("ModelPrediction", ["pred_noise", "pred_x_start"])


def _extract(a, t, ndim):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (ndim - 1)))


class _LossFn(torch.autograd.Function):
    """target + MSE + loss-weight + mean in one kernel that also emits dL/d(model_out).  `desc` is the noise
    descriptor q_sample ran with: the kernel re-derives x0 and eps from it (Philox block or injected tensor)."""

    @staticmethod
    def forward(ctx, model_out, gd, desc, keep):
        import ctypes as C
        acc = torch.zeros(1, device=model_out.device)
        d_out = torch.empty_like(model_out)
        L.call("b200dm_loss_fwd_bwd", C.byref(desc), model_out.data_ptr(), gd.loss_weight.data_ptr(), acc.data_ptr(),
               d_out.data_ptr(), L.OBJECTIVES[gd.objective])
        ctx.save_for_backward(d_out)
        return acc.reshape(())

    @staticmethod
    def backward(ctx, g):
        (d_out,) = ctx.saved_tensors
        return d_out * g, None, None, None


class GaussianDiffusion(nn.Module):
    def __init__(self, model, *, img_size, timesteps=1000, sampling_timesteps=None, objective="pred_v",
                 beta_schedule="sigmoid", schedule_fn_kwargs=dict(), ddim_sampling_eta=0.0,
                 auto_normalize=True, offset_noise_strength=0.0, min_snr_loss_weight=False,
                 min_snr_gamma=5, rng: str = "philox"):
        super().__init__()
        assert not (type(self) == GaussianDiffusion and model.channels != model.out_dim)
        assert not model.random_or_learned_sinusoidal_cond
        self.model = model
        self.channels = model.channels
        self.self_condition = model.self_condition
        self.img_size = img_size
        self.objective = objective
        assert objective in {"pred_noise", "pred_x0", "pred_v"}, \
            "objective must be either pred_noise (predict noise) or pred_x0 (predict image start) or pred_v"
        if beta_schedule not in ("linear", "cosine", "sigmoid"):
            raise ValueError(f"unknown beta schedule {beta_schedule}")
        assert rng in ("philox", "torch")
        # every constructor argument, for EMA's copy (b200dm.ddpm._clone_diffusion)
        self._ctor = dict(img_size=img_size, timesteps=timesteps, sampling_timesteps=sampling_timesteps,
                          objective=objective, beta_schedule=beta_schedule,
                          schedule_fn_kwargs=dict(schedule_fn_kwargs), ddim_sampling_eta=ddim_sampling_eta,
                          auto_normalize=auto_normalize, offset_noise_strength=offset_noise_strength,
                          min_snr_loss_weight=min_snr_loss_weight, min_snr_gamma=min_snr_gamma, rng=rng)
        self.rng = rng
        self._beta_schedule = beta_schedule
        tables = make_buffers(timesteps, beta_schedule, objective, min_snr_loss_weight, min_snr_gamma,
                              schedule_fn_kwargs)
        self.num_timesteps = int(timesteps)
        self.sampling_timesteps = timesteps if sampling_timesteps is None else sampling_timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        self.offset_noise_strength = offset_noise_strength
        self.auto_normalize = auto_normalize
        dev = next(model.parameters()).device
        for k, v in tables.items():
            self.register_buffer(k, v.to(dev))
        self._host = {k: v.clone() for k, v in tables.items()}      # host copies: per-step scalars
        self._calls = 0

    # ---- small helpers ----------------------------------------------------------------------------------
    @property
    def device(self):
        return self.betas.device

    def normalize(self, img):
        return img * 2 - 1 if self.auto_normalize else img

    def unnormalize(self, t):
        return (t + 1) * 0.5 if self.auto_normalize else t

    def _next_stream(self):
        """(seed, stream id) of the next Philox stream: reproducible under torch.manual_seed."""
        self._calls += 1
        return torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._calls

    def _coef(self, t: int):
        h = self._host
        return (h["sqrt_alphas_cumprod"][t].item(), h["sqrt_one_minus_alphas_cumprod"][t].item(),
                h["sqrt_recip_alphas_cumprod"][t].item(), h["sqrt_recipm1_alphas_cumprod"][t].item())

    # ---- closed forms kept for API parity (ddpm.py:673-705) ----------------------------------------------
    def predict_start_from_noise(self, x_t, t, noise):
        return (_extract(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t
                - _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim()) * noise)

    def predict_noise_from_start(self, x_t, t, x0):
        return ((_extract(self.sqrt_recip_alphas_cumprod, t, x_t.dim()) * x_t - x0)
                / _extract(self.sqrt_recipm1_alphas_cumprod, t, x_t.dim()))

    def predict_v(self, x_start, t, noise):
        return (_extract(self.sqrt_alphas_cumprod, t, x_start.dim()) * noise
                - _extract(self.sqrt_one_minus_alphas_cumprod, t, x_start.dim()) * x_start)

    def predict_start_from_v(self, x_t, t, v):
        return (_extract(self.sqrt_alphas_cumprod, t, x_t.dim()) * x_t
                - _extract(self.sqrt_one_minus_alphas_cumprod, t, x_t.dim()) * v)

    def q_posterior(self, x_start, x_t, t):
        mean = (_extract(self.posterior_mean_coef1, t, x_t.dim()) * x_start
                + _extract(self.posterior_mean_coef2, t, x_t.dim()) * x_t)
        return (mean, _extract(self.posterior_variance, t, x_t.dim()),
                _extract(self.posterior_log_variance_clipped, t, x_t.dim()))

    # ---- forward noising + loss ---------------------------------------------------------------------------
    def _noise_desc(self, img, t, noise, normalize, offset_strength=0.0):
        """b200dm_noise_desc over `img` (+ the tensors it points to, which the caller keeps alive)."""
        img = img.contiguous().float()
        B, chw = img.shape[0], img[0].numel()
        keep = [img, t]
        seed, sid = (0, 0) if noise is not None else self._next_stream()
        if noise is not None:
            noise = noise.contiguous().float()
            keep.append(noise)
        offset = None
        if offset_strength and offset_strength > 0.0:
            # one normal per (sample, channel), ddpm.py:889-891 (drawn after the noise, like the reference)
            if self.rng == "torch":
                offset = torch.randn(img.shape[:2], device=img.device)
            else:
                offset = torch.empty(B, img.shape[1], device=img.device)
                n4 = (offset.numel() + 3) // 4 * 4
                tmp = torch.empty(n4, device=img.device)
                oseed, osid = self._next_stream()
                L.call("b200dm_randn", tmp.data_ptr(), n4, oseed, osid, 0)
                offset.copy_(tmp[:offset.numel()].view_as(offset))
            keep.append(offset)
        d = L.NoiseDesc(img=img.data_ptr(), t=t.data_ptr(), noise=L.ptr(noise), offset=L.ptr(offset),
                        sqrt_ac=self.sqrt_alphas_cumprod.data_ptr(),
                        sqrt_1mac=self.sqrt_one_minus_alphas_cumprod.data_ptr(),
                        offset_strength=float(offset_strength or 0.0), normalize=1 if normalize else 0, B=B,
                        chw=chw, hw=img.shape[-1] * img.shape[-2], seed=seed, stream_id=sid, elem_offset=0)
        return d, keep

    def _q_sample_kernel(self, img, t, noise, normalize, want_noise=False, want_x0=False, offset_strength=0.0):
        import ctypes as C
        d, keep = self._noise_desc(img, t, noise, normalize, offset_strength)
        x_t = torch.empty_like(keep[0])
        noise_out = torch.empty_like(x_t) if want_noise else None
        x0_out = torch.empty_like(x_t) if want_x0 else None
        L.call("b200dm_q_sample", C.byref(d), x_t.data_ptr(), L.ptr(noise_out), L.ptr(x0_out))
        return x_t, noise_out, x0_out, d, keep

    def q_sample(self, x_start, t, noise=None):
        if noise is None and self.rng == "torch":
            noise = torch.randn_like(x_start)
        return self._q_sample_kernel(x_start, t, noise, False)[0]

    def p_losses(self, x_start, t, noise=None, offset_noise_strength=None, _normalize=False, _self_cond=None):
        """ddpm.py:878-925.  q_sample and the loss are two kernels around the UNet; the loss kernel regenerates eps
        from the same Philox blocks (or reads the injected tensor), so neither eps nor x0 is stored in between.
        `_self_cond` forces the reference's coin flip `random() < 0.5` (:902) for tests."""
        if offset_noise_strength is None:
            offset_noise_strength = self.offset_noise_strength
        if noise is None and self.rng == "torch":
            noise = torch.randn_like(x_start)
        x_t, _, _, desc, keep = self._q_sample_kernel(x_start, t, noise, _normalize,
                                                      offset_strength=offset_noise_strength)
        x_self_cond = None
        if self.self_condition and (random() < 0.5 if _self_cond is None else _self_cond):
            # ddpm.py:901-905: a first, gradient-free evaluation; its x0 estimate is fed back (inference plan, so the
            # training plan's saved activations are those of the second evaluation)
            with torch.no_grad():
                x_self_cond = self.model_predictions(x_t, t).pred_x_start.detach()
        model_out = self.model(x_t, t, x_self_cond)
        return _LossFn.apply(model_out, self, desc, keep)

    def forward(self, img, *args, **kwargs):
        b, c, h, w = img.shape
        assert h == self.img_size and w == self.img_size, f"height and width of image must be {self.img_size}"
        t = torch.randint(0, self.num_timesteps, (b,), device=img.device).long()
        # normalize (ddpm.py:945) is fused into the q_sample kernel
        return self.p_losses(img, t, *args, _normalize=self.auto_normalize, **kwargs)

    # ---- reverse process ---------------------------------------------------------------------------------------
    def model_predictions(self, x, t, x_self_cond=None, clip_x_start=False, rederive_pred_noise=False):
        out = self.model(x, t, x_self_cond)
        clip = (lambda v: v.clamp(-1.0, 1.0)) if clip_x_start else (lambda v: v)
        if self.objective == "pred_noise":
            pred_noise = out
            x_start = clip(self.predict_start_from_noise(x, t, pred_noise))
            if clip_x_start and rederive_pred_noise:
                pred_noise = self.predict_noise_from_start(x, t, x_start)
        elif self.objective == "pred_x0":
            x_start = clip(out)
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        else:
            x_start = clip(self.predict_start_from_v(x, t, out))
            pred_noise = self.predict_noise_from_start(x, t, x_start)
        return ModelPrediction(pred_noise, x_start)

    def p_mean_variance(self, x, t, x_self_cond=None, clip_denoised=True):
        x_start = self.model_predictions(x, t, x_self_cond).pred_x_start
        if clip_denoised:
            x_start = x_start.clamp(-1.0, 1.0)
        mean, var, logvar = self.q_posterior(x_start=x_start, x_t=x, t=t)
        return mean, var, logvar, x_start

    def _ddpm_step(self, x, model_out, t: int, noise, x_out, x0_out, seed, sid, elem_offset=0):
        h = self._host
        std = (0.5 * h["posterior_log_variance_clipped"][t]).exp().item()
        L.call("b200dm_ddpm_step", x.data_ptr(), model_out.data_ptr(), L.ptr(noise), x_out.data_ptr(),
               L.ptr(x0_out), *self._coef(t), h["posterior_mean_coef1"][t].item(),
               h["posterior_mean_coef2"][t].item(), std, 1 if t > 0 else 0, L.OBJECTIVES[self.objective],
               x.numel(), seed, sid, elem_offset)

    @torch.inference_mode()
    def p_sample(self, x, t: int, x_self_cond=None, noise=None):
        b = x.shape[0]
        bt = torch.full((b,), t, device=x.device, dtype=torch.long)
        x = x.contiguous().float()
        out = self.model(x, bt, x_self_cond)
        if noise is None and self.rng == "torch" and t > 0:
            noise = torch.randn_like(x)
        seed, sid = (0, 0) if noise is not None else self._next_stream()
        img, x0 = torch.empty_like(x), torch.empty_like(x)
        self._ddpm_step(x, out, t, noise, img, x0, seed, sid)
        return img, x0

    def _loop_state(self, shape, init):
        """The sampler state lives in the Unet plan's static input buffer: no per-step copies."""
        unet = self.model
        plan = unet._plan(shape[0], shape[-1], training=False)
        unet._pack.refresh()
        if init is None:
            if self.rng == "torch":
                plan.x_in.copy_(torch.randn(shape, device=self.device))
            else:
                seed, sid = self._next_stream()
                L.call("b200dm_randn", plan.x_in.data_ptr(), plan.x_in.numel(), seed, sid, 0)
        else:
            plan.x_in.copy_(init)
        if plan.xc_in is not None:
            plan.xc_in.zero_()           # self-conditioning starts from "no estimate" (x_start = None, ddpm.py:770,803)
        return unet, plan

    @torch.inference_mode()
    def p_sample_loop(self, shape, return_all_timesteps=False, init_noise=None, step_noise=None):
        """ddpm.py:759-780.  `init_noise` / `step_noise(t)` inject the normals the reference would draw."""
        unet, plan = self._loop_state(shape, init_noise)
        imgs = [plan.x_in.clone()] if return_all_timesteps else None
        seed, sid = self._next_stream()
        for t in reversed(range(self.num_timesteps)):
            plan.t_in.fill_(t)
            unet.run_plan_forward(plan)
            z = None
            if t > 0:
                if step_noise is not None:
                    z = step_noise(t).contiguous().float()
                elif self.rng == "torch":
                    z = torch.randn_like(plan.x_in)
            # self-conditioned nets: the step's clamped x0 is the next evaluation's second input (ddpm.py:773-774)
            self._ddpm_step(plan.x_in, plan.out, t, z, plan.x_in, plan.xc_in, seed, sid * 4096 + t)
            if return_all_timesteps:
                imgs.append(plan.x_in.clone())
        ret = plan.x_in.clone() if not return_all_timesteps else torch.stack(imgs, dim=1)
        return self.unnormalize(ret)

    @torch.inference_mode()
    def ddim_sample(self, shape, return_all_timesteps=False, init_noise=None, step_noise=None):
        """ddpm.py:782-834."""
        eta, h = self.ddim_sampling_eta, self._host
        unet, plan = self._loop_state(shape, init_noise)
        imgs = [plan.x_in.clone()] if return_all_timesteps else None
        seed, sid = self._next_stream()
        for time, time_next in ddim_time_pairs(self.num_timesteps, self.sampling_timesteps):
            plan.t_in.fill_(time)
            unet.run_plan_forward(plan)
            last = time_next < 0
            if last:
                san = c = sigma = 0.0
            else:
                alpha, alpha_next = h["alphas_cumprod"][time], h["alphas_cumprod"][time_next]
                sig_t = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                c = (1 - alpha_next - sig_t ** 2).sqrt().item()
                san, sigma = alpha_next.sqrt().item(), float(sig_t)
            z = None
            if not last:
                if step_noise is not None and sigma != 0.0:
                    z = step_noise(time).contiguous().float()
                elif self.rng == "torch":
                    z = torch.randn_like(plan.x_in)          # consumed even when eta == 0 (ddpm.py:825)
            L.call("b200dm_ddim_step", plan.x_in.data_ptr(), plan.out.data_ptr(), L.ptr(z),
                   plan.x_in.data_ptr(), L.ptr(plan.xc_in), *self._coef(time), san, c, sigma, 1 if last else 0,
                   L.OBJECTIVES[self.objective], plan.x_in.numel(), seed, sid * 4096 + max(time, 0), 0)
            if return_all_timesteps:
                imgs.append(plan.x_in.clone())
        ret = plan.x_in.clone() if not return_all_timesteps else torch.stack(imgs, dim=1)
        return self.unnormalize(ret)

    @torch.inference_mode()
    def sample(self, batch_size=16, return_all_timesteps=False, **kw):
        fn = self.p_sample_loop if not self.is_ddim_sampling else self.ddim_sample
        return fn((batch_size, self.channels, self.img_size, self.img_size),
                  return_all_timesteps=return_all_timesteps, **kw)

    @torch.inference_mode()
    def sample_shard(self, global_batch: int, rank: int, world_size: int, seed: int = 0):
        """Batch-sharded sampling with no communication: rank r generates images
        [r*B/W, (r+1)*B/W) of the global batch.  All noise is Philox keyed by the GLOBAL element index,
        so the union over ranks is identical for every world size."""
        assert global_batch % world_size == 0
        b = global_batch // world_size
        shape = (b, self.channels, self.img_size, self.img_size)
        unet = self.model
        plan = unet._plan(b, self.img_size, training=False)
        unet._pack.refresh()
        n = plan.x_in.numel()
        off = rank * n
        L.call("b200dm_randn", plan.x_in.data_ptr(), n, seed, 1, off)
        if plan.xc_in is not None:
            plan.xc_in.zero_()
        h, eta = self._host, self.ddim_sampling_eta
        if self.is_ddim_sampling:
            for time, time_next in ddim_time_pairs(self.num_timesteps, self.sampling_timesteps):
                plan.t_in.fill_(time)
                unet.run_plan_forward(plan)
                last = time_next < 0
                if last:
                    san = c = sigma = 0.0
                else:
                    alpha, alpha_next = h["alphas_cumprod"][time], h["alphas_cumprod"][time_next]
                    sig_t = eta * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
                    c = (1 - alpha_next - sig_t ** 2).sqrt().item()
                    san, sigma = alpha_next.sqrt().item(), float(sig_t)
                L.call("b200dm_ddim_step", plan.x_in.data_ptr(), plan.out.data_ptr(), None,
                       plan.x_in.data_ptr(), L.ptr(plan.xc_in), *self._coef(time), san, c, sigma, 1 if last else 0,
                       L.OBJECTIVES[self.objective], n, seed, 2 + time, off)
        else:
            for t in reversed(range(self.num_timesteps)):
                plan.t_in.fill_(t)
                unet.run_plan_forward(plan)
                self._ddpm_step(plan.x_in, plan.out, t, None, plan.x_in, plan.xc_in, seed, 2 + t, off)
        return self.unnormalize(plan.x_in.clone())

    @torch.inference_mode()
    def interpolate(self, x1, x2, t=None, lam=0.5):
        """ddpm.py:847-867."""
        b = x1.shape[0]
        t = self.num_timesteps - 1 if t is None else t
        assert x1.shape == x2.shape
        tb = torch.full((b,), t, device=x1.device, dtype=torch.long)
        xt1, xt2 = self.q_sample(x1, tb), self.q_sample(x2, tb)
        img = (1 - lam) * xt1 + lam * xt2
        x_start = None
        for i in reversed(range(0, t)):
            self_cond = x_start if self.self_condition else None          # ddpm.py:864-865
            img, x_start = self.p_sample(img, i, self_cond)
        return img
