"""Fused Adam over the flat parameter arena (torch.optim.Adam semantics, reference
models/generative/diffusion/ddpm.py:1053-1059: Adam(lr, betas), eps 1e-8, no weight decay).

One kernel per step instead of 283 x (several) element-wise launches.  It is a torch.optim.Optimizer,
so Lightning-style loops (`optimizer.step()`, `optimizer.zero_grad()`) work unchanged.
"""
from __future__ import annotations

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, unet, lr=2e-5, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0):
        self.unet = unet
        params = [p for p in unet.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        flat, _ = unet.flat_parameters()
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = 0
        self.grad_scale = 1.0          # set to 1/world_size when gradients are summed across ranks

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self.param_groups[0]
        flat, gflat = self.unet.flat_parameters()
        self.step_count += 1
        L.call("b200dm_adam_step", flat.data_ptr(), gflat.data_ptr(), self.exp_avg.data_ptr(),
               self.exp_avg_sq.data_ptr(), flat.numel(), g["lr"], g["betas"][0], g["betas"][1], g["eps"],
               g["weight_decay"], self.step_count, self.grad_scale)
        self.unet.arena.touch()        # the weight pack re-syncs lazily (no kernel needed for the bump)
        return loss

    def zero_grad(self, set_to_none: bool = True):
        self.unet.zero_grad(set_to_none=set_to_none)

    def state_dict(self):
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.step_count = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
