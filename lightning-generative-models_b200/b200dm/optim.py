"""Fused Adam over the flat parameter arena (torch.optim.Adam semantics, reference
models/generative/diffusion/ddpm.py:1053-1059: Adam(lr, betas), eps 1e-8, no weight decay).

One kernel per step instead of 283 x (several) element-wise launches.  It is a torch.optim.Optimizer,
so Lightning-style loops (`optimizer.step()`, `optimizer.zero_grad()`) work unchanged.

`overlap_with_backward=True` (opt-in; assumes ONE backward per optimiser step, i.e. no gradient accumulation and no
gradient clipping between backward and step): the update of a gradient bucket and the re-pack of its conv weights
are enqueued on a second stream as soon as backward has finished that bucket (after its all-reduce when training
data-parallel), so the HBM-bound optimiser tail overlaps the latency-bound rest of backward; `step()` then only
joins the streams.  Same kernels, same element-wise arithmetic: the trajectory is bit-identical.
"""
from __future__ import annotations

import torch

from . import _lib as L


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, unet, lr=2e-5, betas=(0.9, 0.99), eps=1e-8, weight_decay=0.0,
                 overlap_with_backward: bool = False):
        self.unet = unet
        params = [p for p in unet.parameters() if p.requires_grad]
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay))
        flat, _ = unet.flat_parameters()
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        self.step_count = 0
        self.grad_scale = 1.0          # set to 1/world_size when gradients are summed across ranks
        self.overlap = bool(overlap_with_backward) and flat.is_cuda
        self._done = set()             # gradient buckets already applied during the current backward
        import os
        self.bg_ctas = int(os.environ.get("B200DM_ADAM_BG_CTAS", "4"))
        self._stream = torch.cuda.Stream(device=flat.device) if self.overlap else None
        if self.overlap:
            unet.bucket_hook = self._on_bucket

    def _adam(self, flat, gflat, begin, end, step, ctas_per_sm=16):
        g = self.param_groups[0]
        n = end - begin
        L.call("b200dm_adam_step_bg", flat.data_ptr() + 4 * begin, gflat.data_ptr() + 4 * begin,
               self.exp_avg.data_ptr() + 4 * begin, self.exp_avg_sq.data_ptr() + 4 * begin, n, g["lr"],
               g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], step, self.grad_scale, ctas_per_sm)

    @torch.no_grad()
    def _on_bucket(self, i, rng):
        """Backward has issued everything that writes gradient bucket i (and its all-reduce, if any)."""
        if i in self._done:
            raise RuntimeError("FusedAdam(overlap_with_backward=True) applies the update during backward and "
                               "therefore supports exactly one backward() per step(); use the default mode for "
                               "gradient accumulation")
        unet = self.unet
        flat, gflat = unet.flat_parameters()
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ev)
            if unet.grad_sync is not None and unet.grad_sync.stream is not None:
                self._stream.wait_stream(unet.grad_sync.stream)
            # a bounded grid: the update must share the SMs with backward, not queue in front of it
            self._adam(flat, gflat, rng[0], rng[1], self.step_count + 1, ctas_per_sm=self.bg_ctas)
            if unet._pack is not None:
                unet._pack.refresh_range(rng[0], rng[1])
        self._done.add(i)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        flat, gflat = self.unet.flat_parameters()
        self.step_count += 1
        unet = self.unet
        if self.overlap and unet._buckets is not None and len(self._done) == len(unet._buckets):
            # every bucket was updated (and its convs re-packed) behind backward: join the streams
            torch.cuda.current_stream().wait_stream(self._stream)
            unet.arena.touch()
            if unet._pack is not None:
                unet._pack.version = unet._pack.up_version = unet.arena.version
        else:
            if self._done:             # partially applied (should not happen): finish the remaining buckets
                torch.cuda.current_stream().wait_stream(self._stream)
                for i, rng in enumerate(unet._buckets):
                    if i not in self._done:
                        self._adam(flat, gflat, rng[0], rng[1], self.step_count)
            else:
                self._adam(flat, gflat, 0, flat.numel(), self.step_count)
            unet.arena.touch()         # the weight pack re-syncs lazily (no kernel needed for the bump)
        self._done.clear()
        return loss

    def zero_grad(self, set_to_none: bool = True):
        self.unet.zero_grad(set_to_none=set_to_none)

    # ---- torch.optim.Adam's checkpoint format (what a Lightning checkpoint of the reference holds) --------------------
    _TORCH_ADAM_DEFAULTS = dict(amsgrad=False, maximize=False, foreach=None, capturable=False, differentiable=False,
                                fused=None, decoupled_weight_decay=False)

    def _names(self):
        return [n for n, p in self.unet.named_parameters() if p.requires_grad]

    def state_dict(self):
        """{"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [{..., "params": [0..n-1]}]} with
        per-parameter tensors in the reference (OIHW / [out,in]) layout, cloned out of the flat arenas."""
        arena, names = self.unet.arena, self._names()
        state = {}
        if self.step_count > 0:
            for i, nm in enumerate(names):
                state[i] = {"step": torch.tensor(float(self.step_count)),
                            "exp_avg": arena._logical(self.exp_avg, nm).detach().contiguous().clone(),
                            "exp_avg_sq": arena._logical(self.exp_avg_sq, nm).detach().contiguous().clone()}
        g = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group = dict(self._TORCH_ADAM_DEFAULTS)
        group.update(g)
        group["params"] = list(range(len(names)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        if "state" not in sd:                      # the flat layout of round 1 ("step", "exp_avg", "exp_avg_sq")
            self.step_count = int(sd["step"])
            self.exp_avg.copy_(sd["exp_avg"])
            self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        else:
            arena, names = self.unet.arena, self._names()
            groups = sd["param_groups"]
            ids = [i for g in groups for i in g["params"]]
            if len(ids) != len(names):
                raise ValueError(f"optimizer state holds {len(ids)} parameters, the model has {len(names)}")
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            step = 0
            with torch.no_grad():
                for pos, pid in enumerate(ids):
                    st = sd["state"].get(pid)
                    if st is None:
                        continue
                    nm = names[pos]
                    arena._logical(self.exp_avg, nm).copy_(st["exp_avg"].to(self.exp_avg.device, torch.float32))
                    arena._logical(self.exp_avg_sq, nm).copy_(st["exp_avg_sq"].to(self.exp_avg.device, torch.float32))
                    step = max(step, int(float(st["step"])))
            self.step_count = step
        for g, ng in zip(self.param_groups, sd.get("param_groups", [])):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in ng:
                    g[k] = tuple(ng[k]) if k == "betas" else ng[k]
