"""`Unet` — drop-in for reference models/generative/diffusion/ddpm.py:275-471 running on libb200dm.

Same constructor signature, attributes (`channels`, `out_dim`, `self_condition`,
`random_or_learned_sinusoidal_cond`, `downsample_factor`), `forward(x, time, x_self_cond=None)` and
state_dict keys/shapes as the reference.  Only the configuration the reference's configs use is built
(dim=64, dim_mults=(1,2,4,8), no learned variance / learned sinusoidal embedding; `self_condition=True` is built);
anything else raises NotImplementedError instead of silently falling back.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
from torch import nn

from . import _lib as L
from .engine import Plan, WeightPack
from .params import ParamArena

PRECISIONS = {"bf16": L.BF16, "fp32": L.F32}


def _attach(root: nn.Module, dotted: str, p: nn.Parameter):
    """Register `p` under the reference's dotted name by growing a tree of container modules, so that
    named_parameters()/state_dict() yield exactly the reference keys."""
    parts = dotted.split(".")
    m = root
    for part in parts[:-1]:
        if part not in m._modules:
            m.add_module(part, nn.Module())
        m = m._modules[part]
    m.register_parameter(parts[-1], p)


class _UnetFn(torch.autograd.Function):
    """Whole-network autograd node: forward = the plan's launch list, backward = the mirrored list."""

    @staticmethod
    def forward(ctx, anchor, unet, x, time, x_self_cond=None):
        plan = unet._plan(x.shape[0], x.shape[-1], training=True)
        out = unet._run(plan, x, time, x_self_cond)
        ctx.unet, ctx.plan, ctx.ticket = unet, plan, plan.ticket
        return out

    @staticmethod
    def backward(ctx, grad_out):
        unet, plan = ctx.unet, ctx.plan
        if plan.ticket != ctx.ticket:
            raise RuntimeError("b200dm.Unet: backward() after a newer forward() of the same shape overwrote "
                               "the saved activations; call backward before the next forward")
        unet._backward(plan, grad_out)
        return torch.zeros_like(unet._anchor), None, None, None, None


class Unet(nn.Module):
    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3,
                 self_condition=False, resnet_block_groups=8, learned_variance=False,
                 learned_sinusoidal_cond=False, random_fourier_features=False, learned_sinusoidal_dim=16,
                 sinusoidal_pos_emb_theta=10000, attn_dim_head=32, attn_heads=4, full_attn=None,
                 flash_attn=False, *, precision: str = "bf16", device=None, use_tc: Optional[bool] = None,
                 cuda_graph: Optional[bool] = None):
        super().__init__()
        unsupported = dict(
            dim=dim != 64, init_dim=init_dim not in (None, dim), out_dim=out_dim not in (None, channels),
            dim_mults=tuple(dim_mults) != (1, 2, 4, 8), channels=channels not in (1, 2, 3),
            resnet_block_groups=resnet_block_groups != 8,
            learned_variance=bool(learned_variance), learned_sinusoidal_cond=bool(learned_sinusoidal_cond),
            random_fourier_features=bool(random_fourier_features),
            sinusoidal_pos_emb_theta=sinusoidal_pos_emb_theta != 10000, attn_dim_head=attn_dim_head != 32,
            attn_heads=attn_heads != 4, full_attn=full_attn not in (None, (False, False, False, True)),
            flash_attn=bool(flash_attn))
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(
                f"b200dm.Unet builds the configuration used by configs/diffusion/*.json only; "
                f"non-default {bad} is not implemented")
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {list(PRECISIONS)}")
        self.channels, self.out_dim, self.dim = channels, channels, dim
        self.self_condition = bool(self_condition)              # ddpm.py:300-301: the stem takes 2 * channels
        self.random_or_learned_sinusoidal_cond = False
        self.precision = precision
        self.dt = PRECISIONS[precision]
        self._device = torch.device("cuda" if device is None else device)
        if self._device.type != "cuda":
            raise L.B200dmError("b200dm.Unet runs on CUDA (sm_100a) only; there is no CPU fallback")
        L.load()                                      # fail loudly if the extension is missing
        self.arena = ParamArena(dim, channels, self._device, with_grad=True, self_condition=self.self_condition)
        self._init_default()
        self._params: Dict[str, nn.Parameter] = {}
        for nm, _ in self.arena.spec:
            p = nn.Parameter(self.arena.views[nm])
            self._params[nm] = p
            _attach(self, nm, p)
        self._anchor = torch.zeros((), device=self._device, requires_grad=True)
        self._use_tc = use_tc
        env = os.environ.get("B200DM_CUDA_GRAPH")
        self._cuda_graph = (env != "0") if cuda_graph is None else cuda_graph
        self._pack: Optional[WeightPack] = None
        self._plans: Dict[tuple, Plan] = {}
        self.grad_sync = None        # b200dm.distributed.GradSync when training data-parallel
        self.bucket_hook = None      # callable(i, (begin, end)): gradient bucket i is complete (FusedAdam overlap)
        self._buckets = None
        self._sync_enabled = True    # False inside no_sync(): gradient accumulation micro-batches
        self._reduced = False        # a gradient all-reduce has been issued since the last zero_grad()

    # ---- reference surface ----------------------------------------------------------------------------
    @property
    def downsample_factor(self):
        return 8

    def forward(self, x, time, x_self_cond=None):
        assert all(d % self.downsample_factor == 0 for d in x.shape[-2:]), \
            f"your input dimensions {tuple(x.shape[-2:])} need to be divisible by {self.downsample_factor}, given the unet"
        if x_self_cond is not None and not self.self_condition:
            raise ValueError("x_self_cond given to a Unet built with self_condition=False")
        if x_self_cond is not None and x_self_cond.shape != x.shape:
            raise ValueError(f"x_self_cond shape {tuple(x_self_cond.shape)} != x shape {tuple(x.shape)}")
        if x.shape[-1] != x.shape[-2]:
            raise NotImplementedError("b200dm.Unet is built for square inputs")
        if x.shape[1] != self.channels:
            raise ValueError(f"expected {self.channels} input channels, got {x.shape[1]}")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self._params.values()):
            return _UnetFn.apply(self._anchor, self, x, time, x_self_cond)
        plan = self._plan(x.shape[0], x.shape[-1], training=False)
        return self._run(plan, x, time, x_self_cond)

    # ---- parameters -------------------------------------------------------------------------------------
    def _init_default(self):
        """torch's default init of the reference modules (kaiming_uniform(a=sqrt 5) == U(+-1/sqrt(fan_in))
        for conv/linear weights and biases, ones for norm gains, zeros for GN bias, N(0,1) mem_kv)."""
        a = self.arena
        with torch.no_grad():
            for nm, shape in a.spec:
                v = a.views[nm]
                if nm.endswith("mem_kv"):
                    v.copy_(torch.randn(shape, device=self._device))
                elif nm.endswith(".g") or nm.endswith("norm.weight"):
                    v.fill_(1.0)
                elif nm.endswith("norm.bias"):
                    v.zero_()
                else:
                    wshape = a.shapes[nm[:-5] + ".weight"] if nm.endswith(".bias") else shape
                    fan_in = 1
                    for s in wshape[1:]:
                        fan_in *= s
                    bound = 1.0 / fan_in ** 0.5
                    v.copy_(torch.empty(shape, device=self._device).uniform_(-bound, bound))

    def load_reference_state_dict(self, sd, prefix: str = ""):
        """Load a reference `Unet.state_dict()` (OIHW conv weights, [out,in] linears)."""
        self.arena.load(sd, prefix)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys,
                              unexpected_keys, error_msgs):
        # parameters live in nested container modules; nn.Module recursion handles them (copy_ into the
        # arena views keeps the packed layout).  Nothing extra here.
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys,
                                      unexpected_keys, error_msgs)

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        # hand out reference-layout, storage-independent tensors (conv weights are permuted views)
        for k in list(sd.keys()):
            v = sd[k]
            if isinstance(v, torch.Tensor) and not v.is_contiguous():
                sd[k] = v.detach().contiguous()
        return sd

    def flat_parameters(self):
        return self.arena.flat, self.arena.gflat

    def no_sync(self):
        """DistributedDataParallel.no_sync(): backward passes inside the context only accumulate local gradients;
        the first backward outside it all-reduces the accumulated sum (gradient accumulation under data parallel)."""
        import contextlib

        @contextlib.contextmanager
        def ctx():
            prev, self._sync_enabled = self._sync_enabled, False
            try:
                yield
            finally:
                self._sync_enabled = prev
        return ctx()

    def zero_grad(self, set_to_none: bool = True):
        self._reduced = False
        self.arena.gflat.zero_()
        for nm, p in self._params.items():
            p.grad = None if set_to_none else self.arena.gviews[nm]

    # ---- execution ----------------------------------------------------------------------------------------
    def _tc(self) -> bool:
        if self._use_tc is None:
            self._use_tc = bool(L.load().b200dm_tc_available()) and os.environ.get("B200DM_NO_TC") != "1"
        return self._use_tc and self.dt == L.BF16

    def _plan(self, B: int, S: int, training: bool) -> Plan:
        # buffers must be ordinary tensors even when first requested under torch.inference_mode()
        with torch.inference_mode(False):
            return self._plan_impl(B, S, training)

    def _plan_impl(self, B: int, S: int, training: bool) -> Plan:
        if self._pack is None:
            self._pack = WeightPack(self.arena, self.dt, with_dgrad=True)
            self._plans.clear()
        key = (B, S, training)
        plan = self._plans.get(key)
        if plan is None:
            if len(self._plans) >= 4:                    # bound activation memory
                self._plans.pop(next(iter(self._plans)))
            plan = Plan(self.arena, self._pack, B, S, self.dt, training, self._tc())
            plan.ticket = 0
            plan.graph_fwd = plan.graph_bwd = None
            plan.warm = 0
            self._plans[key] = plan
        return plan

    def _run(self, plan: Plan, x, time, x_self_cond=None) -> torch.Tensor:
        self._pack.refresh()
        plan.x_in.copy_(x)
        if plan.xc_in is not None:       # ddpm.py:434: zeros when no estimate is given
            if x_self_cond is None:
                plan.xc_in.zero_()
            else:
                plan.xc_in.copy_(x_self_cond)
        plan.t_in.copy_(time)
        self.run_plan_forward(plan)
        plan.ticket += 1
        return plan.out.clone()

    def run_plan_forward(self, plan: Plan):
        """Forward launches over the plan's static buffers (x_in, t_in -> out); graph-replayed once warm."""
        if self._cuda_graph and plan.graph_fwd is not None:
            plan.graph_fwd.replay()
            return
        plan.run_forward()
        plan.warm += 1
        if self._cuda_graph and plan.warm == 2 and not torch.cuda.is_current_stream_capturing():
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                plan.run_forward()
            plan.graph_fwd = g

    def enable_data_parallel(self, group=None):
        """Overlap the gradient all-reduce with backward (DDP semantics; see b200dm.distributed)."""
        from .distributed import GradSync, broadcast_parameters
        broadcast_parameters(self.arena, 0, group)
        self.grad_sync = GradSync(self.arena, group)
        return self.grad_sync

    def run_plan_backward(self, plan: Plan):
        """Backward launches, one segment per gradient bucket; each bucket's all-reduce is enqueued on the
        communication stream as soon as its segment has been issued."""
        sync = self.grad_sync if self._sync_enabled else None
        if sync is not None and sync.world > 1:
            if self._reduced:
                # the arena already holds an all-reduced sum: adding local gradients and reducing again would count
                # the first micro-batch world_size times
                raise RuntimeError("b200dm.Unet: second backward() after the gradients were all-reduced; wrap the "
                                   "accumulation micro-batches in `with unet.no_sync():` (all but the last one) or "
                                   "call zero_grad() between steps")
            self._reduced = True
        nseg = len(plan.bwd_segments)
        hook = self.bucket_hook
        reserve = sync.reserved_sms if sync is not None else 0
        if reserve:          # the launches below (eager or captured) size their grids for the SMs NCCL leaves free
            L.load().b200dm_set_reserved_sms(reserve)
        if hook is not None and self._buckets is None:
            from .distributed import buckets
            self._buckets = buckets(self.arena)
            assert len(self._buckets) == nseg
        # launch groups: with a collective to overlap every gradient bucket is its own group (its all-reduce starts
        # as soon as the segment is issued); single-GPU runs fuse the per-level segments of the down path and the
        # early FiLM GEMM into one group (fewer graph launches, no break in the programmatic-launch chain)
        from .distributed import REGIONS
        if sync is not None:
            groups = [[i] for i in range(nseg)]
        else:
            mid = [i for i, r in enumerate(REGIONS) if r.startswith("downs.") or r == "film"]
            groups = [[i] for i in range(mid[0])] + [mid] + [[i] for i in range(mid[-1] + 1, nseg)]
        key = "sync" if sync is not None else "solo"
        graphs = plan.graph_bwd.get(key) if isinstance(plan.graph_bwd, dict) else None

        def after(i):
            if sync is not None:
                sync.reduce_bucket(i)
            if hook is not None:
                hook(i, self._buckets[i])

        if self._cuda_graph and graphs is not None:
            for g, grp in zip(graphs, groups):
                g.replay()
                for i in grp:
                    after(i)
        else:
            for grp in groups:
                for i in grp:
                    plan.run_backward_segment(i)
                for i in grp:
                    after(i)
            if self._cuda_graph and plan.warm >= 2 and not torch.cuda.is_current_stream_capturing():
                # capture for the following steps (stream capture records the launches, it does not run them)
                graphs = []
                for grp in groups:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for i in grp:
                            plan.run_backward_segment(i)
                    graphs.append(g)
                if not isinstance(plan.graph_bwd, dict):
                    plan.graph_bwd = {}
                plan.graph_bwd[key] = graphs
        if reserve:
            L.load().b200dm_set_reserved_sms(0)
        if sync is not None:
            sync.finish()

    def _backward(self, plan: Plan, grad_out: torch.Tensor):
        first = next(iter(self._params.values()))
        if first.grad is None:                 # zero_grad(set_to_none=True) happened: start from zero
            self.arena.gflat.zero_()
        plan.d_out.copy_(grad_out)
        self.run_plan_backward(plan)
        gv = self.arena.gviews
        for nm, p in self._params.items():
            if p.grad is None:
                p.grad = gv[nm]
            elif p.grad.data_ptr() != gv[nm].data_ptr():
                raise RuntimeError("b200dm.Unet: parameter .grad tensors must alias the gradient arena; "
                                   "use zero_grad() instead of assigning new .grad tensors")
