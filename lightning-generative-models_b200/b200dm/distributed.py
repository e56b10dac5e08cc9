"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch via torch.distributed.

The reference delegates this to Lightning's DDPStrategy (reference utils/lightning_utils.py:41-43):
DistributedDataParallel averages the gradients of all ranks with bucketed all-reduces overlapped with
backward, and rank 0's initial weights are broadcast at construction.  Here the gradients already live
in ONE flat fp32 arena, so synchronisation is a handful of all-reduces over contiguous slices:

  * `buckets()` cuts the arena into regions in the order the backward pass finishes them
    (final block -> up path -> middle -> down path level 3 .. 0 -> stem + time embedding), each a contiguous slice;
  * `GradSync.reduce_bucket(i)` enqueues the all-reduce of bucket i on a side stream as soon as the
    compute stream has produced it (event), so communication overlaps the rest of backward;
  * `GradSync.finish()` makes the compute stream wait for the collectives before the optimiser step.

Averaging is folded into the optimiser (`FusedAdam.grad_scale = 1 / world_size`), so the wire carries
plain sums.  Sampling needs no communication at all (`GaussianDiffusion.sample_shard`).
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


# CTAs NCCL may use for the gradient all-reduce = SMs the backward kernels leave free while it runs.  The step needs
# only ~60 GB/s of all-reduce bandwidth to hide 143 MB behind a 2.8 ms backward pass; every NCCL CTA costs a whole SM,
# because a conv CTA (>= 200 KiB of shared memory) cannot share one with it.
COMM_CTAS = int(os.environ.get("B200DM_COMM_CTAS", "8"))


def nccl_options():
    """ProcessGroupNCCL options that cap the collective at COMM_CTAS CTAs (None if this torch cannot express it)."""
    try:
        opts = dist.ProcessGroupNCCL.Options()
        opts.config.max_ctas = COMM_CTAS
        opts.config.min_ctas = min(COMM_CTAS, 4)
        return opts
    except Exception:                                   # noqa: BLE001
        os.environ.setdefault("NCCL_MAX_CTAS", str(COMM_CTAS))
        return None


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
            opts = nccl_options()
            if opts is not None:
                kw["pg_options"] = opts
        dist.init_process_group(backend, **kw)
    return rank, local, world


# Backward-completion order of the gradient regions (= Plan.bwd_segments): the down path is cut per level so that only
# the gradients of the last (level-0) blocks and of the arena head remain to be reduced when backward ends — the big
# level-3 / level-2 tensors of the down path are on the wire while the slow level-0 layers are still computing.
# "film" = the rows of the FiLM projection that belong to every block but the two level-0 blocks of the down path:
# their gradient GEMM runs (second stream) as soon as level 1 of the down path is done.
REGIONS = ("final", "ups", "mid", "downs.3", "downs.2", "downs.1", "film", "downs.0", "head")


def buckets(arena) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) element ranges of the flat arena, listed in backward-completion order (REGIONS)."""
    names = [nm for nm, _ in arena.spec]

    def span(pred):
        offs = [(arena.offset[n], arena.offset[n] + arena._numel(n)) for n in names if pred(n)]
        return min(o[0] for o in offs), max(o[1] for o in offs)

    def region(prefix):
        return span(lambda n: n.startswith(prefix) and ".mlp.1." not in n)

    named = {"final": region("final_"), "ups": region("ups."), "mid": region("mid_")}
    for i in range(4):
        named[f"downs.{i}"] = region(f"downs.{i}.")
    # arena head: [FiLM weight rows of all but the level-0 down blocks | level-0 rows, FiLM biases, stem, time MLP];
    # the second part is finished last
    cut = arena.film_early_cols * arena.time_dim
    named["film"] = (0, cut)
    named["head"] = (cut, span(lambda n: n.startswith("init_conv") or n.startswith("time_mlp"))[1])
    # make the regions tile the arena exactly (alignment padding goes to the following region)
    by_mem = sorted(named.items(), key=lambda kv: kv[1][0])
    tiled, prev = {}, 0
    for k, (b, e) in by_mem:
        tiled[k] = (prev, e)
        prev = e
    last = by_mem[-1][0]
    tiled[last] = (tiled[last][0], arena.numel)
    return [tiled[k] for k in REGIONS]


class GradSync:
    """Bucketed, overlapped gradient all-reduce over the flat gradient arena."""

    def __init__(self, arena, group=None, wire: str | None = None):
        """wire: "fp32" (default; what DistributedDataParallel does for the reference) or "bf16" (opt-in, also
        B200DM_GRAD_WIRE=bf16): every bucket is rounded to bf16 for the all-reduce — half the bytes on NVLink — and
        widened back into the fp32 arena, like torch's bf16_compress_hook."""
        self.arena, self.group = arena, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = buckets(arena)
        self.cuda = arena.gflat.is_cuda
        # high priority: the collective's CTAs are dispatched ahead of pending compute blocks whenever an SM frees up
        # (the conv kernels are persistent one-CTA-per-SM grids launched back to back, which otherwise starve NCCL)
        self.stream = torch.cuda.Stream(priority=-1) if (self.cuda and self.world > 1) else None
        self.handles = []
        # SMs the backward kernels leave to the collective while it is in flight (b200dm_set_reserved_sms)
        self.reserved_sms = COMM_CTAS if (self.cuda and self.world > 1) else 0
        self.wire = wire or os.environ.get("B200DM_GRAD_WIRE", "fp32")
        assert self.wire in ("fp32", "bf16")
        self._wire_buf = None
        if self.wire == "bf16" and self.cuda and self.world > 1:
            self._wire_buf = torch.empty(arena.numel, dtype=torch.bfloat16, device=arena.gflat.device)

    def reduce_bucket(self, i: int):
        """Call once bucket i's gradients have been enqueued on the current stream."""
        if self.world == 1:
            return
        b, e = self.buckets[i]
        g = self.arena.gflat[b:e]
        if self.stream is None:
            dist.all_reduce(g, group=self.group)
            return
        ev = torch.cuda.Event()
        ev.record()
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ev)
            if self._wire_buf is not None and (e - b) % 4 == 0:
                from . import _lib as L
                w = self._wire_buf[b:e]
                L.call("b200dm_cast_f32_bf16", g.data_ptr(), w.data_ptr(), e - b)
                dist.all_reduce(w, group=self.group)
                L.call("b200dm_cast_bf16_f32", w.data_ptr(), g.data_ptr(), e - b)
            else:
                dist.all_reduce(g, group=self.group)

    def reduce_all(self):
        for i in range(len(self.buckets)):
            self.reduce_bucket(i)
        self.finish()

    def finish(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world


def broadcast_parameters(arena, src: int = 0, group=None):
    """DDP construction semantics: every rank starts from rank `src`'s weights."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(arena.flat, src, group=group)
        arena.touch()           # weight packs refresh
