"""Data-parallel plumbing: one process per GPU, NCCL over NVLink/NVSwitch via torch.distributed.

The reference delegates this to Lightning's DDPStrategy (reference utils/lightning_utils.py:41-43):
DistributedDataParallel averages the gradients of all ranks with bucketed all-reduces overlapped with
backward, and rank 0's initial weights are broadcast at construction.  Here the gradients already live
in ONE flat fp32 arena, so synchronisation is a handful of all-reduces over contiguous slices:

  * `buckets()` cuts the arena into regions in the order the backward pass finishes them
    (final block -> up path -> middle -> down path -> stem + time embedding), each a contiguous slice;
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
        dist.init_process_group(backend, **kw)
    return rank, local, world


def buckets(arena) -> List[Tuple[int, int]]:
    """Contiguous [begin, end) element ranges of the flat arena, listed in backward-completion order."""
    names = [nm for nm, _ in arena.spec]

    def span(pred):
        offs = [(arena.offset[n], arena.offset[n] + arena._numel(n)) for n in names if pred(n)]
        return min(o[0] for o in offs), max(o[1] for o in offs)

    named = {
        "final": span(lambda n: n.startswith("final_") and ".mlp.1." not in n),
        "ups": span(lambda n: n.startswith("ups.") and ".mlp.1." not in n),
        "mid": span(lambda n: n.startswith("mid_") and ".mlp.1." not in n),
        "downs": span(lambda n: n.startswith("downs.") and ".mlp.1." not in n),
        # stem, time MLP and the concatenated FiLM projections (arena head) are finished last
        "head": (0, span(lambda n: n.startswith("init_conv") or n.startswith("time_mlp"))[1]),
    }
    # make the regions tile the arena exactly (alignment padding goes to the following region)
    by_mem = sorted(named.items(), key=lambda kv: kv[1][0])
    tiled, prev = {}, 0
    for k, (b, e) in by_mem:
        tiled[k] = (prev, e)
        prev = e
    last = by_mem[-1][0]
    tiled[last] = (tiled[last][0], arena.numel)
    # order in which Plan.bwd_segments completes them
    return [tiled[k] for k in ("final", "ups", "mid", "downs", "head")]


class GradSync:
    """Bucketed, overlapped gradient all-reduce over the flat gradient arena."""

    def __init__(self, arena, group=None):
        self.arena, self.group = arena, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.buckets = buckets(arena)
        self.cuda = arena.gflat.is_cuda
        self.stream = torch.cuda.Stream() if (self.cuda and self.world > 1) else None
        self.handles = []

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
