"""Parameter inventory and flat fp32 arenas for the B200 Unet.

The 283 tensors keep the reference's names and logical shapes (reference
models/generative/diffusion/ddpm.py:304-422, state_dict keys in SURVEY.md §8b) but live in ONE flat
fp32 arena (and one flat gradient arena) so that Adam, EMA, gradient all-reduce and zeroing are single
kernels / collectives.  Memory layout inside the arena is chosen for the kernels:

  * 3x3 / 1x1 GEMM-conv weights are stored tap-major  [kh*kw][Cout][Cin]  — the implicit-GEMM operand
    order — and exposed as the permuted view [Cout, Cin, kh, kw], so state_dict()/load_state_dict()
    see the reference OIHW shape while wgrad writes coalesced rows;
  * the 19 per-block FiLM projections (`*.mlp.1.{weight,bias}`) are contiguous, in forward order, so
    that they form one [8064, 256] GEMM operand without copies;
  * everything else is stored as the reference does.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

HEADS, DIM_HEAD, NUM_MEM_KV = 4, 32, 4
HIDDEN = HEADS * DIM_HEAD


class ConvInfo:
    """A convolution executed by the implicit-GEMM kernels."""
    __slots__ = ("name", "cin", "cout", "ksize", "mode", "bias", "taps", "master_packed")

    def __init__(self, name, cin, cout, ksize, mode=0, bias=True):
        self.name, self.cin, self.cout, self.ksize, self.mode, self.bias = name, cin, cout, ksize, mode, bias
        self.taps = ksize * ksize if mode == 0 else 4
        self.master_packed = mode == 0       # tap-major master; mode 1 keeps the reference [Cout, 4C]


def build_spec(dim: int, channels: int, dim_mults=(1, 2, 4, 8), self_condition: bool = False):
    """Returns (ordered [(name, shape)], {conv name: ConvInfo}, [resblock names in forward order],
    {resblock name: (cin, cout)})."""
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    convs: Dict[str, ConvInfo] = {}
    time_dim = dim * 4
    dims = [dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    n = len(in_out)
    blocks: Dict[str, Tuple[int, int]] = {}

    def conv(name, cin, cout, k, bias=True, mode=0, gemm=True):
        shape_cin = cin * 4 if mode == 1 else cin
        spec.append((name + ".weight", (cout, shape_cin, k, k)))
        if bias:
            spec.append((name + ".bias", (cout,)))
        if gemm:
            convs[name] = ConvInfo(name, cin, cout, k, mode, bias)

    def linear(name, cin, cout):
        spec.append((name + ".weight", (cout, cin)))
        spec.append((name + ".bias", (cout,)))

    def resblock(name, cin, cout):
        blocks[name] = (cin, cout)
        linear(name + ".mlp.1", time_dim, 2 * cout)
        for blk, ci in (("block1", cin), ("block2", cout)):
            conv(f"{name}.{blk}.proj", ci, cout, 3)
            spec.append((f"{name}.{blk}.norm.weight", (cout,)))
            spec.append((f"{name}.{blk}.norm.bias", (cout,)))
        if cin != cout:
            conv(name + ".res_conv", cin, cout, 1)

    def attention(name, c, full):
        mem_shape = (2, HEADS, NUM_MEM_KV, DIM_HEAD) if full else (2, HEADS, DIM_HEAD, NUM_MEM_KV)
        spec.append((name + ".mem_kv", mem_shape))
        spec.append((name + ".norm.g", (1, c, 1, 1)))
        conv(name + ".to_qkv", c, 3 * HIDDEN, 1, bias=False)
        if full:
            conv(name + ".to_out", HIDDEN, c, 1)
        else:
            conv(name + ".to_out.0", HIDDEN, c, 1)
            spec.append((name + ".to_out.1.g", (1, c, 1, 1)))

    conv("init_conv", channels * (2 if self_condition else 1), dim, 7, gemm=False)     # ddpm.py:300-304
    linear("time_mlp.1", dim, time_dim)
    linear("time_mlp.3", time_dim, time_dim)
    for i, (din, dout) in enumerate(in_out):
        last = i == n - 1
        resblock(f"downs.{i}.0", din, din)
        resblock(f"downs.{i}.1", din, din)
        attention(f"downs.{i}.2", din, full=last)
        if last:
            conv(f"downs.{i}.3", din, dout, 3)
        else:
            conv(f"downs.{i}.3.1", din, dout, 1, mode=1)
    for j, (din, dout) in enumerate(reversed(in_out)):
        last = j == n - 1
        resblock(f"ups.{j}.0", dout + din, dout)
        resblock(f"ups.{j}.1", dout + din, dout)
        attention(f"ups.{j}.2", dout, full=(j == 0))
        conv(f"ups.{j}.3" if last else f"ups.{j}.3.1", dout, din, 3)
    mid = dims[-1]
    resblock("mid_block1", mid, mid)
    attention("mid_attn", mid, full=True)
    resblock("mid_block2", mid, mid)
    resblock("final_res_block", 2 * dim, dim)
    conv("final_conv", dim, channels, 1, gemm=False)

    # forward execution order of the residual blocks (defines the FiLM column layout)
    order = []
    for i in range(n):
        order += [f"downs.{i}.0", f"downs.{i}.1"]
    order += ["mid_block1", "mid_block2"]
    for j in range(n):
        order += [f"ups.{j}.0", f"ups.{j}.1"]
    order.append("final_res_block")
    return spec, convs, order, blocks


class ParamArena:
    """Flat fp32 parameter + gradient arenas with reference-named logical views."""

    def __init__(self, dim: int, channels: int, device, with_grad: bool = True, self_condition: bool = False):
        self.dim, self.channels = dim, channels
        self.self_condition = bool(self_condition)
        self.in_channels = channels * (2 if self_condition else 1)       # stem input: [x_self_cond | x]
        self.spec, self.convs, self.block_order, self.blocks = build_spec(dim, channels,
                                                                          self_condition=self_condition)
        self.shapes = dict(self.spec)
        self.time_dim = 4 * dim
        # FiLM column layout: every block's (scale | shift) columns, the two level-0 blocks of the down path LAST.
        # Those two are the last ResnetBlocks of the backward pass, so the rows [0, film_early_cols) of the projection's
        # weight gradient are complete — and their all-reduce can start — while level 0 is still computing.
        self.film_order = ([b for b in self.block_order if not b.startswith("downs.0.")]
                           + [b for b in self.block_order if b.startswith("downs.0.")])
        self.film_off: Dict[str, int] = {}
        col = 0
        for b in self.film_order:
            if b.startswith("downs.0.") and not hasattr(self, "film_early_cols"):
                self.film_early_cols = col
            self.film_off[b] = col
            col += 2 * self.blocks[b][1]
        self.film_cols = col
        # arena placement
        film_w = [b + ".mlp.1.weight" for b in self.film_order]
        film_b = [b + ".mlp.1.bias" for b in self.film_order]
        placed = film_w + film_b
        placed_set = set(placed)
        placed += [nm for nm, _ in self.spec if nm not in placed_set]
        self.offset: Dict[str, int] = {}
        off = 0
        for nm in placed:
            self.offset[nm] = off
            n = 1
            for s in self.shapes[nm]:
                n *= s
            off += (n + 3) // 4 * 4            # keep every tensor 16-byte aligned
        self.numel = off
        self.epoch = 0            # bumped by kernels that update `flat` through raw pointers (Adam, EMA)
        self.flat = torch.zeros(off, dtype=torch.float32, device=device)
        self.gflat = torch.zeros(off, dtype=torch.float32, device=device) if with_grad else None
        self.views = {nm: self._logical(self.flat, nm) for nm, _ in self.spec}
        self.gviews = ({nm: self._logical(self.gflat, nm) for nm, _ in self.spec}
                       if with_grad else None)

    # ------------------------------------------------------------------------------------------
    def _numel(self, nm):
        n = 1
        for s in self.shapes[nm]:
            n *= s
        return n

    def _conv_of(self, nm):
        if nm.endswith(".weight"):
            return self.convs.get(nm[:-7])
        return None

    def _logical(self, flat: torch.Tensor, nm: str) -> torch.Tensor:
        o, n = self.offset[nm], self._numel(nm)
        raw = flat[o:o + n]
        ci = self._conv_of(nm)
        if ci is not None and ci.master_packed:
            co, cin, kh, kw = self.shapes[nm]
            return raw.view(kh, kw, co, cin).permute(2, 3, 0, 1)      # logical OIHW over tap-major memory
        return raw.view(self.shapes[nm])

    def touch(self):
        """Mark the parameters as modified by a raw-pointer kernel (the GEMM weight packs re-sync lazily).  In-place
        torch ops on `flat` or on the parameter views are detected through the tensor version counter instead."""
        self.epoch += 1

    @property
    def version(self):
        return (self.flat._version, self.epoch)

    def ptr(self, nm: str) -> int:
        return self.flat.data_ptr() + 4 * self.offset[nm]

    def gptr(self, nm: str) -> int:
        return self.gflat.data_ptr() + 4 * self.offset[nm]

    @property
    def film_weight_ptr(self):
        return self.ptr(self.film_order[0] + ".mlp.1.weight")

    @property
    def film_bias_ptr(self):
        return self.ptr(self.film_order[0] + ".mlp.1.bias")

    @property
    def film_weight_gptr(self):
        return self.gptr(self.film_order[0] + ".mlp.1.weight")

    @property
    def film_bias_gptr(self):
        return self.gptr(self.film_order[0] + ".mlp.1.bias")

    def load(self, sd: Dict[str, torch.Tensor], prefix: str = ""):
        missing = [nm for nm, _ in self.spec if prefix + nm not in sd]
        if missing:
            raise KeyError(f"missing parameters in state dict: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        with torch.no_grad():
            for nm, shape in self.spec:
                src = sd[prefix + nm]
                if tuple(src.shape) != tuple(shape):
                    raise ValueError(f"shape mismatch for {nm}: {tuple(src.shape)} vs {tuple(shape)}")
                self.views[nm].copy_(src.to(self.flat.device, torch.float32))

    def export(self) -> Dict[str, torch.Tensor]:
        """Reference-layout (contiguous) copies of every parameter."""
        return {nm: self.views[nm].detach().contiguous().clone() for nm, _ in self.spec}

    def export_grads(self) -> Dict[str, torch.Tensor]:
        return {nm: self.gviews[nm].detach().contiguous().clone() for nm, _ in self.spec}
