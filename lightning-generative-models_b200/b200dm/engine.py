"""Static launch plan of the B200 Unet forward / backward.

`Plan` turns one (batch, size, training?) configuration of the reference's `Unet.forward`
(models/generative/diffusion/ddpm.py:428-471) into a fixed list of C-ABI kernel launches over
pre-allocated NHWC buffers, plus the mirrored backward list (the reference gets its backward from
autograd through ~9,000 ATen calls; here it is ~400 launches, CUDA-graph capturable: no allocation,
no host sync, no shape-dependent control flow at run time).

Data-flow decisions:
  * torch.cat never runs: producers write into channel slices of the consumer's concat buffer;
  * every gradient accumulation is an epilogue flag (accumulate / residual pointer) of the kernel that
    produces the contribution, so there are no stand-alone add kernels;
  * FiLM projections of all 19 ResnetBlocks are one GEMM.
"""
from __future__ import annotations

import ctypes as C
import re
from typing import Callable, List, Optional

import torch

from . import _lib as L
from .params import HIDDEN, ParamArena
from .tensor import View

GROUPS = 8
GN_EPS = 1e-5


class WeightPack:
    """GEMM-operand copies of the conv weights in the activation dtype: forward [taps][Cout][Cin] and
    data-gradient [taps'][Cin][Cout] (taps reversed for k x k convs)."""

    def __init__(self, arena: ParamArena, dt: int, with_dgrad: bool):
        self.arena, self.dt = arena, dt
        tdt = torch.bfloat16 if dt == L.BF16 else torch.float32
        self.fwd, self.tr = {}, {}
        total = sum(ci.taps * ci.cout * ci.cin for ci in arena.convs.values())
        dev = arena.flat.device
        self.fbuf = torch.empty(total, dtype=tdt, device=dev)
        self.tbuf = torch.empty(total, dtype=tdt, device=dev) if with_dgrad else None
        off = 0
        for nm, ci in arena.convs.items():
            n = ci.taps * ci.cout * ci.cin
            self.fwd[nm] = self.fbuf[off:off + n]
            if with_dgrad:
                self.tr[nm] = self.tbuf[off:off + n]
            off += n
        self.version = None
        # stem (7x7, C -> dim) as a GEMM over im2col patches: zero-padded bf16 weight rows [dim][KP]
        self.stem_k = arena.in_channels * 49
        self.stem_kp = (self.stem_k + 63) // 64 * 64
        self.stem = torch.zeros(arena.dim * self.stem_kp, dtype=tdt, device=dev) if dt == L.BF16 else None
        # the same weights in filter-row order (7 taps + a zero per 16-byte chunk) for the one-launch stem of inference
        self.stem_rows_kp = (arena.in_channels * 56 + 63) // 64 * 64
        self.stem_rows = (torch.zeros(arena.dim * self.stem_rows_kp, dtype=tdt, device=dev)
                          if dt == L.BF16 else None)
        # Upsample convs (ddpm.py:93-97) for inference: weights of the fused upsample + 3x3 launch (conv mode 3),
        # [4 taps][4 phases][Cout][Cin]; packed on demand (inference plans only)
        self.up = {}
        if dt == L.BF16:
            for nm, ci in arena.convs.items():
                if re.fullmatch(r"ups\.\d+\.3\.1", nm) and ci.mode == 0 and ci.ksize == 3:
                    self.up[nm] = torch.empty(16 * ci.cout * ci.cin, dtype=tdt, device=dev)
        self.up_version = None
        # LinearAttention blocks for inference (b200dm_linattn_block_fwd): to_qkv weights with the PreNorm gain folded
        # in, bf16 [384][C]; registered by the inference plans that use them
        self.la = {}
        self.la_version = None
        # device table for the one-launch batched pack
        entries = (L.PackEntry * len(arena.convs))()
        tile = 0
        for i, (nm, ci) in enumerate(arena.convs.items()):
            if ci.master_packed:
                s_tap, s_co, s_ci = ci.cout * ci.cin, ci.cin, 1
            else:                       # reference layout [Cout][c*4 + tap]
                s_tap, s_co, s_ci = 1, 4 * ci.cin, 4
            e = entries[i]
            e.w, e.wf = arena.ptr(nm + ".weight"), self.fwd[nm].data_ptr()
            e.wt = self.tr[nm].data_ptr() if with_dgrad else None
            e.taps, e.Cout, e.Cin, e.flip = ci.taps, ci.cout, ci.cin, 1 if ci.mode == 0 else 0
            e.s_tap, e.s_co, e.s_ci = s_tap, s_co, s_ci
            e.tiles_ci, e.tiles_co = (ci.cin + 63) // 64, (ci.cout + 63) // 64     # 64x64 tiles (optim.cu PK)
            e.tile_begin = tile
            tile += ci.taps * e.tiles_ci * e.tiles_co
        self.n_entries, self.total_tiles = len(arena.convs), tile
        raw = bytes(entries)
        self.table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(dev)
        # host copy of (arena offset of the weight, tile_begin, tiles) per entry: `refresh_range` re-packs the convs
        # of one gradient bucket (the optimiser-in-backward path)
        self._entry_size = C.sizeof(L.PackEntry)
        self._entry_info = []
        for i, (nm, ci) in enumerate(arena.convs.items()):
            e = entries[i]
            self._entry_info.append((arena.offset[nm + ".weight"], e.tile_begin, ci.taps * e.tiles_ci * e.tiles_co))
        self._range_cache = {}

    def refresh(self, force: bool = False):
        """Re-pack when the master arena changed (in-place updates bump the tensor version)."""
        v = self.arena.version
        if self.up and (force or v != self.up_version):
            for nm, buf in self.up.items():
                ci = self.arena.convs[nm]
                L.call("b200dm_pack_upconv_weight", self.arena.ptr(nm + ".weight"), buf.data_ptr(), ci.cout, ci.cin)
            self.up_version = v
        if self.la and (force or v != self.la_version):
            for nm, buf in self.la.items():
                L.call("b200dm_pack_linattn_qkv", self.arena.ptr(nm + ".to_qkv.weight"), self.arena.ptr(nm + ".norm.g"),
                       buf.data_ptr(), self.arena.convs[nm + ".to_qkv"].cin)
            self.la_version = v
        if not force and v == self.version:
            return
        L.call("b200dm_pack_conv_weights_batched", self.dt, self.table.data_ptr(), self.n_entries,
               self.total_tiles)
        if self.stem is not None:
            L.call("b200dm_pack_stem_weight", self.arena.ptr("init_conv.weight"), self.stem.data_ptr(),
                   self.arena.dim, self.stem_k, self.stem_kp)
            L.call("b200dm_pack_stem_rows", self.arena.ptr("init_conv.weight"), self.stem_rows.data_ptr(),
                   self.arena.dim, self.arena.in_channels, self.stem_rows_kp)
        self.version = v


    def linattn_qkv(self, nm: str) -> torch.Tensor:
        """Folded to_qkv weight of LinearAttention block `nm` (packed at the next refresh)."""
        if nm not in self.la:
            ci = self.arena.convs[nm + ".to_qkv"]
            self.la[nm] = torch.empty(ci.cout * ci.cin, dtype=torch.bfloat16, device=self.arena.flat.device)
            self.la_version = None
        return self.la[nm]

    def range_runs(self, begin: int, end: int):
        """Contiguous runs of table entries (first, n, tile_first, tiles) whose master weights live in arena elements
        [begin, end), and whether the stem weight does."""
        runs = self._range_cache.get((begin, end))
        if runs is None:
            idx = [i for i, (off, _, _) in enumerate(self._entry_info) if begin <= off < end]
            runs, k = [], 0
            while k < len(idx):
                j = k
                while j + 1 < len(idx) and idx[j + 1] == idx[j] + 1:
                    j += 1
                first, n = idx[k], j - k + 1
                tiles = sum(self._entry_info[i][2] for i in range(first, first + n))
                runs.append((first, n, self._entry_info[first][1], tiles))
                k = j + 1
            stem_off = self.arena.offset["init_conv.weight"]
            runs = (runs, self.stem is not None and begin <= stem_off < end)
            self._range_cache[(begin, end)] = runs
        return runs

    def refresh_range(self, begin: int, end: int):
        """Re-pack the convs whose master weights live in arena elements [begin, end) on the current stream (one
        launch per contiguous run of table entries; the stem rides with the range that holds it)."""
        runs, has_stem = self.range_runs(begin, end)
        for first, n, tile_first, tiles in runs:
            L.call("b200dm_pack_conv_weights_range", self.dt, self.table.data_ptr() + first * self._entry_size, n,
                   tile_first, tiles)
        if has_stem:
            L.call("b200dm_pack_stem_weight", self.arena.ptr("init_conv.weight"), self.stem.data_ptr(),
                   self.arena.dim, self.stem_k, self.stem_kp)
            L.call("b200dm_pack_stem_rows", self.arena.ptr("init_conv.weight"), self.stem_rows.data_ptr(),
                   self.arena.dim, self.arena.in_channels, self.stem_rows_kp)
        for nm, buf in self.up.items():
            if begin <= self.arena.offset[nm + ".weight"] < end:
                ci = self.arena.convs[nm]
                L.call("b200dm_pack_upconv_weight", self.arena.ptr(nm + ".weight"), buf.data_ptr(), ci.cout, ci.cin)


class Plan:
    def __init__(self, arena: ParamArena, pack: WeightPack, B: int, S: int, dt: int, training: bool,
                 use_tc: bool):
        assert S % 8 == 0, "H, W must be divisible by 8"
        self.arena, self.pack = arena, pack
        self.B, self.S, self.dt, self.training, self.use_tc = B, S, dt, training, use_tc
        self.dev = arena.flat.device
        self.tdt = torch.bfloat16 if dt == L.BF16 else torch.float32
        self.fwd: List[Callable[[int], None]] = []
        self.fwd_names: List[str] = []
        self.fwd_flops: List[float] = []
        self.bwd_groups: List[List[Callable[[int], None]]] = []
        self.unit_colsums: List[list] = []     # per unit: bias column sums deferred to the end of its segment
        self._cur_colsums: list = []
        self._cur_bwd: Optional[List[Callable[[int], None]]] = None
        self._region = "head"        # which gradient bucket the current unit's parameters belong to
        self.unit_regions: List[str] = []
        self._keep = []              # ctypes structs / tensors referenced by raw pointers
        self._scratch = {}
        self._scratch_ptrs = set()
        self.lib = L.load()
        self.nbytes = 0
        dim, ch = arena.dim, arena.channels
        self._keep.append(pack)
        import os
        self.side_enabled = os.environ.get("B200DM_SIDE_STREAM", "1") != "0"
        self.fuse_gn_stats = os.environ.get("B200DM_FUSE_GN_STATS", "1") != "0"
        self.fuse_upsample = os.environ.get("B200DM_FUSE_UPSAMPLE", "1") != "0"
        self.fuse_gn = os.environ.get("B200DM_FUSE_GN", "1") != "0"       # conv + GroupNorm + FiLM + SiLU in one launch
        self.batch_colsum = os.environ.get("B200DM_BATCH_COLSUM", "1") != "0"   # one bias-gradient launch per bucket
        self.fuse_linattn = os.environ.get("B200DM_FUSE_LINATTN", "1") != "0"   # inference: LinearAttention block fused
        self.fuse_stem = os.environ.get("B200DM_FUSE_STEM", "1") != "0"         # 7x7 stem: patches built in shared memory
        self._la_ws = None
        self._la_descs = []
        self.side_stream = torch.cuda.Stream(device=self.dev) if self.side_enabled else None
        # forward: the second stream only carries the 1x1 residual convs, which the main chain needs a few kernels
        # later -> high priority (measured, B=128 32x32 forward: 1.467 -> 1.440 ms).  In backward equal priorities are
        # best: with either stream preferred the step is slower (main chain first: 4.95 -> 5.27 ms, the parameter
        # gradients pile up at the segment joins; second stream first: 5.13 ms)
        self.side_stream_fwd = torch.cuda.Stream(device=self.dev, priority=-1) if self.side_enabled else None
        self.x_in = torch.zeros(B, ch, S, S, device=self.dev)          # NCHW fp32 boundary
        # self-conditioned net (ddpm.py:433-435): the previous x0 estimate (zeros when absent) is a second input,
        # concatenated in front of x by the first launch; the samplers write x0 straight into xc_in
        self.xc_in = torch.zeros(B, ch, S, S, device=self.dev) if arena.self_condition else None
        self.stem_in = torch.zeros(B, 2 * ch, S, S, device=self.dev) if arena.self_condition else self.x_in
        self.t_in = torch.zeros(B, dtype=torch.long, device=self.dev)
        self.out = torch.zeros(B, ch, S, S, device=self.dev)
        self.d_out = torch.zeros(B, ch, S, S, device=self.dev) if training else None
        self._build(dim, ch)
        # backward segments in gradient-bucket order (b200dm.distributed.buckets): after segment i has run,
        # bucket i of the gradient arena is final and can be all-reduced while the rest of backward runs
        n_unit_ops = sum(len(g) for g in self.bwd_groups)
        self.bwd_segments: List[List[Callable[[int], None]]] = self._finish_segments()
        self.bwd: List[Callable[[int], None]] = [op for seg in self.bwd_segments for op in seg]
        assert len(self.bwd) >= n_unit_ops       # every unit belongs to a region of distributed.REGIONS
        self.bwd_names = [op.kname for op in self.bwd]
        self.bwd_flops = [op.flops for op in self.bwd]
        self.fwd_names = [op.kname for op in self.fwd]
        self.fwd_flops = [op.flops for op in self.fwd]

    # ---- allocation -------------------------------------------------------------------------------
    def buf(self, H, C, dtype=None) -> View:
        dtype = self.tdt if dtype is None else dtype
        t = torch.zeros(self.B, H, H, C, dtype=dtype, device=self.dev)
        self.nbytes += t.numel() * t.element_size()
        self._keep.append(t)          # ops hold raw pointers only: the plan owns every buffer
        return View(t)

    def scratch(self, key, H, C) -> View:
        """Backward scratch shared by all units of the same geometry (backward is sequential)."""
        k = (key, H, C)
        if k not in self._scratch:
            self._scratch[k] = self.buf(H, C)
            self._scratch_ptrs.add(self._scratch[k].buf.data_ptr())
        return self._scratch[k]

    def f32(self, *shape) -> torch.Tensor:
        t = torch.zeros(*shape, dtype=torch.float32, device=self.dev)
        self.nbytes += t.numel() * 4
        self._keep.append(t)
        return t

    # ---- op emission --------------------------------------------------------------------------------
    def _emit(self, lst, name, *args, kname=None, flops=0.0, side=False, reads=(), writes=()):
        fn = getattr(self.lib, name)
        self._keep.append(args)

        def op(stream, fn=fn, args=args, name=name):
            rc = fn(*args, stream)
            if rc != 0:
                raise L.B200dmError(f"{name} failed with code {rc}: {L.last_error()}")
        op.kname = kname or name.replace("b200dm_", "")     # kernel family (bench.py per-kernel table)
        op.flops = flops                                    # algorithmic FLOPs of this launch
        # side = parameter-gradient work (wgrad, bias column sums) that nothing in the backward chain waits
        # for: it runs on the plan's second stream.  reads (side ops) / writes (main ops) name the buffers
        # whose reuse must be ordered across the two streams (write-after-read on shared scratch).
        op.side = side
        op.reads = frozenset(getattr(v, "buf", v).data_ptr() for v in reads if v is not None)
        op.writes = frozenset(getattr(v, "buf", v).data_ptr() for v in writes if v is not None)
        lst.append(op)

    def F(self, name, *args, **kw):
        self._emit(self.fwd, name, *args, **kw)

    def Bk(self, name, *args, **kw):
        if self.training:
            self._emit(self._cur_bwd, name, *args, **kw)

    def begin_unit(self):
        self._cur_bwd = []
        self._cur_colsums = []
        self.bwd_groups.append(self._cur_bwd)
        self.unit_colsums.append(self._cur_colsums)
        self.unit_regions.append(self._region)

    def _finish_segments(self):
        """Backward segments in gradient-bucket order.  The bias column sums deferred by conv_bwd go to the END of
        their segment, at most 16 per launch (b200dm_colsum_batched), on the second stream."""
        from .distributed import REGIONS
        segs = []
        for region in REGIONS:
            seg, items = [], []
            for g, cs, r in zip(reversed(self.bwd_groups), reversed(self.unit_colsums), reversed(self.unit_regions)):
                if r == region:
                    seg.extend(g)
                    items.extend(cs)
            for k in range(0, len(items), 16):
                chunk = items[k:k + 16]
                arr = (L.ColsumItem * len(chunk))()
                for q, (dy, rows, cout, gptr) in zip(arr, chunk):
                    q.x, q.out, q.rows, q.ld, q.C = dy.ptr, gptr, rows, dy.ld, cout
                self._cur_bwd = seg
                self.Bk("b200dm_colsum_batched", self.dt, arr, len(chunk), kname="colsum", side=True,
                        reads=tuple(dy for dy, _, _, _ in chunk))
            segs.append(seg)
        return segs

    @staticmethod
    def _tc_size_ok(H):
        """The tcgen05 kernels tile 128 pixels as whole image rows: H = W must be a power of two in [4, 128]."""
        return 4 <= H <= 128 and (H & (H - 1)) == 0

    def _impl(self, cin, cout, H):
        """1 = tcgen05 kernels, 0 = SIMT kernels (fp32 mode; channel counts that are not multiples of 64; levels whose
        size the tensor-core tiling does not cover, e.g. the 2x2 level of a 16x16 image or any non-power-of-two size)."""
        return 1 if (self.use_tc and self.dt == L.BF16 and cin % 64 == 0 and cout % 64 == 0
                     and self._tc_size_ok(H)) else 0

    def conv_fwd(self, lst_fn, nm, x: View, y: View, *, dgrad=False, res: Optional[View] = None,
                 accumulate=0, bias=True, side=False, gn_part: Optional[torch.Tensor] = None):
        """Forward conv `nm` (x -> y), or with dgrad=True its data gradient (x = dY, y = dX)."""
        ci = self.arena.convs[nm]
        if not dgrad:
            mode, cin, cout, H = ci.mode, ci.cin, ci.cout, y.H
            w = self.pack.fwd[nm].data_ptr()
            b = self.arena.ptr(nm + ".bias") if (ci.bias and bias) else None
        else:
            mode = 0 if ci.mode == 0 else 2
            cin, cout, H = ci.cout, ci.cin, x.H
            w = self.pack.tr[nm].data_ptr()
            b = None
        assert x.C == cin and y.C == cout, (nm, x.C, cin, y.C, cout)
        d = L.ConvDesc(dtype=self.dt, mode=mode, ksize=ci.ksize if ci.mode == 0 else 1,
                       impl=self._impl(cin, cout, H), B=self.B, H=H, W=H, Cin=cin, Cout=cout,
                       x=x.ptr, x_ld=x.ld, w=w, bias=b, y=y.ptr, y_ld=y.ld,
                       res=None if res is None else res.ptr, res_ld=0 if res is None else res.ld,
                       accumulate=accumulate, gn_part=None if gn_part is None else gn_part.data_ptr(),
                       gn_groups=GROUPS if gn_part is not None else 0)
        taps = ci.taps
        flops = 2.0 * self.B * H * H * cout * cin * taps
        fam = ("conv_tc" if d.impl == 1 else "conv_simt") + ("_dgrad" if dgrad else "_fwd")
        lst_fn("b200dm_conv_fwd", C.byref(d), kname=fam, flops=flops, writes=(y,), side=side,
               reads=(x, res) if side else ())
        self._keep.append(d)

    def conv_gn(self, nm, norm, x: View, y: View, *, film_ptr=None, res: Optional[View] = None,
                raw: Optional[View] = None, stats: Optional[torch.Tensor] = None, emit=True, reads=()):
        """Block.forward as ONE launch (b200dm_conv_gn_fwd): y = SiLU(GN(conv(x)+b)*(scale+1)+shift) (+ res); the
        conv accumulators stay in TMEM across the norm.  emit=False only asks whether the layer is supported."""
        a, ci = self.arena, self.arena.convs[nm]
        d = L.ConvDesc(dtype=self.dt, mode=0, ksize=3, impl=self._impl(ci.cin, ci.cout, y.H), B=self.B, H=y.H, W=y.H,
                       Cin=ci.cin, Cout=ci.cout, x=x.ptr, x_ld=x.ld, w=self.pack.fwd[nm].data_ptr(),
                       bias=a.ptr(nm + ".bias"), y=y.ptr, y_ld=y.ld, res=None if res is None else res.ptr,
                       res_ld=0 if res is None else res.ld, accumulate=0, gn_part=None, gn_groups=0)
        g = L.GnDesc(gamma=a.ptr(norm + ".weight"), beta=a.ptr(norm + ".bias"), film=film_ptr,
                     film_ld=a.film_cols if film_ptr else 0, groups=GROUPS, eps=GN_EPS,
                     raw_ld=0 if raw is None else raw.ld, stats=None if stats is None else stats.data_ptr(),
                     raw=None if raw is None else raw.ptr)
        if not emit:
            return self.lib.b200dm_conv_gn_supported(C.byref(d), C.byref(g)) == 1
        self.F("b200dm_conv_gn_fwd", C.byref(d), C.byref(g), kname="conv_gn_fwd",
               flops=2.0 * self.B * y.H * y.H * ci.cout * ci.cin * 9, writes=(y, raw), reads=reads)
        self._keep.append((d, g))
        return True

    def conv_bwd(self, nm, x: View, dy: View, dx: Optional[View], *, dx_acc=0, dx_res: Optional[View] = None,
                 bias_grad=True, dgrad_side=False):
        """wgrad + bias grad + (optional) dgrad of conv `nm` whose forward was x -> y, given dy."""
        if not self.training:
            return
        ci = self.arena.convs[nm]
        H = dy.H
        if ci.master_packed:
            s_tap, s_co, s_ci = ci.cout * ci.cin, ci.cin, 1
        else:
            s_tap, s_co, s_ci = 1, 4 * ci.cin, 4
        d = L.WgradDesc(dtype=self.dt, mode=ci.mode, ksize=ci.ksize if ci.mode == 0 else 1,
                        impl=self._impl(ci.cin, ci.cout, H), B=self.B, H=H, W=H, Cin=ci.cin, Cout=ci.cout,
                        x=x.ptr, x_ld=x.ld, dy=dy.ptr, dy_ld=dy.ld, dw=self.arena.gptr(nm + ".weight"),
                        accumulate=1, s_tap=s_tap, s_co=s_co, s_ci=s_ci)
        self.Bk("b200dm_conv_wgrad", C.byref(d), kname="wgrad_tc" if d.impl == 1 else "wgrad_simt",
                flops=2.0 * self.B * H * H * ci.cout * ci.cin * ci.taps, side=True, reads=(x, dy))
        self._keep.append(d)
        if ci.bias and bias_grad:
            if self.batch_colsum and dy.buf.data_ptr() not in self._scratch_ptrs and ci.cout % 8 == 0:
                # dy is a per-unit gradient buffer, final until the next step: its column sum joins the bucket's one
                # batched launch at the end of the segment (see _finish_segments)
                self._cur_colsums.append((dy, self.B * H * H, ci.cout, self.arena.gptr(nm + ".bias")))
            else:
                self.Bk("b200dm_colsum", self.dt, dy.ptr, dy.ld, self.B * H * H, ci.cout,
                        self.arena.gptr(nm + ".bias"), 1, side=True, reads=(dy,))
        if dx is not None:
            self.conv_fwd(self.Bk, nm, dy, dx, dgrad=True, res=dx_res, accumulate=dx_acc, side=dgrad_side)

    # ---- composite units -----------------------------------------------------------------------------
    def resblock(self, nm, x: View, out: View, gx: Optional[View], gout: Optional[View], gx_prior: bool):
        a = self.arena
        cin, cout = a.blocks[nm]
        H, HW = x.H, x.H * x.H
        has_res_conv = cin != cout
        tr = self.training
        self.begin_unit()
        c1, h1, c2 = self.buf(H, cout), self.buf(H, cout), self.buf(H, cout)
        st1, st2 = self.f32(self.B, GROUPS, 2), self.f32(self.B, GROUPS, 2)
        film_ptr = self.film.data_ptr() + 4 * a.film_off[nm]
        b1, b2 = nm + ".block1", nm + ".block2"
        rc = None
        if has_res_conv:      # issued first: runs on the second stream next to the whole conv/norm chain
            rc = self.buf(H, cout)
            self.conv_fwd(self.F, nm + ".res_conv", x, rc, side=True)
        res = rc if has_res_conv else x
        gs = cout // GROUPS
        fused = (self.fuse_gn_stats and self._impl(cin, cout, H) == 1 and self._impl(cout, cout, H) == 1
                 and gs % 8 == 0)
        one_launch = (fused and self.fuse_gn
                      and self.conv_gn(b1 + ".proj", b1 + ".norm", x, h1, emit=False)
                      and self.conv_gn(b2 + ".proj", b2 + ".norm", h1, out, emit=False))
        if one_launch:
            # conv -> GroupNorm -> FiLM -> SiLU (-> + residual) per launch; the raw conv outputs c1 / c2 and the
            # statistics are only written when the backward pass will read them
            self.conv_gn(b1 + ".proj", b1 + ".norm", x, h1, film_ptr=film_ptr, raw=c1 if tr else None,
                         stats=st1 if tr else None, reads=(self.film,))
            self.conv_gn(b2 + ".proj", b2 + ".norm", h1, out, res=res, raw=c2 if tr else None,
                         stats=st2 if tr else None, reads=(res,))
        elif fused:
            # GroupNorm statistics come out of the conv epilogues as per-slot partial sums; the norm is one pass
            slots = HW // min(32, HW)
            pt1, pt2 = self.f32(self.B, slots, cout // 8, 2), self.f32(self.B, slots, cout // 8, 2)
            self.conv_fwd(self.F, b1 + ".proj", x, c1, gn_part=pt1)
            self.F("b200dm_gn_fwd_pre", self.dt, c1.ptr, c1.ld, pt1.data_ptr(), slots, st1.data_ptr(),
                   a.ptr(b1 + ".norm.weight"), a.ptr(b1 + ".norm.bias"), film_ptr, a.film_cols, None, 0, h1.ptr,
                   h1.ld, self.B, HW, cout, GROUPS, GN_EPS, reads=(self.film,), kname="gn_fwd")
            self.conv_fwd(self.F, b2 + ".proj", h1, c2, gn_part=pt2)
            self.F("b200dm_gn_fwd_pre", self.dt, c2.ptr, c2.ld, pt2.data_ptr(), slots, st2.data_ptr(),
                   a.ptr(b2 + ".norm.weight"), a.ptr(b2 + ".norm.bias"), None, 0, res.ptr, res.ld, out.ptr, out.ld,
                   self.B, HW, cout, GROUPS, GN_EPS, reads=(res,), kname="gn_fwd")
        else:
            self.conv_fwd(self.F, b1 + ".proj", x, c1)
            self.F("b200dm_gn_fwd", self.dt, c1.ptr, c1.ld, st1.data_ptr(), a.ptr(b1 + ".norm.weight"),
                   a.ptr(b1 + ".norm.bias"), film_ptr, a.film_cols, None, 0, h1.ptr, h1.ld, self.B, HW, cout,
                   GROUPS, GN_EPS, reads=(self.film,))
            self.conv_fwd(self.F, b2 + ".proj", h1, c2)
            self.F("b200dm_gn_fwd", self.dt, c2.ptr, c2.ld, st2.data_ptr(), a.ptr(b2 + ".norm.weight"),
                   a.ptr(b2 + ".norm.bias"), None, 0, res.ptr, res.ld, out.ptr, out.ld, self.B, HW, cout, GROUPS,
                   GN_EPS, reads=(res,))
        if not self.training:
            return
        # dc2 / dc1: separate scratch for the two norm gradients, so block1's norm backward does not have to wait
        # for block2's weight gradient (second stream) to finish reading its operand
        dc2, dc, gh1 = self.scratch("dc2", H, cout), self.scratch("dc", H, cout), self.scratch("gh1", H, cout)
        dfilm_ptr = self.dfilm.data_ptr() + 4 * a.film_off[nm]
        if has_res_conv:
            # the 1x1 skip conv's backward only needs gout: issued first, on the second stream, next to the whole
            # norm/conv chain of the block; conv1's data gradient then accumulates into gx
            self.conv_bwd(nm + ".res_conv", x, gout, gx, dx_acc=1 if gx_prior else 0, dgrad_side=True)
        # block2: GN/SiLU backward -> dc2 ; conv2 backward -> gh1
        self.Bk("b200dm_gn_apply_bwd", self.dt, gout.ptr, gout.ld, c2.ptr, c2.ld, st2.data_ptr(),
                a.ptr(b2 + ".norm.weight"), a.ptr(b2 + ".norm.bias"), None, 0, dc2.ptr, dc2.ld,
                a.gptr(b2 + ".norm.weight"), a.gptr(b2 + ".norm.bias"), None, a.gptr(b2 + ".proj.bias"),
                self.sums.data_ptr(), self.gmeans.data_ptr(), self.B, HW, cout, GROUPS, writes=(dc2,))
        self.conv_bwd(b2 + ".proj", h1, dc2, gh1, bias_grad=False)
        # block1: GN/FiLM/SiLU backward -> dc ; conv1 backward -> gx (+ identity-skip gradient)
        self.Bk("b200dm_gn_apply_bwd", self.dt, gh1.ptr, gh1.ld, c1.ptr, c1.ld, st1.data_ptr(),
                a.ptr(b1 + ".norm.weight"), a.ptr(b1 + ".norm.bias"), film_ptr, a.film_cols, dc.ptr, dc.ld,
                a.gptr(b1 + ".norm.weight"), a.gptr(b1 + ".norm.bias"), dfilm_ptr, a.gptr(b1 + ".proj.bias"),
                self.sums.data_ptr(), self.gmeans.data_ptr(), self.B, HW, cout, GROUPS, writes=(dc,))
        self.conv_bwd(b1 + ".proj", x, dc, gx, dx_acc=1 if (gx_prior or has_res_conv) else 0,
                      dx_res=None if has_res_conv else gout, bias_grad=False)

    def attention(self, nm, x: View, out: View, gx: Optional[View], gout: Optional[View], full: bool):
        a = self.arena
        Cc, H = x.C, x.H
        n, rows = H * H, self.B * H * H
        self.begin_unit()
        if not full and not self.training and self.fuse_linattn and self.dt == L.BF16 and self.use_tc:
            # inference: RMSNorm -> to_qkv -> LinearAttention -> to_out -> RMSNorm -> + x in three launches that keep
            # q, k, v on chip (csrc/linattn_tc.cu)
            d = L.LinAttnBlockDesc(B=self.B, n=n, C=Cc, x_ld=x.ld, y_ld=out.ld, x=x.ptr, y=out.ptr,
                                   wqkv=self.pack.linattn_qkv(nm).data_ptr(),
                                   wout=self.pack.fwd[nm + ".to_out.0"].data_ptr(), bout=a.ptr(nm + ".to_out.0.bias"),
                                   gout=a.ptr(nm + ".to_out.1.g"), mem_kv=a.ptr(nm + ".mem_kv"))
            # below 1024 pixels per sample the four launches' fixed costs outweigh the saved traffic (measured at
            # batch 256: n = 256 fused 127 us, unfused chain ~75 us; n = 1024 fused 136 / 208 us, unfused ~190 / 235)
            if n >= 1024 and self.lib.b200dm_linattn_block_supported(C.byref(d)) == 1:
                need = self.lib.b200dm_linattn_block_ws_floats(self.B, n, Cc)
                if self._la_ws is None or self._la_ws.numel() < need:
                    self._la_ws = self.f32(need)        # the blocks run one after the other: one scratch for all
                    for dd in self._la_descs:
                        dd.ws = self._la_ws.data_ptr()
                d.ws = self._la_ws.data_ptr()
                self._la_descs.append(d)
                self.F("b200dm_linattn_block_fwd", C.byref(d), kname="linattn_block", writes=(out,),
                       flops=2.0 * rows * (Cc * 3 * HIDDEN + 2 * HIDDEN * 32 + HIDDEN * Cc))
                self._keep.append(d)
                return
        xn, qkv, ao = self.buf(H, Cc), self.buf(H, 3 * HIDDEN), self.buf(H, HIDDEN)
        self.F("b200dm_rmsnorm_fwd", self.dt, x.ptr, x.ld, a.ptr(nm + ".norm.g"), None, 0, xn.ptr, xn.ld, rows, Cc)
        self.conv_fwd(self.F, nm + ".to_qkv", xn, qkv)
        if full:
            self.F("b200dm_attn_fwd", self.dt, qkv.ptr, qkv.ld, a.ptr(nm + ".mem_kv"), ao.ptr, ao.ld, self.B, n)
            self.conv_fwd(self.F, nm + ".to_out", ao, out, res=x)
        else:
            ctx, kstat = self.f32(self.B, 4, 32, 32), self.f32(self.B, 4, 32, 2)
            to = self.buf(H, Cc)
            self.F("b200dm_linattn_fwd", self.dt, qkv.ptr, qkv.ld, a.ptr(nm + ".mem_kv"), ctx.data_ptr(),
                   kstat.data_ptr(), ao.ptr, ao.ld, self.B, n)
            self.conv_fwd(self.F, nm + ".to_out.0", ao, to)
            self.F("b200dm_rmsnorm_fwd", self.dt, to.ptr, to.ld, a.ptr(nm + ".to_out.1.g"), x.ptr, x.ld,
                   out.ptr, out.ld, rows, Cc)
        if not self.training:
            return
        dao, dqkv, dxn = self.scratch("dao", H, HIDDEN), self.scratch("dqkv", H, 3 * HIDDEN), self.scratch("dxn", H, Cc)
        if full:
            self.conv_bwd(nm + ".to_out", ao, gout, dao)
            self.Bk("b200dm_attn_bwd", self.dt, dao.ptr, dao.ld, qkv.ptr, qkv.ld, a.ptr(nm + ".mem_kv"),
                    dqkv.ptr, dqkv.ld, a.gptr(nm + ".mem_kv"), self.B, n, writes=(dqkv,))
        else:
            dto = self.buf(H, Cc)      # per unit (not shared scratch): its column sum is deferred to the segment end
            self.Bk("b200dm_rmsnorm_bwd", self.dt, gout.ptr, gout.ld, to.ptr, to.ld, a.ptr(nm + ".to_out.1.g"),
                    None, 0, dto.ptr, dto.ld, a.gptr(nm + ".to_out.1.g"), rows, Cc, writes=(dto,))
            self.conv_bwd(nm + ".to_out.0", ao, dto, dao)
            self.Bk("b200dm_linattn_bwd", self.dt, dao.ptr, dao.ld, qkv.ptr, qkv.ld, a.ptr(nm + ".mem_kv"),
                    ctx.data_ptr(), kstat.data_ptr(), self.dctx.data_ptr(), dqkv.ptr, dqkv.ld,
                    a.gptr(nm + ".mem_kv"), self.B, n, writes=(dqkv,))
        self.conv_bwd(nm + ".to_qkv", xn, dqkv, dxn)
        # gx = rmsnorm'(dxn) + gout   (the `attn(x) + x` skip, ddpm.py:449,455,464)
        self.Bk("b200dm_rmsnorm_bwd", self.dt, dxn.ptr, dxn.ld, x.ptr, x.ld, a.ptr(nm + ".norm.g"),
                gout.ptr, gout.ld, gx.ptr, gx.ld, a.gptr(nm + ".norm.g"), rows, Cc, writes=(gx,))

    def plain_conv(self, nm, x: View, out: View, gx, gout, gx_prior: bool):
        self.begin_unit()
        self.conv_fwd(self.F, nm, x, out)
        if self.training:
            self.conv_bwd(nm, x, gout, gx, dx_acc=1 if gx_prior else 0)

    def upsample_conv(self, nm, x: View, out: View, gx, gout):
        self.begin_unit()
        ci = self.arena.convs[nm]
        if nm in self.pack.up and self._impl(ci.cin, ci.cout, x.H) == 1 and self.fuse_upsample:
            # nearest-2x upsample + 3x3 conv as ONE launch over the low-resolution tensor (conv mode 3: four 2x2
            # convs, one per output phase; 16 instead of 36 multiply-adds per output, no upsampled copy on the chain)
            d = L.ConvDesc(dtype=self.dt, mode=3, ksize=3, impl=1, B=self.B, H=x.H, W=x.H, Cin=ci.cin, Cout=ci.cout,
                           x=x.ptr, x_ld=x.ld, w=self.pack.up[nm].data_ptr(), bias=self.arena.ptr(nm + ".bias"),
                           y=out.ptr, y_ld=out.ld, res=None, res_ld=0, accumulate=0, gn_part=None, gn_groups=0)
            self.F("b200dm_conv_fwd", C.byref(d), kname="conv_tc_fwd",
                   flops=2.0 * self.B * x.H * x.H * ci.cout * ci.cin * 16, writes=(out,))
            self._keep.append(d)
            if self.training:
                # backward is the reference's: the weight gradient wants the upsampled tensor as its operand, so it is
                # materialised there, on the side stream next to the weight-gradient kernel that reads it
                xu = self.buf(2 * x.H, x.C)
                self.Bk("b200dm_upsample2x_fwd", self.dt, x.ptr, x.ld, xu.ptr, xu.ld, self.B, x.H, x.H, x.C,
                        side=True, reads=(x,), writes=(xu,))
                gxu = self.scratch("gxu", 2 * x.H, x.C)
                self.conv_bwd(nm, xu, gout, gxu)
                self.Bk("b200dm_upsample2x_bwd", self.dt, gxu.ptr, gxu.ld, gx.ptr, gx.ld, self.B, x.H, x.H, x.C,
                        writes=(gx,))
            return
        xu = self.buf(2 * x.H, x.C)
        self.F("b200dm_upsample2x_fwd", self.dt, x.ptr, x.ld, xu.ptr, xu.ld, self.B, x.H, x.H, x.C)
        self.conv_fwd(self.F, nm, xu, out)
        if self.training:
            gxu = self.scratch("gxu", 2 * x.H, x.C)
            self.conv_bwd(nm, xu, gout, gxu)
            self.Bk("b200dm_upsample2x_bwd", self.dt, gxu.ptr, gxu.ld, gx.ptr, gx.ld, self.B, x.H, x.H, x.C,
                    writes=(gx,))

    # ---- the network -----------------------------------------------------------------------------------
    def _build(self, dim, ch):
        a, B, S, tr = self.arena, self.B, self.S, self.training
        dims = [dim, dim, 2 * dim, 4 * dim, 8 * dim]
        td = a.time_dim
        # time embedding + all FiLM projections (ddpm.py:440, :191-194)
        self.emb, self.pre1, self.h = self.f32(B, dim), self.f32(B, td), self.f32(B, td)
        self.pre3, self.tact = self.f32(B, td), self.f32(B, td)
        self.film = self.f32(B, a.film_cols)
        if tr:
            self.dfilm, self.dtact, self.dh = self.f32(B, a.film_cols), self.f32(B, td), self.f32(B, td)
            ws = max(self.lib.b200dm_gn_bwd_ws_floats(B, (S >> lv) ** 2, cc)
                     for lv in range(4) for cc in (dim << lv, dim << min(lv + 1, 3), dim))
            self.sums, self.gmeans = self.f32(int(ws)), self.f32(B, GROUPS, 2)
            self.dctx = self.f32(B, 4, 33, 32)
        self.begin_unit()
        # the time-embedding chain only feeds the FiLM scale/shift of the first norm: second stream, next to the stem
        self.F("b200dm_sinusoidal", self.t_in.data_ptr(), self.emb.data_ptr(), B, dim, 10000.0, side=True)
        self.F("b200dm_linear_fwd", self.emb.data_ptr(), a.ptr("time_mlp.1.weight"), a.ptr("time_mlp.1.bias"),
               self.h.data_ptr(), self.pre1.data_ptr(), B, td, dim, 1, side=True)
        self.F("b200dm_linear_fwd", self.h.data_ptr(), a.ptr("time_mlp.3.weight"), a.ptr("time_mlp.3.bias"),
               self.tact.data_ptr(), self.pre3.data_ptr(), B, td, td, 2, side=True)
        self.F("b200dm_linear_fwd", self.tact.data_ptr(), a.film_weight_ptr, a.film_bias_ptr,
               self.film.data_ptr(), None, B, a.film_cols, td, 0, side=True, writes=(self.film,))
        if tr:  # runs LAST in backward (dfilm is complete once every block has run): the FiLM rows of the two level-0
            # down blocks (+ the bias gradient of all rows); the other rows were done behind level 1 (see below)
            ec = a.film_early_cols
            self.Bk("b200dm_linear_bwd_cols", self.tact.data_ptr(), a.film_weight_ptr + 4 * ec * td,
                    self.dfilm.data_ptr() + 4 * ec, a.film_cols, self.dtact.data_ptr(), 1,
                    a.film_weight_gptr + 4 * ec * td, B, a.film_cols - ec, td)
            self.Bk("b200dm_colsum", L.F32, self.dfilm.data_ptr(), a.film_cols, B, a.film_cols, a.film_bias_gptr, 1)
            self.Bk("b200dm_linear_bwd", self.h.data_ptr(), a.ptr("time_mlp.3.weight"), self.pre3.data_ptr(),
                    self.dtact.data_ptr(), self.dh.data_ptr(), a.gptr("time_mlp.3.weight"),
                    a.gptr("time_mlp.3.bias"), B, td, td, 2)
            self.Bk("b200dm_linear_bwd", self.emb.data_ptr(), a.ptr("time_mlp.1.weight"), self.pre1.data_ptr(),
                    self.dh.data_ptr(), None, a.gptr("time_mlp.1.weight"), a.gptr("time_mlp.1.bias"),
                    B, td, dim, 1)

        res = [S, S // 2, S // 4, S // 8]
        G = (lambda H, Cc: self.buf(H, Cc)) if tr else (lambda H, Cc: None)
        # concat buffers of the up path, [x_up | skip]; level i of the down path feeds up level 3-i
        catB, catA, gcatB, gcatA = {}, {}, {}, {}
        for i in range(4):
            d_in, d_out = dims[i], dims[i + 1]
            catB[i], catA[i] = self.buf(res[i], d_out + d_in), self.buf(res[i], d_out + d_in)
            gcatB[i], gcatA[i] = G(res[i], d_out + d_in), G(res[i], d_out + d_in)
        catF, gcatF = self.buf(S, 2 * dim), G(S, 2 * dim)
        sl = (lambda g, off, Cc: None if g is None else g.slice(off, Cc))

        # init conv -> r (= second half of the final concat), ddpm.py:437-438, :468
        self.begin_unit()
        r, gr = catF.slice(dim, dim), sl(gcatF, dim, dim)
        chi = a.in_channels
        if a.self_condition:     # torch.cat((x_self_cond, x), dim=1), ddpm.py:435
            self.F("b200dm_concat2_nchw", self.xc_in.data_ptr(), self.x_in.data_ptr(), self.stem_in.data_ptr(), B,
                   ch * S * S, ch * S * S)
        if self.use_tc and self.dt == L.BF16 and self.pack.stem is not None and self._tc_size_ok(S):
            # stem on the tensor cores: im2col patches (kept for the weight gradient) + 1x1 GEMM conv
            KP, K = self.pack.stem_kp, self.pack.stem_k
            # inference: ONE launch, patches built in shared memory (csrc/stem_tc.cu).  Training needs the patch matrix
            # for the weight gradient anyway and measured 0.03 ms slower with it rebuilt in backward: it keeps the GEMM path.
            KR = self.pack.stem_rows_kp
            fused_stem = (self.fuse_stem and not tr and dim == 64
                          and self.lib.b200dm_stem7_supported(B, chi, S, S, KR, r.ld) == 1)
            P = None if fused_stem else self.buf(S, KP)
            if fused_stem:
                self.F("b200dm_stem7_fwd", self.stem_in.data_ptr(), self.pack.stem_rows.data_ptr(),
                       a.ptr("init_conv.bias"), r.ptr, r.ld, B, chi, S, S, KR, kname="stem7_fwd",
                       flops=2.0 * B * S * S * dim * K, writes=(r,))
            else:
                self.F("b200dm_im2col7", self.stem_in.data_ptr(), P.ptr, B, chi, S, S, KP)
                d = L.ConvDesc(dtype=self.dt, mode=0, ksize=1, impl=1, B=B, H=S, W=S, Cin=KP, Cout=dim, x=P.ptr,
                               x_ld=P.ld, w=self.pack.stem.data_ptr(), bias=a.ptr("init_conv.bias"), y=r.ptr,
                               y_ld=r.ld, res=None, res_ld=0, accumulate=0)
                self.F("b200dm_conv_fwd", C.byref(d), kname="conv_tc_fwd", flops=2.0 * B * S * S * dim * K)
                self._keep.append(d)
            if tr:
                wd = L.WgradDesc(dtype=self.dt, mode=0, ksize=1, impl=1, B=B, H=S, W=S, Cin=KP, Cout=dim,
                                 x=P.ptr, x_ld=P.ld, dy=gr.ptr, dy_ld=gr.ld, dw=a.gptr("init_conv.weight"),
                                 accumulate=1, cin_valid=K, s_tap=dim * K, s_co=K, s_ci=1)
                self.Bk("b200dm_conv_wgrad", C.byref(wd), kname="wgrad_tc", flops=2.0 * B * S * S * dim * K,
                        side=True, reads=(P, gr))
                self._keep.append(wd)
                self.Bk("b200dm_colsum", self.dt, gr.ptr, gr.ld, B * S * S, dim, a.gptr("init_conv.bias"), 1,
                        side=True, reads=(gr,))
        else:
            self.F("b200dm_init_conv_fwd", self.dt, self.stem_in.data_ptr(), a.ptr("init_conv.weight"),
                   a.ptr("init_conv.bias"), r.ptr, r.ld, B, chi, S, S, dim)
            if tr:
                self.Bk("b200dm_init_conv_wgrad", self.dt, self.stem_in.data_ptr(), gr.ptr, gr.ld,
                        a.gptr("init_conv.weight"), B, chi, S, S, dim)
                self.Bk("b200dm_colsum", self.dt, gr.ptr, gr.ld, B * S * S, dim, a.gptr("init_conv.bias"), 1)

        x, gx = r, gr
        x_prior = True        # gcatF[dim:] also receives the final block's gradient first
        for i in range(4):
            self._region = f"downs.{i}"          # one gradient bucket per level of the down path
            d_in, d_out, H = dims[i], dims[i + 1], res[i]
            last = i == 3
            x1, gx1 = catA[i].slice(d_out, d_in), sl(gcatA[i], d_out, d_in)
            self.resblock(f"downs.{i}.0", x, x1, gx, gx1, gx_prior=x_prior)
            x2, gx2 = self.buf(H, d_in), G(H, d_in)
            self.resblock(f"downs.{i}.1", x1, x2, gx1, gx2, gx_prior=True)
            x3, gx3 = catB[i].slice(d_out, d_in), sl(gcatB[i], d_out, d_in)
            self.attention(f"downs.{i}.2", x2, x3, gx2, gx3, full=last)
            Hn = H if last else H // 2
            x4, gx4 = self.buf(Hn, d_out), G(Hn, d_out)
            self.plain_conv(f"downs.{i}.3" if last else f"downs.{i}.3.1", x3, x4, gx3, gx4, gx_prior=True)
            x, gx, x_prior = x4, gx4, False
            if i == 0 and tr:
                # backward reaches this unit when levels 3..1 of the down path (and everything above) are done: all
                # FiLM gradients except those of the two level-0 blocks are final -> their share of the projection's
                # weight gradient and of d(time activation) now, on the second stream, and its bucket on the wire
                self._region = "film"
                self.begin_unit()
                ec = a.film_early_cols
                self.Bk("b200dm_linear_bwd_cols", self.tact.data_ptr(), a.film_weight_ptr, self.dfilm.data_ptr(),
                        a.film_cols, self.dtact.data_ptr(), 0, a.film_weight_gptr, B, ec, td, side=True,
                        reads=(self.dfilm,), writes=(self.dtact,))

        mid = dims[4]
        Hm = res[3]
        self._region = "mid"
        m1, gm1 = self.buf(Hm, mid), G(Hm, mid)
        self.resblock("mid_block1", x, m1, gx, gm1, gx_prior=False)
        m2, gm2 = self.buf(Hm, mid), G(Hm, mid)
        self.attention("mid_attn", m1, m2, gm1, gm2, full=True)
        u0, gu0 = catB[3].slice(0, mid), sl(gcatB[3], 0, mid)
        self.resblock("mid_block2", m2, u0, gm2, gu0, gx_prior=False)

        self._region = "ups"
        for j in range(4):
            i = 3 - j
            d_in, d_out, H = dims[i], dims[i + 1], res[i]
            last = j == 3
            u1, gu1 = catA[i].slice(0, d_out), sl(gcatA[i], 0, d_out)
            self.resblock(f"ups.{j}.0", catB[i], u1, gcatB[i], gu1, gx_prior=False)
            u2, gu2 = self.buf(H, d_out), G(H, d_out)
            self.resblock(f"ups.{j}.1", catA[i], u2, gcatA[i], gu2, gx_prior=False)
            u3, gu3 = self.buf(H, d_out), G(H, d_out)
            self.attention(f"ups.{j}.2", u2, u3, gu2, gu3, full=(j == 0))
            if last:
                nxt, gnxt = catF.slice(0, dim), sl(gcatF, 0, dim)
                self.plain_conv(f"ups.{j}.3", u3, nxt, gu3, gnxt, gx_prior=False)
            else:
                nxt, gnxt = catB[i - 1].slice(0, d_in), sl(gcatB[i - 1], 0, d_in)
                self.upsample_conv(f"ups.{j}.3.1", u3, nxt, gu3, gnxt)

        self._region = "final"
        yb, gy = self.buf(S, dim), G(S, dim)
        self.resblock("final_res_block", catF, yb, gcatF, gy, gx_prior=False)
        self.begin_unit()
        self.F("b200dm_final_conv_fwd", self.dt, yb.ptr, yb.ld, a.ptr("final_conv.weight"),
               a.ptr("final_conv.bias"), self.out.data_ptr(), B, S * S, dim, ch)
        if tr:
            # data gradient on the main chain; the (pixel-reduction) parameter gradients overlap it on the side stream
            self.Bk("b200dm_final_conv_bwd", self.dt, yb.ptr, yb.ld, a.ptr("final_conv.weight"),
                    self.d_out.data_ptr(), gy.ptr, gy.ld, None, None, B, S * S, dim, ch, writes=(gy,))
            self.Bk("b200dm_final_conv_bwd", self.dt, yb.ptr, yb.ld, a.ptr("final_conv.weight"),
                    self.d_out.data_ptr(), None, 0, a.gptr("final_conv.weight"),
                    a.gptr("final_conv.bias"), B, S * S, dim, ch, side=True, reads=(yb,))

    # ---- execution --------------------------------------------------------------------------------------
    def run_forward(self):
        self._run_ops(self.fwd)

    def run_backward(self):
        st = L.stream_ptr()
        for op in self.bwd:
            op(st)

    def run_backward_segment(self, i: int):
        """Launch backward segment i (see `_run_ops`); the gradient bucket of the segment is complete, in stream
        order, when this returns."""
        self._run_ops(self.bwd_segments[i])

    def _run_ops(self, ops):
        """Launch a list of ops on two streams.  Ops marked `side` (parameter-gradient kernels in backward, the
        1x1 residual convs in forward) go to the plan's second stream and overlap the main chain, which is
        latency-bound at the benchmark batch.  Ordering across the streams is by events, only where needed:
          * a side op starts after everything issued on the main stream so far (its inputs' producers);
          * a main op that writes a buffer a side op is still reading, or touches a buffer a side op writes,
            waits for that side op;
        and the streams join at the end, so everything is complete (in stream order) when this returns.  Works
        the same under CUDA-graph capture: the side stream forks from and rejoins the capturing stream."""
        if not self.side_enabled or not any(op.side for op in ops):
            st = L.stream_ptr()
            for op in ops:
                op(st)
            return
        main = torch.cuda.current_stream()
        side = self.side_stream_fwd if ops is self.fwd else self.side_stream
        main_ptr, side_ptr = main.cuda_stream, side.cuda_stream
        pending_r, pending_w = {}, {}      # buffer -> event of the last side kernel reading / writing it
        dirty = True
        for op in ops:
            if op.side:
                if dirty:                  # everything issued on the main stream so far (incl. the producers)
                    fork = torch.cuda.Event()
                    fork.record(main)
                    side.wait_event(fork)
                    dirty = False
                op(side_ptr)
                if op.reads or op.writes:
                    ev = torch.cuda.Event()
                    ev.record(side)
                    for b in op.reads:
                        pending_r[b] = ev
                    for b in op.writes:
                        pending_w[b] = ev
            else:
                waited = set()
                for b in op.writes:
                    for pend in (pending_r, pending_w):
                        ev = pend.pop(b, None)
                        if ev is not None and id(ev) not in waited:
                            main.wait_event(ev)
                            waited.add(id(ev))
                for b in op.reads:
                    ev = pending_w.pop(b, None)
                    if ev is not None and id(ev) not in waited:
                        main.wait_event(ev)
                        waited.add(id(ev))
                op(main_ptr)
                dirty = True
        main.wait_stream(side)
