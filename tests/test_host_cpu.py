"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, schedules match
the oracle bit-for-bit, the ctypes structs match the C layout."""
import ctypes
import os
import re

import torch

from b200dm import _lib as L
from b200dm import schedule as S
from oracle import ddpm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200dm.h")).read()
    declared = sorted(set(re.findall(r"\b(b200dm_[a-z0-9_]+)\s*\(", header)))
    assert declared == L.ALL_SYMBOLS, set(declared) ^ set(L.ALL_SYMBOLS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200dm_version() == 100
    assert lib.b200dm_launch_count() == 0


def test_struct_layouts_match_c():
    # 9 int32, then pointer-aligned fields: offsets follow the C ABI on LP64
    assert ctypes.sizeof(L.ConvDesc) == 120 and L.ConvDesc.x.offset == 40 and L.ConvDesc.accumulate.offset == 100
    assert L.ConvDesc.gn_part.offset == 104 and L.ConvDesc.gn_groups.offset == 112
    assert ctypes.sizeof(L.PackEntry) == 80
    assert ctypes.sizeof(L.WgradDesc) == 112 and L.WgradDesc.dw.offset == 72 and L.WgradDesc.s_tap.offset == 88


def test_schedules_bit_equal_to_oracle():
    for sched in ("linear", "cosine", "sigmoid"):
        for obj in ("pred_noise", "pred_x0", "pred_v"):
            a, b = S.make_buffers(1000, sched, obj), O.make_buffers(1000, sched, obj)
            assert a.keys() == b.keys()
            for k in a:
                assert torch.equal(a[k], b[k]), (sched, obj, k)
    a, b = S.make_buffers(8, "sigmoid", "pred_v"), O.make_buffers(8, "sigmoid", "pred_v")
    assert all(torch.equal(a[k], b[k]) for k in a)
    o = O.DiffusionOracle({}, img_size=32, sampling_timesteps=50)
    assert S.ddim_time_pairs(1000, 50) == o.ddim_time_pairs()
