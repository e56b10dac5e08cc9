#!/usr/bin/env python
"""Per-layer timing of the tcgen05 conv kernels at the BASELINE shapes (every distinct conv of
SURVEY.md Appendix A at B=128, S=32), through the C ABI.  Launches are issued back-to-back behind a
device-side sleep so CUDA events measure GPU time, not host launch rate.

  python scripts/conv_microbench.py [--what fwd|wgrad|both|gn] [--reps 20] [--batch 128] [--size 32] [--only i]

--what gn: the fused conv + GroupNorm + FiLM + SiLU launch (b200dm_conv_gn_fwd) next to the two launches it replaces
(b200dm_conv_fwd with statistics in the epilogue + b200dm_gn_fwd_pre), for every 3x3 layer the fused kernel supports.
"""
import ctypes
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch  # noqa: E402

from b200dm import _lib as L  # noqa: E402
from b200dm.tensor import View  # noqa: E402

SHAPES = [  # (level, Cin, Cout, k)
    (0, 64, 64, 3), (0, 128, 64, 3), (0, 64, 384, 1), (0, 128, 64, 1),
    (1, 64, 64, 3), (1, 128, 128, 3), (1, 192, 128, 3), (1, 256, 128, 3), (1, 128, 384, 1),
    (2, 128, 128, 3), (2, 256, 256, 3), (2, 384, 256, 3), (2, 512, 256, 3), (2, 256, 384, 1),
    (3, 256, 256, 3), (3, 256, 512, 3), (3, 512, 512, 3), (3, 768, 512, 3), (3, 512, 384, 1),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="both")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--only", type=int, default=-1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = "cuda"
    B = a.batch
    rows = []
    for idx, (lvl, cin, cout, k) in enumerate(SHAPES):
        if a.only >= 0 and idx != a.only:
            continue
        H = a.size >> lvl
        x = View(torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16))
        dy = View(torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16))
        y = View.zeros(B, H, H, cout, torch.bfloat16, dev)
        w = (torch.randn(k * k, cout, cin, device=dev) / (cin * k * k) ** 0.5).to(torch.bfloat16)
        bias = torch.randn(cout, device=dev)
        dw = torch.zeros(k * k, cout, cin, device=dev)
        flops = 2.0 * B * H * H * cin * cout * k * k
        cd = L.ConvDesc(dtype=L.BF16, mode=0, ksize=k, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr,
                        x_ld=x.ld, w=w.data_ptr(), bias=bias.data_ptr(), y=y.ptr, y_ld=y.ld, res=None,
                        res_ld=0, accumulate=0)
        wd = L.WgradDesc(dtype=L.BF16, mode=0, ksize=k, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr,
                         x_ld=x.ld, dy=dy.ptr, dy_ld=dy.ld, dw=dw.data_ptr(), accumulate=1)
        rec = {"level": lvl, "H": H, "Cin": cin, "Cout": cout, "k": k, "gflop": flops / 1e9}
        if a.what == "gn":
            gamma, beta = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
            film = torch.randn(B, 2 * cout, device=dev) * 0.1
            stats = torch.zeros(B, 8, 2, device=dev)
            h = View.zeros(B, H, H, cout, torch.bfloat16, dev)
            gd = L.GnDesc(gamma=gamma.data_ptr(), beta=beta.data_ptr(), film=film.data_ptr(), film_ld=2 * cout,
                          groups=8, eps=1e-5, raw_ld=0, stats=None, raw=None)
            gd_tr = L.GnDesc(gamma=gamma.data_ptr(), beta=beta.data_ptr(), film=film.data_ptr(), film_ld=2 * cout,
                             groups=8, eps=1e-5, raw_ld=y.ld, stats=stats.data_ptr(), raw=y.ptr)
            cg = L.ConvDesc(dtype=L.BF16, mode=0, ksize=k, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr,
                            x_ld=x.ld, w=w.data_ptr(), bias=bias.data_ptr(), y=h.ptr, y_ld=h.ld, res=None,
                            res_ld=0, accumulate=0)
            if k != 3 or not L.load().b200dm_conv_gn_supported(ctypes.byref(cg), ctypes.byref(gd)):
                continue
            slots = H * H // min(32, H * H)
            part = torch.zeros(B, slots, cout // 8, 2, device=dev)
            cd.gn_part, cd.gn_groups = part.data_ptr(), 8

            def two():
                L.call("b200dm_conv_fwd", cd)
                L.call("b200dm_gn_fwd_pre", L.BF16, y.ptr, y.ld, part.data_ptr(), slots, stats.data_ptr(),
                       gamma.data_ptr(), beta.data_ptr(), film.data_ptr(), 2 * cout, None, 0, h.ptr, h.ld, B, H * H,
                       cout, 8, 1e-5)
            for nm, fn in (("conv+gn", two), ("fused_infer", lambda: L.call("b200dm_conv_gn_fwd", cg, gd)),
                           ("fused_train", lambda: L.call("b200dm_conv_gn_fwd", cg, gd_tr))):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(4_000_000)
                e0.record()
                for _ in range(a.reps):
                    fn()
                e1.record()
                torch.cuda.synchronize()
                rec[nm + "_us"] = round(e0.elapsed_time(e1) * 1e3 / a.reps, 2)
            rows.append(rec)
            print(json.dumps(rec), flush=True)
            continue
        for what, name, desc in (("fwd", "b200dm_conv_fwd", cd), ("wgrad", "b200dm_conv_wgrad", wd)):
            if a.what not in (what, "both"):
                continue
            for _ in range(3):
                L.call(name, desc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(4_000_000)
            e0.record()
            for _ in range(a.reps):
                L.call(name, desc)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / a.reps
            rec[what + "_us"] = round(us, 2)
            rec[what + "_tflops"] = round(flops / us / 1e6, 1)
        rows.append(rec)
        print(json.dumps(rec), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
