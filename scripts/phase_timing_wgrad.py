#!/usr/bin/env python
"""Phase timeline of CTA 0 of the 3x3 weight-gradient halo kernel (debug build, see scripts/phase_timing.py)."""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightning-generative-models_b200")
os.environ.setdefault("B200DM_LIB", os.path.join(PKG, "b200dm", "libb200dm_timing.so"))
sys.path.insert(0, PKG)
import torch
from b200dm import _lib as L
from b200dm.tensor import View
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128); ap.add_argument("--size", type=int, default=32)
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=64)
a = ap.parse_args()
dev = "cuda"; B, H, cin, cout = a.batch, a.size, a.cin, a.cout
x = View(torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)); dy = View(torch.randn(B, H, H, cout, device=dev).to(torch.bfloat16))
dw = torch.zeros(9, cout, cin, device=dev)
wd = L.WgradDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr, x_ld=x.ld, dy=dy.ptr,
                 dy_ld=dy.ld, dw=dw.data_ptr(), accumulate=1)
lib = L.load(); lib.b200dm_debug_set_timing_buf.argtypes = [ctypes.c_void_p]
buf = torch.zeros(512, dtype=torch.int64, device=dev)
for _ in range(3): L.call("b200dm_conv_wgrad", wd)
torch.cuda.synchronize(); assert lib.b200dm_debug_set_timing_buf(buf.data_ptr()) == 0
L.call("b200dm_conv_wgrad", wd); L.call("b200dm_conv_wgrad", wd); torch.cuda.synchronize()
t = buf.cpu().tolist(); t0 = t[300]
us = lambda i: (t[i] - t0) / 1965.0 if t[i] else float("nan")
print(f"wgrad3x3 {cin}->{cout} @{H}x{H} batch {B}: setup {us(301):.2f} | dependency {us(302):.2f} | last MMA issued {us(303):.2f} | "
      f"accumulators complete {us(304):.2f} | reductions issued {us(305):.2f} | CTA done {us(306):.2f}")
print("tiles (operands landed, MMAs issued): " + " ".join(f"[{us(310+2*i):.2f} {us(311+2*i):.2f}]" for i in range(8)))
