#!/usr/bin/env python
"""Phase timeline of CTA 0 of the fused conv + GroupNorm launch (debug build, see scripts/phase_timing.py)."""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightning-generative-models_b200")
os.environ.setdefault("B200DM_LIB", os.path.join(PKG, "b200dm", "libb200dm_timing.so"))
sys.path.insert(0, PKG)
import torch
from b200dm import _lib as L
from b200dm.tensor import View
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256); ap.add_argument("--size", type=int, default=64)
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=64)
a = ap.parse_args()
dev = "cuda"; B, H, cin, cout = a.batch, a.size, a.cin, a.cout
x = View(torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16)); y = View.zeros(B, H, H, cout, torch.bfloat16, dev)
w = (torch.randn(9, cout, cin, device=dev) / (cin * 9) ** 0.5).to(torch.bfloat16); bias = torch.randn(cout, device=dev)
gamma, beta = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
cd = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr, x_ld=x.ld,
                w=w.data_ptr(), bias=bias.data_ptr(), y=y.ptr, y_ld=y.ld, res=None, res_ld=0, accumulate=0)
gd = L.GnDesc(gamma=gamma.data_ptr(), beta=beta.data_ptr(), film=None, film_ld=0, groups=8, eps=1e-5, raw_ld=0, stats=None, raw=None)
lib = L.load(); lib.b200dm_debug_set_timing_buf.argtypes = [ctypes.c_void_p]
buf = torch.zeros(512, dtype=torch.int64, device=dev)
for _ in range(3): L.call("b200dm_conv_gn_fwd", cd, gd)
torch.cuda.synchronize(); assert lib.b200dm_debug_set_timing_buf(buf.data_ptr()) == 0
L.call("b200dm_conv_gn_fwd", cd, gd); L.call("b200dm_conv_gn_fwd", cd, gd); torch.cuda.synchronize()
t = buf.cpu().tolist(); t0 = t[0]
us = lambda i: (t[i] - t0) / 1965.0 if t[i] else float("nan")
print(f"conv_gn {cin}->{cout} @{H}x{H} batch {B}: setup {us(1):.2f} dep {us(2):.2f} done {us(4):.2f}")
for si in range(4):
    e = 10 + si * 40
    print(f"sample {si}: start {us(e):.2f} | stats units " + " ".join(f"{us(e+1+k):.2f}" for k in range(4)) +
          f" | all-warps {us(e+5):.2f} published {us(e+6):.2f} peers {us(e+7):.2f} sums {us(e+8):.2f} coef {us(e+9):.2f} | apply units " +
          " ".join(f"{us(e+10+k):.2f}" for k in range(4)))
    print("   mma: " + " ".join(f"[{us(200+si*16+2*k):.2f} {us(200+si*16+2*k+1):.2f}]" for k in range(8)))
