#!/usr/bin/env python
"""Per-tap error of the halo weight-gradient kernel against torch autograd (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
import torch.nn.functional as F
from b200dm import _lib as L
from b200dm.tensor import View

torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
B, S, Cin, Cout = 2, 32, 64, 64
g = torch.Generator().manual_seed(1)
x = torch.randn(B, Cin, S, S, generator=g).to(dev).bfloat16().float()
dy = torch.randn(B, Cout, S, S, generator=g).to(dev).bfloat16().float()
w = torch.zeros(Cout, Cin, 3, 3, device=dev, requires_grad=True)
F.conv2d(x, w, None, padding=1).backward(dy)
ref = w.grad.permute(2, 3, 0, 1).reshape(9, Cout, Cin)
xv = View.zeros(B, S, S, Cin, torch.bfloat16, dev).from_nchw(x)
dyv = View.zeros(B, S, S, Cout, torch.bfloat16, dev).from_nchw(dy)
dw = torch.zeros(9, Cout, Cin, device=dev)
d = L.WgradDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=S, W=S, Cin=Cin, Cout=Cout, x=xv.ptr, x_ld=xv.ld,
                dy=dyv.ptr, dy_ld=dyv.ld, dw=dw.data_ptr(), accumulate=1)
L.call("b200dm_conv_wgrad", d)
torch.cuda.synchronize()
for t in range(9):
    e = (dw[t] - ref[t]).norm() / ref[t].norm()
    # which reference tap does the result resemble most?
    best = min(range(9), key=lambda u: (dw[t] - ref[u]).norm().item())
    bt = min(range(9), key=lambda u: (dw[t] - ref[u].t()).norm().item()) if Cin == Cout else -1
    print(f"tap {t} (dy={t // 3},dx={t % 3}): rel err {e.item():.3e}  closest ref tap {best} "
          f"({((dw[t] - ref[best]).norm() / ref[best].norm()).item():.2e}), closest transposed {bt} "
          f"({((dw[t] - ref[bt].t()).norm() / ref[bt].norm()).item():.2e})  |dw| {dw[t].norm().item():.2f} |ref| {ref[t].norm().item():.2f}")
