"""Which filter taps does the halo kernel get right?  One-hot tap weights, compared with torch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch, torch.nn.functional as F
from b200dm import _lib as L
from b200dm.tensor import View
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
B, S, Cin, Cout = 2, 32, 64, 64
g = torch.Generator().manual_seed(0)
x = torch.randn(B, Cin, S, S, generator=g).to(dev).to(torch.bfloat16).float()
wfull = (torch.randn(Cout, Cin, 3, 3, generator=g) / 24).to(dev).to(torch.bfloat16).float()
xv = View.zeros(B, S, S, Cin, torch.bfloat16, dev).from_nchw(x)
for tap in list(range(9)) + [-1]:
    w = torch.zeros_like(wfull)
    if tap >= 0:
        w[:, :, tap // 3, tap % 3] = wfull[:, :, tap // 3, tap % 3]
    else:
        w = wfull
    ref = F.conv2d(x, w, None, padding=1)
    wp = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous().to(torch.bfloat16)
    yv = View.zeros(B, S, S, Cout, torch.bfloat16, dev)
    d = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=S, W=S, Cin=Cin, Cout=Cout, x=xv.ptr, x_ld=xv.ld,
                   w=wp.data_ptr(), bias=None, y=yv.ptr, y_ld=yv.ld, res=None, res_ld=0, accumulate=0)
    L.call("b200dm_conv_fwd", d)
    out = yv.to_nchw()
    err = ((out - ref).norm() / ref.norm()).item()
    # per-position error map summary: which (y%16, x%8) rows are wrong
    e = (out - ref).abs().amax(dim=(0, 1))          # [S,S]
    bad = (e > 0.05 * ref.abs().max()).float()
    print(f"tap {tap} (dy={tap//3},dx={tap%3}) rel {err:.4f}  bad-pixel fraction {bad.mean().item():.3f} "
          f"bad cols(x%8) {[int(bad[:, c::8].sum().item()) for c in range(8)]} bad rows(y%16) {[int(bad[r::16, :].sum().item()) for r in range(16)]}")
