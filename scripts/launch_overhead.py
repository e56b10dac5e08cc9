#!/usr/bin/env python
"""Per-launch cost of back-to-back kernels, eager stream vs CUDA-graph replay (null conv kernel via
B200DM_HALO_DEBUG=8, and a 1-element fill)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
from b200dm import _lib as L
from b200dm.tensor import View

dev = "cuda"
B, H, cin, cout = 128, 32, 64, 64
x = View(torch.randn(B, H, H, cin, device=dev).to(torch.bfloat16))
y = View.zeros(B, H, H, cout, torch.bfloat16, dev)
w = torch.randn(9, cout, cin, device=dev).to(torch.bfloat16)
bias = torch.randn(cout, device=dev)
cd = L.ConvDesc(dtype=L.BF16, mode=0, ksize=3, impl=1, B=B, H=H, W=H, Cin=cin, Cout=cout, x=x.ptr, x_ld=x.ld,
                w=w.data_ptr(), bias=bias.data_ptr(), y=y.ptr, y_ld=y.ld, res=None, res_ld=0, accumulate=0)
buf = torch.zeros(16, device=dev)
N = 200

def conv():
    L.call("b200dm_conv_fwd", cd)

def fill():
    L.call("b200dm_fill_f32", buf.data_ptr(), 1, 1.0)

def timeit(fn, label):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(8_000_000)
    e0.record()
    for _ in range(N):
        fn()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) * 1e3 / N
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(N):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: eager {eager:.2f} us/launch, graph {e0.elapsed_time(e1) * 1e3 / N:.2f} us/launch", flush=True)

timeit(fill, "fill(1 elem)")
timeit(conv, f"conv halo dbg={os.environ.get('B200DM_HALO_DEBUG', '0')}")
