#!/usr/bin/env python
"""tcgen05.mma issue-rate ceiling per SM (see b200dm_debug_umma_rate): cycles per MMA and the TFLOP/s it implies."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
from b200dm import _lib as L

dev = torch.device("cuda", 0)
for ctas in (1, 148):
    for mode in (0, 1, 2):
        for n in (64, 128, 256):
            out = torch.zeros(ctas, dtype=torch.int64, device=dev)
            iters = 512
            for _ in range(2):
                L.call("b200dm_debug_umma_rate", n, iters, mode, ctas, out.data_ptr())
            torch.cuda.synchronize()
            cyc = out.float().mean().item() / (iters * 8)
            fl = 2 * 128 * n * 16
            print(f"ctas {ctas:3d} mode {mode} N {n:3d}: {cyc:7.1f} cycles/MMA -> {fl / cyc:7.0f} FLOP/cycle/SM "
                  f"({fl / cyc * 148 * 1.9e9 / 1e12:6.0f} TFLOP/s at 1.9 GHz x 148 SMs)", flush=True)
