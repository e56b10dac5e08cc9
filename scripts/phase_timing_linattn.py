#!/usr/bin/env python
"""Phase timeline of CTA 0 of the fused LinearAttention passes (debug build: make -C csrc timing)."""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightning-generative-models_b200")
os.environ.setdefault("B200DM_LIB", os.path.join(PKG, "b200dm", "libb200dm_timing.so"))
sys.path.insert(0, ROOT); sys.path.insert(0, PKG); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from b200dm import _lib as L
ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=256); ap.add_argument("--S", type=int, default=64); ap.add_argument("--C", type=int, default=64)
a = ap.parse_args()
import test_kernels_gpu as T
lib = L.load(); lib.b200dm_debug_set_la_timing_buf.argtypes = [ctypes.c_void_p]
y, ref, ws, d, keep = T._linattn_block_case(a.B, a.S, a.C, seed=1)
L.call("b200dm_linattn_block_fwd", ctypes.byref(d)); torch.cuda.synchronize()
buf = torch.zeros(512, dtype=torch.int64, device="cuda")
assert lib.b200dm_debug_set_la_timing_buf(buf.data_ptr()) == 0
L.call("b200dm_linattn_block_fwd", ctypes.byref(d)); torch.cuda.synchronize()
t = buf.cpu().tolist()
t0 = t[0]
c = lambda i: (t[i] - t0) if t[i] else -1
print("pass 2 (cycles since the transform warps start): per tile j: T1[wait-start, dq ready, done]  E[wait-start, dy ready, pre-bar, post-bar, done]")
for j in range(8):
    print(j, [c(16 + j * 8 + k) for k in range(8)], " mma[wait q, Y issued, Q(j+2) issued]", [c(100 + j * 4 + k) for k in range(3)])
print("pass 2, softmax warp, first head of tile j: [dq ready, TMEM loaded, max done, exp done, stored]  then [before, after] fence.proxy.async")
for j in range(8):
    print(j, [c(16 + j * 8 + 1)] + [c(300 + j * 4 + k) for k in range(4)], [c(340 + j * 2 + k) for k in range(2)])
t1 = t[200]
c1 = lambda i: (t[i] - t1) if t[i] else -1
print("pass 1: per 64-px tile: [wait-start, d1 ready, slot free, done]")
for j in range(16):
    print(j, [c1(200 + j * 4 + k) for k in range(4)])
