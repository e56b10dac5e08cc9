#!/usr/bin/env python
"""GroupNorm forward/backward launches at the small (deep-level) shapes of the benchmark step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
from b200dm import _lib as L
from b200dm.tensor import View
dev = torch.device("cuda", 0)
for B, S, Cc in ((128, 4, 512), (128, 8, 256), (128, 16, 128), (128, 32, 64)):
    x = View(torch.randn(B, S, S, Cc, device=dev).to(torch.bfloat16))
    y = View.zeros(B, S, S, Cc, torch.bfloat16, dev)
    dx = View.zeros(B, S, S, Cc, torch.bfloat16, dev)
    stats = torch.empty(B, 8, 2, device=dev)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    film = torch.randn(B, 2 * Cc, device=dev) * 0.1
    dg, db, dbias = (torch.zeros(Cc, device=dev) for _ in range(3))
    dfilm = torch.zeros(B, 2 * Cc, device=dev)
    ws = torch.empty(max(1, L.load().b200dm_gn_bwd_ws_floats(B, S * S, Cc)), device=dev)
    gm = torch.empty(B, 8, 2, device=dev)
    def fwd():
        L.call("b200dm_gn_fwd", L.BF16, x.ptr, x.ld, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), film.data_ptr(),
               2 * Cc, None, 0, y.ptr, y.ld, B, S * S, Cc, 8, 1e-5)
    def bwd():
        L.call("b200dm_gn_apply_bwd", L.BF16, y.ptr, y.ld, x.ptr, x.ld, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
               film.data_ptr(), 2 * Cc, dx.ptr, dx.ld, None if os.environ.get("NOATOM") else dg.data_ptr(), db.data_ptr(), dfilm.data_ptr(), dbias.data_ptr(),
               ws.data_ptr(), gm.data_ptr(), B, S * S, Cc, 8)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda._sleep(2_000_000)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"gn_{name} B={B} {S}x{S} C={Cc}: {e0.elapsed_time(e1) * 50:.2f} us", flush=True)
