#!/usr/bin/env python
"""Achieved HBM bandwidth of the memory-bound kernels against their ALGORITHMIC bytes, at the
benchmark shape and at a bandwidth-saturating shape (tensors >= 256 MB, larger than the 126 MB L2).

  python scripts/hbm_microbench.py [--out gpurun_out/hbm.json]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
from b200dm import _lib as L
from b200dm.schedule import make_buffers
from b200dm.tensor import View

dev = torch.device("cuda", 0)
PEAK = 6553.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(p):
    PEAK = json.load(open(p))["hbm_gbs"]


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(4_000_000)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps       # us


rows = []


def report(name, shape, nbytes, us):
    gbs = nbytes / us / 1e3
    rows.append({"kernel": name, "shape": shape, "algorithmic_MB": round(nbytes / 1e6, 1), "us": round(us, 2),
                 "GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / PEAK, 3)})
    print(json.dumps(rows[-1]), flush=True)


buf = {k: v.to(dev) for k, v in make_buffers(1000, "sigmoid", "pred_v").items()}
for tag, B, C, S in (("bench C2", 128, 3, 32), ("saturating", 8192, 3, 64)):
    n, chw = B * C * S * S, C * S * S
    img = torch.rand(B, C, S, S, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    xt, eps, x0, out, dout = (torch.empty_like(img) for _ in range(5))
    out.normal_()
    acc = torch.zeros(1, device=dev)
    import ctypes
    nd = L.NoiseDesc(img=img.data_ptr(), t=t.data_ptr(), noise=None, offset=None,
                     sqrt_ac=buf["sqrt_alphas_cumprod"].data_ptr(),
                     sqrt_1mac=buf["sqrt_one_minus_alphas_cumprod"].data_ptr(), offset_strength=0.0, normalize=1,
                     B=B, chw=chw, hw=S * S, seed=1234, stream_id=1, elem_offset=0)
    us = timeit(lambda: L.call("b200dm_q_sample", ctypes.byref(nd), xt.data_ptr(), None, None))
    report("q_sample (normalize + Philox + q_sample: read img, write x_t)", f"{tag} [{B},{C},{S},{S}] fp32", 8 * n, us)
    us = timeit(lambda: L.call("b200dm_loss_fwd_bwd", ctypes.byref(nd), out.data_ptr(), buf["loss_weight"].data_ptr(),
                               acc.data_ptr(), dout.data_ptr(), 2))
    report("loss_fwd_bwd (Philox eps regenerated: read out + img, write grad)", f"{tag} [{B},{C},{S},{S}] fp32",
           12 * n, us)
    us = timeit(lambda: L.call("b200dm_ddim_step", xt.data_ptr(), out.data_ptr(), None, x0.data_ptr(), None, 0.7, 0.7,
                               1.4, 1.0, 0.8, 0.6, 0.0, 0, 2, n, 1234, 1, 0))
    report("ddim_step", f"{tag} [{B},{C},{S},{S}] fp32", 12 * n, us)
    us = timeit(lambda: L.call("b200dm_ddpm_step", xt.data_ptr(), out.data_ptr(), None, x0.data_ptr(), None, 0.7, 0.7,
                               1.4, 1.0, 0.5, 0.5, 0.1, 1, 2, n, 1234, 1, 0))
    report("ddpm_step (Philox in registers)", f"{tag} [{B},{C},{S},{S}] fp32", 12 * n, us)

for tag, B, S, Cc in (("bench C2 level 0", 128, 32, 64), ("saturating", 256, 64, 64)):
    x = View(torch.randn(B, S, S, Cc, device=dev).to(torch.bfloat16))
    r = View(torch.randn(B, S, S, Cc, device=dev).to(torch.bfloat16))
    y = View.zeros(B, S, S, Cc, torch.bfloat16, dev)
    dx = View.zeros(B, S, S, Cc, torch.bfloat16, dev)
    stats = torch.empty(B, 8, 2, device=dev)
    gamma, beta = torch.ones(Cc, device=dev), torch.zeros(Cc, device=dev)
    film = torch.randn(B, 2 * Cc, device=dev) * 0.1
    dg, db, dbias = (torch.zeros(Cc, device=dev) for _ in range(3))
    dfilm = torch.zeros(B, 2 * Cc, device=dev)
    ws = torch.empty(max(1, L.load().b200dm_gn_bwd_ws_floats(B, S * S, Cc)), device=dev)
    gm = torch.empty(B, 8, 2, device=dev)
    nel = B * S * S * Cc
    us = timeit(lambda: L.call("b200dm_gn_fwd", L.BF16, x.ptr, x.ld, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                               film.data_ptr(), 2 * Cc, r.ptr, r.ld, y.ptr, y.ld, B, S * S, Cc, 8, 1e-5))
    report("gn_fwd (stats + FiLM + SiLU + residual, 1 launch)", f"{tag} [{B},{S},{S},{Cc}] bf16", 6 * nel, us)
    us = timeit(lambda: L.call("b200dm_gn_apply_bwd", L.BF16, y.ptr, y.ld, x.ptr, x.ld, stats.data_ptr(), gamma.data_ptr(),
                               beta.data_ptr(), film.data_ptr(), 2 * Cc, dx.ptr, dx.ld, dg.data_ptr(), db.data_ptr(),
                               dfilm.data_ptr(), dbias.data_ptr(), ws.data_ptr(), gm.data_ptr(), B, S * S, Cc, 8))
    report("gn_apply_bwd (reductions + dx, 1 launch)", f"{tag} [{B},{S},{S},{Cc}] bf16", 6 * nel, us)
    g1 = torch.ones(Cc, device=dev)
    us = timeit(lambda: L.call("b200dm_rmsnorm_fwd", L.BF16, x.ptr, x.ld, g1.data_ptr(), r.ptr, r.ld, y.ptr, y.ld,
                               B * S * S, Cc))
    report("rmsnorm_fwd (+ residual)", f"{tag} [{B},{S},{S},{Cc}] bf16", 6 * nel, us)

n = 35_719_555
pp, g, m, v = (torch.randn(n, device=dev) * 0.01 for _ in range(4))
v.abs_()
us = timeit(lambda: L.call("b200dm_adam_step", pp.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), n, 2e-5, 0.9, 0.99,
                           1e-8, 0.0, 3, 1.0))
report("adam_step (flat arena, 35.7 M params)", "fp32 p,g,m,v", 28 * n, us)

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
a = ap.parse_args()
if a.out:
    json.dump({"hbm_peak_GBps": PEAK, "rows": rows}, open(a.out, "w"), indent=1)
