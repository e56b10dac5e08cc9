"""CPU tests of the host side: C-ABI library loads and exports every declared symbol, schedules match
the oracle bit-for-bit, the ctypes structs match the C layout."""
import ctypes
import json
import os
import re

import pytest
import torch

from b200dm import _lib as L
from b200dm import schedule as S
from oracle import ddpm_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightning-generative-models_b200")


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "b200dm.h")).read()
    declared = sorted(set(re.findall(r"\b(b200dm_[a-z0-9_]+)\s*\(", header)))
    assert declared == L.ALL_SYMBOLS, set(declared) ^ set(L.ALL_SYMBOLS)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.b200dm_version() == 100
    assert lib.b200dm_launch_count() == 0


def test_struct_layouts_match_c():
    # 9 int32, then pointer-aligned fields: offsets follow the C ABI on LP64
    assert ctypes.sizeof(L.ConvDesc) == 120 and L.ConvDesc.x.offset == 40 and L.ConvDesc.accumulate.offset == 100
    assert L.ConvDesc.gn_part.offset == 104 and L.ConvDesc.gn_groups.offset == 112
    assert ctypes.sizeof(L.PackEntry) == 80
    assert ctypes.sizeof(L.WgradDesc) == 112 and L.WgradDesc.dw.offset == 72 and L.WgradDesc.s_tap.offset == 88
    assert ctypes.sizeof(L.NoiseDesc) == 104 and L.NoiseDesc.offset_strength.offset == 48
    assert L.NoiseDesc.chw.offset == 64 and L.NoiseDesc.seed.offset == 80
    # b200dm_colsum_item {ptr, ptr, int64, int32, int32}; b200dm_linattn_block_desc {6 int32, 8 pointers}
    assert ctypes.sizeof(L.ColsumItem) == 32 and L.ColsumItem.rows.offset == 16 and L.ColsumItem.C.offset == 28
    assert ctypes.sizeof(L.LinAttnBlockDesc) == 88 and L.LinAttnBlockDesc.x.offset == 24
    assert L.LinAttnBlockDesc.ws.offset == 80


def test_schedules_bit_equal_to_oracle():
    for sched in ("linear", "cosine", "sigmoid"):
        for obj in ("pred_noise", "pred_x0", "pred_v"):
            a, b = S.make_buffers(1000, sched, obj), O.make_buffers(1000, sched, obj)
            assert a.keys() == b.keys()
            for k in a:
                assert torch.equal(a[k], b[k]), (sched, obj, k)
    a, b = S.make_buffers(8, "sigmoid", "pred_v"), O.make_buffers(8, "sigmoid", "pred_v")
    assert all(torch.equal(a[k], b[k]) for k in a)
    o = O.DiffusionOracle({}, img_size=32, sampling_timesteps=50)
    assert S.ddim_time_pairs(1000, 50) == o.ddim_time_pairs()


def test_upsample_conv_phase_decomposition_identity():
    """The algebra behind conv_fwd mode 3 (b200dm_pack_upconv_weight): a 3x3 'same' conv over a nearest-2x upsampled
    image equals, per output phase (oy&1, ox&1), a 2x2 conv over the SOURCE image whose taps are sums of the 3x3 taps
    that read the same source pixel (reference Upsample, ddpm.py:93-97).  Exact in fp64."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(7)
    B, Cin, Cout, H = 2, 5, 4, 6
    x = torch.randn(B, Cin, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(Cout, Cin, 3, 3, generator=g, dtype=torch.float64)
    bias = torch.randn(Cout, generator=g, dtype=torch.float64)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, bias, padding=1)
    sel = ([[0], [1, 2]], [[0, 1], [2]])          # phase -> 2x2 tap -> 3x3 taps landing on that source pixel
    out = torch.empty_like(ref)
    xp = F.pad(x, (1, 1, 1, 1))                    # zero padding == TMA out-of-bounds fill
    for a in range(2):
        for b in range(2):
            acc = bias.view(1, Cout, 1, 1).expand(B, Cout, H, H).clone()
            for r in range(2):
                for c in range(2):
                    wsum = sum(w[:, :, ky, kx] for ky in sel[a][r] for kx in sel[b][c])          # [Cout, Cin]
                    dy, dx = r + a - 1, c + b - 1                                               # source offset
                    patch = xp[:, :, 1 + dy:1 + dy + H, 1 + dx:1 + dx + H]
                    acc = acc + torch.einsum("oc,bchw->bohw", wsum, patch)
            out[:, :, a::2, b::2] = acc
    assert (out - ref).abs().max().item() < 1e-12


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs) prints one JSON line with the contract's keys."""
    import subprocess
    import sys as _sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([_sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-batch", "2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_groupnorm_film_silu_backward_closed_form():
    """The closed forms gn_bwd_cluster_kernel evaluates (norm.cu): with z = (gamma*xn + beta)*(scale+1) + shift,
    y = silu(z), dz = dy*silu'(z), per-channel S1 = sum dz, S2 = sum dz*xn, S0 = sum x and per-group means
    M1, M2 of a*S1, a*S2 (a = (scale+1)*gamma):  dx = rstd*a*dz - rstd*M1 - rstd^2*M2*(x - mean); the parameter
    gradients and the gradient of the producing conv's bias follow from the same sums.  Checked against autograd of
    Block.forward (ddpm.py:164-173) in fp64."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(11)
    B, C, G, H = 3, 16, 4, 5
    HW, gs = H * H, C // G
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    conv, bias = rnd(B, C, H, H), rnd(C).requires_grad_(True)
    gamma, beta = rnd(C).requires_grad_(True), rnd(C).requires_grad_(True)
    film = rnd(B, 2 * C).requires_grad_(True)
    dy = rnd(B, C, H, H)
    x = (conv + bias.view(1, C, 1, 1))
    x.retain_grad()
    sc, sh = film[:, :C].view(B, C, 1, 1), film[:, C:].view(B, C, 1, 1)
    y = F.silu(F.group_norm(x, G, gamma, beta, eps=1e-5) * (sc + 1) + sh)
    y.backward(dy)
    with torch.no_grad():
        xg = x.view(B, G, gs * HW)
        mean = xg.mean(-1).repeat_interleave(gs, 1).view(B, C, 1, 1)
        rstd = (xg.var(-1, unbiased=False) + 1e-5).rsqrt().repeat_interleave(gs, 1).view(B, C, 1, 1)
        xn = (x - mean) * rstd
        scale1 = sc + 1
        z = (gamma.view(1, C, 1, 1) * xn + beta.view(1, C, 1, 1)) * scale1 + sh
        sg = torch.sigmoid(z)
        dz = dy * sg * (1 + z * (1 - sg))
        S1, S2, S0 = dz.sum((2, 3)), (dz * xn).sum((2, 3)), x.sum((2, 3))            # [B, C]
        a = (scale1 * gamma.view(1, C, 1, 1)).view(B, C)
        M1 = (a * S1).view(B, G, gs).sum(-1) / (gs * HW)
        M2 = (a * S2).view(B, G, gs).sum(-1) / (gs * HW)
        M1c, M2c = M1.repeat_interleave(gs, 1), M2.repeat_interleave(gs, 1)           # [B, C]
        rs, mu = rstd.view(B, C), mean.view(B, C)
        P, R = rs * a, -rs * rs * M2c
        Q = -rs * M1c - R * mu
        dx = P.view(B, C, 1, 1) * dz + Q.view(B, C, 1, 1) + R.view(B, C, 1, 1) * x
        assert (dx - x.grad).abs().max().item() < 1e-10
        assert ((scale1.view(B, C) * S2).sum(0) - gamma.grad).abs().max().item() < 1e-10
        assert ((scale1.view(B, C) * S1).sum(0) - beta.grad).abs().max().item() < 1e-10
        dfilm = torch.cat((gamma * S2 + beta * S1, S1), 1)
        assert (dfilm - film.grad).abs().max().item() < 1e-10
        sum_xn = (S0 - HW * mu) * rs
        dbias = (rs * (a * S1 - HW * M1c - M2c * sum_xn)).sum(0)
        assert (dbias - bias.grad).abs().max().item() < 1e-10
        # forward: silu(z) = h + h*tanh(h) with h = z/2 (one MUFU per element in the bf16 kernels)
        assert (z / 2 + z / 2 * torch.tanh(z / 2) - F.silu(z)).abs().max().item() < 1e-12


def test_linear_attention_partial_merge_and_backward_closed_form():
    """Algebra of the linear-attention kernels (attention.cu) against LinearAttention (ddpm.py:222-238) in fp64:
    forward — every CTA of a cluster owns a pixel range and produces (column max M_r, exp-sums L_r, unnormalised
    context C_r); the merge  ctx = sum_r e^{M_r-M} C_r / sum_r e^{M_r-M} L_r  equals the softmax over all n;
    backward — dctx = scale * softmax_d(q) dout^T, dq = scale * P .* (ctx dout - colsum(P .* ctx dout)),
    dv = ks^T-weighted dctx, dk = ks .* (v dctx^T - rowsum(dctx .* ctx))."""
    g = torch.Generator().manual_seed(13)
    d, n, nm, scale = 8, 37, 4, 8 ** -0.5
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    q, k, v = rnd(d, n).requires_grad_(True), rnd(d, n).requires_grad_(True), rnd(d, n).requires_grad_(True)
    mk, mv = rnd(d, nm).requires_grad_(True), rnd(d, nm).requires_grad_(True)
    dout = rnd(d, n)
    kk, vv = torch.cat((mk, k), 1), torch.cat((mv, v), 1)            # memory key/values are prepended (ddpm.py:225-226)
    ks = kk.softmax(-1)                                              # over the n + 4 positions
    qs = q.softmax(0) * scale                                        # over the channels, times dim_head^-0.5
    ctx = ks @ vv.t()                                                # [d, e]
    out = ctx.t() @ qs                                               # [e, n]
    out.backward(dout)
    with torch.no_grad():
        # ---- forward: three "CTAs" with uneven pixel ranges, rank 0 also owns the memory rows
        bounds, parts = [0, 4 + 10, 4 + 25, 4 + n], []
        for r in range(3):
            kr, vr = kk[:, bounds[r]:bounds[r + 1]], vv[:, bounds[r]:bounds[r + 1]]
            M = kr.max(1).values
            p = torch.exp(kr - M[:, None])
            parts.append((M, p.sum(1), p @ vr.t()))
        Mg = torch.stack([p[0] for p in parts]).max(0).values
        L = sum(torch.exp(M - Mg) * l for M, l, _ in parts)
        Cm = sum(torch.exp(M - Mg)[:, None] * c for M, _, c in parts) / L[:, None]
        assert (Cm - ctx).abs().max().item() < 1e-12
        # ---- backward from (ctx, M, L) as the kernel keeps them
        P = q.softmax(0)
        dctx = scale * (P @ dout.t())                                # [d, e]
        dqs = ctx @ dout                                             # [d, n]
        dq = scale * P * (dqs - (P * dqs).sum(0, keepdim=True))
        ksr = torch.exp(kk - Mg[:, None]) / L[:, None]
        dvv = (ksr.t() @ dctx).t()                                   # [e, n + 4]
        dks = dctx @ vv                                              # [d, n + 4]
        Dd = (dctx * ctx).sum(1)
        dkk = ksr * (dks - Dd[:, None])
        assert (dq - q.grad).abs().max().item() < 1e-12
        assert (dkk[:, nm:] - k.grad).abs().max().item() < 1e-12 and (dkk[:, :nm] - mk.grad).abs().max().item() < 1e-12
        assert (dvv[:, nm:] - v.grad).abs().max().item() < 1e-12 and (dvv[:, :nm] - mv.grad).abs().max().item() < 1e-12


def test_fused_linear_attention_block_algebra():
    """fp64 restatement of what csrc/linattn_tc.cu computes (per-pixel 1/|x| applied to the accumulators with the
    RMSNorm gain folded into the projection, softmax over pixels against the exact row maximum with partial sums over
    pixel ranges, memory key/values added once, to_out folded into a per-sample operand Mb) against the oracle's
    LinearAttention + skip (ddpm.py:205-238, :449)."""
    g = torch.Generator().manual_seed(3)
    b, c, h = 2, 64, 16
    n = h * h
    dd = torch.float64
    x = torch.randn(b, c, h, h, generator=g, dtype=dd)
    sd = {"a.norm.g": 1 + 0.2 * torch.randn(1, c, 1, 1, generator=g, dtype=dd),
          "a.to_qkv.weight": torch.randn(384, c, 1, 1, generator=g, dtype=dd) / c ** 0.5,
          "a.mem_kv": torch.randn(2, 4, 32, 4, generator=g, dtype=dd),
          "a.to_out.0.weight": torch.randn(c, 128, 1, 1, generator=g, dtype=dd) / 128 ** 0.5,
          "a.to_out.0.bias": 0.1 * torch.randn(c, generator=g, dtype=dd),
          "a.to_out.1.g": 1 + 0.2 * torch.randn(1, c, 1, 1, generator=g, dtype=dd)}
    want = O.linear_attention(sd, "a", x, O.Emu(None)) + x
    # ---- the kernels' formulation
    X = x.permute(0, 2, 3, 1).reshape(b, n, c)                                  # NHWC rows
    Wf = sd["a.to_qkv.weight"].view(384, c) * sd["a.norm.g"].view(1, c) * c ** 0.5   # b200dm_pack_linattn_qkv
    rn = 1.0 / X.norm(dim=2).clamp_min(1e-12)                                   # pass 0
    q = (X @ Wf[:128].T) * rn[..., None]
    k = (X @ Wf[128:256].T) * rn[..., None]
    v = (X @ Wf[256:].T) * rn[..., None]
    mk, mv = sd["a.mem_kv"][0].reshape(128, 4), sd["a.mem_kv"][1].reshape(128, 4)
    split = 4
    kmax = torch.stack([k[:, i * n // split:(i + 1) * n // split].max(dim=1).values for i in range(split)]).max(0).values
    m = torch.maximum(kmax, mk.max(dim=1).values[None])                         # [b, 128]
    ctx = torch.zeros(b, 4, 32, 32, dtype=dd)
    ssum = torch.zeros(b, 128, dtype=dd)
    for i in range(split):                                                      # pass 1, partial sums per pixel range
        sl = slice(i * n // split, (i + 1) * n // split)
        p = torch.exp(k[:, sl] - m[:, None])
        ssum += p.sum(dim=1)
        ctx += torch.einsum("bnhd,bnhe->bhde", p.view(b, -1, 4, 32), v[:, sl].reshape(b, -1, 4, 32))
    pm = torch.exp(mk[None] - m[..., None])                                     # mid: memory key/values
    ssum += pm.sum(dim=2)
    ctx += torch.einsum("bhdj,hej->bhde", pm.view(b, 4, 32, 4), mv.view(4, 32, 4))
    ctxn = ctx * (32 ** -0.5) / ssum.view(b, 4, 32, 1)
    Wout = sd["a.to_out.0.weight"].view(c, 4, 32)
    Mb = torch.einsum("che,bhde->bchd", Wout, ctxn).reshape(b, c, 128)          # [C][(h, d)]
    qs = q.view(b, n, 4, 32).softmax(dim=3).reshape(b, n, 128)                  # pass 2
    y = qs @ Mb.transpose(1, 2) + sd["a.to_out.0.bias"]
    y = y / y.norm(dim=2, keepdim=True).clamp_min(1e-12) * sd["a.to_out.1.g"].view(1, 1, c) * c ** 0.5 + X
    got = y.reshape(b, h, h, c).permute(0, 3, 1, 2)
    assert (got - want).abs().max().item() < 1e-10


def test_softmax_attention_backward_closed_form():
    """attn_bwd_tc_kernel (attention.cu) against Attention + Attend's math branch (ddpm.py:255-271,
    models/modules/attend.py:111-126) in fp64: P = softmax(scale Q K^T), dP = dO V^T,
    dS = P .* (dP - rowsum(P .* dP)), dQ = scale dS K, dK = scale dS^T Q, dV = P^T dO (memory key/values are
    the first four rows of K and V)."""
    g = torch.Generator().manual_seed(17)
    n, nm, d, scale = 9, 4, 8, 8 ** -0.5
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    q, k, v = rnd(n, d).requires_grad_(True), rnd(n, d).requires_grad_(True), rnd(n, d).requires_grad_(True)
    mk, mv = rnd(nm, d).requires_grad_(True), rnd(nm, d).requires_grad_(True)
    do = rnd(n, d)
    K, V = torch.cat((mk, k)), torch.cat((mv, v))
    out = (q @ K.t() * scale).softmax(-1) @ V
    out.backward(do)
    with torch.no_grad():
        P = (q @ K.t() * scale).softmax(-1)
        dP = do @ V.t()
        dS = P * (dP - (P * dP).sum(-1, keepdim=True))
        dQ, dK, dV = scale * dS @ K, scale * dS.t() @ q, P.t() @ do
        assert (dQ - q.grad).abs().max().item() < 1e-12
        assert (dK[nm:] - k.grad).abs().max().item() < 1e-12 and (dK[:nm] - mk.grad).abs().max().item() < 1e-12
        assert (dV[nm:] - v.grad).abs().max().item() < 1e-12 and (dV[:nm] - mv.grad).abs().max().item() < 1e-12


def test_bench_clock_sampler_reports_only_the_timed_region():
    """bench.py's nvidia-smi sampler: samples before mark() are ignored, throttle reasons inside the region are kept."""
    import importlib.util
    import time
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)

    class _Proc:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            pass
    c = b.ClockSampler(0)
    c.proc = _Proc()
    t = time.perf_counter()
    idle = ["1200", "1965", "100", "Not Active", "Not Active", "Not Active", "Not Active", "3"]
    capped = ["1950", "1965", "700", "Not Active", "Not Active", "Not Active", "Active", "99"]
    full = ["1965", "1965", "700", "Not Active", "Not Active", "Not Active", "Not Active", "20"]
    c.rows = [(t - 1, idle), (t + 1, capped), (t + 2, full)]
    c.t_mark = t
    r = c.stop()
    assert r["sm_mhz"] == 1957.5 and r["sm_max_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"]
    assert r["samples_under_load"] == 2 and r["samples"] == 3


def test_train_entry_args_loader_and_datamodule(tmp_path):
    """N3 host logic (reference train.py:24-90, utils/loader.py:47-86, data/datamodule.py:33): argument surface,
    precision mapping, config sanity checks, per-rank batch rule and disjoint rank shards."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("b200_train", os.path.join(PKG, "train.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    cfg_path = os.path.join(PKG, "configs", "diffusion", "ddpm.json")
    a = tr.parse_args(["--config_path", cfg_path, "--max_steps", "20", "--precision", "bf16-mixed", "--gpus", "8",
                       "--accumulate_grad_batches", "2"])
    assert a.max_steps == 20 and a.gpus == 8 and a.accumulate_grad_batches == 2 and a.check_val_every_n_epoch == 5
    assert tr.parse_args(["--config_path", cfg_path]).max_epochs == 1000
    assert tr.precision_of(None) == "fp32" and tr.precision_of("32-true") == "fp32"
    assert tr.precision_of("bf16-mixed") == "bf16"
    with pytest.raises(ValueError):
        tr.precision_of("16-mixed")
    from utils.loader import load_config, load_model
    cfg = load_config(cfg_path)
    assert cfg["model"]["name"] == "DDPM" and cfg["model"]["args"]["img_size"] == cfg["dataset"]["img_size"]
    bad = tmp_path / "bad.json"
    cfg2 = json.loads(json.dumps(cfg))
    cfg2["dataset"]["img_size"] = 64
    bad.write_text(json.dumps(cfg2))
    with pytest.raises(ValueError, match="img_size"):
        load_config(str(bad))
    with pytest.raises(FileNotFoundError):
        load_config(str(tmp_path / "missing.json"))
    with pytest.raises(ValueError, match="Failed to import"):
        load_model({"name": "NoSuchModel", "args": {}})
    from data.datamodule import DataModule
    dms = [DataModule(**cfg["dataset"], num_images=256, world_size=2, rank=r, pin_memory=False) for r in range(2)]
    assert dms[0].batch_size == cfg["dataset"]["batch_size"] // 2            # datamodule.py:33
    seen = []
    for dm in dms:
        batches = list(dm.train_dataloader(epoch=3))
        assert len(batches) == dm.steps_per_epoch() and len(batches) > 0
        x, y = batches[0]
        assert x.shape == (16, 3, 32, 32) and x.dtype == torch.float32 and y.shape == (16,)
        assert -1.0 <= x.min().item() and x.max().item() <= 1.0
        seen.append(torch.cat([b[0] for b in batches]).flatten(1).sum(1))
    assert not set(seen[0].tolist()) & set(seen[1].tolist())                 # disjoint shards of one permutation
    again = torch.cat([b[0] for b in dms[0].train_dataloader(epoch=3)]).flatten(1).sum(1)
    assert torch.equal(again, seen[0])                                       # seeded: same epoch, same order
