"""End-to-end GPU parity: b200dm.Unet / GaussianDiffusion / DDPM against the oracle (oracle/ddpm_oracle.py,
itself pinned to the reference by tests/golden) and against the committed golden fixtures.

Tolerances (BASELINE.json north_star / BASELINE.md §3):
  fp32 mode : UNet output and loss rel <= 1e-4, gradients rel-L2 <= 1e-3 per tensor
  bf16 mode : UNet output rel-L2 <= 1e-2 and loss rel <= 1e-2 vs the fp32 reference (every case, 1-channel included).
              The yardstick is the reference's own bf16 arithmetic: the oracle run on this GPU under
              torch.autocast(bf16) (tests/_refcuda.py); its distance to fp32 is reported beside ours.
  samplers  : DDIM / DDPM images in [0,1]: fp32 PSNR >= 80 dB; bf16 PSNR >= 40 dB and L-inf <= 0.05.  One fixture
              (c3s64: pred_noise + linear schedule, 4 DDIM steps from t=999) multiplies the UNet error by
              sqrt(1/abar - 1) ~ 158 before the clamp, so NO bf16 evaluation meets 40 dB there, the reference's own
              included; `sampler_gate` therefore demands the absolute bound whenever the autocast reference meets
              it, and otherwise "at least as close to fp32 as the autocast reference is" (-1 dB, x1.25 L-inf).
"""
import json
import math
import os

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O
from _refcuda import cuda_oracle, precision_ctx

pytestmark = pytest.mark.gpu
DEV = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "parity_report.jsonl")


def report(**kw):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def psnr(a, b):
    mse = (a.detach().float().cpu() - b.detach().float().cpu()).pow(2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


def sampler_gate(p, li, yard_p, yard_li, what):
    """bf16 sampler bound: absolute (>= 40 dB, L-inf <= 0.05) whenever the reference's own bf16-autocast run meets
    it; otherwise no further from the fp32 reference than that run is."""
    if yard_p >= 40 and yard_li <= 0.05:
        assert p >= 40 and li <= 0.05, (what, p, li, yard_p, yard_li)
    else:
        assert p >= yard_p - 1.0 and li <= 1.25 * yard_li + 0.01, (what, p, li, yard_p, yard_li)


def seeded_inputs(b, c, s, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, c, s, s, generator=g)
    t = torch.randint(0, 1000, (b,), generator=g)
    noise = torch.randn(b, c, s, s, generator=g)
    init = torch.randn(b, c, s, s, generator=g)
    return x, t, noise, init


_SD = {}


def synth(ch):
    if ch not in _SD:
        _SD[ch] = O.synth_state_dict(64, ch, seed=10)
    return _SD[ch]


def build(ch, size, precision, objective="pred_v", schedule="sigmoid", timesteps=1000, sampling=4, **kw):
    from b200dm import GaussianDiffusion, Unet
    unet = Unet(dim=64, channels=ch, precision=precision, **kw)
    unet.load_reference_state_dict(synth(ch))
    gd = GaussianDiffusion(unet, img_size=size, timesteps=timesteps, sampling_timesteps=sampling,
                           objective=objective, beta_schedule=schedule)
    return unet, gd


CASES = {"c3s32": (3, 32, 2, "pred_v", "sigmoid"), "c1s32": (1, 32, 2, "pred_v", "sigmoid"),
         "c3s64": (3, 64, 1, "pred_noise", "linear"), "c3s32_x0": (3, 32, 2, "pred_x0", "cosine")}


def test_state_dict_surface():
    from b200dm import Unet
    inv = json.load(open(os.path.join(GOLD, "state_dict_inventory.json")))
    unet = Unet(dim=64, channels=3)
    sd = unet.state_dict()
    assert [[k, list(v.shape)] for k, v in sd.items()] == inv["3"]
    assert [n for n, _ in unet.named_parameters()] == [k for k, _ in inv["3"]]
    assert sum(p.numel() for p in unet.parameters()) == 35719555
    ref = synth(3)
    unet.load_state_dict(ref)                                 # nn.Module path
    out = unet.state_dict()
    assert all(torch.equal(out[k].cpu(), ref[k]) for k in ref)
    assert unet.channels == 3 and unet.out_dim == 3 and unet.downsample_factor == 8
    assert not unet.self_condition and not unet.random_or_learned_sinusoidal_cond
    with pytest.raises(NotImplementedError):
        Unet(dim=32)
    with pytest.raises(AssertionError):
        unet(torch.zeros(1, 3, 28, 28, device=DEV), torch.zeros(1, dtype=torch.long, device=DEV))


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_forward_vs_golden_and_oracle(name, precision):
    ch, s, b, objective, sched = CASES[name]
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    x, t, noise, init = seeded_inputs(b, ch, s)
    unet, gd = build(ch, s, precision, objective, sched)
    with torch.no_grad():
        out = unet((x * 2 - 1).to(DEV), t.to(DEV))
    ref = torch.from_numpy(gold["unet_out"])
    r = rel(out, ref)
    report(test="unet_fwd", case=name, precision=precision, rel_vs_reference=r)
    if precision == "fp32":
        assert r <= 1e-4, r
    else:
        assert r <= 1e-2, r
        orc = cuda_oracle(ch, s)
        with torch.no_grad(), precision_ctx("autocast"):
            ac = orc.model((x * 2 - 1).to(DEV), t.to(DEV)).float()
        r2, r3 = rel(out, ac), rel(ac, ref)
        report(test="unet_fwd_autocast", case=name, rel_vs_autocast=r2, autocast_vs_reference=r3)
        assert r2 <= 1e-2 + r3, (r2, r3)


@pytest.mark.parametrize("s", [16, 24, 40])
def test_bf16_sizes_outside_the_tensor_core_tiling(s):
    """Images whose levels the tcgen05 kernels do not tile (2x2 innermost level, non-power-of-two sizes) run those levels
    on the SIMT kernels in bf16: loss, UNet output and gradients against the oracle, DDIM-4 against its samples."""
    ch, b = 3, 2
    x, t, noise, init = seeded_inputs(b, ch, s)
    unet, gd = build(ch, s, "bf16")
    orc = cuda_oracle(ch, s, grads=True, sampling_timesteps=4)
    xd, td, nd = x.to(DEV), t.to(DEV), noise.to(DEV)
    loss = gd.p_losses(xd, td, noise=nd, _normalize=True)
    loss.backward()
    with precision_ctx("fp32"):
        want, _, out_ref, _ = orc.p_losses(xd * 2 - 1, td, nd, return_parts=True)
    want.backward()
    plan = unet._plan(b, s, training=True)
    r_out = rel(plan.out, out_ref)
    num = sum((p.grad.double().cpu() - orc.sd[n].grad.double().cpu()).pow(2).sum().item() for n, p in unet.named_parameters())
    den = sum(orc.sd[n].grad.double().pow(2).sum().item() for n, _ in unet.named_parameters())
    r_grad = (num / den) ** 0.5
    with torch.no_grad(), precision_ctx("fp32"):
        img_ref = orc.sample(init.to(DEV))
    img = gd.sample(batch_size=b, init_noise=init.to(DEV))
    p = psnr(img, img_ref)
    report(test="bf16_odd_sizes", size=s, loss_rel=abs(loss.item() - want.item()) / abs(want.item()), out_rel=r_out,
           grad_rel=r_grad, ddim4_psnr=p)
    assert abs(loss.item() - want.item()) <= 1e-2 * abs(want.item())
    assert r_out <= 1.25e-2 and r_grad <= 5e-2 and p >= 38, (r_out, r_grad, p)


@pytest.mark.parametrize("name", ["c3s32", "c1s32", "c3s64", "c3s32_x0"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_loss_and_all_gradients(name, precision):
    ch, s, b, objective, sched = CASES[name]
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    x, t, noise, init = seeded_inputs(b, ch, s)
    unet, gd = build(ch, s, precision, objective, sched)
    # GaussianDiffusion.forward semantics with injected t / noise: normalize + p_losses
    loss = gd.p_losses(x.to(DEV), t.to(DEV), noise=noise.to(DEV), _normalize=True)
    loss.backward()
    gl = float(gold["loss"])
    lrel = abs(loss.item() - gl) / abs(gl)
    # oracle gradients (fp32 autograd on the CPU restatement)
    sd = {k: v.clone().requires_grad_(True) for k, v in synth(ch).items()}
    orc = O.DiffusionOracle(sd, img_size=s, channels=ch, objective=objective, beta_schedule=sched)
    orc.forward(x, t, noise).backward()
    worst, worst_name, num, den = 0.0, "", 0.0, 0.0
    grads = {n: p.grad for n, p in unet.named_parameters()}
    for k, _ in O.unet_param_spec(64, ch):
        g, gr = grads[k].detach().float().cpu(), sd[k].grad
        assert g.shape == gr.shape, k
        e = (g - gr).norm().item()
        num += e * e
        den += gr.norm().item() ** 2
        rk = e / max(gr.norm().item(), 1e-12)
        if rk > worst and gr.norm().item() > 1e-6:
            worst, worst_name = rk, k
    total = math.sqrt(num / den)
    report(test="loss_grads", case=name, precision=precision, loss_rel=lrel, grad_rel_total=total,
           grad_rel_worst=worst, worst_tensor=worst_name)
    gn = np.array([grads[k].norm().item() for k, _ in O.unet_param_spec(64, ch)], np.float32)
    if precision == "fp32":
        assert lrel <= 1e-4, lrel
        assert total <= 1e-3 and worst <= 5e-3, (total, worst, worst_name)
        np.testing.assert_allclose(gn, gold["grad_norms"], rtol=5e-3, atol=1e-7)
    else:
        assert lrel <= 1e-2, lrel
        # yardstick: the reference's bf16-autocast gradients on this GPU against the same fp32 gradients
        orc = cuda_oracle(ch, s, grads=True, objective=objective, beta_schedule=sched)
        with precision_ctx("autocast"):
            la = orc.forward(x.to(DEV), t.to(DEV), noise.to(DEV))
        la.backward()
        ynum = yworst = 0.0
        for k, _ in O.unet_param_spec(64, ch):
            gr, ga = sd[k].grad, orc.sd[k].grad.float().cpu()
            e = (ga - gr).norm().item()
            ynum += e * e
            if gr.norm().item() > 1e-6:
                yworst = max(yworst, e / gr.norm().item())
        ytotal = math.sqrt(ynum / den)
        report(test="loss_grads_autocast_yardstick", case=name, grad_rel_total=ytotal, grad_rel_worst=yworst)
        assert total <= max(2e-2, 1.5 * ytotal) and worst <= max(5e-2, 2.0 * yworst), \
            (total, worst, worst_name, ytotal, yworst)


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_samplers_vs_golden(name, precision):
    ch, s, b, objective, sched = CASES[name]
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    x, t, noise, init = seeded_inputs(b, ch, s)
    unet, gd = build(ch, s, precision, objective, sched)
    img = gd.sample(batch_size=b, init_noise=init.to(DEV))
    ref = torch.from_numpy(gold["ddim4"])
    p, linf = psnr(img, ref), (img.cpu() - ref).abs().max().item()
    report(test="ddim4", case=name, precision=precision, psnr=p, linf=linf)
    if precision == "fp32":
        assert p >= 80 and linf <= 1e-3, (p, linf)
    else:
        orc = cuda_oracle(ch, s, sampling_timesteps=4, objective=objective, beta_schedule=sched)
        with torch.no_grad(), precision_ctx("autocast"):
            ac = orc.sample(init.to(DEV)).float().cpu()
        yp, yl = psnr(ac, ref), (ac - ref).abs().max().item()
        report(test="ddim4_autocast_yardstick", case=name, psnr=yp, linf=yl)
        sampler_gate(p, linf, yp, yl, "ddim4 " + name)
    for tt in (999, 500, 0):
        im, x0 = gd.p_sample(init.to(DEV), tt, noise=noise.to(DEV))
        d = (im.cpu() - torch.from_numpy(gold[f"p_sample_{tt}"])).abs()
        d0 = (x0.cpu() - torch.from_numpy(gold[f"p_sample_{tt}_x0"])).abs()
        report(test="p_sample", case=name, precision=precision, t=tt, linf=d.max().item(),
               linf_x0=d0.max().item(), mean=d.mean().item(), mean_x0=d0.mean().item())
        if precision == "fp32":
            assert d.max().item() <= 2e-3 and d0.max().item() <= 2e-3, (tt, d.max().item(), d0.max().item())
        else:
            # x0 = sqrt(1/abar) x - ... amplifies the UNet error by up to ~160x at t=999 (then clamps): bounded by
            # what the reference's own bf16-autocast step does on the same inputs (x1.5, + 0.01 absolute)
            with torch.no_grad(), precision_ctx("autocast"):
                ai, a0 = orc.p_sample(init.to(DEV), tt, noise.to(DEV))
            yd = (ai.float().cpu() - torch.from_numpy(gold[f"p_sample_{tt}"])).abs()
            yd0 = (a0.float().cpu() - torch.from_numpy(gold[f"p_sample_{tt}_x0"])).abs()
            report(test="p_sample_autocast_yardstick", case=name, t=tt, linf=yd.max().item(),
                   linf_x0=yd0.max().item(), mean=yd.mean().item(), mean_x0=yd0.mean().item())
            assert d.mean().item() <= 1.5 * yd.mean().item() + 1e-3, (tt, d.mean().item(), yd.mean().item())
            assert d0.mean().item() <= 1.5 * yd0.mean().item() + 1e-3, (tt, d0.mean().item(), yd0.mean().item())
            assert d.max().item() <= 1.5 * yd.max().item() + 0.01, (tt, d.max().item(), yd.max().item())
            assert d0.max().item() <= 1.5 * yd0.max().item() + 0.01, (tt, d0.max().item(), yd0.max().item())
    # model_predictions (ddpm.py:707-734) against the reference's values
    mp = gd.model_predictions(init.to(DEV), t.to(DEV), clip_x_start=True, rederive_pred_noise=True)
    assert type(mp).__name__ == "ModelPrediction" and mp._fields == ("pred_noise", "pred_x_start")
    gn_, g0_ = torch.from_numpy(gold["mp_noise"]), torch.from_numpy(gold["mp_x0"])
    dn, d0 = (mp.pred_noise.float().cpu() - gn_).abs(), (mp.pred_x_start.float().cpu() - g0_).abs()
    report(test="model_predictions", case=name, precision=precision, noise_linf=dn.max().item(),
           x0_linf=d0.max().item(), noise_rel=rel(mp.pred_noise, gn_), x0_rel=rel(mp.pred_x_start, g0_))
    if precision == "fp32":
        assert rel(mp.pred_noise, gn_) <= 1e-4 and d0.max().item() <= 2e-3, (rel(mp.pred_noise, gn_), d0.max().item())
    else:
        with torch.no_grad(), precision_ctx("autocast"):
            amp = orc.model_predictions(init.to(DEV), t.to(DEV), clip_x_start=True, rederive_pred_noise=True)
        yn, y0 = rel(amp.pred_noise.float(), gn_), rel(amp.pred_x_start.float(), g0_)
        report(test="model_predictions_autocast_yardstick", case=name, noise_rel=yn, x0_rel=y0)
        assert rel(mp.pred_noise, gn_) <= max(1e-2, 1.25 * yn), (rel(mp.pred_noise, gn_), yn)
        assert rel(mp.pred_x_start, g0_) <= max(1e-2, 1.25 * y0), (rel(mp.pred_x_start, g0_), y0)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ddpm_ancestral_chain_vs_golden(precision):
    gold = np.load(os.path.join(GOLD, "golden_ddpm8.npz"))
    noises = [torch.from_numpy(n).to(DEV) for n in gold["noises"]]
    step = {t: noises[1 + (7 - t)] for t in range(7, 0, -1)}
    unet, gd = build(3, 32, precision, timesteps=8, sampling=None)
    assert not gd.is_ddim_sampling
    img = gd.sample(batch_size=2, init_noise=noises[0], step_noise=lambda t: step[t])
    p = psnr(img, torch.from_numpy(gold["img"]))
    report(test="ddpm8", precision=precision, psnr=p)
    assert p >= (80 if precision == "fp32" else 40), p


def test_config1_train_step_and_ddim10_fullsize():
    """BASELINE config 1 at its real size: [64,1,32,32], one training step + DDIM-10, CPU oracle vs B200."""
    ch, s, b = 1, 32, 64
    x, t, noise, init = seeded_inputs(b, ch, s, seed=10)
    for precision in ("fp32", "bf16"):
        unet, gd = build(ch, s, precision, sampling=10)
        loss = gd.p_losses(x.to(DEV), t.to(DEV), noise=noise.to(DEV), _normalize=True)
        loss.backward()
        orc = O.DiffusionOracle(synth(ch), img_size=s, channels=ch, sampling_timesteps=10)
        with torch.no_grad():
            lref = orc.forward(x, t, noise).item()
        lrel = abs(loss.item() - lref) / abs(lref)
        img = gd.sample(batch_size=b, init_noise=init.to(DEV))
        with torch.no_grad():
            iref = orc.sample(init)
        p, linf = psnr(img, iref), (img.cpu() - iref).abs().max().item()
        report(test="config1", precision=precision, loss_rel=lrel, ddim10_psnr=p, ddim10_linf=linf)
        assert lrel <= (1e-4 if precision == "fp32" else 1e-2)
        assert p >= (80 if precision == "fp32" else 40) and linf <= (1e-3 if precision == "fp32" else 0.05)


def test_batch_shard_invariance_and_philox_sampling():
    """Sampling shards by batch with no communication: the union of 2 shards equals the
    1-shard result because Philox is keyed by the global element index."""
    unet, gd = build(3, 32, "bf16", sampling=3)
    full = gd.sample_shard(8, 0, 1, seed=42)
    halves = torch.cat([gd.sample_shard(8, r, 2, seed=42) for r in range(2)], 0)
    assert full.shape == (8, 3, 32, 32) and torch.isfinite(full).all()
    assert (full - halves).abs().max().item() <= 2e-3
    assert 0.0 <= full.min().item() and full.max().item() <= 1.0
    a = gd.sample(batch_size=4)
    assert a.shape == (4, 3, 32, 32) and torch.isfinite(a).all()


def test_cuda_graph_replay_matches_eager():
    x, t, _, _ = seeded_inputs(4, 3, 32)
    outs = []
    for graph in (False, True):
        unet, gd = build(3, 32, "bf16", cuda_graph=graph)
        with torch.no_grad():
            for _ in range(4):            # 3rd call onwards replays the captured graph
                out = unet((x * 2 - 1).to(DEV), t.to(DEV))
        outs.append(out)
    assert torch.equal(outs[0], outs[1])


def test_ddpm_module_training_loop_fused_adam_and_ema():
    from b200dm import DDPM
    torch.manual_seed(0)
    m = DDPM(img_channels=3, img_size=32, dim=64, diffusion_timesteps=1000, sampling_timesteps=5,
             lr=2e-4, betas=(0.9, 0.99), ema_update_every=2, ema_decay=0.995)
    m.train()
    opt = m.configure_optimizers()
    g = torch.Generator().manual_seed(3)
    data = torch.rand(8, 3, 32, 32, generator=g).to(DEV)
    batch = (data, torch.zeros(8, dtype=torch.long, device=DEV))
    unet = m.ema.model.model
    # reference optimiser on a copy of the parameters, fed with the same gradients
    ref_p = unet.arena.flat.clone().requires_grad_(True)
    ref_opt = torch.optim.Adam([ref_p], lr=2e-4, betas=(0.9, 0.99))
    losses = []
    for step in range(6):
        opt.zero_grad()
        loss = m.training_step(batch)
        assert loss.requires_grad and loss.dim() == 0
        loss.backward()
        ref_p.grad = unet.arena.gflat.clone()
        opt.step()
        ref_opt.step()
        m.on_train_batch_end(None, batch, step)
        losses.append(loss.item())
    assert all(math.isfinite(v) for v in losses)
    assert (unet.arena.flat - ref_p.detach()).abs().max().item() < 1e-6
    # EMA: steps 0,2,4 <= update_after_step -> plain copies of the online weights at those steps
    assert m.ema._step_host == 6
    m.eval()
    with torch.no_grad():
        v = m.validation_step(batch)
    assert math.isfinite(v.item())
    imgs = m.sample(batch_size=4)
    assert imgs.shape == (4, 3, 32, 32) and torch.isfinite(imgs).all()
    report(test="ddpm_module", losses=losses)


def test_optimizer_overlapped_with_backward():
    """FusedAdam(overlap_with_backward=True): per-bucket Adam + weight re-pack on a second stream behind backward.
    Checked against torch.optim.Adam fed with the same gradients (the weight-gradient reductions are not
    bit-reproducible run to run, so two separate trainings cannot be compared bit for bit) and against a forced
    full re-pack of the final weights."""
    from b200dm import DDPM
    torch.manual_seed(0)
    m = DDPM(img_channels=3, img_size=32, dim=64, lr=2e-4, ema_update_every=2, overlap_optimizer=True)
    m.train()
    opt = m.configure_optimizers()
    assert opt.overlap
    g = torch.Generator().manual_seed(5)
    data = torch.rand(8, 3, 32, 32, generator=g).to(DEV)
    batch = (data, torch.zeros(8, dtype=torch.long, device=DEV))
    unet = m.ema.model.model
    ref_p = unet.arena.flat.clone().requires_grad_(True)
    ref_opt = torch.optim.Adam([ref_p], lr=2e-4, betas=(0.9, 0.99))
    for step in range(6):                 # the backward is graph-replayed from the 3rd step on
        opt.zero_grad()
        loss = m.training_step(batch)
        loss.backward()
        assert len(opt._done) == 9        # every bucket was applied behind backward
        ref_p.grad = unet.arena.gflat.clone()
        opt.step()
        ref_opt.step()
        m.on_train_batch_end(None, batch, step)
        assert math.isfinite(loss.item())
    assert opt.step_count == 6
    assert (unet.arena.flat - ref_p.detach()).abs().max().item() < 1e-6
    pack = unet._pack
    assert pack.version == unet.arena.version          # the next forward will not re-pack
    got = [pack.fbuf.clone(), pack.tbuf.clone(), pack.stem.clone()] + [pack.up[k].clone() for k in sorted(pack.up)]
    assert len(pack.up) == 3 and pack.up_version == unet.arena.version
    pack.refresh(force=True)
    now = [pack.fbuf, pack.tbuf, pack.stem] + [pack.up[k] for k in sorted(pack.up)]
    assert all(torch.equal(a, b) for a, b in zip(got, now))


def test_reference_configs_load_through_loader_convention():
    import importlib
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lightning-generative-models_b200")
    for cfg_name in ("ddpm", "ddim", "ddpm_flowers102"):
        cfg = json.load(open(os.path.join(pkg, "configs", "diffusion", cfg_name + ".json")))
        assert cfg["model"]["args"]["img_size"] == cfg["dataset"]["img_size"]
        mod = importlib.import_module(f"models.generative.diffusion.{cfg['model']['name'].lower()}")
        model = getattr(mod, cfg["model"]["name"])(**cfg["model"]["args"])
        assert model.channels == 3 and model.img_size == cfg["model"]["args"]["img_size"]
        assert model.ema.model.is_ddim_sampling == (cfg_name == "ddim")
        del model
        torch.cuda.empty_cache()
