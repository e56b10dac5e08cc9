"""GPU parity at the BASELINE.json configuration sizes (VERDICT r1 "next round" item 1).

The persistent-CTA tile loops, split-K over 148 CTAs, multi-wave grids and the CUDA-graph replay only see their
real tile counts at these sizes, so every config is checked here through the public API against the oracle run on
the same GPU (tests/_refcuda.py): in true fp32 and under torch.autocast(bf16) — the reference's own bf16 path.

Gates (north_star / BASELINE.md §3, written out):
  fp32 mode : UNet output / loss rel <= 1e-4 vs the fp32 reference; gradient arena total rel-L2 <= 1e-3.
  bf16 mode : UNet output rel-L2 <= 1e-2 and loss rel <= 1e-2 vs the fp32 reference, AND no further from it than
              the reference's own bf16-autocast evaluation is (x1.1) — the two bf16 evaluations round at different
              points, so their errors are independent and their mutual distance is ~sqrt(a^2+b^2); it is reported,
              and bounded by the triangle 1e-2 + d(autocast, fp32).
  samplers  : images in [0,1]: fp32 PSNR >= 80 dB; bf16 PSNR >= 40 dB and L-inf <= 0.05 vs the autocast reference
              AND vs the fp32 reference.
"""
import math

import pytest
import torch

from _refcuda import DEV, cuda_oracle, grad_distance, linf, precision_ctx, psnr, rel, report, synth
from oracle import ddpm_oracle as O

pytestmark = pytest.mark.gpu


def build(ch, size, precision, *, timesteps=1000, sampling=None, objective="pred_v", schedule="sigmoid", **kw):
    from b200dm import GaussianDiffusion, Unet
    unet = Unet(dim=64, channels=ch, precision=precision, **kw)
    unet.load_reference_state_dict(synth(ch))
    gd = GaussianDiffusion(unet, img_size=size, timesteps=timesteps, sampling_timesteps=sampling,
                           objective=objective, beta_schedule=schedule)
    return unet, gd


def inputs(b, c, s, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, c, s, s, generator=g)
    t = torch.randint(0, 1000, (b,), generator=g)
    noise = torch.randn(b, c, s, s, generator=g)
    init = torch.randn(b, c, s, s, generator=g)
    return x.to(DEV), t.to(DEV), noise.to(DEV), init.to(DEV)


def reference_loss_and_grads(ch, s, x, t, noise, mode):
    orc = cuda_oracle(ch, s, grads=True)
    with precision_ctx(mode):
        loss, xt, out, target = orc.p_losses(x * 2 - 1, t, noise, return_parts=True)
    loss.backward()
    grads = {k: v.grad.detach().float() for k, v in orc.sd.items()}
    return loss.item(), out.detach().float(), grads


def _train_parity(tag, ch, s, b, steps_replayed=3):
    """Loss, UNet output and the whole gradient arena of one training step at batch `b`, in both precisions.
    The step is repeated so that the third pass runs the CUDA-graph replay (same inputs: identical results up to the
    fp32 atomics of the split-K weight gradients)."""
    x, t, noise, _ = inputs(b, ch, s, seed=10)
    spec = O.unet_param_spec(64, ch)
    l32, o32, g32 = reference_loss_and_grads(ch, s, x, t, noise, "fp32")
    lac, oac, gac = reference_loss_and_grads(ch, s, x, t, noise, "autocast")
    yard_out, yard_loss = rel(oac, o32), abs(lac - l32) / abs(l32)
    yard_grad, _, _ = grad_distance(gac, g32, spec)
    for precision in ("fp32", "bf16"):
        unet, gd = build(ch, s, precision)
        for it in range(steps_replayed):
            unet.zero_grad()
            loss = gd.p_losses(x, t, noise=noise, _normalize=True)
            loss.backward()
        plan = unet._plan(b, s, training=True)
        assert plan.graph_fwd is not None and plan.graph_bwd is not None     # the replayed path was checked
        out = plan.out.detach().float()
        grads = {n: p.grad for n, p in unet.named_parameters()}
        lrel = abs(loss.item() - l32) / abs(l32)
        orel, orel_ac = rel(out, o32), rel(out, oac)
        gtot, gworst, gname = grad_distance(grads, g32, spec)
        gtot_ac, _, _ = grad_distance(grads, gac, spec)
        report(test="config_train", config=tag, precision=precision, batch=b, loss_rel_vs_fp32=lrel,
               out_rel_vs_fp32=orel, out_rel_vs_autocast=orel_ac, grad_rel_vs_fp32=gtot,
               grad_rel_vs_autocast=gtot_ac, grad_worst=gworst, grad_worst_tensor=gname,
               ref_autocast_vs_fp32=dict(out=yard_out, loss=yard_loss, grad=yard_grad))
        assert math.isfinite(loss.item())
        if precision == "fp32":
            assert lrel <= 1e-4 and orel <= 1e-4, (lrel, orel)
            assert gtot <= 1e-3, (gtot, gworst, gname)
        else:
            assert lrel <= 1e-2, lrel
            assert orel <= 1e-2, (orel, yard_out)
            assert orel <= 1.1 * yard_out + 1e-4, (orel, yard_out)           # no worse than the reference's bf16 mode
            assert orel_ac <= 1e-2 + yard_out, (orel_ac, yard_out)
            assert gtot <= max(3e-2, 1.5 * yard_grad), (gtot, yard_grad, gworst, gname)
        del unet, gd, plan
        torch.cuda.empty_cache()


def test_config2_train_step_batch128_32px():
    """BASELINE configs[1]: 3x32x32, batch 128."""
    _train_parity("C2", 3, 32, 128)


def test_config5_train_step_batch64_64px():
    """BASELINE configs[4] per GPU: 3x64x64, 64 images."""
    _train_parity("C5", 3, 64, 64)


def _sample_reference(ch, s, init, mode, **kw):
    orc = cuda_oracle(ch, s, **kw)
    with torch.no_grad(), precision_ctx(mode):
        return orc.sample(init).float()


def test_config3_ddim50_64px_shard_of_32():
    """BASELINE configs[2]: DDIM-50, 3x64x64, reference defaults (pred_v, sigmoid, eta=0); 32 images = the 8-way
    shard of the batch of 256 (each image's trajectory is independent), all 50 steps."""
    ch, s, b = 3, 64, 32
    _, _, _, init = inputs(b, ch, s, seed=33)
    r32 = _sample_reference(ch, s, init, "fp32", sampling_timesteps=50)
    rac = _sample_reference(ch, s, init, "autocast", sampling_timesteps=50)
    yard = dict(psnr=psnr(rac, r32), linf=linf(rac, r32))
    for precision in ("fp32", "bf16"):
        unet, gd = build(ch, s, precision, sampling=50)
        img = gd.sample(batch_size=b, init_noise=init)
        p32, l32, pac, lac = psnr(img, r32), linf(img, r32), psnr(img, rac), linf(img, rac)
        report(test="config_ddim50", config="C3", precision=precision, batch=b, psnr_vs_fp32=p32, linf_vs_fp32=l32,
               psnr_vs_autocast=pac, linf_vs_autocast=lac, ref_autocast_vs_fp32=yard)
        assert torch.isfinite(img).all() and 0.0 <= img.min().item() and img.max().item() <= 1.0
        if precision == "fp32":
            assert p32 >= 80 and l32 <= 1e-3, (p32, l32)
        else:
            assert p32 >= 40 and l32 <= 0.05, (p32, l32, yard)
            assert pac >= 40 and lac <= 0.05, (pac, lac, yard)
        del unet, gd
        torch.cuda.empty_cache()


def test_config4_ancestral_batch1024_32px_short_chain():
    """BASELINE configs[3] at its batch (1024 images, 3x32x32) on a T=8 chain: all eight ancestral steps with
    injected noise (the full 1000-step chain is the bench; the per-step arithmetic is identical)."""
    ch, s, b, T = 3, 32, 1024, 8
    g = torch.Generator().manual_seed(44)
    noises = [torch.randn(b, ch, s, s, generator=g).to(DEV) for _ in range(T + 1)]
    step = {t: noises[1 + (T - 1 - t)] for t in range(T - 1, 0, -1)}
    refs = {}
    for mode in ("fp32", "autocast"):
        orc = cuda_oracle(ch, s, timesteps=T)
        with torch.no_grad(), precision_ctx(mode):
            refs[mode] = orc.p_sample_loop(noises[0], lambda t: step[t]).float()
    yard = dict(psnr=psnr(refs["autocast"], refs["fp32"]), linf=linf(refs["autocast"], refs["fp32"]))
    for precision in ("fp32", "bf16"):
        unet, gd = build(ch, s, precision, timesteps=T)
        img = gd.sample(batch_size=b, init_noise=noises[0], step_noise=lambda t: step[t])
        p32, pac = psnr(img, refs["fp32"]), psnr(img, refs["autocast"])
        report(test="config_ddpm_chain", config="C4", precision=precision, batch=b, psnr_vs_fp32=p32,
               psnr_vs_autocast=pac, linf_vs_fp32=linf(img, refs["fp32"]), ref_autocast_vs_fp32=yard)
        assert p32 >= (80 if precision == "fp32" else 40), (p32, yard)
        if precision == "bf16":
            assert pac >= 40, (pac, yard)
        del unet, gd
        torch.cuda.empty_cache()


def test_bench_step_loss_decreases_on_fixed_batch():
    """The bench's own training loop (DDPM module, fused Adam behind backward, graph replay) on ONE fixed batch with
    fixed t / noise must make the loss fall — guards the un-checked timed region of bench.py."""
    from b200dm import DDPM
    torch.manual_seed(3)
    m = DDPM(img_channels=3, img_size=32, dim=64, lr=1e-4, overlap_optimizer=True)
    m.train()
    opt = m.configure_optimizers()
    gd = m.ema.model
    x, t, noise, _ = inputs(32, 3, 32, seed=5)
    losses = []
    for _ in range(12):
        opt.zero_grad()
        loss = gd.p_losses(x, t, noise=noise, _normalize=True)
        loss.backward()
        opt.step()
        m.on_train_batch_end(None, None, 0)
        losses.append(loss.item())
    report(test="loss_decreases", losses=losses)
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < 0.9 * losses[0], losses
