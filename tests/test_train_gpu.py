"""N3 / N2 on the GPU: the train.py-compatible launcher end to end, checkpoints in the Lightning layout, EMA resume,
gradient accumulation, the optional GaussianDiffusion flags (N4)."""
import json
import math
import os
import subprocess
import sys

import pytest
import torch

from _refcuda import DEV, cuda_oracle, precision_ctx, rel, report, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "lightning-generative-models_b200")


def test_train_py_runs_and_loss_decreases(tmp_path):
    cfg = json.load(open(os.path.join(PKG, "configs", "diffusion", "ddpm.json")))
    cfg["model"]["args"]["lr"] = 1e-3
    cfg["model"]["args"]["sampling_timesteps"] = 5         # keep the step-0 sample(64) short
    p = tmp_path / "ddpm.json"
    p.write_text(json.dumps(cfg))
    exp = tmp_path / "exp"
    r = subprocess.run([sys.executable, os.path.join(PKG, "train.py"), "--config_path", str(p), "--max_steps", "200",
                        "--precision", "bf16-mixed", "--log_every", "1", "--experiment_dir", str(exp),
                        "--num_images", "256"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    rows = [json.loads(l) for l in open(exp / "train_log.jsonl")]
    losses = [x["train_loss"] for x in rows if x["event"] == "train"]
    assert len(losses) == 200 and all(math.isfinite(v) for v in losses)
    # t and the noise are redrawn every step, so single losses scatter by +-30 %: compare 50-step means
    first, last = sum(losses[:50]) / 50, sum(losses[-50:]) / 50
    report(test="train_py", first50=first, last50=last)
    assert last < 0.97 * first, (first, last)
    assert any(x["event"] == "sample" and x["shape"] == [64, 3, 32, 32] for x in rows)
    ck = torch.load(exp / "last.ckpt", map_location="cpu", weights_only=False)
    assert ck["global_step"] == 200 and "optimizer_states" in ck
    keys = set(ck["state_dict"])
    assert {"ema.initted", "ema.step", "ema.online_model.model.init_conv.weight",
            "ema.ema_model.model.final_conv.bias", "ema.online_model.betas"} <= keys


def _mk(lr=2e-4, **kw):
    from b200dm import DDPM
    m = DDPM(img_channels=3, img_size=32, dim=64, lr=lr, ema_update_every=2, sampling_timesteps=4, **kw)
    m.train()
    return m, m.configure_optimizers()


def _steps(m, opt, batch, n):
    out = []
    for _ in range(n):
        opt.zero_grad()
        loss = m.training_step(batch)
        loss.backward()
        opt.step()
        m.on_train_batch_end(None, batch, 0)
        out.append(loss.item())
    return out


def test_checkpoint_roundtrip_ema_resume_and_torch_adam_format(tmp_path):
    g = torch.Generator().manual_seed(3)
    batch = (torch.rand(8, 3, 32, 32, generator=g).to(DEV), torch.zeros(8, dtype=torch.long, device=DEV))
    m, opt = _mk()
    m.ema.update_after_step = 2                      # reach the lerp regime within a few steps
    _steps(m, opt, batch, 7)
    path = tmp_path / "last.ckpt"
    m.save_checkpoint(str(path), opt, epoch=1)
    ck = torch.load(path, map_location="cpu", weights_only=False)
    # torch.optim.Adam can load the optimiser state as is (same parameter order, OIHW shapes)
    ref_params = [torch.nn.Parameter(p.detach().clone().cpu()) for p in m.ema.model.model.parameters()]
    ref_opt = torch.optim.Adam(ref_params, lr=1.0)
    ref_opt.load_state_dict(ck["optimizer_states"][0])
    assert ref_opt.param_groups[0]["lr"] == 2e-4 and int(ref_opt.state[ref_params[0]]["step"]) == 7
    assert ref_opt.state[ref_params[0]]["exp_avg"].shape == ref_params[0].shape
    # resume into a fresh module: weights, EMA weights, EMA counters and Adam moments all come back
    m2, opt2 = _mk(lr=1e-3)
    m2.ema.update_after_step = 2
    m2.load_checkpoint(str(path), opt2)
    assert opt2.param_groups[0]["lr"] == 2e-4 and opt2.step_count == 7
    assert m2.ema._step_host == 7 and m2.ema._initted_host and m2.global_step == 7
    a, b = m.ema.model.model.arena, m2.ema.model.model.arena
    assert torch.equal(a.flat, b.flat)
    assert torch.equal(m.ema.ema_model.model.arena.flat, m2.ema.ema_model.model.arena.flat)
    assert torch.equal(opt.exp_avg, opt2.exp_avg) and torch.equal(opt.exp_avg_sq, opt2.exp_avg_sq)
    # the next EMA update after the resume is a lerp, not a copy of the online weights
    ema_before = m2.ema.ema_model.model.arena.flat.clone()
    m2.ema.update()                                  # step 7: not an update step (7 % 2 != 0)
    m2.ema.update()                                  # step 8: lerp
    online = m2.ema.model.model.arena.flat
    ema_after = m2.ema.ema_model.model.arena.flat
    assert not torch.equal(ema_after, online)
    d = m2.ema.get_current_decay()
    assert 0.0 < d < 1.0
    # a torch-Adam-format state produced by torch itself loads too (reference checkpoint direction)
    opt3_sd = ref_opt.state_dict()
    m3, opt3 = _mk()
    opt3.load_state_dict(opt3_sd)
    assert torch.equal(opt3.exp_avg.cpu(), opt.exp_avg.cpu()) and opt3.step_count == 7


def test_reference_checkpoint_keys_load(tmp_path):
    """A checkpoint with the reference's key set (inventory generated from the reference, tests/golden) loads."""
    inv = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_inventory.json")))
    from b200dm.schedule import make_buffers
    sd_u = synth(3)
    bufs = make_buffers(1000, "sigmoid", "pred_v", False, 5, {})
    assert sorted(bufs) == sorted(inv["diffusion_buffers"])
    sd = {"ema.initted": torch.tensor(True), "ema.step": torch.tensor(1234)}
    for side in ("online_model", "ema_model"):
        for k, v in sd_u.items():
            sd[f"ema.{side}.model.{k}"] = v if side == "online_model" else v * 0.5
        for k, v in bufs.items():
            sd[f"ema.{side}.{k}"] = v
    from b200dm import DDPM
    m = DDPM(img_channels=3, img_size=32, dim=64)
    res = m.load_checkpoint({"state_dict": sd, "global_step": 1234})
    assert not res.missing_keys and not res.unexpected_keys
    assert m.ema._step_host == 1234 and m.ema._initted_host
    out = m.ema.model.model.state_dict()
    assert all(torch.equal(out[k].cpu(), sd_u[k]) for k in sd_u)
    assert torch.allclose(m.ema.ema_model.model.state_dict()["init_conv.weight"].cpu(), sd_u["init_conv.weight"] * 0.5)
    assert set(m.state_dict()) == set(sd)


def test_gradient_accumulation_matches_one_big_batch():
    """accumulate_grad_batches=2: two micro-batches (first under no_sync) == one batch of twice the size."""
    from b200dm import GaussianDiffusion, Unet
    g = torch.Generator().manual_seed(9)
    x = torch.rand(8, 3, 32, 32, generator=g).to(DEV)
    t = torch.randint(0, 1000, (8,), generator=g).to(DEV)
    noise = torch.randn(8, 3, 32, 32, generator=g).to(DEV)
    unet = Unet(dim=64, channels=3, precision="fp32")
    unet.load_reference_state_dict(synth(3))
    gd = GaussianDiffusion(unet, img_size=32)
    unet.zero_grad()
    gd.p_losses(x, t, noise=noise, _normalize=True).backward()
    full = unet.arena.gflat.clone()
    unet.zero_grad()
    with unet.no_sync():
        (gd.p_losses(x[:4], t[:4], noise=noise[:4], _normalize=True) / 2).backward()
    (gd.p_losses(x[4:], t[4:], noise=noise[4:], _normalize=True) / 2).backward()
    assert rel(unet.arena.gflat, full) <= 1e-5


def test_min_snr_offset_noise_and_interpolate():
    """N4: min_snr_loss_weight, offset_noise_strength (ddpm.py:889-891), interpolate (:847-867) vs the oracle."""
    from b200dm import GaussianDiffusion, Unet
    from oracle import ddpm_oracle as O
    g = torch.Generator().manual_seed(21)
    x = torch.rand(4, 3, 32, 32, generator=g).to(DEV)
    t = torch.tensor([5, 300, 650, 990], device=DEV)
    noise = torch.randn(4, 3, 32, 32, generator=g).to(DEV)
    unet = Unet(dim=64, channels=3, precision="fp32")
    unet.load_reference_state_dict(synth(3))
    # min-SNR loss weighting
    gd = GaussianDiffusion(unet, img_size=32, min_snr_loss_weight=True, min_snr_gamma=5)
    loss = gd.p_losses(x, t, noise=noise, _normalize=True)
    orc = cuda_oracle(3, 32)
    orc.buf = {k: v.to(DEV) for k, v in O.make_buffers(1000, "sigmoid", "pred_v", True, 5).items()}
    with precision_ctx("fp32"), torch.no_grad():
        ref = orc.forward(x, t, noise)
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    # offset noise with torch's generator: the reference's RNG order is noise (injected), then randn(b, c)
    gd = GaussianDiffusion(unet, img_size=32, offset_noise_strength=0.1, rng="torch")
    torch.manual_seed(77)
    loss = gd.p_losses(x * 2 - 1, t, noise=noise.clone())
    torch.manual_seed(77)
    off = torch.randn(4, 3, device=DEV)
    orc = cuda_oracle(3, 32)
    with precision_ctx("fp32"), torch.no_grad():
        ref = orc.p_losses(x * 2 - 1, t, noise + 0.1 * off.view(4, 3, 1, 1))
    assert abs(loss.item() - ref.item()) <= 1e-4 * abs(ref.item()), (loss.item(), ref.item())
    # interpolate: q_sample both images at t, mix, run the ancestral chain down from t (T = 6 here)
    gd = GaussianDiffusion(unet, img_size=32, timesteps=6, rng="torch")
    torch.manual_seed(5)
    out = gd.interpolate(x[:2] * 2 - 1, x[2:] * 2 - 1, t=4, lam=0.3)
    torch.manual_seed(5)
    orc = cuda_oracle(3, 32, timesteps=6)
    with precision_ctx("fp32"), torch.no_grad():
        tb = torch.full((2,), 4, device=DEV)
        xt1 = orc.q_sample(x[:2] * 2 - 1, tb, torch.randn_like(x[:2]))
        xt2 = orc.q_sample(x[2:] * 2 - 1, tb, torch.randn_like(x[2:]))
        img = 0.7 * xt1 + 0.3 * xt2
        for i in reversed(range(0, 4)):
            z = torch.randn_like(img) if i > 0 else None
            img, _ = orc.p_sample(img, i, z)
    assert rel(out, img) <= 1e-4, rel(out, img)
