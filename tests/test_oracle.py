"""CPU tests: the oracle (oracle/ddpm_oracle.py) against the reference-generated golden fixtures
(tests/golden/make_golden.py) and, when /root/reference is present, against the live reference."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import ddpm_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = {  # name: channels, size, batch, objective, schedule
    "c3s32": (3, 32, 2, "pred_v", "sigmoid"),
    "c1s32": (1, 32, 2, "pred_v", "sigmoid"),
    "c3s64": (3, 64, 1, "pred_noise", "linear"),
    "c3s32_x0": (3, 32, 2, "pred_x0", "cosine"),
}


def seeded_inputs(b, c, s, seed=1234):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, c, s, s, generator=g)
    t = torch.randint(0, 1000, (b,), generator=g)
    noise = torch.randn(b, c, s, s, generator=g)
    init = torch.randn(b, c, s, s, generator=g)
    return x, t, noise, init


def test_schedule_kats():
    kat = json.load(open(os.path.join(GOLD, "kat.json")))
    for name, fn in (("linear", O.linear_beta_schedule), ("cosine", O.cosine_beta_schedule),
                     ("sigmoid", O.sigmoid_beta_schedule)):
        b = fn(1000)
        ac = torch.cumprod(1 - b, 0)
        assert [b[i].item() for i in (0, 499, 998, 999)] == kat[name]["betas"]
        assert [ac[i].item() for i in (0, 499, 999)] == kat[name]["alphas_cumprod"]
    # SURVEY §4 KATs
    b = O.sigmoid_beta_schedule(1000)
    assert abs(b[0].item() - 3.0027919741e-04) < 1e-13 and b[999].item() == 0.999
    buf = O.make_buffers(1000, "sigmoid", "pred_v")
    assert abs(buf["posterior_log_variance_clipped"][0].item() - (-46.0517006)) < 1e-5
    assert abs(buf["loss_weight"][499].item() - 0.5) < 1e-6
    e = O.sinusoidal_pos_emb(torch.tensor([1.0]), 64)[0]
    np.testing.assert_allclose(e.numpy(), np.array(kat["sinusoidal_t1"], np.float32), rtol=0, atol=1e-7)


def test_param_inventory():
    inv = json.load(open(os.path.join(GOLD, "state_dict_inventory.json")))
    for ch in (1, 3):
        spec = O.unet_param_spec(64, ch)
        assert [[k, list(s)] for k, s in spec] == inv[str(ch)]
        assert len(spec) == 283
    assert sum(int(np.prod(s)) for _, s in O.unet_param_spec(64, 3)) == 35719555
    assert sorted(O.make_buffers().keys()) == sorted(inv["diffusion_buffers"])


def test_ddim_time_grid():
    o = O.DiffusionOracle({}, img_size=32, sampling_timesteps=50)
    times = [a for a, _ in o.ddim_time_pairs()] + [o.ddim_time_pairs()[-1][1]]
    assert times[:3] == [999, 979, 959] and times[-3:] == [39, 19, -1] and len(times) == 51
    o = O.DiffusionOracle({}, img_size=32, sampling_timesteps=10)
    assert [a for a, _ in o.ddim_time_pairs()] == list(range(999, 98, -100))


@pytest.mark.parametrize("name", list(CASES))
def test_golden_forward_loss_grads(name):
    ch, s, b, objective, sched = CASES[name]
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    sd = O.synth_state_dict(64, ch, seed=10)
    for v in sd.values():
        v.requires_grad_(True)
    x, t, noise, init = seeded_inputs(b, ch, s)
    orc = O.DiffusionOracle(sd, img_size=s, channels=ch, sampling_timesteps=4,
                            objective=objective, beta_schedule=sched)
    with torch.no_grad():
        out = orc.model(x * 2 - 1, t)
        xt = orc.q_sample(x * 2 - 1, t, noise)
    np.testing.assert_allclose(out.numpy(), gold["unet_out"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(xt.numpy(), gold["x_t"], rtol=0, atol=1e-6)
    loss = orc.forward(x, t, noise)
    assert abs(loss.item() - float(gold["loss"])) <= 2e-6 * max(1.0, abs(float(gold["loss"])))
    loss.backward()
    gn = np.array([sd[k].grad.norm().item() for k, _ in O.unet_param_spec(64, ch)], np.float32)
    np.testing.assert_allclose(gn, gold["grad_norms"], rtol=2e-3, atol=1e-7)
    for key in gold.files:
        if key.startswith("grad:"):
            g = sd[key[5:]].grad.numpy()
            ref = gold[key]
            assert np.abs(g - ref).max() <= 1e-3 * np.abs(ref).max() + 1e-8, key


@pytest.mark.parametrize("name", list(CASES))
def test_golden_samplers(name):
    ch, s, b, objective, sched = CASES[name]
    gold = np.load(os.path.join(GOLD, f"golden_{name}.npz"))
    sd = O.synth_state_dict(64, ch, seed=10)
    x, t, noise, init = seeded_inputs(b, ch, s)
    orc = O.DiffusionOracle(sd, img_size=s, channels=ch, sampling_timesteps=4,
                            objective=objective, beta_schedule=sched)
    with torch.no_grad():
        img = orc.sample(init)
        np.testing.assert_allclose(img.numpy(), gold["ddim4"], rtol=0, atol=5e-5)
        for tt in (999, 500, 0):
            im, x0 = orc.p_sample(init, tt, noise)
            # x0 = sqrt(1/abar) x - ... amplifies fp32 round-off by up to ~160x at t = 999
            np.testing.assert_allclose(im.numpy(), gold[f"p_sample_{tt}"], rtol=0, atol=3e-4)
            np.testing.assert_allclose(x0.numpy(), gold[f"p_sample_{tt}_x0"], rtol=0, atol=3e-4)
        mp = orc.model_predictions(init, t, clip_x_start=True, rederive_pred_noise=True)
        # pred_noise is divided by sqrt(1/abar - 1) which is tiny at small t: compare relatively
        scale = np.abs(gold["mp_noise"]).max()
        assert np.abs(mp.pred_noise.numpy() - gold["mp_noise"]).max() <= 1e-4 * scale
        np.testing.assert_allclose(mp.pred_x_start.numpy(), gold["mp_x0"], rtol=0, atol=3e-4)


def test_golden_ddpm_chain():
    gold = np.load(os.path.join(GOLD, "golden_ddpm8.npz"))
    sd = O.synth_state_dict(64, 3, seed=10)
    orc = O.DiffusionOracle(sd, img_size=32, timesteps=8)
    noises = [torch.from_numpy(n) for n in gold["noises"]]
    # reference order: randn(init), then one randn_like per step t = 7..1 (none at t == 0)
    step = {t: noises[1 + (7 - t)] for t in range(7, 0, -1)}
    with torch.no_grad():
        img = orc.p_sample_loop(noises[0], lambda t: step[t])
    np.testing.assert_allclose(img.numpy(), gold["img"], rtol=0, atol=5e-5)


def test_golden_self_conditioning():
    """Self-conditioned UNet / losses / samplers against the reference (tests/golden/make_golden_selfcond.py)."""
    gold = np.load(os.path.join(GOLD, "golden_selfcond.npz"))
    for name, ch, s, b in (("c1s32", 1, 32, 2), ("c3s16", 3, 16, 2)):
        sd = O.synth_state_dict(64, ch, seed=10, self_condition=True)
        assert sd["init_conv.weight"].shape == (64, 2 * ch, 7, 7)
        orc = O.DiffusionOracle(sd, img_size=s, channels=ch, sampling_timesteps=4)
        assert orc.self_condition
        x, t, noise, init = seeded_inputs(b, ch, s, seed=4321)
        cond = torch.from_numpy(gold[f"{name}:cond"])
        with torch.no_grad():
            np.testing.assert_allclose(orc.model(x * 2 - 1, t, cond).numpy(), gold[f"{name}:unet_out_cond"],
                                       rtol=0, atol=2e-5)
            np.testing.assert_allclose(orc.model(x * 2 - 1, t).numpy(), gold[f"{name}:unet_out_nocond"],
                                       rtol=0, atol=2e-5)
        for flag, sc in (("sc", True), ("nosc", False)):
            sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            og = O.DiffusionOracle(sdg, img_size=s, channels=ch)
            loss = og.p_losses(x * 2 - 1, t, noise, self_cond=sc)
            loss.backward()
            assert abs(loss.item() - float(gold[f"{name}:loss_{flag}"])) <= 2e-6 * max(1.0, abs(loss.item()))
            gn = np.array([sdg[k].grad.norm().item() for k, _ in O.unet_param_spec(64, ch, self_condition=True)])
            np.testing.assert_allclose(gn, gold[f"{name}:grad_norms_{flag}"], rtol=2e-3, atol=1e-7)
        with torch.no_grad():
            np.testing.assert_allclose(orc.sample(init).numpy(), gold[f"{name}:ddim4"], rtol=0, atol=5e-5)
            img, x0 = orc.p_sample(init, 500, noise, x_self_cond=cond)
            np.testing.assert_allclose(img.numpy(), gold[f"{name}:p_sample_500"], rtol=0, atol=5e-5)
            np.testing.assert_allclose(x0.numpy(), gold[f"{name}:p_sample_500_x0"], rtol=0, atol=5e-5)
    sd = O.synth_state_dict(64, 1, seed=10, self_condition=True)
    orc = O.DiffusionOracle(sd, img_size=32, channels=1, timesteps=6)
    noises = [torch.from_numpy(a) for a in gold["ddpm6:noises"]]
    step = {t: noises[1 + (5 - t)] for t in range(5, 0, -1)}
    with torch.no_grad():
        img = orc.p_sample_loop(noises[0], lambda t: step[t])
    np.testing.assert_allclose(img.numpy(), gold["ddpm6:img"], rtol=0, atol=5e-5)


def test_input_size_assert():
    sd = O.synth_state_dict(64, 1, seed=10)
    with pytest.raises(AssertionError):
        O.unet_forward(sd, torch.zeros(1, 1, 28, 28), torch.zeros(1, dtype=torch.long))


def test_oracle_vs_live_reference():
    from _ref_import import import_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present (GPU box)")
    ref = import_reference()
    sd = O.synth_state_dict(64, 3, seed=3)
    m = ref.Unet(dim=64, channels=3)
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 3, 32, 32, generator=g)
    t = torch.tensor([321])
    with torch.no_grad():
        assert (m(x, t) - O.unet_forward(sd, x, t)).abs().max().item() < 2e-5
