"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py
Outputs tests/golden/*.npz / *.json.  Weights are NOT stored: they are regenerated at test time
from `synth_state_dict(dim, channels, seed)` (numpy MT19937, platform-stable) and loaded into the
reference with load_state_dict here, so every fixture is the reference's own arithmetic on
reproducible inputs.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from _ref_import import import_reference  # noqa: E402
from oracle.ddpm_oracle import synth_state_dict  # noqa: E402  (weight generator only)


def seeded_inputs(b, c, s, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(b, c, s, s, generator=g)                 # images in [0,1)
    t = torch.randint(0, 1000, (b,), generator=g)
    noise = torch.randn(b, c, s, s, generator=g)
    init = torch.randn(b, c, s, s, generator=g)
    return x, t, noise, init


class PatchedRandn:
    """Feeds a fixed list of tensors to torch.randn / torch.randn_like (the samplers have no
    noise-injection argument, SURVEY hard part 7)."""

    def __init__(self, tensors):
        self.q = list(tensors)

    def __enter__(self):
        self._randn, self._randn_like = torch.randn, torch.randn_like
        torch.randn = lambda *a, **k: self.q.pop(0).clone()
        torch.randn_like = lambda *a, **k: self.q.pop(0).clone()
        return self

    def __exit__(self, *exc):
        torch.randn, torch.randn_like = self._randn, self._randn_like


def main():
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    out = {}

    # ---- schedule KATs (ddpm.py:491-529, :577-662) -------------------------------------------
    kat = {}
    for name, fn in (("linear", ref.linear_beta_schedule), ("cosine", ref.cosine_beta_schedule),
                     ("sigmoid", ref.sigmoid_beta_schedule)):
        b = fn(1000)
        ac = torch.cumprod(1 - b, 0)
        kat[name] = {"betas": [b[i].item() for i in (0, 499, 998, 999)],
                     "alphas_cumprod": [ac[i].item() for i in (0, 499, 999)]}
    kat["sinusoidal_t1"] = ref.SinusoidalPosEmb(64)(torch.tensor([1.0]))[0].tolist()
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- state-dict inventory ------------------------------------------------------------------
    inv = {}
    for ch in (1, 3):
        m = ref.Unet(dim=64, channels=ch)
        inv[str(ch)] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
    gd = ref.GaussianDiffusion(ref.Unet(dim=64, channels=3), img_size=32)
    inv["diffusion_buffers"] = [k for k in gd.state_dict().keys() if not k.startswith("model.")]
    with open(os.path.join(HERE, "state_dict_inventory.json"), "w") as f:
        json.dump(inv, f)

    # ---- forward / loss / grads / samplers -------------------------------------------------------
    cases = [  # name, channels, size, batch, objective, schedule
        ("c3s32", 3, 32, 2, "pred_v", "sigmoid"),
        ("c1s32", 1, 32, 2, "pred_v", "sigmoid"),
        ("c3s64", 3, 64, 1, "pred_noise", "linear"),
        ("c3s32_x0", 3, 32, 2, "pred_x0", "cosine"),
    ]
    for name, ch, s, b, objective, sched in cases:
        sd = synth_state_dict(64, ch, seed=10)
        unet = ref.Unet(dim=64, channels=ch)
        unet.load_state_dict(sd)
        gd = ref.GaussianDiffusion(unet, img_size=s, timesteps=1000, sampling_timesteps=4,
                                   objective=objective, beta_schedule=sched)
        x, t, noise, init = seeded_inputs(b, ch, s, seed=1234)
        rec = {}
        with torch.no_grad():
            rec["unet_out"] = unet(x * 2 - 1, t).numpy()
            rec["x_t"] = gd.q_sample(x * 2 - 1, t, noise=noise.clone()).numpy()
        unet.zero_grad()
        with PatchedRandn([noise]):
            # exercise GaussianDiffusion.forward (normalize + p_losses) with t forced
            _ri = torch.randint
            torch.randint = lambda *a, **k: t.clone()
            try:
                loss = gd(x)
            finally:
                torch.randint = _ri
        loss.backward()
        rec["loss"] = np.float32(loss.item())
        rec["grad_norms"] = np.array([p.grad.norm().item() for p in unet.parameters()], np.float32)
        named = dict(unet.named_parameters())
        for k in ("final_conv.weight", "final_conv.bias", "init_conv.bias",
                  "downs.0.0.block1.norm.weight", "downs.0.0.block1.norm.bias",
                  "mid_attn.mem_kv", "downs.0.2.mem_kv", "downs.0.2.to_out.1.g",
                  "ups.3.1.res_conv.bias", "time_mlp.3.bias", "downs.1.0.mlp.1.bias",
                  "mid_block1.block2.proj.bias"):
            rec["grad:" + k] = named[k].grad.numpy().copy()
        with torch.no_grad():
            # DDIM with 4 steps (eta = 0; randn_like still consumed each step, ddpm.py:825)
            with PatchedRandn([init] + [torch.zeros_like(init)] * 8):
                rec["ddim4"] = gd.sample(batch_size=b).numpy()
            # single ancestral steps at three timesteps, injected noise
            for tt in (999, 500, 0):
                with PatchedRandn([noise]):
                    img, x0 = gd.p_sample(init.clone(), tt)
                rec[f"p_sample_{tt}"] = img.numpy()
                rec[f"p_sample_{tt}_x0"] = x0.numpy()
            mp = gd.model_predictions(init, t, clip_x_start=True, rederive_pred_noise=True)
            rec["mp_noise"], rec["mp_x0"] = mp.pred_noise.numpy(), mp.pred_x_start.numpy()
        np.savez_compressed(os.path.join(HERE, f"golden_{name}.npz"), **rec)
        out[name] = float(loss.item())
        print(name, "loss", loss.item())

    # a short full ancestral chain: T=8 timesteps, 8 steps
    sd = synth_state_dict(64, 3, seed=10)
    unet = ref.Unet(dim=64, channels=3)
    unet.load_state_dict(sd)
    gd = ref.GaussianDiffusion(unet, img_size=32, timesteps=8, sampling_timesteps=None)
    g = torch.Generator().manual_seed(77)
    noises = [torch.randn(2, 3, 32, 32, generator=g) for _ in range(9)]
    with torch.no_grad(), PatchedRandn(noises):
        img = gd.sample(batch_size=2)
    np.savez_compressed(os.path.join(HERE, "golden_ddpm8.npz"), img=img.numpy(),
                        noises=torch.stack(noises).numpy())
    print("done", out)


if __name__ == "__main__":
    main()
