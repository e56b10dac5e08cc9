"""Golden fixture for self-conditioning (reference Unet(self_condition=True), ddpm.py:300-304, :433-435, :773, :807,
:901-905) from the UNMODIFIED reference.  Build container only:   python tests/golden/make_golden_selfcond.py
Writes tests/golden/golden_selfcond.npz.  Weights come from synth_state_dict(..., self_condition=True)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from _ref_import import import_reference  # noqa: E402
from make_golden import PatchedRandn, seeded_inputs  # noqa: E402
from oracle.ddpm_oracle import synth_state_dict  # noqa: E402  (weight generator only)


def main():
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    rec = {}
    for name, ch, s, b in (("c1s32", 1, 32, 2), ("c3s16", 3, 16, 2)):
        sd = synth_state_dict(64, ch, seed=10, self_condition=True)
        unet = ref.Unet(dim=64, channels=ch, self_condition=True)
        unet.load_state_dict(sd)
        gd = ref.GaussianDiffusion(unet, img_size=s, timesteps=1000, sampling_timesteps=4)
        assert gd.self_condition
        x, t, noise, init = seeded_inputs(b, ch, s, seed=4321)
        g = torch.Generator().manual_seed(99)
        cond = torch.rand(b, ch, s, s, generator=g) * 2 - 1
        with torch.no_grad():
            rec[f"{name}:unet_out_cond"] = unet(x * 2 - 1, t, cond).numpy()
            rec[f"{name}:unet_out_nocond"] = unet(x * 2 - 1, t).numpy()
        rec[f"{name}:cond"] = cond.numpy()
        for flag, coin in (("sc", 0.0), ("nosc", 0.9)):       # random() < 0.5 decides (ddpm.py:902)
            unet.zero_grad()
            _coin, _ri = ref.random, torch.randint
            ref.random = lambda: coin
            torch.randint = lambda *a, **k: t.clone()
            try:
                with PatchedRandn([noise]):
                    loss = gd(x)
            finally:
                ref.random, torch.randint = _coin, _ri
            loss.backward()
            rec[f"{name}:loss_{flag}"] = np.float32(loss.item())
            rec[f"{name}:grad_norms_{flag}"] = np.array([p.grad.norm().item() for p in unet.parameters()], np.float32)
        with torch.no_grad():
            with PatchedRandn([init] + [torch.zeros_like(init)] * 8):
                rec[f"{name}:ddim4"] = gd.sample(batch_size=b).numpy()
            with PatchedRandn([noise]):
                img, x0 = gd.p_sample(init.clone(), 500, cond)
            rec[f"{name}:p_sample_500"], rec[f"{name}:p_sample_500_x0"] = img.numpy(), x0.numpy()
        print(name, float(rec[f"{name}:loss_sc"]), float(rec[f"{name}:loss_nosc"]))
    # a short ancestral chain with self-conditioning: T = 6
    sd = synth_state_dict(64, 1, seed=10, self_condition=True)
    unet = ref.Unet(dim=64, channels=1, self_condition=True)
    unet.load_state_dict(sd)
    gd = ref.GaussianDiffusion(unet, img_size=32, timesteps=6, sampling_timesteps=None)
    g = torch.Generator().manual_seed(78)
    noises = [torch.randn(2, 1, 32, 32, generator=g) for _ in range(7)]
    with torch.no_grad(), PatchedRandn(noises):
        rec["ddpm6:img"] = gd.sample(batch_size=2).numpy()
    rec["ddpm6:noises"] = torch.stack(noises).numpy()
    np.savez_compressed(os.path.join(HERE, "golden_selfcond.npz"), **rec)
    print("done")


if __name__ == "__main__":
    main()
