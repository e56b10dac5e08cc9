#!/usr/bin/env python
"""Eager (no CUDA graph) training steps or UNet evaluations of the benchmark workload, for ncu.

  python scripts/profile_step.py [--workload train|eval] [--steps 3] [--batch 128] [--size 32]

Every step issues the same launch sequence, so `ncu -s <launches of the first steps> -c <launches of one
step>` isolates one warm step.  The script prints the launches per step (libb200dm's own counter).
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
os.environ["B200DM_CUDA_GRAPH"] = "0"
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="train", choices=["train", "eval"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--size", type=int, default=32)
    a = ap.parse_args()
    from b200dm import DDPM, _lib as L
    dev = torch.device("cuda", 0)
    torch.manual_seed(10)
    B, S = a.batch, a.size
    model = DDPM(img_channels=3, img_size=S, dim=64, precision="bf16", device=dev)
    lib = L.load()
    x = torch.rand(B, 3, S, S, device=dev)
    labels = torch.zeros(B, dtype=torch.long, device=dev)
    if a.workload == "train":
        model.train()
        opt = model.configure_optimizers()
        for i in range(a.steps):
            lib.b200dm_reset_launch_count()
            opt.zero_grad()
            loss = model.training_step((x, labels))
            loss.backward()
            opt.step()
            model.on_train_batch_end(None, None, 0)
            torch.cuda.synchronize()
            print(f"step {i}: loss {loss.item():.5f}, {int(lib.b200dm_launch_count())} libb200dm launches", flush=True)
    else:
        unet = model.ema.model.model
        t = torch.randint(0, 1000, (B,), device=dev)
        with torch.no_grad():
            for i in range(a.steps):
                lib.b200dm_reset_launch_count()
                out = unet(x, t)
                torch.cuda.synchronize()
                print(f"eval {i}: |out| {out.abs().mean().item():.5f}, {int(lib.b200dm_launch_count())} launches", flush=True)


if __name__ == "__main__":
    main()
