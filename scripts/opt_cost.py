#!/usr/bin/env python
"""What the optimiser tail costs inside the training loop: ms/step with the fused Adam after backward, with it applied
per gradient bucket behind backward, and (analysis only, not a valid training step) without any optimiser step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
import torch
from b200dm import DDPM

dev = torch.device("cuda", 0)
B, S = 128, 32
x = [torch.rand(B, 3, S, S, device=dev) for _ in range(4)]
labels = torch.zeros(B, dtype=torch.long, device=dev)


def run(overlap, do_step, steps=40):
    torch.manual_seed(0)
    m = DDPM(img_channels=3, img_size=S, dim=64, precision="bf16", device=dev, overlap_optimizer=overlap)
    m.train()
    opt = m.configure_optimizers()

    def step(i):
        opt.zero_grad()
        loss = m.training_step((x[i % 4], labels))
        loss.backward()
        if do_step:
            opt.step()
        else:
            opt._done.clear()
        m.on_train_batch_end(None, None, 0)
    for i in range(6):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del m, opt
    torch.cuda.empty_cache()
    return ms


for name, ov, st in (("adam after backward", False, True), ("adam per bucket behind backward", True, True),
                     ("no optimiser step (analysis only)", False, False)):
    print(f"{name:40s} {run(ov, st):.3f} ms/step", flush=True)
