"""Measurement only: what do the second-stream kernels of backward (weight gradients, bias column sums) cost the
training step?  Times the bench's training step (C2 by default) three ways: as shipped; with the second stream's
backward kernels dropped (a WRONG step — gradients missing — that shows the main chain alone); and with everything
on one stream.  Monkeypatches Plan._run_ops from outside; the product code has no such switch.

    python scripts/side_cost.py [--size 32 --batch 128]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))


def time_steps(fn, n=30, reps=5):
    out = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(n):
            fn(i)
        b.record()
        torch.cuda.synchronize()
        out.append(a.elapsed_time(b) / n)
    return sorted(out)[len(out) // 2]


def build(B, S):
    from b200dm import DDPM
    torch.manual_seed(10)
    model = DDPM(img_channels=3, img_size=S, dim=64, lr=2e-5, betas=(0.9, 0.99), precision="bf16",
                 overlap_optimizer=True)
    model.train()
    opt = model.configure_optimizers()
    xs = [torch.rand(B, 3, S, S, device="cuda") for _ in range(4)]
    labels = torch.zeros(B, dtype=torch.long, device="cuda")

    def step(i):
        opt.zero_grad()
        loss = model.training_step((xs[i % 4], labels))
        loss.backward()
        opt.step()
        model.on_train_batch_end(None, None, 0)

    def fwd_only(i):
        model.training_step((xs[i % 4], labels))

    return step, fwd_only


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=32)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--only-shipped", action="store_true")
    a = ap.parse_args()
    from b200dm import engine
    res = {"size": a.size, "batch": a.batch}
    orig = engine.Plan._run_ops

    def measure(tag, with_fwd=False):
        step, fwd_only = build(a.batch, a.size)
        for i in range(6):
            step(i)
        res[tag] = round(time_steps(step), 4)
        if with_fwd:
            res[tag + "_forward_only"] = round(time_steps(fwd_only), 4)
        torch.cuda.empty_cache()

    measure("as_shipped", with_fwd=True)
    if a.only_shipped:
        print(json.dumps(res))
        return

    def no_side_bwd(self, ops):
        if ops is not self.fwd:
            ops = [op for op in ops if not op.side]
        return orig(self, ops)

    engine.Plan._run_ops = no_side_bwd
    measure("main_chain_only")

    for drop in ("colsum", "wgrad_tc", "linear_bwd_cols", "final_conv_bwd"):
        def no_kind(self, ops, drop=drop):
            if ops is not self.fwd:
                ops = [op for op in ops if not (op.side and op.kname == drop)]
            return orig(self, ops)
        engine.Plan._run_ops = no_kind
        measure("without_" + drop)

    def one_stream(self, ops):
        st = engine.L.stream_ptr()
        for op in ops:
            op(st)

    engine.Plan._run_ops = one_stream
    measure("one_stream")
    engine.Plan._run_ops = orig
    print(json.dumps(res))


if __name__ == "__main__":
    main()
