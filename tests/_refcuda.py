"""Yardsticks for the GPU parity tests: the oracle (oracle/ddpm_oracle.py — the reference's algorithm as plain
torch ops, pinned to the reference by tests/golden) run ON THE GPU, in fp32 (TF32 off) and under
`torch.autocast("cuda", torch.bfloat16)`, which is what the reference's `--precision bf16-mixed` does through
Lightning (reference train.py:40,132; `q_sample` stays fp32, ddpm.py:869).  BASELINE.md §3 defines the bf16 gate
with this arithmetic as the yardstick.  Test infrastructure only.
"""
import contextlib
import json
import math
import os

import torch

from oracle import ddpm_oracle as O

DEV = "cuda"
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out",
                      "parity_report.jsonl")
_SD = {}


def report(**kw):
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        with open(REPORT, "a") as f:
            f.write(json.dumps(kw) + "\n")
    except OSError:
        pass


def synth(ch):
    if ch not in _SD:
        _SD[ch] = O.synth_state_dict(64, ch, seed=10)
    return _SD[ch]


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def psnr(a, b):
    mse = (a.detach().double().cpu() - b.detach().double().cpu()).pow(2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


def linf(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


@contextlib.contextmanager
def precision_ctx(mode):
    """mode 'fp32': true fp32 (TF32 off); 'autocast': torch.autocast(cuda, bf16)."""
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        if mode == "autocast":
            with torch.autocast("cuda", dtype=torch.bfloat16):
                yield
        else:
            with torch.autocast("cuda", enabled=False):
                yield
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b


def cuda_oracle(ch, size, *, grads=False, **kw):
    """DiffusionOracle on the GPU over the synthetic reference weights (optionally requiring grad)."""
    sd = {k: v.to(DEV).clone().requires_grad_(grads) for k, v in synth(ch).items()}
    return O.DiffusionOracle(sd, img_size=size, channels=ch, **kw).to(DEV)


def grad_distance(grads_a, grads_b, spec):
    """(total rel-L2 over the whole gradient arena, worst per-tensor rel-L2, its name)."""
    num = den = 0.0
    worst, worst_name = 0.0, ""
    for k, _ in spec:
        ga, gb = grads_a[k].detach().double().cpu(), grads_b[k].detach().double().cpu()
        assert ga.shape == gb.shape, k
        e, n = (ga - gb).norm().item(), gb.norm().item()
        num += e * e
        den += n * n
        if n > 1e-6 and e / n > worst:
            worst, worst_name = e / n, k
    return math.sqrt(num / max(den, 1e-300)), worst, worst_name
