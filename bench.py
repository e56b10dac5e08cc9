#!/usr/bin/env python
"""Benchmark of the B200 diffusion hot path (contract: see the task statement / DESIGN.md §6).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|ddim|ddpm|train64] [--impl reference]

Workload at N=1 (BASELINE.json configs[1], the configuration the metric is quoted on):
  DDPM training step, UNet(dim=64) on 3x32x32, batch 128 per GPU, bf16 activations / fp32 accumulate:
  normalize+q_sample (Philox) -> UNet forward -> loss -> full backward -> fused Adam -> EMA bookkeeping.
N > 1 (torchrun): the same per-GPU work on every rank (weak scaling) with an NCCL all-reduce of the
flat gradient arena every step.  `--workload ddim` times DDIM-50 sampling at 3x64x64 (configs[2]),
batch-sharded with no communication; `--workload ddpm` times the 1000-step ancestral sampler at 3x32x32,
global batch 1024 (configs[3]; one "step" is a whole 1000-evaluation chain, so K is capped at 3);
`--workload train64` is the training step at 3x64x64 with 64 images per GPU (configs[4]: global batch 512 on 8 GPUs).

One JSON line is printed by rank 0.  `value` = device-resident throughput, `e2e` = the same step driven
through the public API from pinned host memory (H2D of the batch and D2H of the loss inside the timed
region).  `--impl reference` times the CPU oracle (the reference's algorithm on the host cores).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "lightning-generative-models_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_FWD_32 = 3.651e9       # algorithmic FLOPs of one UNet forward per image @3x32x32
F_FWD_64 = 14.594e9      # @3x64x64
TRAIN_B, TRAIN_S = 128, 32
DDIM_B, DDIM_S, DDIM_STEPS = 256, 64, 50
F_FWD = {32: F_FWD_32, 64: F_FWD_64}
TRAIN_CFG = {"train": (128, 32), "train64": (64, 64)}                  # per-GPU batch, image size
SAMPLE_CFG = {"ddim": (256, 64, 50), "ddpm": (1024, 32, 1000)}         # global batch, image size, UNet evaluations


def profiled_traffic(family):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of a kernel family from the committed
    `ncu --set full` captures (profiles/r2_roofline_traffic.json, scripts/make_profiles_r2.py; the conv / weight-gradient
    families at the training shape, the fused conv+GroupNorm launch at the DDIM shape); None if absent."""
    p = os.path.join(ROOT, "profiles", "r2_roofline_traffic.json")
    if not os.path.isfile(p):
        return None
    d = json.load(open(p))
    names = {"wgrad_tc": ("wgrad3x3_halo_kernel", "wgrad_tc_kernel"),
             "conv_tc_fwd": ("conv3x3_halo_kernel", "conv_tc_kernel"),
             "conv_tc_dgrad": ("conv3x3_halo_kernel", "conv_tc_kernel"),
             "conv_gn_fwd": ("conv3x3_gn_kernel",)}.get(family, (family,))
    n = sum(d[k]["launches"] for k in names if k in d)
    if n == 0:
        return None
    return round(sum(d[k]["dram_MB"] for k in names if k in d) * 1e6 / n)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Start of the timed region: only samples taken after this (and under load) are reported."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, sm_all, mx, reasons = [], [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        t_mark = getattr(self, "t_mark", 0.0)
        for ts, r in self.rows:
            try:
                mx = float(r[1])
                sm_all.append(float(r[0]))
                if ts < t_mark:
                    continue
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
                sm.append(float(r[0]))           # taken between mark() and stop(): inside the timed regions
            except (ValueError, IndexError):
                continue
        use = sm if sm else sm_all
        return {"sm_mhz": statistics.median(use) if use else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples_under_load": len(sm), "samples": len(sm_all)}


# --------------------------------------------------------------------------------------------------------
# Baseline arms: the reference itself (baseline/_ref, stub-imported, UNMODIFIED) when it travelled to this box,
# else the oracle port.  CPU (all host threads) and eager PyTorch on the same GPU (cuDNN / cuBLAS / ATen).
# --------------------------------------------------------------------------------------------------------
os.environ.setdefault("TQDM_DISABLE", "1")       # the reference samplers wrap their loops in tqdm


def _reference_module():
    try:
        from baseline import ref_import as R
        if R.reference_available():
            return R.import_reference()
    except Exception as e:                         # noqa: BLE001  (the port is the documented fallback)
        print(f"bench: reference module unavailable ({e}); timing the oracle port", file=sys.stderr)
    return None


class RefArm:
    """One training step / one UNet evaluation / one DDIM-50 chain of the reference algorithm on `device`."""

    def __init__(self, S, device, autocast=False):
        self.S, self.dev, self.autocast = S, torch.device(device), autocast
        self.ref = _reference_module()
        self.kind = "reference" if self.ref is not None else "port"
        torch.manual_seed(10)                      # the reference's seed (train.py:20)
        if self.ref is not None:
            self.unet = self.ref.Unet(dim=64, channels=3).to(self.dev)
            self.sd = None
        else:
            from oracle import ddpm_oracle as O
            self.O = O
            self.sd = {k: v.to(self.dev) for k, v in O.synth_state_dict(64, 3, seed=10).items()}
        self.g = torch.Generator(device=self.dev).manual_seed(10)

    def ctx(self):
        if self.dev.type == "cuda":
            return torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.autocast)
        import contextlib
        return contextlib.nullcontext()

    def train_step_fn(self, batch):
        S, dev = self.S, self.dev
        if self.ref is not None:
            gd = self.ref.GaussianDiffusion(self.unet, img_size=S).to(dev)
            params = list(self.unet.parameters())
        else:
            for v in self.sd.values():
                v.requires_grad_(True)
            gd = self.O.DiffusionOracle(self.sd, img_size=S, channels=3).to(dev)
            params = list(self.sd.values())
        opt = torch.optim.Adam(params, lr=2e-5, betas=(0.9, 0.99))
        x = torch.rand(batch, 3, S, S, generator=self.g, device=dev)

        def step():
            opt.zero_grad()
            with self.ctx():
                if self.ref is not None:
                    loss = gd(x)                   # GaussianDiffusion.forward: randint t, randn noise, p_losses
                else:
                    t = torch.randint(0, 1000, (batch,), generator=self.g, device=dev)
                    noise = torch.randn(batch, 3, S, S, generator=self.g, device=dev)
                    loss = gd.forward(x, t, noise)
            loss.backward()
            opt.step()
            return loss
        return step

    def eval_fn(self, batch):
        S, dev = self.S, self.dev
        x = torch.randn(batch, 3, S, S, generator=self.g, device=dev)
        t = torch.full((batch,), 500, dtype=torch.long, device=dev)

        def step():
            with torch.no_grad(), self.ctx():
                out = self.unet(x, t) if self.ref is not None else self.O.unet_forward(self.sd, x, t)
            return out
        return step

    def chain_fn(self, batch, steps):
        """Whole sampling chain through the reference's own `sample()` (ddpm.py:836-845)."""
        S, dev = self.S, self.dev
        st = None if steps >= 1000 else steps
        if self.ref is not None:
            gd = self.ref.GaussianDiffusion(self.unet, img_size=S, sampling_timesteps=st).to(dev)

            def step():
                with self.ctx():
                    return gd.sample(batch_size=batch)
        else:
            orc = self.O.DiffusionOracle(self.sd, img_size=S, channels=3, sampling_timesteps=st).to(dev)

            def step():
                init = torch.randn(batch, 3, S, S, generator=self.g, device=dev)
                with torch.no_grad(), self.ctx():
                    return orc.sample(init, (lambda t: torch.randn(batch, 3, S, S, generator=self.g, device=dev))
                                      if st is None else None)
        return step


def cpu_baseline(workload, budget_s=20.0):
    """Bounded sample of the same workload on the host cores (the unmodified reference when available)."""
    torch.set_num_threads(os.cpu_count())
    if workload in TRAIN_CFG:
        S = TRAIN_CFG[workload][1]
        b = 16 if S == 32 else 4
        arm = RefArm(S, "cpu")
        step = arm.train_step_fn(b)
        step()
        t0, n = time.perf_counter(), 0
        while True:
            step()
            n += 1
            if time.perf_counter() - t0 > budget_s or n >= 8:
                break
        dt = time.perf_counter() - t0
        return {"value": b * n / dt, "unit": "img/s", "cores": os.cpu_count(), "kind": arm.kind,
                "sample": f"{n} fp32 training steps (fwd+bwd+Adam) at batch {b}, 3x{S}x{S}, torch CPU {os.cpu_count()} threads"}
    _, S, evals = SAMPLE_CFG[workload]
    b = 2 if S == 64 else 8
    arm = RefArm(S, "cpu")
    step = arm.eval_fn(b)
    step()
    t0, n = time.perf_counter(), 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 6:
            break
    dt = time.perf_counter() - t0
    return {"value": b * n / dt / evals, "unit": "img/s", "cores": os.cpu_count(), "kind": arm.kind,
            "sample": f"{n} fp32 UNet evaluations at batch {b}, 3x{S}x{S}, extrapolated x{evals} steps per image"}


def _cuda_time(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def gpu_eager_baseline(workload):
    """The reference algorithm as eager PyTorch on THIS GPU (cuDNN / cuBLAS / ATen kernels), fp32 (TF32 off) and
    under torch.autocast(bf16) like the reference's `--precision bf16-mixed`: the kernel-for-kernel bar of SURVEY 8d.
    Full batch of the config; training = 10 timed steps, sampling = one whole chain through the reference's
    `sample()`."""
    out = {"unit": "img/s", "device": torch.cuda.get_device_name(0)}
    tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        for name, amp in (("fp32", False), ("autocast_bf16", True)):
            if workload in TRAIN_CFG:
                B, S = TRAIN_CFG[workload]
                arm = RefArm(S, "cuda", autocast=amp)
                step = arm.train_step_fn(B)
                for _ in range(3):
                    step()
                ms = _cuda_time(step, 10)
                out[name] = {"value": B / (ms / 1e3), "ms_per_step": ms}
                out["sample"] = f"10 timed training steps (fwd+bwd+Adam) at batch {B}, 3x{S}x{S}"
            else:
                B, S, evals = SAMPLE_CFG[workload]
                arm = RefArm(S, "cuda", autocast=amp)
                if evals > 100:                    # DDPM-1000: time 20 evaluations, extrapolate
                    step = arm.eval_fn(B)
                    step()
                    ms = _cuda_time(step, 20) * evals
                    out["sample"] = f"20 UNet evaluations at batch {B}, 3x{S}x{S}, x{evals // 20}"
                else:
                    arm.chain_fn(8, evals)()       # warm-up chain at a small batch
                    step = arm.chain_fn(B, evals)
                    ms = _cuda_time(step, 1)
                    out["sample"] = f"one DDIM-{evals} chain through sample(batch_size={B}), 3x{S}x{S}"
                out[name] = {"value": B / (ms / 1e3), "ms_per_step": ms}
            out["kind"] = arm.kind
            del arm, step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
    return out


def run_reference(args):
    """`--impl reference`: the reference's own implementation of the path on the host cores (all threads), on a
    bounded per-step sample; `--ref-device cuda` (informational) runs it eagerly on the GPU instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    if args.ref_device == "cuda":
        g = gpu_eager_baseline(args.workload)
        r = g["autocast_bf16" if args.ref_autocast else "fp32"]
        line = {"impl": "reference", "metric": metric_name(args.workload), "value": r["value"], "unit": "img/s",
                "n_gpus": 1, "steps": K, "warmup": W, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "bf16-autocast" if args.ref_autocast else "f32",
                "data": "synthetic", "config": {"workload": g["sample"], "device": g["device"], "kind": g["kind"]},
                "cpu_baseline": None, "e2e": None, "gpu_eager_baseline": g}
        print(json.dumps(line), flush=True)
        return
    torch.set_num_threads(os.cpu_count())
    if args.workload in TRAIN_CFG:
        TRAIN_B, TRAIN_S = TRAIN_CFG[args.workload]
        arm = RefArm(TRAIN_S, "cpu")
        probe = arm.train_step_fn(4)
        probe()
        t0 = time.perf_counter()
        probe()
        per_img = (time.perf_counter() - t0) / 4
        b = int(max(1, min(TRAIN_B, 150.0 / max(1, K + W) / per_img)))
        if args.ref_batch:
            b = min(b, args.ref_batch)
        step = arm.train_step_fn(b)
        for _ in range(max(W, 1)):                 # the first step at a new batch size pays oneDNN primitive creation
            step()
        t0 = time.perf_counter()
        for _ in range(K):
            step()
        dt = time.perf_counter() - t0
        value = b * K / dt
        sample = f"each step = one fp32 training step (fwd+bwd+Adam) on {b} of the {TRAIN_B} images of the batch"
        cfg = {"workload": f"DDPM train step UNet(dim=64) 3x{TRAIN_S}x{TRAIN_S} batch {TRAIN_B}/GPU (CPU sample batch {b})"}
    else:
        DDIM_B, DDIM_S, DDIM_STEPS = SAMPLE_CFG[args.workload]
        b = 1
        arm = RefArm(DDIM_S, "cpu")
        step = arm.eval_fn(b)
        for _ in range(min(W, 1)):
            step()
        n = max(1, min(K, 8))
        t0 = time.perf_counter()
        for _ in range(n):
            step()
        dt = time.perf_counter() - t0
        value = b * n / dt / DDIM_STEPS
        K = n
        sample = f"each step = one fp32 UNet evaluation of 1 image 3x{DDIM_S}x{DDIM_S}; img/s = evals/s / {DDIM_STEPS}"
        kind = "DDIM" if args.workload == "ddim" else "DDPM"
        cfg = {"workload": f"{kind}-{DDIM_STEPS} sampling 3x{DDIM_S}x{DDIM_S} batch {DDIM_B} (CPU sample: single evaluations)"}
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "img/s",
            "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": dt / K * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": os.cpu_count(), "kind": arm.kind, "sample": sample},
            "e2e": {"value": value, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def metric_name(workload):
    if workload in TRAIN_CFG:
        return "DDPM train img/s"
    return f"DDIM-{DDIM_STEPS} sample img/s" if workload == "ddim" else "DDPM-1000 sample img/s"


# --------------------------------------------------------------------------------------------------------
# per-kernel timing of one step (roofline of the dominant kernel)
# --------------------------------------------------------------------------------------------------------
def profile_plan(plan, passes=3, backward=True):
    """CUDA-event time of every launch of one forward(+backward) pass, in program order and natural cache
    state.  A leading device-side sleep lets the host run ahead so gaps between events are GPU time only."""
    ops = list(plan.fwd) + (list(plan.bwd) if backward else [])
    names = list(plan.fwd_names) + (list(plan.bwd_names) if backward else [])
    flops = list(plan.fwd_flops) + (list(plan.bwd_flops) if backward else [])
    tot = [0.0] * len(ops)
    from b200dm import _lib as L
    for _ in range(passes):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ops) + 1)]
        torch.cuda._sleep(20_000_000)                      # ~10 ms head start for the host
        st = L.stream_ptr()
        evs[0].record()
        for i, op in enumerate(ops):
            op(st)
            evs[i + 1].record()
        torch.cuda.synchronize()
        for i in range(len(ops)):
            tot[i] += evs[i].elapsed_time(evs[i + 1]) / passes
    profile_plan.detail = [{"i": i, "kernel": n, "us": round(ms * 1e3, 2), "gflop": round(fl / 1e9, 3)}
                           for i, (n, ms, fl) in enumerate(zip(names, tot, flops))]
    fam = {}
    for n, ms, fl in zip(names, tot, flops):
        f = fam.setdefault(n, {"ms": 0.0, "launches": 0, "flops": 0.0})
        f["ms"] += ms
        f["launches"] += 1
        f["flops"] += fl
    return fam, sum(tot)


# --------------------------------------------------------------------------------------------------------
REPEATS = 5          # timed regions of exactly K steps each; the median region is reported


class Ctx:
    """Process-wide state of one bench run (rank, device, barrier + max-over-ranks CUDA-event timing)."""

    def __init__(self):
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            from b200dm.distributed import nccl_options
            opts = nccl_options()               # caps the collective's CTAs (the backward kernels leave those SMs free)
            kw = {"pg_options": opts} if opts is not None else {}
            dist.init_process_group("nccl", device_id=self.dev, **kw)
            self.dist = dist
        self.peaks = measured_peaks()

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        """EXACTLY `steps` calls bracketed by barrier + synchronize on both sides; CUDA events; max over ranks."""
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return ms.item()

    def timed_median(self, fn, steps, repeats=REPEATS):
        all_ms = [self.timed(fn, steps) for _ in range(repeats)]
        return statistics.median(all_ms), [round(m / steps, 4) for m in all_ms]


def roofline_of(ctx, plan, is_train, step_flops, ms_per_step, profile_out, workload_desc):
    """Per-kernel CUDA-event table of one pass and the roofline entry of the dominant tensor-bound family."""
    peaks = ctx.peaks
    fam, total_ms = profile_plan(plan, passes=3, backward=is_train)
    table = {k: {"ms": round(v["ms"], 4), "launches": v["launches"], "share": round(v["ms"] / total_ms, 4),
                 "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] and v["ms"] > 0 else None}
             for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    k, v = max(((k, v) for k, v in fam.items() if v["flops"] > 0), key=lambda kv: kv[1]["ms"])
    achieved = v["flops"] / (v["ms"] * 1e-3) / 1e12
    roof = {"kernel": k, "bound": "tensor", "achieved": round(achieved, 1), "peak": peaks["tf_burst"],
            "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_burst"], 4), "traffic": profiled_traffic(k),
            "peak_source": peaks["source"] + " bf16_tflops (burst: kernel timed per launch with CUDA events)",
            "launches_per_pass": v["launches"], "avg_launch_us": round(v["ms"] * 1e3 / v["launches"], 2),
            "share_of_step": round(v["ms"] / total_ms, 4),
            "step_mfu": round(step_flops / (ms_per_step * 1e-3) / 1e12 / peaks["tf_sustained"], 4)}
    if profile_out:
        with open(profile_out, "w") as f:
            json.dump({"kernels": table, "sum_ms": total_ms, "workload": workload_desc,
                       "launches": getattr(profile_plan, "detail", None)}, f, indent=1)
    return roof, table


def run_train(ctx, workload, K, W, profile_out=None):
    """DDPM training step (configs[1] / configs[4]): q_sample -> UNet fwd -> loss -> bwd -> fused Adam -> EMA."""
    from b200dm import DDPM, _lib as L
    dev, rank, world, dist = ctx.dev, ctx.rank, ctx.world, ctx.dist
    clocks = ClockSampler(ctx.local)
    clocks.start()
    torch.manual_seed(10 + rank)
    B, S = TRAIN_CFG[workload]
    model = DDPM(img_channels=3, img_size=S, dim=64, diffusion_timesteps=1000, sampling_timesteps=None,
                 lr=2e-5, betas=(0.9, 0.99), ema_update_every=10, ema_decay=0.995, precision="bf16",
                 device=dev, overlap_optimizer=os.environ.get("B200DM_OVERLAP_OPT", "1") != "0")
    model.train()
    unet = model.ema.model.model
    opt = model.configure_optimizers()
    sync = None
    if dist is not None:       # DDP semantics: broadcast rank 0's weights, bucketed all-reduce overlapped with backward
        sync = unet.grad_sync or unet.enable_data_parallel()
        model.ema.ema_model.model.arena.flat.copy_(unet.arena.flat)
        opt.grad_scale = sync.grad_scale
    g = torch.Generator().manual_seed(10 + rank)
    host = [torch.rand(B, 3, S, S, generator=g).pin_memory() for _ in range(4)]
    labels = torch.zeros(B, dtype=torch.long, device=dev)
    dev_batches = [h.to(dev) for h in host]
    loss_host = torch.zeros(1).pin_memory()
    losses = []

    def step_device(i):
        opt.zero_grad()
        loss = model.training_step((dev_batches[i % 4], labels))
        loss.backward()                                    # includes the overlapped gradient all-reduce
        opt.step()
        model.on_train_batch_end(None, None, 0)

    copy_stream = torch.cuda.Stream(device=dev)

    def step_e2e(i):
        # public-API step from pinned host memory.  The loss of THIS step is read back on the host before the
        # next step starts; the read is issued as soon as the loss exists (after the forward pass, on a second
        # stream) so that the host does not wait for backward + Adam before it can enqueue the next step.
        data = host[i % 4].to(dev, non_blocking=True)       # H2D from pinned memory
        opt.zero_grad()
        loss = model.training_step((data, labels))
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ready)
            loss_host.copy_(loss.detach().reshape(1), non_blocking=True)   # D2H of the step's result
            done = torch.cuda.Event()
            done.record()
        loss.backward()
        opt.step()
        model.on_train_batch_end(None, None, 0)
        done.synchronize()                                  # the host holds this step's loss (loss.item())
        losses.append(float(loss_host[0]))

    # launches per step, counted on an eager (non-graph-replayed) step
    L.load().b200dm_reset_launch_count()
    step_device(0)
    torch.cuda.synchronize()
    launches_per_step = int(L.load().b200dm_launch_count())
    for i in range(W):
        step_device(i)
    clocks.mark()
    ms, all_ms = ctx.timed_median(step_device, K)
    exposed = None
    if sync is not None:
        # exposed (non-overlapped) all-reduce time: the same K steps with the collectives switched off
        unet.grad_sync = None
        step_device(0)
        ms_local, _ = ctx.timed_median(step_device, K, repeats=3)
        unet.grad_sync = sync
        from b200dm.distributed import broadcast_parameters
        broadcast_parameters(unet.arena, 0)                 # the ranks diverged during the local-only steps
        exposed = round((ms - ms_local) / K, 4)
    for i in range(2):
        step_e2e(i)
    losses.clear()
    ms_e2e, all_e2e = ctx.timed_median(step_e2e, K)
    clk = clocks.stop()
    # the timed region is not un-checked: every e2e step's loss was read back
    import math
    assert all(math.isfinite(v) for v in losses), "non-finite training loss inside the timed region"
    q = max(1, len(losses) // 4)
    loss_first, loss_last = sum(losses[:q]) / q, sum(losses[-q:]) / q
    # (t and the noise are redrawn every step, so quarter means scatter by ~15 %: the bound only catches divergence;
    #  tests/test_parity_configs_gpu.py::test_bench_step_loss_decreases_on_fixed_batch checks the decrease itself)
    assert loss_last < 2.0 * loss_first, (loss_first, loss_last)
    imgs = B * world * K
    plan = unet._plan(B, S, training=True)
    cfg = {"workload": f"DDPM train step UNet(dim=64) 3x{S}x{S} batch {B}/GPU, objective pred_v, sigmoid schedule "
                       "(reference defaults), fwd+loss+bwd+fused Adam+EMA",
           "global_batch": B * world, "parallelism": f"dp{world}",
           "l2": f"per-step working set {plan.nbytes / 1e9:.2f} GB of activations > 126 MB L2 (no flush needed)",
           "cuda_graph": bool(unet._cuda_graph), "repeats": REPEATS,
           "optimizer": "fused Adam + weight re-pack per gradient bucket, overlapped with backward"
           if opt.overlap else "fused Adam after backward"}
    res = {"metric": metric_name(workload), "value": imgs / (ms / 1e3), "unit": "img/s", "ms_per_step": ms / K,
           "ms_per_step_repeats": all_ms, "scaling": "weak", "config": cfg, "clocks": clk,
           "e2e": {"value": imgs / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": B * 3 * S * S * 4,
                   "d2h_bytes_per_step": 4, "ms_per_step_repeats": all_e2e},
           "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
           "loss": {"first_quarter_mean": round(loss_first, 5), "last_quarter_mean": round(loss_last, 5),
                    "steps_read_back": len(losses)}}
    if exposed is not None:
        res["exposed_allreduce_ms_per_step"] = exposed
    if rank == 0:
        res["roofline"], res["kernels"] = roofline_of(ctx, plan, True, 3 * F_FWD[S] * B, ms / K, profile_out,
                                                      cfg["workload"])
    del model, opt, plan
    torch.cuda.empty_cache()
    return res


def run_sample(ctx, workload, K, profile_out=None):
    """DDIM-50 (configs[2]) / DDPM-1000 (configs[3]) sampling, batch-sharded with no communication.  One "step" is
    one whole chain.  `value` = strong scaling of the BASELINE global batch; `weak` = the same per-GPU batch on
    every rank (N>1 only)."""
    from b200dm import GaussianDiffusion, Unet, _lib as L
    dev, rank, world = ctx.dev, ctx.rank, ctx.world
    clocks = ClockSampler(ctx.local)
    clocks.start()
    GB, S, STEPS = SAMPLE_CFG[workload]
    K = max(1, min(K, 3 if workload == "ddpm" else 4))      # a chain is 50 / 1000 UNet evaluations
    W = 3                                                   # three untimed chains (timing rules)
    repeats = 1 if workload == "ddpm" else 3
    unet = Unet(dim=64, channels=3, precision="bf16", device=dev)
    gd = GaussianDiffusion(unet, img_size=S, timesteps=1000, sampling_timesteps=STEPS if workload == "ddim" else None)

    def runner(global_batch):
        B = global_batch // world
        out_host = torch.zeros(B, 3, S, S).pin_memory()

        def step_device(i):
            gd.sample_shard(global_batch, rank, world, seed=i)

        def step_e2e(i):
            img = gd.sample_shard(global_batch, rank, world, seed=i)
            out_host.copy_(img)                                 # D2H of the images (the step's result)
            assert bool(torch.isfinite(out_host).all())
        return B, step_device, step_e2e

    B, step_device, step_e2e = runner(GB)
    step_device(0)
    torch.cuda.synchronize()
    # kernels per chain: graph replays bypass the library's launch counter, so one UNet evaluation is counted on an
    # eager pass of the plan and multiplied out (evaluations x (UNet + scheduler-step kernel) + the initial randn)
    L.load().b200dm_reset_launch_count()
    unet._plan(B, S, training=False).run_forward()
    torch.cuda.synchronize()
    launches_per_eval = int(L.load().b200dm_launch_count())
    launches_per_step = STEPS * (launches_per_eval + 1) + 1
    for i in range(W if workload == "ddim" else 1):
        step_device(i)
    clocks.mark()
    ms, all_ms = ctx.timed_median(step_device, K, repeats)
    ms_e2e, all_e2e = ctx.timed_median(step_e2e, K, repeats)
    weak = None
    if world > 1 and workload == "ddim":
        _, wdev, _ = runner(GB * world)
        wdev(0)
        wms, _ = ctx.timed_median(wdev, K, repeats)
        weak = {"value": GB * world * K / (wms / 1e3), "unit": "img/s", "global_batch": GB * world,
                "ms_per_step": wms / K, "scaling": "weak"}
    clk = clocks.stop()
    plan = unet._plan(B, S, training=False)
    kind = "DDIM-%d sampling (eta=0)" % STEPS if workload == "ddim" else "DDPM-1000 ancestral sampling"
    cfg = {"workload": f"{kind} UNet(dim=64) 3x{S}x{S} global batch {GB}, "
                       f"batch-sharded over {world} GPU(s) with no communication",
           "global_batch": GB, "parallelism": f"shard{world}",
           "l2": f"per-evaluation working set {plan.nbytes / 1e9:.2f} GB > 126 MB L2 (no flush needed)",
           "cuda_graph": bool(unet._cuda_graph), "repeats": repeats}
    res = {"metric": metric_name(workload), "value": GB * K / (ms / 1e3), "unit": "img/s", "steps": K, "warmup": W,
           "ms_per_step": ms / K, "ms_per_step_repeats": all_ms, "scaling": "strong", "config": cfg, "clocks": clk,
           "e2e": {"value": GB * K / (ms_e2e / 1e3), "unit": "img/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": B * 3 * S * S * 4, "ms_per_step_repeats": all_e2e},
           "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step}
    if weak is not None:
        res["weak"] = weak
    if rank == 0:
        res["roofline"], res["kernels"] = roofline_of(ctx, plan, False, F_FWD[S] * B * STEPS, ms / K, profile_out,
                                                      cfg["workload"])
    del unet, gd, plan
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["train", "ddim", "ddpm", "train64"], default="train")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip the eager-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--no-secondary", action="store_true",
                    help="default workload only: skip the DDIM-50 half of the BASELINE metric")
    ap.add_argument("--ref-device", choices=["cpu", "cuda"], default="cpu",
                    help="with --impl reference: cuda = the reference as eager PyTorch on the GPU (informational)")
    ap.add_argument("--ref-autocast", action="store_true", help="with --ref-device cuda: torch.autocast(bf16)")
    ap.add_argument("--ref-batch", type=int, default=0, help="with --impl reference: cap the CPU sample batch")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel time table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    ctx = Ctx()
    K, W = args.steps, args.warmup
    is_train = args.workload in TRAIN_CFG
    if is_train:
        res = run_train(ctx, args.workload, K, W, args.profile_out)
    else:
        res = run_sample(ctx, args.workload, K, args.profile_out)
    # BASELINE.json's metric is "DDPM train img/s & DDIM-50 sample img/s": the default run measures both halves
    secondary = None
    if args.workload == "train" and not args.no_secondary:
        prof2 = (args.profile_out[:-5] + "_ddim.json") if (args.profile_out or "").endswith(".json") else None
        secondary = run_sample(ctx, "ddim", K, prof2)

    if ctx.rank == 0:
        solo = ctx.world == 1
        line = {"metric": res["metric"], "value": res["value"], "unit": "img/s", "n_gpus": ctx.world,
                "steps": res.get("steps", K), "warmup": res.get("warmup", W), "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": res["scaling"], "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic"}
        for k in ("config", "clocks", "e2e", "gpu_launches", "launches_per_step", "ms_per_step_repeats", "loss",
                  "exposed_allreduce_ms_per_step", "weak", "roofline"):
            if k in res:
                line[k] = res[k]
        line["cpu_baseline"] = cpu_baseline(args.workload) if solo and not args.no_cpu_baseline else None
        line["gpu_eager_baseline"] = gpu_eager_baseline(args.workload) if solo and not args.no_gpu_baseline else None
        if secondary is not None:
            sec = {k: secondary[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "scaling",
                                             "config", "clocks", "e2e", "gpu_launches", "launches_per_step",
                                             "ms_per_step_repeats", "roofline") if k in secondary}
            sec["higher_is_better"], sec["dtype"], sec["data"] = True, "bf16", "synthetic"
            if "weak" in secondary:
                sec["weak"] = secondary["weak"]
            sec["cpu_baseline"] = cpu_baseline("ddim") if solo and not args.no_cpu_baseline else None
            sec["gpu_eager_baseline"] = gpu_eager_baseline("ddim") if solo and not args.no_gpu_baseline else None
            sec["kernels"] = secondary.get("kernels")
            line["secondary"] = sec
        line["kernels"] = res.get("kernels")
        print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
