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
    `ncu --set full` capture (profiles/r1_roofline_traffic.json, scripts/make_profiles.py); None if absent."""
    p = os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")
    if not os.path.isfile(p):
        return None
    d = json.load(open(p))
    names = {"wgrad_tc": ("wgrad3x3_halo_kernel", "wgrad_tc_kernel"),
             "conv_tc_fwd": ("conv3x3_halo_kernel", "conv_tc_kernel"),
             "conv_tc_dgrad": ("conv3x3_halo_kernel", "conv_tc_kernel")}.get(family, (family,))
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
# CPU arm: the oracle (reference algorithm) on the host cores
# --------------------------------------------------------------------------------------------------------
def cpu_train_step_fn(batch, TRAIN_S=TRAIN_S):
    from oracle import ddpm_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = {k: v.requires_grad_(True) for k, v in O.synth_state_dict(64, 3, seed=10).items()}
    orc = O.DiffusionOracle(sd, img_size=TRAIN_S, channels=3)
    opt = torch.optim.Adam(list(sd.values()), lr=2e-5, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(10)

    def step(b=batch):
        x = torch.rand(b, 3, TRAIN_S, TRAIN_S, generator=g)
        t = torch.randint(0, 1000, (b,), generator=g)
        noise = torch.randn(b, 3, TRAIN_S, TRAIN_S, generator=g)
        opt.zero_grad()
        loss = orc.forward(x, t, noise)
        loss.backward()
        opt.step()
        return loss.item()
    return step


def cpu_ddim_eval_fn(batch, DDIM_S=DDIM_S):
    from oracle import ddpm_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = O.synth_state_dict(64, 3, seed=10)
    g = torch.Generator().manual_seed(10)
    x = torch.randn(batch, 3, DDIM_S, DDIM_S, generator=g)
    t = torch.full((batch,), 500, dtype=torch.long)

    def step():
        with torch.no_grad():
            return O.unet_forward(sd, x, t).sum().item()
    return step


def cpu_baseline(workload, budget_s=20.0):
    """Bounded sample of the same workload on the host cores (oracle = port of the reference)."""
    if workload in TRAIN_CFG:
        S = TRAIN_CFG[workload][1]
        b = 16 if S == 32 else 4
        step = cpu_train_step_fn(b, S)
        step()
        t0, n = time.perf_counter(), 0
        while True:
            step()
            n += 1
            if time.perf_counter() - t0 > budget_s or n >= 8:
                break
        dt = time.perf_counter() - t0
        return {"value": b * n / dt, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
                "sample": f"{n} fp32 training steps (fwd+bwd+Adam) at batch {b}, 3x{S}x{S}, torch CPU {os.cpu_count()} threads"}
    _, S, evals = SAMPLE_CFG[workload]
    b = 2 if S == 64 else 8
    step = cpu_ddim_eval_fn(b, S)
    step()
    t0, n = time.perf_counter(), 0
    while True:
        step()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 6:
            break
    dt = time.perf_counter() - t0
    return {"value": b * n / dt / evals, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{n} fp32 UNet evaluations at batch {b}, 3x{S}x{S}, extrapolated x{evals} steps per image"}


def run_reference_gpu(args):
    """`--impl reference --ref-device cuda` (informational, not part of the driver's contract): the same oracle —
    the reference's algorithm as plain PyTorch ops — run eagerly on the GPU (cuDNN / cuBLAS / ATen kernels), fp32 or
    under torch.autocast(bf16) like the reference's `--precision bf16-mixed`.  This is the kernel-for-kernel bar on the
    same B200; full batch, K timed steps."""
    from oracle import ddpm_oracle as O
    dev = torch.device("cuda", 0)
    K, W = args.steps, max(args.warmup, 3)
    amp = args.ref_autocast
    ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if amp else (lambda: torch.autocast("cuda", enabled=False))
    g = torch.Generator(device=dev).manual_seed(10)
    if args.workload in TRAIN_CFG:
        B, S = TRAIN_CFG[args.workload]
        sd = {k: v.to(dev).requires_grad_(True) for k, v in O.synth_state_dict(64, 3, seed=10).items()}
        orc = O.DiffusionOracle(sd, img_size=S, channels=3)
        orc.buf = {k: v.to(dev) for k, v in orc.buf.items()}
        opt = torch.optim.Adam(list(sd.values()), lr=2e-5, betas=(0.9, 0.99))
        x = torch.rand(B, 3, S, S, generator=g, device=dev)

        def step():
            t = torch.randint(0, 1000, (B,), generator=g, device=dev)
            noise = torch.randn(B, 3, S, S, generator=g, device=dev)
            opt.zero_grad()
            with ctx():
                loss = orc.forward(x, t, noise)
            loss.backward()
            opt.step()
        imgs_per_step, what = B, f"DDPM train step 3x{S}x{S} batch {B}"
    else:
        B, S, evals = SAMPLE_CFG[args.workload]
        sd = {k: v.to(dev) for k, v in O.synth_state_dict(64, 3, seed=10).items()}
        x = torch.randn(B, 3, S, S, generator=g, device=dev)
        t = torch.full((B,), 500, dtype=torch.long, device=dev)

        def step():
            with torch.no_grad(), ctx():
                O.unet_forward(sd, x, t)
        imgs_per_step, what = B / evals, f"UNet evaluation 3x{S}x{S} batch {B}; img/s = evals/s x batch / {evals}"
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    value = imgs_per_step / (ms / 1e3)
    kind = "eager PyTorch on the same GPU, " + ("torch.autocast(bf16)" if amp else "fp32 (TF32 off)")
    line = {"impl": "reference", "metric": metric_name(args.workload), "value": value, "unit": "img/s", "n_gpus": 1,
            "steps": K, "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16-autocast" if amp else "f32", "data": "synthetic",
            "config": {"workload": what, "device": torch.cuda.get_device_name(0), "kind": kind},
            "cpu_baseline": None, "e2e": None}
    print(json.dumps(line), flush=True)


def run_reference(args):
    """`--impl reference`: the reference's own CPU path (oracle port), bounded per-step sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.ref_device == "cuda":
        return run_reference_gpu(args)
    K, W = args.steps, args.warmup
    if args.workload in TRAIN_CFG:
        TRAIN_B, TRAIN_S = TRAIN_CFG[args.workload]
        probe = cpu_train_step_fn(4, TRAIN_S)
        probe()
        t0 = time.perf_counter()
        probe()
        per_img = (time.perf_counter() - t0) / 4
        b = int(max(1, min(TRAIN_B, 150.0 / max(1, K + W) / per_img)))
        step = cpu_train_step_fn(b, TRAIN_S)
        for _ in range(W):
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
        step = cpu_ddim_eval_fn(b, DDIM_S)
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
            "cpu_baseline": {"value": value, "unit": "img/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", choices=["train", "ddim", "ddpm", "train64"], default="train")
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-device", choices=["cpu", "cuda"], default="cpu",
                    help="with --impl reference: cuda = the oracle as eager PyTorch on the GPU (informational)")
    ap.add_argument("--ref-autocast", action="store_true", help="with --ref-device cuda: torch.autocast(bf16)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel time table (JSON) here")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # nvidia-smi needs a few hundred ms before its first line: start it now, samples before mark() are ignored
    clocks = ClockSampler(local)
    clocks.start()
    from b200dm import DDPM, _lib as L
    peaks = measured_peaks()
    K, W = args.steps, args.warmup
    dev = torch.device("cuda", local)
    torch.manual_seed(10 + rank)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if dist is not None:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    is_train = args.workload in TRAIN_CFG
    if is_train:
        B, S = TRAIN_CFG[args.workload]
        model = DDPM(img_channels=3, img_size=S, dim=64, diffusion_timesteps=1000, sampling_timesteps=None,
                     lr=2e-5, betas=(0.9, 0.99), ema_update_every=10, ema_decay=0.995, precision="bf16",
                     device=dev, overlap_optimizer=os.environ.get("B200DM_OVERLAP_OPT", "1") != "0")
        model.train()
        unet = model.ema.model.model
        opt = model.configure_optimizers()
        if dist is not None:       # DDP semantics: broadcast rank 0's weights, bucketed all-reduce overlapped with backward
            sync = unet.enable_data_parallel()
            model.ema.ema_model.model.arena.flat.copy_(unet.arena.flat)
            opt.grad_scale = sync.grad_scale
        g = torch.Generator().manual_seed(10 + rank)
        host = [torch.rand(B, 3, S, S, generator=g).pin_memory() for _ in range(4)]
        labels = torch.zeros(B, dtype=torch.long, device=dev)
        dev_batches = [h.to(dev) for h in host]
        loss_host = torch.zeros(1).pin_memory()

        def step_core(data):
            opt.zero_grad()
            loss = model.training_step((data, labels))
            loss.backward()                                    # includes the overlapped gradient all-reduce
            opt.step()
            model.on_train_batch_end(None, None, 0)
            return loss

        def step_device(i):
            step_core(dev_batches[i % 4])

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
            assert loss_host[0] == loss_host[0]                 # touch the value (and catch a NaN)

        # launches per step, counted on an eager (non-graph-replayed) step
        L.load().b200dm_reset_launch_count()
        step_device(0)
        torch.cuda.synchronize()
        launches_per_step = int(L.load().b200dm_launch_count())
        for i in range(W):
            step_device(i)
        clocks.mark()
        ms = timed(step_device, K)
        for i in range(2):
            step_e2e(i)
        ms_e2e = timed(step_e2e, K)
        clk = clocks.stop()
        imgs = B * world * K
        value, e2e_value = imgs / (ms / 1e3), imgs / (ms_e2e / 1e3)
        h2d, d2h = B * 3 * S * S * 4, 4
        plan = unet._plan(B, S, training=True)
        flop_per_img = 3 * F_FWD[S]
        cfg = {"workload": f"DDPM train step UNet(dim=64) 3x{S}x{S} batch {B}/GPU, objective pred_v, sigmoid schedule "
                           "(reference defaults), fwd+loss+bwd+fused Adam+EMA",
               "global_batch": B * world, "parallelism": f"dp{world}",
               "l2": f"per-step working set {plan.nbytes / 1e9:.2f} GB of activations > 126 MB L2 (no flush needed)",
               "cuda_graph": bool(unet._cuda_graph),
               "optimizer": "fused Adam + weight re-pack per gradient bucket, overlapped with backward"
               if opt.overlap else "fused Adam after backward"}
        step_flops = flop_per_img * B
    else:
        from b200dm import GaussianDiffusion, Unet
        DDIM_B, DDIM_S, DDIM_STEPS = SAMPLE_CFG[args.workload]
        if args.workload == "ddpm":
            K = min(K, 3)                 # one step = a whole 1000-evaluation ancestral chain
        B, S = DDIM_B // world, DDIM_S
        unet = Unet(dim=64, channels=3, precision="bf16", device=dev)
        gd = GaussianDiffusion(unet, img_size=S, timesteps=1000,
                               sampling_timesteps=DDIM_STEPS if args.workload == "ddim" else None)
        out_host = torch.zeros(B, 3, S, S).pin_memory()

        def step_device(i):
            gd.sample_shard(DDIM_B, rank, world, seed=i)

        def step_e2e(i):
            img = gd.sample_shard(DDIM_B, rank, world, seed=i)
            out_host.copy_(img)                                 # D2H of the images (the step's result)

        L.load().b200dm_reset_launch_count()
        step_device(0)
        torch.cuda.synchronize()
        launches_per_step = int(L.load().b200dm_launch_count())
        W = 3                          # a chain is 50 / 1000 UNet evaluations; three untimed chains (timing rules)
        for i in range(W):
            step_device(i)
        clocks.mark()
        ms = timed(step_device, K)
        ms_e2e = timed(step_e2e, K)
        clk = clocks.stop()
        imgs = DDIM_B * K
        value, e2e_value = imgs / (ms / 1e3), imgs / (ms_e2e / 1e3)
        h2d, d2h = 0, B * 3 * S * S * 4
        plan = unet._plan(B, S, training=False)
        kind = "DDIM-%d sampling (eta=0)" % DDIM_STEPS if args.workload == "ddim" else "DDPM-1000 ancestral sampling"
        cfg = {"workload": f"{kind} UNet(dim=64) 3x{S}x{S} global batch {DDIM_B}, "
                           f"batch-sharded over {world} GPU(s) with no communication",
               "global_batch": DDIM_B, "parallelism": f"shard{world}",
               "l2": f"per-evaluation working set {plan.nbytes / 1e9:.2f} GB > 126 MB L2 (no flush needed)",
               "cuda_graph": bool(unet._cuda_graph)}
        step_flops = F_FWD[S] * B * DDIM_STEPS

    # per-kernel table and roofline of the dominant kernel (rank 0)
    roof, table = None, None
    if rank == 0:
        fam, total_ms = profile_plan(plan, passes=3, backward=is_train)
        table = {k: {"ms": round(v["ms"], 4), "launches": v["launches"], "share": round(v["ms"] / total_ms, 4),
                     "tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1) if v["flops"] and v["ms"] > 0 else None}
                 for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
        top = max(((k, v) for k, v in fam.items() if v["flops"] > 0), key=lambda kv: kv[1]["ms"])
        k, v = top
        achieved = v["flops"] / (v["ms"] * 1e-3) / 1e12
        roof = {"kernel": k, "bound": "tensor", "achieved": round(achieved, 1), "peak": peaks["tf_burst"],
                "unit": "TFLOP/s", "frac": round(achieved / peaks["tf_burst"], 4), "traffic": profiled_traffic(k),
                "peak_source": peaks["source"] + " bf16_tflops (burst: kernel timed per launch with CUDA events)",
                "launches_per_pass": v["launches"], "avg_launch_us": round(v["ms"] * 1e3 / v["launches"], 2),
                "share_of_step": round(v["ms"] / total_ms, 4),
                "step_mfu": round(step_flops / (ms / K * 1e-3) / 1e12 / peaks["tf_sustained"], 4)}
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                json.dump({"kernels": table, "sum_ms": total_ms, "workload": cfg["workload"],
                           "launches": getattr(profile_plan, "detail", None)}, f, indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.workload)

    if rank == 0:
        line = {"metric": metric_name(args.workload), "value": value, "unit": "img/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak"
                if is_train else "strong", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": cfg, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": "img/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "gpu_launches": launches_per_step * K, "launches_per_step": launches_per_step,
                "roofline": roof, "cpu_baseline": cpu, "kernels": table}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
