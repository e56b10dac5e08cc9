#!/usr/bin/env python
"""`train.py`-compatible entry point for the diffusion path (reference train.py:24-141), without Lightning / W&B /
torchvision datasets:

    python lightning-generative-models_b200/train.py --config_path configs/diffusion/ddpm.json \\
        [--gpus N] [--precision bf16-mixed] [--max_steps K] [--max_epochs E] [--accumulate_grad_batches A] \\
        [--ckpt_path last.ckpt]

JSON -> `load_config` -> `load_model` (the loader convention resolves to b200dm.DDPM) -> one process per GPU
(`--gpus N` re-launches this script under torch.distributed.run; NCCL over NVLink) -> the Lightning fit loop the
reference relies on, restated: training_step -> backward (gradient all-reduce overlapped, `no_sync` on accumulation
micro-batches) -> optimizer step -> on_train_batch_end (EMA) -> every `--sample_every` steps rank 0 samples 64
images from the EMA model (ddpm.py:1025-1042; the grid is saved instead of sent to W&B) -> validation every
`--check_val_every_n_epoch` epochs -> `last.ckpt` in the Lightning layout.
`--precision`: None / "32" / "32-true" = fp32 (the reference's default), "bf16-mixed" / "bf16" = the tensor-core path.
"""
import argparse
import json
import os
import subprocess
import sys
import time
from datetime import datetime

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def parse_args(argv=None):
    p = argparse.ArgumentParser("Train script")
    p.add_argument("--config_path", type=str, required=True, help="Path to configs")
    p.add_argument("--num_workers", type=int, default=0)
    p.add_argument("--check_val_every_n_epoch", type=int, default=5)
    p.add_argument("--max_epochs", type=int, default=-1)
    p.add_argument("--max_steps", type=int, default=-1)
    p.add_argument("--accumulate_grad_batches", type=int, default=1)
    p.add_argument("--precision", type=str, default=None)
    p.add_argument("--ckpt_path", type=str, default=None)
    p.add_argument("--experiment_name", type=str, default=datetime.now().strftime("%Y-%m-%d_%H:%M"))
    p.add_argument("--experiment_dir", type=str, default=None)
    # B200 launcher options (the reference picks DDP over all visible GPUs through Lightning)
    p.add_argument("--gpus", type=int, default=1, help="processes / GPUs on this node")
    p.add_argument("--master_port", type=int, default=29517)
    p.add_argument("--sample_every", type=int, default=1000, help="global steps between sample(64) on rank 0")
    p.add_argument("--log_every", type=int, default=50)
    p.add_argument("--num_images", type=int, default=2048, help="size of the synthetic dataset")
    p.add_argument("--lr", type=float, default=None, help="override the config's learning rate")
    args = p.parse_args(argv)
    if args.max_epochs < 0 and args.max_steps < 0:
        args.max_epochs = 1000                      # Lightning's default when both are unset
    return args


def precision_of(flag):
    if flag in (None, "32", "32-true", "fp32"):
        return "fp32"
    if flag in ("bf16-mixed", "bf16", "bf16-true"):
        return "bf16"
    raise ValueError(f"unsupported --precision {flag!r} (32-true | bf16-mixed)")


def relaunch_multi_gpu(args, argv):
    """`--gpus N` outside a torchrun worker: one process per GPU on this node."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
           "--master-addr", "127.0.0.1", "--master-port", str(args.master_port), os.path.abspath(__file__)] + argv
    return subprocess.call(cmd)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    args = parse_args(argv)
    if args.gpus > 1 and "RANK" not in os.environ:
        return relaunch_multi_gpu(args, argv)

    import torch
    from b200dm.distributed import init_from_env
    from data.datamodule import DataModule
    from utils.loader import load_config, load_model

    rank, local, world = init_from_env()
    torch.manual_seed(10)                           # seed_everything(seed=10), train.py:20
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg = load_config(args.config_path)
    if args.lr is not None:
        cfg["model"]["args"]["lr"] = args.lr
    exp_dir = args.experiment_dir or os.path.join(HERE, "experiments", cfg["model"]["name"], args.experiment_name)
    if rank == 0:
        os.makedirs(exp_dir, exist_ok=True)
        with open(os.path.join(exp_dir, os.path.basename(args.config_path)), "w") as f:
            json.dump({"args": {k: v for k, v in vars(args).items()}, "config": cfg}, f, indent=1, default=str)

    model = load_model(cfg["model"], precision=precision_of(args.precision), device=dev)
    dm = DataModule(**cfg["dataset"], num_workers=args.num_workers, pin_memory=True, num_images=args.num_images,
                    world_size=world, rank=rank, device=dev)
    dm.setup()
    model.train()
    opt = model.configure_optimizers()              # also enables the bucketed gradient all-reduce when world > 1
    unet = model.ema.model.model
    epoch0 = 0
    if args.ckpt_path:
        ck = torch.load(args.ckpt_path, map_location="cpu", weights_only=False)
        model.load_checkpoint(ck, opt)
        epoch0 = int(ck.get("epoch", 0))
    acc = max(1, args.accumulate_grad_batches)
    log = open(os.path.join(exp_dir, "train_log.jsonl"), "a") if rank == 0 else None

    def emit(**kw):
        if log is not None:
            log.write(json.dumps(kw) + "\n")
            log.flush()
            print(json.dumps(kw), flush=True)

    def sample_grid(step):
        imgs = model.sample(batch_size=64)          # EMA model, ddpm.py:1029-1031
        path = os.path.join(exp_dir, f"samples_step{step:07d}.pt")
        torch.save(imgs.cpu(), path)
        try:
            from torchvision.utils import save_image
            save_image(imgs, path[:-3] + ".png", nrow=8)
        except Exception:                           # noqa: BLE001  (torchvision is optional here)
            pass
        emit(event="sample", step=step, shape=list(imgs.shape), min=float(imgs.min()), max=float(imgs.max()))

    step, micro, done = model.global_step, 0, False
    t0 = time.perf_counter()
    epoch = epoch0
    opt.zero_grad()
    while not done and (args.max_epochs < 0 or epoch < args.max_epochs):
        model.train()
        for batch in dm.train_dataloader(epoch):
            last_micro = (micro + 1) % acc == 0
            if step % args.sample_every == 0 and micro % acc == 0 and rank == 0:
                sample_grid(step)                   # `global_step % 1000 == 0 and is_master_process()`, ddpm.py:1025
                model.train()
            if last_micro:
                loss = model.training_step(batch)
                (loss / acc if acc > 1 else loss).backward()
            else:
                with unet.no_sync():
                    loss = model.training_step(batch)
                    (loss / acc).backward()
            micro += 1
            if not last_micro:
                continue
            opt.step()
            opt.zero_grad()
            model.on_train_batch_end(None, batch, step)
            step += 1
            if step % args.log_every == 0 or step == 1:
                synced = model.logged.get("train_loss", loss) if hasattr(model, "logged") else loss
                torch.cuda.synchronize()
                emit(event="train", step=step, epoch=epoch, train_loss=float(synced),
                     img_per_s=round(step * dm.batch_size * world * acc / (time.perf_counter() - t0), 1))
            if 0 <= args.max_steps <= step:
                done = True
                break
        epoch += 1
        if not done and epoch % max(1, args.check_val_every_n_epoch) == 0:
            model.eval()
            tot, n = 0.0, 0
            with torch.no_grad():
                for batch in dm.val_dataloader():
                    tot += float(model.validation_step(batch))
                    n += 1
            emit(event="val", step=step, epoch=epoch, val_loss=tot / max(n, 1))
            if rank == 0:
                model.save_checkpoint(os.path.join(exp_dir, "last.ckpt"), opt, epoch)
    torch.cuda.synchronize()
    if rank == 0:
        model.save_checkpoint(os.path.join(exp_dir, "last.ckpt"), opt, epoch)
        emit(event="done", step=step, epoch=epoch, seconds=round(time.perf_counter() - t0, 2),
             checkpoint=os.path.join(exp_dir, "last.ckpt"))
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
