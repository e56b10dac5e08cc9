"""2-GPU data-parallel parity (NCCL): two ranks with half of the batch each, bucketed all-reduce
overlapped with backward, must reproduce the single-GPU gradients of the full batch, and batch-sharded
sampling must reproduce the single-GPU samples.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _inputs():
    g = torch.Generator().manual_seed(4321)
    x = torch.rand(4, 3, 32, 32, generator=g)
    t = torch.randint(0, 1000, (4,), generator=g)
    noise = torch.randn(4, 3, 32, 32, generator=g)
    return x, t, noise


def _build(dev):
    from b200dm import GaussianDiffusion, Unet
    from oracle import ddpm_oracle as O
    unet = Unet(dim=64, channels=3, precision="fp32", device=dev)
    unet.load_reference_state_dict(O.synth_state_dict(64, 3, seed=10))
    return unet, GaussianDiffusion(unet, img_size=32, sampling_timesteps=3)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    unet, gd = _build(dev)
    sync = unet.enable_data_parallel()
    x, t, noise = _inputs()
    sl = slice(rank * 2, rank * 2 + 2)
    for _ in range(3):                      # 3rd iteration replays the captured per-bucket graphs
        unet.zero_grad()
        loss = gd.p_losses(x[sl].to(dev), t[sl].to(dev), noise=noise[sl].to(dev), _normalize=True)
        loss.backward()
    torch.cuda.synchronize()
    grads = (unet.arena.gflat * sync.grad_scale).cpu()
    # gradient accumulation under data parallel: micro-batches inside no_sync() only accumulate locally,
    # the last one all-reduces the sum; a second synchronising backward without zero_grad() is refused
    unet.zero_grad()
    xs, ts, ns = x[sl].to(dev), t[sl].to(dev), noise[sl].to(dev)
    with unet.no_sync():
        (gd.p_losses(xs[:1], ts[:1], noise=ns[:1], _normalize=True) / 2).backward()
    (gd.p_losses(xs[1:], ts[1:], noise=ns[1:], _normalize=True) / 2).backward()
    torch.cuda.synchronize()
    acc = (unet.arena.gflat * sync.grad_scale).cpu()
    refused = False
    try:
        (gd.p_losses(xs[1:], ts[1:], noise=ns[1:], _normalize=True) / 2).backward()
    except RuntimeError:
        refused = True
    # opt-in bf16 wire: same sums within bf16 rounding
    from b200dm.distributed import GradSync
    unet.grad_sync = GradSync(unet.arena, wire="bf16")
    unet.zero_grad()
    gd.p_losses(xs, ts, noise=ns, _normalize=True).backward()
    torch.cuda.synchronize()
    g16 = (unet.arena.gflat * sync.grad_scale).cpu()
    unet.grad_sync = sync
    imgs = gd.sample_shard(4, rank, world, seed=7).cpu()
    q.put((rank, grads, imgs, acc, refused, g16))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradients_and_sharded_samples_match_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    dev = torch.device("cuda", 0)
    unet, gd = _build(dev)
    x, t, noise = _inputs()
    gd.p_losses(x.to(dev), t.to(dev), noise=noise.to(dev), _normalize=True).backward()
    ref = unet.arena.gflat.cpu()
    for rank, grads, _, acc, refused, g16 in res:
        assert ((grads - ref).norm() / ref.norm()).item() < 1e-5, rank
        assert ((acc - ref).norm() / ref.norm()).item() < 1e-5, rank          # accumulation == one big batch
        assert refused, rank
        assert ((g16 - ref).norm() / ref.norm()).item() < 5e-3, rank          # bf16 on the wire
    full = gd.sample_shard(4, 0, 1, seed=7).cpu()
    both = torch.cat([res[0][2], res[1][2]], 0)
    assert (full - both).abs().max().item() < 1e-5
