"""World-size-2 gloo tests (CPU) of the data-parallel host logic: bucket layout, bucketed all-reduce ==
gradient averaging (what DDP does in the reference, utils/lightning_utils.py:41-43), weight broadcast,
and the batch-shard bookkeeping of the samplers."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from b200dm.distributed import GradSync, broadcast_parameters, buckets
from b200dm.params import ParamArena


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_buckets_tile_the_arena_in_backward_order():
    a = ParamArena(64, 3, "cpu")
    bs = buckets(a)
    assert len(bs) == 9
    covered = sorted(bs)
    assert covered[0][0] == 0 and covered[-1][1] == a.numel
    assert all(covered[i][1] == covered[i + 1][0] for i in range(len(covered) - 1))
    # first bucket to be reduced holds the final block, the last one the stem / time MLP / FiLM weights
    fb, fe = bs[0]
    assert fb <= a.offset["final_conv.weight"] < fe and fb <= a.offset["final_res_block.block1.proj.weight"] < fe
    hb, he = bs[-1]
    assert a.offset["init_conv.weight"] < he and a.offset["time_mlp.3.bias"] < he
    # FiLM projections live at the arena head: the rows of the two level-0 down blocks (the last ResnetBlocks of the
    # backward pass) in the last bucket, all other rows in the "film" bucket that is reduced behind level 1
    assert hb <= a.offset["downs.0.0.mlp.1.weight"] < he and hb <= a.offset["downs.0.1.mlp.1.weight"] < he
    fb2, fe2 = bs[6]
    assert fb2 == 0 and fe2 == hb == a.film_early_cols * a.time_dim
    assert all(fb2 <= a.offset[b + ".mlp.1.weight"] < fe2 for b in a.film_order[:-2])
    # bucket i must hold exactly the parameters whose gradients Plan.bwd_segments[i] produces
    for i, prefix in ((0, "final_"), (1, "ups."), (2, "mid_"), (3, "downs.3."), (4, "downs.2."), (5, "downs.1."),
                      (7, "downs.0.")):
        b, e = bs[i]
        for nm, _ in a.spec:
            if nm.startswith(prefix) and ".mlp.1." not in nm:
                assert b <= a.offset[nm] and a.offset[nm] + a._numel(nm) <= e, (prefix, nm)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    a = ParamArena(64, 1, "cpu")
    a.flat.copy_(torch.randn(a.numel))
    broadcast_parameters(a, 0)
    w0 = a.flat.clone()
    g_local = torch.randn(a.numel)
    a.gflat.copy_(g_local)
    sync = GradSync(a)
    for i in range(len(sync.buckets)):        # as the backward pass would: bucket by bucket
        sync.reduce_bucket(i)
    sync.finish()
    gathered = [torch.empty(a.numel) for _ in range(world)]
    dist.all_gather(gathered, g_local)
    ref = torch.stack(gathered).sum(0)
    ok_grad = torch.allclose(a.gflat, ref, atol=1e-6)
    gw = [torch.empty(a.numel) for _ in range(world)]
    dist.all_gather(gw, w0)
    ok_bcast = all(torch.equal(gw[0], t) for t in gw)
    # sampler sharding bookkeeping: rank r owns images [r*B/W, (r+1)*B/W) and Philox offset r*n
    B, n_img = 8, 3 * 32 * 32
    per = B // world
    offs = (rank * per * n_img, (rank + 1) * per * n_img)
    q.put((rank, ok_grad, ok_bcast, sync.grad_scale, offs))
    dist.barrier()
    dist.destroy_process_group()


def test_weight_pack_ranges_follow_the_gradient_buckets():
    """The optimiser-in-backward path re-packs one bucket's conv weights at a time: the per-bucket runs of pack-table
    entries must cover every conv exactly once, own all tiles of the batched pack, and carry the stem once."""
    from b200dm import _lib as L
    from b200dm.engine import WeightPack
    a = ParamArena(64, 3, "cpu")
    pack = WeightPack(a, L.BF16, with_dgrad=True)
    seen, tiles, stems = [], 0, 0
    for b, e in buckets(a):
        runs, has_stem = pack.range_runs(b, e)
        stems += int(has_stem)
        for first, n, tile_first, nt in runs:
            assert tile_first == pack._entry_info[first][1]
            seen.extend(range(first, first + n))
            tiles += nt
    assert sorted(seen) == list(range(pack.n_entries)) and tiles == pack.total_tiles and stems == 1
    # buckets are contiguous arena ranges in table order: a handful of launches per step, not one per conv
    assert sum(len(pack.range_runs(b, e)[0]) for b, e in buckets(a)) <= 12


def test_gloo_world2_bucketed_allreduce_and_broadcast():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_grad, ok_bcast, scale, offs in res:
        assert ok_grad and ok_bcast and scale == 0.5
    assert res[0][4][1] == res[1][4][0]           # shards are contiguous and disjoint
