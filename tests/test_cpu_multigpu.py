"""not-gpu: the N>1 host logic (page sharding + single gather of result arenas) with world_size 2
over gloo on the CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ctd_b200 import multigpu


def test_shard_range_covers_all_pages():
    for n in (0, 1, 7, 16, 128, 129):
        for world in (1, 2, 3, 8):
            spans = [multigpu.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, H, W = 2, 64, 64
    lay = multigpu.arena_layout(B, H, W)
    arena = np.zeros(lay["total"], np.uint8)
    lo, hi = multigpu.shard_range(4, rank, world)
    # fake per-page results that encode the global page id
    for i, page in enumerate(range(lo, hi)):
        arena[i * H * W:(i + 1) * H * W] = page + 1
        det = arena[lay["det"]:lay["det"] + B * 300 * 6 * 4].view(np.float32).reshape(B, 300, 6)
        det[i, :page + 1] = page
        arena[lay["det_count"]:lay["det_count"] + B * 4].view(np.int32)[i] = page + 1
        arena[lay["n_labels"]:lay["n_labels"] + B * 4].view(np.int32)[i] = 10 * page
    out = multigpu.gather_arenas(torch.from_numpy(arena), dist, rank, world, dst=0)
    if rank == 0:
        res = []
        for r in range(world):
            u = multigpu.unpack_arena(out[r].numpy(), lay, B, H, W)
            for i in range(B):
                res.append((int(u["mask"][i, 0, 0]), len(u["det"][i]), int(u["n_labels"][i])))
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(1, 1, 0), (2, 2, 10), (3, 3, 20), (4, 4, 30)]
