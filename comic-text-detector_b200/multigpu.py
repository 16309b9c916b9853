"""Page sharding + result gather for one-process-per-GPU runs.

Pages are independent (reference inference.py:141-178 is pure per image), so the only
"communication" is a gather of fixed-size per-rank result arenas to rank 0.  The functions are
backend-agnostic (`nccl` on GPUs, `gloo` in the CPU tests)."""
import numpy as np


def shard_range(n_pages, rank, world):
    """Contiguous shard [lo, hi) of `n_pages` for `rank`; shards differ by at most one page."""
    base, rem = divmod(int(n_pages), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def arena_layout(max_batch, h, w):
    """Byte offsets inside an engine's result arena (mirrors engine.cu: 256-byte aligned fields)."""
    al = lambda v: (v + 255) // 256 * 256
    o_det = al(max_batch * h * w)
    o_cnt = o_det + al(max_batch * 300 * 6 * 4)
    o_nl = o_cnt + al(max_batch * 4)
    o_lb = o_nl + al(max_batch * 4)
    o_ls = o_lb + al(max_batch * 1000 * 8 * 2)
    o_lc = o_ls + al(max_batch * 1000 * 4)
    return dict(mask=0, det=o_det, det_count=o_cnt, n_labels=o_nl, line_boxes=o_lb, line_scores=o_ls, line_count=o_lc,
                total=o_lc + al(max_batch * 4))


def gather_arenas(local_arena, dist, rank, world, dst=0):
    """One collective: every rank contributes its uint8 arena tensor; rank `dst` gets the list."""
    import torch
    out = [torch.empty_like(local_arena) for _ in range(world)] if rank == dst else None
    dist.gather(local_arena, gather_list=out, dst=dst)
    return out


def unpack_arena(arena_u8, max_batch, n, h, w):
    """arena bytes (numpy uint8) -> dict(mask u8 [n,h,w], det list of [k,6] f32, n_labels i32 [n])."""
    lay = arena_layout(max_batch, h, w)
    a = np.asarray(arena_u8)
    mask = a[lay["mask"]:lay["mask"] + n * h * w].reshape(n, h, w)
    det = a[lay["det"]:lay["det"] + n * 300 * 6 * 4].view(np.float32).reshape(n, 300, 6)
    cnt = a[lay["det_count"]:lay["det_count"] + n * 4].view(np.int32)
    nl = a[lay["n_labels"]:lay["n_labels"] + n * 4].view(np.int32)
    lb = a[lay["line_boxes"]:lay["line_boxes"] + n * 1000 * 16].view(np.int16).reshape(n, 1000, 4, 2)
    ls = a[lay["line_scores"]:lay["line_scores"] + n * 1000 * 4].view(np.float32).reshape(n, 1000)
    lc = a[lay["line_count"]:lay["line_count"] + n * 4].view(np.int32)
    return dict(mask=mask, det=[det[i, :cnt[i]] for i in range(n)], n_labels=nl,
                line_boxes=[lb[i, :lc[i]] for i in range(n)], line_scores=[ls[i, :lc[i]] for i in range(n)])
