"""Row sharding of one frame over the ranks of a torchrun job + the single gather.

One process per GPU.  Rank r renders rows r, r+G, r+2G, ... (row-cyclic: the cost of a ray depends
only on its distance from the image centre, so contiguous row blocks would be badly imbalanced);
every rank holds the full scene struct and a replica of the texture (<= 24.9 MB at 4K).  There is no
exchange during integration; the only collective is ONE all_gather of the packed per-rank block
(rgb ++ cls, 4 B/pixel) over NCCL (NVLink/NVSwitch) — or gloo in the CPU tests — followed by the
row un-permutation.  Optional per-ray diagnostics (final_q, ic, ...) are gathered only on request.
"""
import os

import numpy as np


def dist_info():
    """(rank, world_size) of the current torch.distributed job, (0, 1) when not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def shard_rows(height, rank, world):
    """Rows of rank `rank`: (row_begin, row_end, row_stride) of the cyclic distribution."""
    return rank, height, world


def rows_of(height, rank, world):
    return len(range(rank, height, world))


def unpermute_rows(blocks, height):
    """blocks[r] holds rows r, r+G, ... of a (height, ...) array; returns the array in natural row order."""
    world = len(blocks)
    out = np.empty((height,) + blocks[0].shape[1:], blocks[0].dtype)
    for r, b in enumerate(blocks):
        out[r::world] = b[:rows_of(height, r, world)]
    return out


def gather_row_blocks(local, height, world, rank, device=None):
    """all_gather equally-padded per-rank row blocks and return the un-permuted (height, ...) array.

    `local`: numpy array (rows_of(rank), ...) or a torch tensor on the current device.
    One collective; rank blocks are padded to ceil(height/world) rows so the message sizes match."""
    import torch
    import torch.distributed as dist
    max_rows = (height + world - 1) // world
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    if t.shape[0] < max_rows:
        pad = torch.zeros((max_rows - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        t = torch.cat([t, pad], 0)
    full = torch.empty((world * max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(full, t.contiguous())          # rank blocks concatenated along dim 0
    full = full.view((world, max_rows) + tuple(t.shape[1:])).cpu().numpy()
    return unpermute_rows([full[r] for r in range(world)], height)


def render_frame(scene, tex, extras=False, handle=None):
    """Render the whole frame of `scene`; sharded + gathered when torch.distributed is initialised.

    Returns dict(rgb (h,w,3) u8, cls (h,w) u8, q0 (4,)) plus, with extras, final_q (n,4), ic (n,9),
    hit (n,2), uv (n,2), nsteps (n,) in natural pixel order."""
    from .native import default_handle
    h, w = scene.height, scene.width
    rank, world = dist_info()
    hd = handle or default_handle()
    rb, re, rs = shard_rows(h, rank, world)
    out = hd.render(scene, tex, rb, re, rs, extras=extras)
    q0, _ = hd.initial_conditions_px(np.array([scene.obs[0], scene.obs[1], scene.obs[2]]),
                                     np.zeros((1, 3)), scene.mass)
    if world == 1:
        res = dict(rgb=out["rgb"], cls=out["cls"], q0=q0[0])
        if extras:
            res.update({k: out[k] for k in ("final_q", "ic", "hit", "uv", "nsteps")})
        return res
    import torch
    dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else None
    nloc = out["rgb"].shape[0]
    packed = np.concatenate([out["rgb"].reshape(nloc, w * 3), out["cls"].reshape(nloc, w)], axis=1)   # (rows, 4w) u8
    full = gather_row_blocks(packed, h, world, rank, dev)
    res = dict(rgb=full[:, :3 * w].reshape(h, w, 3).copy(), cls=full[:, 3 * w:].copy(), q0=q0[0])
    if extras:
        for k, width in (("final_q", 4), ("ic", 9), ("hit", 2), ("uv", 2), ("nsteps", 1)):
            blk = out[k].reshape(nloc, w * width)
            res[k] = gather_row_blocks(blk, h, world, rank, dev).reshape(h * w, width) if width > 1 else \
                gather_row_blocks(blk, h, world, rank, dev).reshape(h * w)
    return res
