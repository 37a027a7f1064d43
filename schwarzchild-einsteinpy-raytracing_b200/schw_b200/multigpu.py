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


def _torch_stream_ptr():
    """cudaStream_t of torch's current stream; the legacy default stream is passed as cudaStreamLegacy (0x1)
    because the C ABI reads a NULL stream as "the handle's own stream"."""
    import torch
    p = torch.cuda.current_stream().cuda_stream
    return p if p else 1


def render_frame(scene, tex, extras=False, handle=None):
    """Render the whole frame of `scene`; sharded + gathered when torch.distributed is initialised.

    world == 1: the host-buffer C-ABI call.  world > 1: the texture is uploaded once per rank, the rank's rows
    are rendered into device buffers (schw_render_dev on torch's current stream), the packed (rows, 4w) u8
    blocks are all_gather'ed on the device (NCCL) and copied to the host once.

    Returns dict(rgb (h,w,3) u8, cls (h,w) u8, q0 (4,)) plus, with extras, final_q (n,4), ic (n,9),
    hit (n,2), uv (n,2), nsteps (n,) in natural pixel order."""
    from .native import default_handle
    h, w = scene.height, scene.width
    rank, world = dist_info()
    hd = handle or default_handle()
    rb, re, rs = shard_rows(h, rank, world)
    q0, _ = hd.initial_conditions_px(np.array([scene.obs[0], scene.obs[1], scene.obs[2]]),
                                     np.zeros((1, 3)), scene.mass)
    if world == 1:
        out = hd.render(scene, tex, rb, re, rs, extras=extras)
        res = dict(rgb=out["rgb"], cls=out["cls"], q0=q0[0])
        if extras:
            res.update({k: out[k] for k in ("final_q", "ic", "hit", "uv", "nsteps")})
        return res

    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    max_rows = (h + world - 1) // world
    nloc = rows_of(h, rank, world)
    d_tex = torch.from_numpy(np.ascontiguousarray(tex, np.uint8)).to(dev) if tex is not None else None
    d_rgb = torch.zeros((max_rows, w, 3), dtype=torch.uint8, device=dev)
    d_cls = torch.zeros((max_rows, w), dtype=torch.uint8, device=dev)
    ex = {}
    if extras:
        ex = dict(final_q=torch.zeros((max_rows, w, 4), dtype=torch.float64, device=dev),
                  ic=torch.zeros((max_rows, w, 9), dtype=torch.float64, device=dev),
                  hit=torch.zeros((max_rows, w, 2), dtype=torch.float64, device=dev),
                  uv=torch.zeros((max_rows, w, 2), dtype=torch.int32, device=dev),
                  nsteps=torch.zeros((max_rows, w), dtype=torch.int32, device=dev))
    hd.render_dev(scene, d_tex.data_ptr() if d_tex is not None else 0, rb, re, rs, d_rgb.data_ptr(), d_cls.data_ptr(),
                  stream=_torch_stream_ptr(),
                  d_final_q=ex["final_q"].data_ptr() if extras else 0, d_ic=ex["ic"].data_ptr() if extras else 0,
                  d_hit=ex["hit"].data_ptr() if extras else 0, d_texel=ex["uv"].data_ptr() if extras else 0,
                  d_nsteps=ex["nsteps"].data_ptr() if extras else 0)
    packed = torch.cat([d_rgb.view(max_rows, 3 * w), d_cls], dim=1)                   # (max_rows, 4w) u8
    full = torch.empty((world * max_rows, 4 * w), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(full, packed)                                          # THE collective of a frame
    full = full.view(world, max_rows, 4 * w).cpu().numpy()
    img = unpermute_rows([full[r] for r in range(world)], h)
    res = dict(rgb=img[:, :3 * w].reshape(h, w, 3).copy(), cls=img[:, 3 * w:].copy(), q0=q0[0])
    for k, t in ex.items():                                                            # diagnostics, on request only
        g = torch.empty((world * max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(g, t)
        g = g.view((world, max_rows) + tuple(t.shape[1:])).cpu().numpy()
        a = unpermute_rows([g[r] for r in range(world)], h)
        res[k] = a.reshape(h * w, -1) if a.ndim == 3 else a.reshape(h * w)
    return res
