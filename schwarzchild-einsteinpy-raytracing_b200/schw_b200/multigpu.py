"""Row sharding of one frame over the ranks of a torchrun job + the single gather.

One process per GPU.  Rank r renders rows r, r+G, r+2G, ... (row-cyclic: the cost of a ray depends
only on its distance from the image centre, so contiguous row blocks would be badly imbalanced);
every rank holds the full scene struct and a replica of the texture (<= 24.9 MB at 4K).  There is no
exchange during integration; the only collective is ONE all_gather of the packed per-rank block
(rgb ++ cls, 4 B/pixel) over NCCL (NVLink/NVSwitch) — or gloo in the CPU tests — followed by the
row un-permutation.  Optional per-ray diagnostics (final_q, ic, ...) are gathered only on request.

Opt-in de-duplicated render on N ranks: the ranks share the integration of the DISTINCT
camera rays instead of splitting the rows (`dedup_layout`): every rank plans the whole frame, integrates every
world-th distinct ray, one all_gather of their final states + step counts, then the per-pixel epilogue of its
own rows and the same frame gather as above.
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


def broadcast_object(obj, src=0):
    """Small Python object from rank `src` to all ranks (host-side bookkeeping, e.g. the sampled pixel list)."""
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


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


def dedup_layout(n_unique, world):
    """Slot layout of the distinct rays of a de-duplicated frame over `world` ranks (mirrors schw_dedup_plan_dev):
    distinct ray u (in order of camera angle) -> slot (u % world) * chunk + u // world, chunk = ceil(n_unique / world).
    Rank r integrates the contiguous slots [r*chunk, r*chunk + counts[r]) = the distinct rays u = r (mod world):
    neighbouring angles cost the same number of steps, so the shares are balanced.  Returns (chunk, counts)."""
    chunk = (n_unique + world - 1) // world
    counts = [len(range(r, n_unique, world)) for r in range(world)]
    return chunk, counts


def dedup_slot(u, world, chunk):
    """Slot of distinct ray(s) u (int or ndarray) in the layout of `dedup_layout`."""
    return (u % world) * chunk + u // world


def gather_ray_slots(local, world):
    """all_gather of every rank's `chunk` slots (numpy (chunk, ...) or torch tensor) -> (world*chunk, ...), ONE collective."""
    import torch
    import torch.distributed as dist
    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    full = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(full, t.contiguous())
    return full if isinstance(local, torch.Tensor) else full.numpy()


def _torch_stream_ptr():
    """cudaStream_t of torch's current stream; the legacy default stream is passed as cudaStreamLegacy (0x1)
    because the C ABI reads a NULL stream as "the handle's own stream"."""
    import torch
    p = torch.cuda.current_stream().cuda_stream
    return p if p else 1


def render_frame(scene, tex, extras=False, handle=None, root=None, dedup=False):
    """Render the whole frame of `scene`; sharded + gathered when torch.distributed is initialised.

    world == 1: the host-buffer C-ABI call.  world > 1: every rank uploads the texture on a side stream while its rows
    are already being integrated (schw_render_dev_pitched on torch's current stream; the kernel waits for the upload only
    at its first texel lookup) and writes rgb and class side by side into its (rows, 4w) u8 block, the blocks are
    exchanged on the device with ONE NCCL collective and copied to page-locked host arrays.  root=None: all_gather, every
    rank returns the frame; root=r: gather, only rank r receives and returns it (the other ranks get rgb = cls = None) --
    the frame then crosses PCIe once, not once per rank.
    dedup=True (opt-in): the de-duplicated render, same output bits.  world == 1: schw_render_dedup.  world > 1: the
    ranks share the integration of the distinct rays of the frame (see `dedup_layout`) -- one extra all_gather of their
    final states and step counts -- and each runs the per-pixel epilogue of its own rows.

    Returns dict(rgb (h,w,3) u8, cls (h,w) u8, q0 (4,)) plus, with extras, final_q (n,4), ic (n,9),
    hit (n,2), uv (n,2), nsteps (n,) in natural pixel order."""
    from .native import default_handle, observer_q0
    h, w = scene.height, scene.width
    rank, world = dist_info()
    hd = handle or default_handle()
    rb, re, rs = shard_rows(h, rank, world)
    q0 = observer_q0(scene)
    if world == 1:
        out = hd.render(scene, tex, rb, re, rs, extras=extras, dedup=dedup)
        res = dict(rgb=out["rgb"], cls=out["cls"], q0=q0)
        if extras:
            res.update({k: out[k] for k in ("final_q", "ic", "hit", "uv", "nsteps")})
        return res

    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", torch.cuda.current_device())
    max_rows = (h + world - 1) // world
    B = _buffers(dev, h, w, world, extras, tex is not None)
    main = torch.cuda.current_stream()
    stream = _torch_stream_ptr()
    d_flag = 0
    if tex is not None:
        # texture upload on the side stream, flag raised behind it (schw_set_tex_ready_flag)
        B["flag"].zero_()
        B["side"].wait_stream(main)
        with torch.cuda.stream(B["side"]):
            B["tex"].copy_(torch.from_numpy(np.ascontiguousarray(tex, np.uint8)), non_blocking=True)
            B["flag"].copy_(B["one"], non_blocking=True)
        d_flag = B["flag"].data_ptr()
    ex = B["extras"]
    xp = dict(d_final_q=ex["final_q"].data_ptr(), d_ic=ex["ic"].data_ptr(), d_hit=ex["hit"].data_ptr(),
              d_texel=ex["uv"].data_ptr(), d_nsteps=ex["nsteps"].data_ptr()) if extras else {}
    d_tex = B["tex"].data_ptr() if tex is not None else 0
    # rgb and class side by side in the (max_rows, 4w) message of the frame gather: no packing pass afterwards
    d_rgb = B["packed"].data_ptr()
    d_cls = d_rgb + 3 * w
    if not dedup:
        hd.render_dev_pitched(scene, d_tex, rb, re, rs, d_rgb, 4 * w, d_cls, 4 * w, stream=stream, d_tex_ready=d_flag, **xp)
        if tex is not None:
            main.wait_stream(B["side"])             # the upload belongs to this frame (and B["tex"] is reused by the next)
    else:
        n_unique, chunk = hd.dedup_plan_dev(scene, world, stream=stream)          # identical on every rank
        fq_all = torch.empty((world * chunk, 4), dtype=torch.float64, device=dev)
        ns_all = torch.empty((world * chunk,), dtype=torch.int32, device=dev)
        hd.dedup_integrate_dev(scene, rank, fq_all.data_ptr(), ns_all.data_ptr(), stream=stream)
        # THE extra exchange of the de-duplicated frame: final states + step counts of every rank's share
        dist.all_gather_into_tensor(fq_all, fq_all[rank * chunk:(rank + 1) * chunk].clone())
        dist.all_gather_into_tensor(ns_all, ns_all[rank * chunk:(rank + 1) * chunk].clone())
        if tex is not None:
            main.wait_stream(B["side"])             # the epilogue kernel has no upload flag to wait on
        hd.dedup_epilogue_dev(scene, d_tex, rb, re, rs, d_rgb, 4 * w, d_cls, 4 * w, fq_all.data_ptr(), ns_all.data_ptr(),
                              stream=stream, **xp)
    mine = root is None or rank == root

    def exchange(full, part):                       # THE collective of a frame (one more per diagnostic array)
        if root is None:
            dist.all_gather_into_tensor(full, part)
        else:
            dist.gather(part, list(full.view((world,) + tuple(part.shape)).unbind(0)) if mine else None, dst=root)

    exchange(B["gathered"] if mine else None, B["packed"])
    res = dict(rgb=None, cls=None, q0=q0)
    if mine:
        # rank blocks -> natural row order on the device: row i = r + world*k lives at [r, k]
        nat = B["gathered"].view(world, max_rows, 4 * w).permute(1, 0, 2).reshape(world * max_rows, 4 * w)[:h]
        # fresh arrays like the reference's np.zeros, but page-locked (torch's caching host allocator recycles the
        # blocks): the D2H of the frame runs at PCIe rate instead of being staged through the driver's bounce buffers
        rgb_t = torch.empty((h, w, 3), dtype=torch.uint8, pin_memory=True)
        cls_t = torch.empty((h, w), dtype=torch.uint8, pin_memory=True)
        rgb_t.view(h, 3 * w).copy_(nat[:, :3 * w], non_blocking=True)                  # D2H straight into the results
        cls_t.copy_(nat[:, 3 * w:], non_blocking=True)
        res.update(rgb=rgb_t.numpy(), cls=cls_t.numpy())
    for k, t in ex.items():                                                            # diagnostics, on request only
        g = torch.empty((world * max_rows,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) if mine else None
        exchange(g, t)
        if not mine:
            res[k] = None
            continue
        g = g.view((world, max_rows) + tuple(t.shape[1:])).transpose(0, 1).reshape((world * max_rows,) + tuple(t.shape[1:]))[:h]
        a = g.cpu().numpy()
        res[k] = a.reshape(h * w, -1) if a.ndim == 3 else a.reshape(h * w)
    hd.last_status()          # waits for this rank's work; raises what its device code flagged (schw_last_status)
    main.synchronize()
    return res


_BUF = {}


def _buffers(dev, h, w, world, extras, has_tex):
    """Per-(device, frame shape) device/pinned buffers, reused across frames (sweeps render many frames)."""
    import torch
    key = (str(dev), h, w, world)
    B = _BUF.get(key)
    max_rows = (h + world - 1) // world
    if B is None:
        B = dict(tex=torch.empty((h, w, 3), dtype=torch.uint8, device=dev),
                 packed=torch.zeros((max_rows, 4 * w), dtype=torch.uint8, device=dev),
                 gathered=torch.empty((world * max_rows, 4 * w), dtype=torch.uint8, device=dev), extras={},
                 side=torch.cuda.Stream(device=dev), flag=torch.zeros(1, dtype=torch.int32, device=dev),
                 one=torch.ones(1, dtype=torch.int32).pin_memory())
        _BUF.clear()                      # keep one frame shape resident
        _BUF[key] = B
    if extras and not B["extras"]:
        B["extras"] = dict(final_q=torch.zeros((max_rows, w, 4), dtype=torch.float64, device=dev),
                           ic=torch.zeros((max_rows, w, 9), dtype=torch.float64, device=dev),
                           hit=torch.zeros((max_rows, w, 2), dtype=torch.float64, device=dev),
                           uv=torch.zeros((max_rows, w, 2), dtype=torch.int32, device=dev),
                           nsteps=torch.zeros((max_rows, w), dtype=torch.int32, device=dev))
    if not extras:
        return dict(B, extras={})
    return B
