"""world_size-2 (and 3) CPU test of the multi-GPU host logic over gloo: row-cyclic sharding, the single
all_gather of packed rgb+class blocks, un-permutation.  The per-rank row blocks are produced by the CPU
oracle standing in for the kernel (there is no GPU here); the gathered frame must equal the oracle's
whole-frame render bit for bit on every rank."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, q):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from schw_b200.multigpu import dist_info, gather_row_blocks, shard_rows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert dist_info() == (rank, world)
        rng = np.random.default_rng(5)
        tex = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        sc = O.make_scene(height=h, width=w, delta=0.2)
        rb, re, rs = shard_rows(h, rank, world)
        loc = O.render(sc, tex, rb, re, rs, nthreads=1)
        nloc = loc["rgb"].shape[0]
        packed = np.concatenate([loc["rgb"].reshape(nloc, 3 * w), loc["cls"].reshape(nloc, w)], axis=1)
        full = gather_row_blocks(packed, h, world, rank)
        fq = gather_row_blocks(loc["final_q"].reshape(nloc, w * 4), h, world, rank)
        ref = O.render(sc, tex, nthreads=1)
        ok = (np.array_equal(full[:, :3 * w].reshape(h, w, 3), ref["rgb"]) and np.array_equal(full[:, 3 * w:], ref["cls"])
              and np.array_equal(fq.reshape(-1, 4), ref["final_q"], equal_nan=True))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w", [(2, 13, 20), (3, 10, 16)])
def test_row_cyclic_gather_over_gloo(world, h, w):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]


def _dedup_worker(rank, world, port, h, w, q):
    """The multi-rank de-duplicated render (schw_b200.multigpu.render_frame(dedup=True)) with the CPU oracle standing in
    for the three device stages: same slot layout (dedup_layout / dedup_slot), same all_gather of the shares."""
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import oracle as O
    from schw_b200.multigpu import dedup_layout, dedup_slot, gather_ray_slots, gather_row_blocks, shard_rows
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(11)
        tex = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        sc = O.make_scene(height=h, width=w, delta=0.2)
        # plan (every rank, whole frame): initial conditions of every pixel, distinct rays by the bits of p0 (a function
        # of the camera angle alone), in ascending order of the key like the device's radix sort
        ij = np.array([(i, j) for i in range(h) for j in range(w)], np.int32)
        q0s, ic, rc = O.initial_conditions(np.array([30.0, 0, 0]), O.camera_pixels(sc, ij), 1.0)
        keys = np.ascontiguousarray(ic[:, 0:4]).view(np.uint64)
        _, first, inverse = np.unique(keys, axis=0, return_index=True, return_inverse=True)
        n_unique = len(first)
        chunk, counts = dedup_layout(n_unique, world)
        slot_of_pixel = dedup_slot(inverse.reshape(-1), world, chunk)
        # integrate: this rank's slots = the distinct rays u = rank (mod world)
        mine = np.arange(rank, n_unique, world)
        assert len(mine) == counts[rank]
        fq, ns = O.integrate_batch(q0s[first[mine]], ic[first[mine], 0:4], sc.steps, sc.delta, 2.0, sc.r_max, 1.0, nthreads=1)
        share_q = np.zeros((chunk, 4)); share_q[:len(mine)] = fq
        share_n = np.zeros((chunk,), np.int64); share_n[:len(mine)] = ns
        fq_all = gather_ray_slots(share_q, world)
        ns_all = gather_ray_slots(share_n, world)
        # epilogue of this rank's rows from the gathered final states, then the frame gather
        rb, re, rs = shard_rows(h, rank, world)
        rows = np.arange(rb, re, rs)
        pix = (rows[:, None] * w + np.arange(w)[None, :]).reshape(-1)
        sl = slot_of_pixel[pix]
        loc = O.classify(sc, fq_all[sl], ic[pix, 4], ic[pix, 5], tex)
        packed = np.concatenate([loc["rgb"].reshape(len(rows), 3 * w), loc["cls"].reshape(len(rows), w)], axis=1)
        full = gather_row_blocks(packed, h, world, rank)
        ref = O.render(sc, tex, nthreads=1)
        ok = (np.array_equal(full[:, :3 * w].reshape(h, w, 3), ref["rgb"]) and np.array_equal(full[:, 3 * w:], ref["cls"])
              and np.array_equal(fq_all[slot_of_pixel], ref["final_q"], equal_nan=True)
              and np.array_equal(ns_all[slot_of_pixel], ref["nsteps"]) and n_unique < h * w // 2)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w", [(2, 12, 18), (3, 9, 14)])
def test_shared_dedup_over_gloo(world, h, w):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dedup_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    got = sorted(q.get(timeout=5) for _ in range(world))
    assert got == [(r, True) for r in range(world)]
