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
