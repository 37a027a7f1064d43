"""Two real ranks (one process per GPU, NCCL): `render_frame` — row-cyclic shards, one device-side
all_gather, un-permutation — must reproduce the single-GPU frame bit for bit, diagnostics included; so must the
opt-in de-duplicated render, in which the ranks share the integration of the distinct camera rays.
Skipped on a one-GPU box (NCCL cannot put two ranks on one device); CPU coverage of the same host
logic is tests/test_multigpu_gloo.py."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import PKG, ROOT

pytestmark = pytest.mark.gpu


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
    import torch
    import torch.distributed as dist
    import schw_b200
    from schw_b200.multigpu import render_frame
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rng = np.random.default_rng(5)
        tex = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        sc = schw_b200.make_scene(height=h, width=w, delta=0.05)
        hd = schw_b200.Handle(rank)
        got = render_frame(sc, tex, extras=True, handle=hd)
        ok = True
        if rank == 0:
            ref = hd.render(sc, tex, extras=True)
            ok = (np.array_equal(got["rgb"], ref["rgb"]) and np.array_equal(got["cls"], ref["cls"])
                  and np.array_equal(got["final_q"], ref["final_q"], equal_nan=True)
                  and np.array_equal(got["ic"], ref["ic"]) and np.array_equal(got["uv"], ref["uv"])
                  and np.array_equal(got["nsteps"], ref["nsteps"]) and got["rgb"].any())
        # root=0: a gather instead of the all_gather -- rank 0 gets the same frame, the others nothing
        got0 = render_frame(sc, tex, extras=True, handle=hd, root=0)
        if rank == 0:
            ok = ok and all(np.array_equal(got0[k], got[k], equal_nan=(k == "final_q"))
                            for k in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"))
        else:
            ok = ok and got0["rgb"] is None and got0["cls"] is None and got0["final_q"] is None
        # opt-in de-duplicated render: the ranks share the integration of the distinct rays -- every output array must
        # be the one-ray-per-pixel frame bit for bit (all_gather and root gather), with several times fewer steps integrated
        for root in (None, 0):
            dd = render_frame(sc, tex, extras=True, handle=hd, root=root, dedup=True)
            integrated = hd.last_total_steps()
            if root is None or rank == 0:
                ok = ok and all(np.array_equal(np.ascontiguousarray(dd[k]).view(np.uint8), np.ascontiguousarray(got[k]).view(np.uint8))
                                for k in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"))
            else:
                ok = ok and dd["rgb"] is None
            ok = ok and 0 < integrated < int(got["nsteps"].astype(np.int64).sum()) // (2 * world) if got["nsteps"] is not None else ok
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_render_frame_two_ranks_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (one rank per GPU)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    world, h, w = 2, 77, 130          # odd row count: the ranks get 39 and 38 rows
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(world)) == [(0, True), (1, True)]
