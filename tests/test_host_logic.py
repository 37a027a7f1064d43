"""CPU tests of the host-side mirror of the reference interface (no GPU calls)."""
import numpy as np
import pytest

from conftest import relerr


REFERENCE_DEFAULTS = dict(                      # config.py:6-32 of the reference (types matter: 31 and 30 are ints)
    size=200, fov=80.0, background='images/backgrounds/milky-way-equirec.jpg', steps=200000, delta=0.01, omega=0.01,
    rtol=1e-2, atol=1e-2, order=2, suppress_warnings=False, cuda=True, bh_mass=1.0, boundary_radius=31,
    observer_distance=30, bg_patch_center_theta=90, bg_patch_center_phi=180, bg_patch_center_theta_relobs=0,
    bg_patch_center_phi_relobs=0, bg_patch_size_theta=180, bg_patch_size_phi=360, bg_flip_theta=False,
    bg_flip_phi=False, no_flat_trajectories=False)


def test_config_flags_and_defaults_match_reference():
    import config
    a = vars(config.parse_args([]))
    for k, v in REFERENCE_DEFAULTS.items():
        assert a[k] == v and type(a[k]) is type(v), k
    b = config.parse_args(["--size", "400", "--bg-flip-phi", "--boundary-radius", "21", "--order", "4"])
    assert b.size == 400 and b.bg_flip_phi and b.boundary_radius == 21.0 and b.order == 4
    with pytest.raises(SystemExit):
        config.parse_args(["--order", "3"])
    # additive flags default to the reference's behaviour (square frame, one integration per pixel)
    assert a["width"] is None and a["height"] is None and a["dedup"] is False
    assert config.parse_args(["--dedup"]).dedup is True


def test_scene_classes():
    from simulation.blackhole import BlackHole, Observer
    bh = BlackHole(mass=1.5)
    assert bh.rs == 3.0 and bh.position.shape == (3,)
    ob = Observer(position=[30, 0, 0], fov=1.0, image_size=(4, 6))
    assert ob.image_size == (4, 6) and ob.position.dtype.kind == "i"


def test_null_momentum_helpers_vs_reference_golden(golden):
    """angles_to_p_sph + build_null_4momentum_ep_sph reproduce the reference's p0 from its own alpha0."""
    from simulation.utils import angles_to_p_sph, build_null_4momentum_ep_sph
    g = golden("ic_golden.npz")["table"]
    for row in g[::37]:
        obs_x, mass = row[0], row[4]
        p0_ref, alpha0 = row[11:15], row[15]
        p_sp = angles_to_p_sph(alpha0, 0.0, obs_x, mass_bh=mass)
        p0 = build_null_4momentum_ep_sph(p_sp, np.array([obs_x, np.pi / 2, 0.0]), mass_bh=mass, future=True)
        # alpha0 = acos(cos alpha) loses ~1e-16/sin(alpha) near the image centre, times r_obs in p_phi
        assert relerr(p0, p0_ref, 1.0).max() < 1e-11
        assert np.signbit(p0[2])
        pm = build_null_4momentum_ep_sph(p_sp, np.array([obs_x, np.pi / 2, 0.0]), mass_bh=mass, future=False)
        assert pm[0] == -p0[0]
    with pytest.raises(ValueError):
        build_null_4momentum_ep_sph([1, 0, 0], [1.9, np.pi / 2, 0.0])


def test_apply_relative_offsets():
    from simulation.utils import _apply_relative_offsets
    th, ph = _apply_relative_offsets(90, 180, 0, 0)
    assert th == np.pi / 2 and ph == np.pi
    th, ph = _apply_relative_offsets(90, 180, 100, 200)          # clip to [0, pi], wrap to [0, 2 pi)
    assert th == np.pi and abs(ph - np.deg2rad(20)) < 1e-15
    th, ph = _apply_relative_offsets(10, 10, -30, -25)
    assert th == 0.0 and abs(ph - np.deg2rad(345)) < 1e-15


def test_row_sharding_helpers():
    from schw_b200.multigpu import rows_of, shard_rows, unpermute_rows
    for h in (1, 7, 16, 2160):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                rb, re, rs = shard_rows(h, r, world)
                rows = list(range(rb, re, rs))
                assert len(rows) == rows_of(h, r, world)
                seen += rows
            assert sorted(seen) == list(range(h))
            full = np.arange(h * 3).reshape(h, 3)
            max_rows = (h + world - 1) // world
            blocks = []
            for r in range(world):
                b = np.zeros((max_rows, 3), full.dtype)
                b[:rows_of(h, r, world)] = full[r::world]
                blocks.append(b)
            assert np.array_equal(unpermute_rows(blocks, h), full)


def test_run_manual_simulation_quirks_without_gpu(tmp_path, monkeypatch):
    """Reference quirks reproduced on the host side: n_samples=0 returns an all-black frame (App. D-3);
    the CPU branch does not exist (App. D-2)."""
    from simulation.blackhole import BlackHole, Observer
    from simulation.raytracing import run_manual_simulation
    monkeypatch.chdir(tmp_path)
    bh = BlackHole(1.0); ob = Observer(np.array([30, 0, 0]), np.radians(80), (6, 8))
    img = run_manual_simulation(bh, ob, steps=10, delta=0.1, use_cuda=True, boundary_radius=31, n_samples=0)
    assert img.shape == (6, 8, 3) and img.dtype == np.uint8 and not img.any()
    # without use_cuda the reference integrates nothing: NameError at raytracing.py:131 when trajectories are sampled,
    # the black frame + "Summary counts not available" otherwise (raytracing.py:301-314)
    with pytest.raises(NameError):
        run_manual_simulation(bh, ob, steps=10, delta=0.1, use_cuda=False, boundary_radius=31, n_samples=2)
    img = run_manual_simulation(bh, ob, steps=10, delta=0.1, use_cuda=False, n_samples=0)
    assert img.shape == (6, 8, 3) and not img.any()


def test_dedup_flag_reaches_the_render_call(tmp_path, monkeypatch):
    """`main.py --dedup` -> run_manual_simulation(dedup=True) -> render_frame(dedup=True) -> Handle.render(dedup=True)
    (plumbing only: the library call is intercepted, there is no GPU here)."""
    import schw_b200
    from schw_b200 import multigpu
    from simulation.blackhole import BlackHole, Observer
    from simulation.raytracing import run_manual_simulation
    monkeypatch.chdir(tmp_path)
    seen = {}

    class Stop(Exception):
        pass

    class FakeHandle:
        def initial_conditions_px(self, obs, pixels, mass):
            return np.array([[0.0, 30.0, np.pi / 2, 0.0]]), np.zeros((1, 9))

        def render(self, scene, tex, rb, re, rs, extras=False, dedup=False):
            seen.update(extras=extras, dedup=dedup, rows=(rb, re, rs), shape=(scene.height, scene.width))
            raise Stop()

    monkeypatch.setattr(schw_b200.native, "default_handle", lambda device=-1: FakeHandle())
    bh = BlackHole(1.0); ob = Observer(np.array([30, 0, 0]), np.radians(80), (6, 8))
    for flag in (True, False):
        with pytest.raises(Stop):
            run_manual_simulation(bh, ob, steps=10, delta=0.1, use_cuda=True, boundary_radius=31, n_samples=2, dedup=flag)
        assert seen == dict(extras=True, dedup=flag, rows=(0, 6, 1), shape=(6, 8))
    assert multigpu.render_frame.__defaults__[-1] is False          # opt-in


def test_pinned_frame_buffers_are_plain_arrays_and_recycled(monkeypatch):
    """render() hands out page-locked frame buffers (schw_host_alloc) as plain ndarrays; a block returns to the pool
    when the last view of it dies.  The allocator is replaced by malloc here (no GPU in this container)."""
    import ctypes
    import gc
    from schw_b200 import native
    libc = ctypes.CDLL(None)
    libc.malloc.restype = ctypes.c_void_p; libc.malloc.argtypes = [ctypes.c_size_t]; libc.free.argtypes = [ctypes.c_void_p]
    allocs, frees = [], []
    monkeypatch.setattr(native, "_host_alloc", lambda n: (allocs.append(n), libc.malloc(n))[1])
    monkeypatch.setattr(native, "_host_free", lambda a: (frees.append(a), libc.free(a)))
    monkeypatch.setattr(native, "_PINNED_POOL", {})
    a = native.pinned_empty((4, 5, 3), np.uint8)
    assert type(a) is np.ndarray and a.flags.writeable and a.flags.c_contiguous and a.shape == (4, 5, 3)
    a[:] = 7
    v = a[1:3]
    del a; gc.collect()
    assert native._PINNED_POOL.get(60, []) == [] and v.sum() == 7 * 30        # a view keeps the block alive
    del v; gc.collect()
    assert len(native._PINNED_POOL[60]) == 1
    b = native.pinned_empty((60,), np.uint8)
    assert allocs == [60] and native._PINNED_POOL[60] == []                    # recycled, not re-allocated
    many = [native.pinned_empty((60,), np.uint8) for _ in range(7)]
    del many, b; gc.collect()
    assert len(native._PINNED_POOL[60]) == native._PINNED_POOL_MAX and len(frees) == 8 - native._PINNED_POOL_MAX
    assert native.pinned_empty((0, 3), np.uint8).shape == (0, 3)
    for addr in native._PINNED_POOL.pop(60):
        libc.free(addr)


def test_camera_rays_repeat_bitwise_premise_of_the_dedup_render(orc):
    """Premise of the opt-in de-duplicated render (schw_render_dedup), checked with the oracle's restatement
    of get_initial_conditions: the 4-momentum of a camera ray is a function of its camera angle alone and the pixel grid
    is mirror symmetric, so a frame holds several times fewer DISTINCT (q0, p0) than pixels -- bit for bit."""
    for (h, w), min_ratio in (((400, 400), 10.0), ((108, 192), 6.5), ((135, 135), 8.0)):
        sc = orc.make_scene(height=h, width=w)
        ii, jj = np.mgrid[0:h, 0:w]
        px = orc.camera_pixels(sc, np.stack([ii.ravel(), jj.ravel()], 1).astype(np.int32))
        q0s, ic, rc = orc.initial_conditions(np.array([30.0, 0.0, 0.0]), px, 1.0)
        assert rc == 0 and len(np.unique(q0s.view(np.int64), axis=0)) == 1
        p0 = np.ascontiguousarray(ic[:, 0:4])
        distinct = len(np.unique(p0.view(np.int64), axis=0))
        assert h * w / distinct >= min_ratio, (h, w, distinct)
        # most mirror images of a pixel carry the same momentum bit for bit -- not all: (j + 0.5) / w - 0.5 does not round
        # symmetrically -- which is why the render keys on the bits themselves instead of assuming the symmetry
        P = p0.reshape(h, w, 4).view(np.int64)
        assert 0.9 < (P == P[::-1]).all(-1).mean() < 1.0 and 0.9 < (P == P[:, ::-1]).all(-1).mean() < 1.0


def test_dedup_is_exact_in_the_references_own_arithmetic(orc):
    """The de-duplicated render as an algorithm, run entirely with the CPU restatement of the reference: integrate each
    DISTINCT (q0, p0) once, give every pixel the final state of its geodesic and its own beta, classify -- and obtain
    the frame the one-ray-per-pixel loop of run_manual_simulation produces, bit for bit (class, rgb, texel, final state).
    The GPU pipeline (schw_render_dedup) has the same structure; tests/test_gpu_dedup.py checks it against the GPU render."""
    from helpers import synthetic_sky
    h, w = 60, 96
    sc = orc.make_scene(height=h, width=w, delta=0.05, patch_center_phi=2.5, patch_size_phi=4.0, flip_phi=True)
    tex = synthetic_sky(h, w, seed=3)
    ref = orc.render(sc, tex, nthreads=2)
    p0 = np.ascontiguousarray(ref["ic"][:, 0:4])
    uniq, inv = np.unique(p0.view(np.int64), axis=0, return_inverse=True)
    inv = inv.reshape(-1)
    assert len(uniq) < h * w / 3
    q0 = np.tile([0.0, 30.0, np.pi / 2, 0.0], (len(uniq), 1))
    fq_u, ns_u = orc.integrate_batch(q0, uniq.view(np.float64), sc.steps, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0)
    fq = fq_u[inv]
    got = orc.classify(sc, fq, ref["ic"][:, 4], ref["ic"][:, 5], tex)
    assert np.array_equal(fq.view(np.int64), ref["final_q"].view(np.int64)) and np.array_equal(ns_u[inv], ref["nsteps"])
    assert np.array_equal(got["cls"], ref["cls"].reshape(-1)) and np.array_equal(got["rgb"], ref["rgb"].reshape(-1, 3))
    assert np.array_equal(got["uv"], ref["uv"]) and np.array_equal(got["hit"].view(np.int64), ref["hit"].view(np.int64))
