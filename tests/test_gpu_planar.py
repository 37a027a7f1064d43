"""The equatorial ("planar") fast path of the step loop (csrc/schw_device.cuh, "Planar steps") must be an exact
specialisation: every output of every entry point is bit-identical with SCHW_PLANAR=0, the path is actually taken on
camera rays, and it is left (not mis-applied) on rays with polar motion, odd parameters and NaNs."""
import numpy as np
import pytest

from helpers import random_rays, synthetic_sky

pytestmark = pytest.mark.gpu


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view(np.uint8)


def _render(handle, monkeypatch, planar, sc, tex, **kw):
    monkeypatch.setenv("SCHW_PLANAR", "1" if planar else "0")
    out = handle.render(sc, tex, extras=True, **kw)
    out["counters"] = handle.last_counters()
    return out


@pytest.mark.parametrize("kw", [
    dict(height=270, width=480),                                     # default scene
    dict(height=135, width=135),                                     # odd size: centre pixel is exactly radial (p_phi == 0)
    dict(height=120, width=200, delta=0.005),                        # the sweep's finer step
    dict(height=96, width=160, delta=0.05, r_max=60.0, obs=(45.0, 0.0, 0.0)),
    dict(height=90, width=120, mass=0.5, obs=(20.0, 0.0, 0.0)),      # rs = 1 (metric derivative keeps the literal 2.0)
    dict(height=64, width=96, omega=0.01, delta=0.03, steps=4000),   # weak coupling, step cap reached by some rays
    dict(height=48, width=64, obs=(30.0, 0.0, 4.0)),                 # observer off the equatorial plane: general path only
], ids=["default", "odd", "fine", "far", "halfmass", "weak", "offplane"])
def test_render_is_bit_identical_with_and_without_the_fast_path(handle, monkeypatch, kw):
    import schw_b200
    sc = schw_b200.make_scene(**kw)
    tex = synthetic_sky(sc.height, sc.width, seed=2)
    a = _render(handle, monkeypatch, True, sc, tex)
    b = _render(handle, monkeypatch, False, sc, tex)
    for key in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"):
        assert np.array_equal(_bits(a[key]), _bits(b[key])), key
    steps_a, iters_a, planar_a = a["counters"]
    steps_b, iters_b, planar_b = b["counters"]
    assert steps_a == steps_b == int(a["nsteps"].astype(np.int64).sum())
    assert planar_b == 0
    if kw.get("obs", (30.0, 0.0, 0.0))[2] != 0.0:
        assert planar_a == 0                                          # theta != th_eq: never entered
    elif kw.get("delta", 0.01) <= 0.01:
        assert planar_a > 0.9 * iters_a, (planar_a, iters_a)          # camera rays live on the fast path
    else:
        # a coarse step leaves less room below half an ulp of theta: the bound T on |p_theta| shrinks (R1, R3) and rays
        # are handed to the general flow once cot(th_eq) has fed their p_theta past it
        assert 0 < planar_a <= iters_a
    assert steps_a <= 32 * iters_a


def test_integrate_batch_is_bit_identical_with_and_without_the_fast_path(handle, monkeypatch):
    # camera-like equatorial rays first, then tilted rays, in one queue: warps start on the fast path, meet the odd rays
    # below and the tilted ones as their lanes refill, and switch paths lane group by lane group
    qa, pa = random_rays(6000, seed=21)
    eq = qa[:, 2] == np.pi / 2
    order = np.concatenate([np.flatnonzero(eq), np.flatnonzero(~eq)])
    q0s, p0s = qa[order].copy(), pa[order].copy()
    # some rays that sit on theta == th_eq but violate the other conditions of the fast path
    m = 1500
    q0s[m:m + 64, 2] = np.pi / 2
    p0s[m:m + 16, 2] = 1e-13            # |p_theta| above the bound T
    p0s[m + 16:m + 32, 3] = 1e-9        # |p_phi| below 2^28 T
    p0s[m + 32:m + 48, 2] = -0.0
    p0s[m + 48:m + 56, 3] = 0.0         # exactly radial
    q0s[m + 56:m + 60, 1] = np.nan
    p0s[m + 60:m + 64, 2] = np.nan
    res = {}
    for planar in (True, False):
        # "2": the fast-path kernel even though the host-side sample of this batch is not (nearly) all planar
        monkeypatch.setenv("SCHW_PLANAR", "2" if planar else "0")
        out, ns = handle.integrate_batch(q0s, p0s, 3000, 0.02, 2.0, 45.0, 1.0, want_steps=True)
        res[planar] = (out, ns, handle.last_counters())
    assert np.array_equal(_bits(res[True][0]), _bits(res[False][0]))
    assert np.array_equal(res[True][1], res[False][1])
    assert res[False][2][2] == 0 and res[True][2][0] == res[False][2][0]
    assert 0 < res[True][2][2] < res[True][2][1]          # some warps on the fast path, some off it
    # default: the host looks at the batch and picks the plain kernel for it
    monkeypatch.setenv("SCHW_PLANAR", "1")
    out = handle.integrate_batch(q0s, p0s, 3000, 0.02, 2.0, 45.0, 1.0)
    assert handle.last_counters()[2] == 0 and np.array_equal(_bits(out), _bits(res[False][0]))
    # all-equatorial batch: fast path taken; all-tilted batch: never
    qe, pe = random_rays(2048, seed=5, equatorial_frac=1.0)
    handle.integrate_batch(qe, pe, 12000, 0.005, 2.0, 45.0, 1.0)
    _, iters, planar_iters = handle.last_counters()
    # (rays with a large angular momentum L collect p_theta = cot(th_eq) L dphi beyond the bound T and leave the path)
    assert planar_iters > 0.5 * iters, (planar_iters, iters)
    qt, pt = random_rays(2048, seed=6, equatorial_frac=0.0)
    monkeypatch.setenv("SCHW_PLANAR", "2")
    handle.integrate_batch(qt, pt, 3000, 0.02, 2.0, 45.0, 1.0)
    assert handle.last_counters()[2] == 0


def test_fast_path_constants(handle):
    """What the fast path hard-wires about theta = th_eq: 1/sin == 1.0 exactly, cot == cos(th_eq) as glibc rounds it."""
    r, s, c = handle.math_probe(np.array([1.0, np.pi / 2]))
    assert r[0] == 1.0                                   # fast_rcp(1.0)
    assert s[1] == 1.0 and c[1] == 6.123233995736766e-17
    assert np.float64(np.pi / 2).view(np.int64) == 0x3FF921FB54442D18


@pytest.mark.parametrize("kw,rows", [
    (dict(height=270, width=480), (0, None, 1)),                                  # default scene, width a multiple of 32
    (dict(height=135, width=135), (0, None, 1)),                                  # ragged tiles at the row ends
    (dict(height=77, width=45, delta=0.05), (1, 77, 3)),                          # a strided shard, rows shorter than two tiles
    (dict(height=64, width=96, omega=0.01, delta=0.03, steps=4000), (0, None, 1)),  # step cap ends some rays
    (dict(height=40, width=70, delta=0.05, steps=0), (0, None, 1)),               # every ray ends before its first step
    (dict(height=40, width=70, delta=0.05, r_max=25.0), (0, None, 1)),            # observer outside r_max: ends at once
    (dict(height=33, width=65, delta=0.05, has_texture=False), (0, None, 1)),
], ids=["default", "ragged", "shard", "cap", "steps0", "outside", "notex"])
def test_all_render_kernels_are_bit_identical(handle, monkeypatch, kw, rows):
    """The three schedulers of the render -- per-lane refill (persistent_rays), whole tiles in the same kernel, and the
    whole-tile kernel with one or two rays per thread (persistent_tiles<R>, the shipped path) -- integrate the same rays
    with the same arithmetic: every output array agrees bit for bit, so do the step totals."""
    import schw_b200
    sc = schw_b200.make_scene(**kw)
    tex = synthetic_sky(sc.height, sc.width, seed=6) if kw.get("has_texture", True) else None
    rb, re, rs = rows
    outs = {}
    for mode, rays in (("0", "1"), ("1", "1"), ("2", "1"), ("2", "2")):
        monkeypatch.setenv("SCHW_TILE_MODE", mode)
        monkeypatch.setenv("SCHW_TILE_RAYS", rays)
        out = handle.render(sc, tex, rb, re, rs, extras=True)
        out["counters"] = handle.last_counters()
        outs[(mode, rays)] = out
    ref = outs[("0", "1")]
    assert ref["counters"][0] == int(ref["nsteps"].astype(np.int64).sum())
    for key, out in outs.items():
        for name in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"):
            assert np.array_equal(_bits(ref[name]), _bits(out[name])), (key, name)
        assert out["counters"][0] == ref["counters"][0], key                     # steps integrated
        assert out["counters"][0] <= 32 * out["counters"][1] or out["counters"][0] == 0
