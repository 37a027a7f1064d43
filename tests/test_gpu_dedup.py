"""Opt-in geodesic de-duplication (schw_render_dedup): pixels whose camera angle agrees bitwise share one
integration.  It must reproduce the one-ray-per-pixel render in every output bit, while integrating several times
fewer rays."""
import numpy as np
import pytest

from helpers import synthetic_sky

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("kw,rows,max_share", [
    (dict(height=270, width=480), (0, None, 1), 0.5),
    (dict(height=135, width=135), (0, None, 1), 0.5),                       # odd: exactly radial centre ray
    (dict(height=200, width=200, delta=0.05, steps=300), (0, None, 1), 0.5),  # step cap -> in_domain
    (dict(height=96, width=160, patch_center_phi=2.0, patch_size_phi=1.5, patch_size_theta=1.0, flip_phi=True), (0, None, 1), 0.5),
    # one rank's row-cyclic shard: the vertical mirror image of its rows belongs to another rank, the horizontal one stays
    (dict(height=216, width=384), (1, 216, 4), 0.75),
], ids=["default", "odd", "cap", "patch", "shard"])
def test_dedup_render_is_bit_identical(handle, kw, rows, max_share):
    import schw_b200
    sc = schw_b200.make_scene(**kw)
    tex = synthetic_sky(sc.height, sc.width, seed=4)
    rb, re, rs = rows
    a = handle.render(sc, tex, rb, re, rs, extras=True)
    steps_full = handle.last_total_steps()
    b = handle.render(sc, tex, rb, re, rs, extras=True, dedup=True)
    steps_dedup = handle.last_total_steps()
    unique, pixel_steps = handle.last_dedup_stats()
    for key in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"):
        assert np.array_equal(_bits(a[key]), _bits(b[key])), key
    n = a["cls"].size
    assert pixel_steps == steps_full == int(a["nsteps"].astype(np.int64).sum())
    assert 0 < unique < max_share * n and steps_dedup < max_share * steps_full, (unique, n, steps_dedup, steps_full)
    # without the diagnostics too
    c = handle.render(sc, tex, rb, re, rs, dedup=True)
    assert np.array_equal(c["rgb"], a["rgb"]) and np.array_equal(c["cls"], a["cls"])


def test_dedup_reuses_the_geodesics_when_only_the_sky_changes(handle, monkeypatch):
    """An azimuth sweep changes the sky patch, not the camera: the geodesics of the previous frame are reused
    ("integrate once, N lookups") and the frame is still the one-ray-per-pixel frame, bit for bit."""
    import schw_b200
    base = dict(height=180, width=320)
    tex = synthetic_sky(180, 320, seed=9)
    first = handle.render(schw_b200.make_scene(**base), tex, extras=True, dedup=True)
    steps_first = handle.last_total_steps()
    assert steps_first > 0
    for k, extra in enumerate((dict(patch_center_phi=np.deg2rad(217.0)),
                               dict(patch_center_phi=1.0, patch_size_phi=2.0, patch_size_theta=1.2, flip_theta=True, flip_phi=True))):
        sc = schw_b200.make_scene(**base, **extra)
        tex_k = synthetic_sky(180, 320, seed=10 + k)
        got = handle.render(sc, tex_k, extras=True, dedup=True)
        assert handle.last_total_steps() == 0                          # nothing integrated
        uniq, pixel_steps = handle.last_dedup_stats()
        want = handle.render(sc, tex_k, extras=True)
        assert pixel_steps == handle.last_total_steps() and uniq > 0
        for key in ("rgb", "cls", "final_q", "ic", "hit", "uv", "nsteps"):
            assert np.array_equal(_bits(got[key]), _bits(want[key])), (k, key)
        assert np.array_equal(got["final_q"], first["final_q"], equal_nan=True)
    # a different step size is a different frame: integrated again
    sc = schw_b200.make_scene(**base, delta=0.02)
    got = handle.render(sc, tex, dedup=True)
    assert handle.last_total_steps() > 0
    assert np.array_equal(got["rgb"], handle.render(sc, tex)["rgb"])
    # and the switch
    monkeypatch.setenv("SCHW_DEDUP_REUSE", "0")
    handle.render(sc, tex, dedup=True)
    assert handle.last_total_steps() > 0
