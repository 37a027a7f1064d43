"""Pins the CPU oracle (oracle/schw_oracle.c) against outputs of the UNMODIFIED reference.

Golden fixtures come from tests/golden/make_golden.py (reference modules imported with
NUMBA_ENABLE_CUDASIM=1) and from the two genuine EinsteinPy trajectories the reference ships.
CPU-only: runs in the build container and on the GPU box alike.
"""
import numpy as np

from conftest import relerr
from helpers import scene_kwargs_from_cfg


def _null_pt(q0, p_sph, mass=1.0, future=False):
    r, th = q0[1], q0[2]
    f = 1.0 - 2.0 * mass / r
    A = -1.0 / f
    C = f * p_sph[0] * p_sph[0] + (1.0 / (r * r)) * p_sph[1] * p_sph[1] + (1.0 / (r * r * np.sin(th) ** 2)) * p_sph[2] * p_sph[2]
    pt = np.sqrt(-4.0 * A * C) / (2.0 * (-A))
    return pt if future else -pt


def test_stepper_default_scene_bit_exact(orc, golden):
    """cuda_geodesic.py:579-602 on 24 default-scene camera rays (incl. horizon blow-up rays): bit-for-bit."""
    g = golden("stepper_default.npz")
    out, ns = orc.integrate_batch(g["q0s"], g["p0s"], int(g["steps"]), float(g["delta"]), 2.0, float(g["r_max"]), float(g["omega"]))
    assert np.array_equal(out, g["out_qs"])
    # step counts quoted in SURVEY App. E for the first twelve pixels
    assert list(ns[:12]) == [3085, 3144, 3303, 3773, 7469, 7238, 7112, 6968, 6665, 6125, 5139, 4271]


def test_stepper_general_rays_bit_exact(orc, golden):
    g = golden("stepper_general.npz")
    for tag, omega in (("out_w1", 1.0), ("out_w001", 0.01)):
        out, ns = orc.integrate_batch(g["q0s"], g["p0s"], int(g["steps"]), float(g["delta"]), 2.0, float(g["r_max"]), omega)
        assert np.array_equal(out, g[tag]), tag
        assert ns.min() > 100 and ns.max() < int(g["steps"])


def test_single_ray_k3_bit_exact(orc, golden):
    """single_ray_cuda_test.py:250-321 through integrate_batch_full: 6665 stored rows, bit-for-bit."""
    k = golden("single_ray_k3.npz")
    traj, rows = orc.integrate_batch_full(k["q0"][None], k["p0"][None], int(k["steps"]), float(k["delta"]), 2.0,
                                          float(k["r_max"]), float(k["omega"]))
    assert rows[0] == 6665 == k["traj"].shape[0]
    assert np.array_equal(traj[0, :6665], k["traj"])
    assert not traj[0, 6665:].any()                       # zero rows after termination (np.zeros, :649)
    assert k["traj"][-1, 1] <= 2.2 and k["traj"][-2, 1] > 2.2


def test_kat_k1_einsteinpy_trajectory(orc, golden):
    """tests/einsteinpy-geodesic-test/trajectory.csv (genuine EinsteinPy FANTASY output), rows 1..2868."""
    k = golden("kat_k1_einsteinpy_trajectory.npz")
    q0 = k["q0"]; p0 = np.array([_null_pt(q0, k["p_sph"]), *k["p_sph"]])
    traj, mom, rows = orc.integrate_batch_full(q0[None], p0[None], 2900, float(k["delta"]), 2.0, 1e6, float(k["omega"]),
                                               with_momenta=True)
    T, M = traj[0][k["step"]], mom[0][k["step"]]
    x = T[:, 1] * np.cos(T[:, 3]) * np.sin(T[:, 2]); y = T[:, 1] * np.sin(T[:, 3]) * np.sin(T[:, 2]); z = T[:, 1] * np.cos(T[:, 2])
    mine = np.stack([T[:, 0], x, y, z, M[:, 0], M[:, 1], M[:, 2], M[:, 3]], 1)
    assert relerr(mine, k["vals"]).max() < 5e-14          # SURVEY §4 quotes 1.8e-14


def test_kat_k2_einsteinpy_notebook(orc, golden):
    """x/y/z arrays embedded in einsteinpy_ray.ipynb cell 9 (10 000 steps, delta=0.001)."""
    k = golden("kat_k2_einsteinpy_notebook.npz")
    q0 = k["q0"]; p0 = np.array([_null_pt(q0, k["p_sph"]), *k["p_sph"]])
    traj, rows = orc.integrate_batch_full(q0[None], p0[None], 10001, float(k["delta"]), 2.0, 1e6, float(k["omega"]))
    T = traj[0][k["step"]]
    xyz = np.stack([T[:, 1] * np.cos(T[:, 3]) * np.sin(T[:, 2]), T[:, 1] * np.sin(T[:, 3]) * np.sin(T[:, 2]), T[:, 1] * np.cos(T[:, 2])], 1)
    assert relerr(xyz, k["xyz"]).max() < 5e-15            # SURVEY §4 quotes 4.1e-16


def test_initial_conditions_vs_reference(orc, golden):
    """get_initial_conditions (utils.py:91-141) on 1681 reference camera pixels.
    Tolerance 1e-14 relative: NumPy's SVML atan2/acos/tan/hypot and OpenBLAS dot differ from glibc in the last bit."""
    g = golden("ic_golden.npz")["table"]
    for key in np.unique(g[:, :5], axis=0):
        sel = np.all(g[:, :5] == key, axis=1)
        obs_x, fov, h, w, mass = key
        sc = orc.make_scene(obs=(obs_x, 0, 0), fov=fov, height=int(h), width=int(w), mass=mass)
        ij = g[sel][:, 5:7].astype(np.int32)
        px = orc.camera_pixels(sc, ij)
        assert relerr(px, g[sel][:, 20:23], floor=1.0).max() < 1e-15
        q0s, ic, rc = orc.initial_conditions((obs_x, 0, 0), px, mass)
        assert rc == 0
        assert np.array_equal(q0s, g[sel][:, 7:11])
        ref = np.concatenate([g[sel][:, 11:15], g[sel][:, 15:16], g[sel][:, 19:20], g[sel][:, 16:19]], axis=1)
        assert relerr(ic, ref, floor=1.0).max() < 1e-14
        assert np.all(np.signbit(ic[:, 2]))               # p_theta = -0.0 (utils.py:172,184)


def _check_render(orc, golden, name):
    from PIL import Image
    import os
    from helpers import GOLDEN, near_rounding_boundary, texel_coords
    g = golden(f"render_{name}.npz")
    kw, c = scene_kwargs_from_cfg(g["cfg"], g["patch_center"])
    tex = np.array(Image.open(os.path.join(GOLDEN, "texture_card.png")).convert("RGB").resize((c["w"], c["h"]), Image.LANCZOS))
    assert np.array_equal(tex, g["bg_array"])             # same PIL resize as when the golden was made
    out = orc.render(orc.make_scene(**kw), tex)
    assert out["rc"] == 0
    cls = out["cls"]
    assert np.array_equal(cls, g["cls"]), f"class mask differs at {np.argwhere(cls != g['cls'])[:5]}"
    # images: identical except where the continuous texel coordinate sits on a rounding boundary
    diff = np.any(out["rgb"] != g["img"], axis=2).reshape(-1)
    if diff.any():
        fu, fv = texel_coords(kw, g["final_th"], g["final_ph"])
        assert np.all(near_rounding_boundary(fu, fv)[diff])
    # final states: bit-exact stepper, ulp-level ICs => 1e-9 outside the bh mask (band H is inside it)
    esc = (g["cls"].reshape(-1) != 0)
    assert relerr(out["final_q"][esc, 1], g["final_r"][esc]).max() < 1e-9
    assert np.abs(out["hit"][esc] - np.stack([g["final_th"], g["final_ph"]], 1)[esc]).max() < 1e-9
    assert relerr(out["ic"][:, 4], g["alpha0"], floor=1.0).max() < 1e-14
    return g, out


def test_render_R1_default_patch(orc, golden):
    g, out = _check_render(orc, golden, "R1")
    assert set(np.unique(g["cls"])) == {0, 2}


def test_render_R2_flips_small_patch_offsets(orc, golden):
    g, out = _check_render(orc, golden, "R2")
    assert {0, 2, 3} <= set(np.unique(g["cls"]))          # bh, escape_bg and escape_no_patch all occur


def test_render_R3_numerical_error_class(orc, golden):
    g, out = _check_render(orc, golden, "R3")
    assert 1 in np.unique(g["cls"])
    assert np.all(g["img"][g["cls"] == 1] == (255, 0, 0))


def test_render_R4_step_cap_in_domain(orc, golden):
    g, out = _check_render(orc, golden, "R4")
    assert 4 in np.unique(g["cls"])
    assert (out["nsteps"] == 120).all()


def test_render_R5_second_scene(orc, golden):
    _check_render(orc, golden, "R5")


def test_k3_conditioning(orc, golden):
    """Documents why config 2 cannot be matched at 1e-9 over its whole length by ANY re-implementation (including
    the reference's own FMA-contracting GPU build): the omega=0.01 two-copy trajectory amplifies a 1-ulp change of
    p0 exponentially.  The GPU test's per-window tolerances are derived from these numbers."""
    k = golden("single_ray_k3.npz")
    p = k["p0"].copy(); p2 = p * (1 + 2.2e-16)
    a, ra = orc.integrate_batch_full(k["q0"][None], p[None], 7000, float(k["delta"]), 2.0, float(k["r_max"]), float(k["omega"]))
    b, rb = orc.integrate_batch_full(k["q0"][None], p2[None], 7000, float(k["delta"]), 2.0, float(k["r_max"]), float(k["omega"]))
    n = int(min(ra[0], rb[0]))
    d = (np.abs(a[0, :n] - b[0, :n]) / np.maximum(np.abs(a[0, :n]), 1.0)).max(1)
    assert d[:2500].max() < 1e-12
    assert 1e-11 < d[4000:5000].max() < 1e-7
    assert 1e-9 < d[6000:6600].max() < 1e-5


def test_k3_reference_error_vs_truth(orc, golden):
    """The binary128 truth (oracle/schw_oracle_q.c) on config 2: how far the REFERENCE's own FP64 rows are from the
    exact result of the discrete map -- the measured backing of every tolerance above 1e-9 in the GPU tests."""
    k = golden("single_ray_k3.npz")
    n = k["traj"].shape[0]
    exact, rows = orc.exact_integrate_batch_full(k["q0"][None], k["p0"][None], n, float(k["delta"]), 2.0, float(k["r_max"]),
                                                 float(k["omega"]), stop=False)
    err = (np.abs(k["traj"] - exact[0]) / np.maximum(np.abs(exact[0]), 1.0)).max(1)
    assert err[:2000].max() < 1e-11 and err[:3000].max() < 2.5e-10               # the reference is 1e-9-accurate up to ~row 3000
    assert 1e-9 < err[4500] < 1e-7 and 1e-7 < err[6600] < 1e-4                   # and loses six more digits by the end
    # with the termination test the exact trajectory stops within a row or two of the reference's 6665 rows
    _, rows_stop = orc.exact_integrate_batch_full(k["q0"][None], k["p0"][None], n + 50, float(k["delta"]), 2.0, float(k["r_max"]),
                                                  float(k["omega"]))
    assert abs(int(rows_stop[0]) - n) <= 3


def test_truth_oracle_is_pinned_by_the_einsteinpy_kats(orc, golden):
    """The binary128 stepper is the same map as EinsteinPy's FANTASY order 2: on the stable K2 trajectory (10 000 steps) and
    on the first 1200 rows of K1 it agrees with the genuine EinsteinPy outputs to FP64 rounding (they are FP64 results)."""
    k2 = golden("kat_k2_einsteinpy_notebook.npz")
    q0 = k2["q0"]; p0 = np.array([_null_pt(q0, k2["p_sph"]), *k2["p_sph"]])
    T, _ = orc.exact_integrate_batch_full(q0[None], p0[None], 10001, float(k2["delta"]), 2.0, 1e6, float(k2["omega"]), stop=False)
    T = T[0][k2["step"]]
    xyz = np.stack([T[:, 1] * np.cos(T[:, 3]) * np.sin(T[:, 2]), T[:, 1] * np.sin(T[:, 3]) * np.sin(T[:, 2]), T[:, 1] * np.cos(T[:, 2])], 1)
    assert relerr(xyz, k2["xyz"], 1.0).max() < 1e-10           # measured 5e-12: FP64 rounding accumulated over 10 000 steps
    k = golden("kat_k1_einsteinpy_trajectory.npz")
    q0 = k["q0"]; p0 = np.array([_null_pt(q0, k["p_sph"]), *k["p_sph"]])
    T, M, _ = orc.exact_integrate_batch_full(q0[None], p0[None], 2900, float(k["delta"]), 2.0, 1e6, float(k["omega"]), stop=False,
                                             with_momenta=True)
    sel = k["step"] < 1200
    T, M = T[0][k["step"][sel]], M[0][k["step"][sel]]
    mine = np.stack([T[:, 0], T[:, 1] * np.cos(T[:, 3]) * np.sin(T[:, 2]), T[:, 1] * np.sin(T[:, 3]) * np.sin(T[:, 2]),
                     T[:, 1] * np.cos(T[:, 2]), M[:, 0], M[:, 1], M[:, 2], M[:, 3]], 1)
    assert relerr(mine, k["vals"][sel], 1.0).max() < 1e-10
    # and with the double oracle on well-conditioned rays to a few ulps times the step count
    from helpers import random_rays
    q0s, p0s = random_rays(64, seed=2)
    a, na = orc.integrate_batch(q0s, p0s, 300, 0.05, 2.0, 1e6, 1.0)
    b, nb = orc.exact_integrate_batch(q0s, p0s, 300, 0.05, 2.0, 1e6, 1.0, force_steps=na)
    assert np.median(relerr(a, b, 1.0).max(1)) < 1e-13


def test_flat_render_vs_reference(orc, golden):
    """N1: save_no_gravity_image_with_background (background.py:19-141).  F1 is the main.py default, which the
    reference renders all black (phi0 % 2pi == phi1 % 2pi == 0 for a 360-degree patch, background.py:8-17)."""
    from helpers import flat_scene_kwargs
    for name in ("F1", "F2", "F3"):
        g = golden(f"flat_{name}.npz")
        img = orc.flat_render(orc.make_scene(**flat_scene_kwargs(g["cfg"])), g["bg_array"])
        mism = np.any(img != g["img"], axis=2).mean()
        assert mism <= 2e-3, (name, mism)          # truncating texel map: a pixel may flip on an integer boundary
    assert not golden("flat_F1.npz")["img"].any()
