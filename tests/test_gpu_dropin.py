"""GPU tests of the drop-in mirror of the reference's Python interface (`simulation.*`, `main.py`):
same names, argument meaning, return values, side effects and error behaviour as the reference, with the
results checked against the reference's own outputs (tests/golden) and the oracle."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, relerr
from helpers import flat_mismatch_explained, flat_scene_kwargs, in_band_g, near_rounding_boundary, scene_kwargs_from_cfg, state_close, texel_coords

pytestmark = pytest.mark.gpu


def test_cuda_schwarzschild_integrator_api(golden):
    """`from simulation.cuda_geodesic import CUDASchwarzschildIntegrator` as single_ray_cuda_test.py:34,298-299 uses it."""
    from simulation.cuda_geodesic import CUDASchwarzschildIntegrator
    integ = CUDASchwarzschildIntegrator()                              # defaults of cuda_geodesic.py:632
    assert (integ.steps, integ.delta, integ.rs, integ.omega, integ.r_max) == (500, 0.2, 2.0, 1.0, 1e6)
    g = golden("stepper_default.npz")
    integ = CUDASchwarzschildIntegrator(steps=200000, delta=0.01, mass=1.0, r_max=31)
    out, none = integ.integrate_batch(g["q0s"], g["p0s"])
    assert none is None and out.shape == (24, 4) and out.dtype == np.float64
    bh_angle = np.arcsin(3 * np.sqrt(3) * 2.0 / 30.0) / 2
    esc = g["alpha0"] > bh_angle                                         # outside the analytic shadow mask (band H lives inside it)
    assert esc.sum() >= 15 and (g["out_qs"][esc, 1] >= 31.0).all()
    assert state_close(out[esc], g["out_qs"][esc]).all()
    k = golden("single_ray_k3.npz")
    integ = CUDASchwarzschildIntegrator(steps=200_000, delta=0.03, mass=1.0, omega=0.01, r_max=50)
    traj = integ.integrate_batch_full(k["q0"][np.newaxis, :], k["p0"][np.newaxis, :])[0]
    assert traj.shape == (200000, 4)
    safe = traj[:, 1] > 1.1 * 2.0                                        # the script's own truncation (:307-309)
    n = int(np.argmax(~safe))
    assert abs(n - 6664) <= 2
    assert state_close(traj[:2500], k["traj"][:2500]).all()


def test_get_initial_conditions_single_ray(golden):
    from simulation.utils import get_initial_conditions
    row = golden("ic_golden.npz")["table"][4]
    q0, p0, alpha0, h_r, h_theta, h_phi, beta = get_initial_conditions(np.array([row[0], 0, 0]), row[20:23], mass_bh=row[4])
    assert np.array_equal(q0, row[7:11])
    assert relerr(np.array([*p0, alpha0, h_r, h_theta, h_phi, beta]), row[11:20], 1.0).max() < 1e-13
    with pytest.raises(ValueError):
        get_initial_conditions(np.array([1.5, 0, 0]), np.array([1.2, 0.1, 0.0]))


@pytest.mark.parametrize("name", ["R2", "R5"])
def test_run_manual_simulation_vs_reference_outputs(golden, tmp_path, monkeypatch, name):
    """The reference's own entry point, same arguments as when the golden was made: image, photon_data.csv and
    sampled_rays.csv are compared with what the reference wrote."""
    import pandas as pd
    import random
    from simulation.blackhole import BlackHole, Observer
    from simulation.raytracing import run_manual_simulation
    g = golden(f"render_{name}.npz")
    c = json.loads(str(g["cfg"]))
    monkeypatch.chdir(tmp_path)
    bh = BlackHole(mass=c["mass"])
    obs = Observer(position=np.array([c["obs_x"], 0, 0]), fov=np.radians(c["fov"]), image_size=(c["h"], c["w"]))
    random.seed(0)                                                       # same sampled pixels as the golden run
    img, trajs = run_manual_simulation(
        bh, obs, real_time=False, update_every=1, steps=c["steps"], delta=c["delta"], omega=0.01, rtol=1e-2, atol=1e-2,
        order=2, suppress_warnings=False, background_path=os.path.join(GOLDEN, "texture_card.png"), use_cuda=True,
        boundary_radius=c["r_max"], patch_center_theta=float(g["patch_center"][0]), patch_center_phi=float(g["patch_center"][1]),
        patch_size_theta=np.deg2rad(c["pst"]), patch_size_phi=np.deg2rad(c["psp"]), flip_theta=c["flt"], flip_phi=c["flp"],
        n_samples=2)
    assert img.shape == g["img"].shape and img.dtype == np.uint8
    kw, _ = scene_kwargs_from_cfg(g["cfg"], g["patch_center"])
    fu, fv = texel_coords(kw, g["final_th"], g["final_ph"])
    diff = np.any(img != g["img"], axis=2).reshape(-1)
    assert np.all(near_rounding_boundary(fu, fv)[diff]) and diff.mean() < 5e-3
    # files the reference writes (raytracing.py:282-298)
    assert os.path.exists("images/manual_output.png")
    df = pd.read_csv("photon_data.csv")
    assert list(df.columns) == ['i', 'j', 'final_r', 'final_th', 'final_ph', 'collision', 'h_r', 'h_theta', 'h_phi',
                                'p0_t', 'p0_r', 'p0_th', 'p0_ph', 'alpha0']
    classes = str(g["classes"]).split("|")
    assert np.array_equal(df["collision"].map({k: i for i, k in enumerate(classes)}).to_numpy(), g["cls"].reshape(-1))
    nb = (g["cls"].reshape(-1) != 0) & ~in_band_g(g["p0"])
    assert state_close(df["final_r"].to_numpy()[nb], g["final_r"][nb]).all()
    assert relerr(df[["p0_t", "p0_r", "p0_th", "p0_ph"]].to_numpy(), g["p0"], 1.0).max() < 1e-13
    assert relerr(df[["h_r", "h_theta", "h_phi"]].to_numpy(), g["h_dir"], 1.0).max() < 1e-13
    # sampled trajectories: same sampled pixels (seeded), decimated rows, rotated back by beta
    assert len(trajs) == 2
    sr = pd.read_csv("sampled_rays.csv")
    assert " ".join(sr.columns) == str(g["sampled_rays_cols"])
    for t, ref in zip(trajs, (g["sampled_traj0"], g["sampled_traj1"])):
        assert t.shape == ref.shape
        assert (np.abs(t - ref) <= 1e-9 * np.maximum(np.abs(ref), 1.0)).all()
    assert sr.shape == g["sampled_rays"].shape
    assert (np.abs(sr.to_numpy(np.float64) - g["sampled_rays"]) <= 1e-9 * np.maximum(np.abs(g["sampled_rays"]), 1.0)).all()


def test_flat_render_vs_reference_and_oracle(handle, orc, golden, tmp_path):
    """N1 (background.py): kernel vs the reference's no_gravity.png and vs the oracle; entry-point side effects."""
    import schw_b200
    from simulation.background import save_no_gravity_image_with_background
    from simulation.blackhole import Observer
    from PIL import Image
    for name in ("F1", "F2", "F3"):
        g = golden(f"flat_{name}.npz")
        kw = flat_scene_kwargs(g["cfg"])
        img = handle.flat_render(schw_b200.make_scene(**kw), g["bg_array"])
        # pixels may differ from the reference's image only ON a truncation boundary of its int() texel map (or a patch
        # edge), where the last bit of acos/atan2 decides -- checked per pixel, not assumed
        for other in (g["img"], orc.flat_render(orc.make_scene(**kw), g["bg_array"])):
            ndiff, unexplained = flat_mismatch_explained(kw, img, other)
            assert unexplained == 0 and ndiff <= 2e-3 * img.shape[0] * img.shape[1], (name, ndiff, unexplained)
    # bigger frame vs the oracle
    from helpers import synthetic_sky
    tex = synthetic_sky(270, 480, 3)
    kw = dict(obs=(20.0, 0, 0), fov=np.radians(90), height=270, width=480, r_max=21.0, patch_center_theta=np.deg2rad(100),
              patch_center_phi=np.deg2rad(205), patch_size_theta=np.deg2rad(126), patch_size_phi=np.deg2rad(224), flip_phi=True)
    a = handle.flat_render(schw_b200.make_scene(**kw), tex); b = orc.flat_render(orc.make_scene(**kw), tex)
    ndiff, unexplained = flat_mismatch_explained(kw, a, b)
    assert unexplained == 0 and ndiff < 1e-3 * a.shape[0] * a.shape[1] and a.any(), (ndiff, unexplained)
    # the entry point with the reference's signature and files
    c = json.loads(str(golden("flat_F2.npz")["cfg"]))
    obs = Observer(position=np.array([c["obs_x"], 0, 0]), fov=np.radians(c["fov"]), image_size=(c["h"], c["w"]))
    out = tmp_path / "images" / "no_gravity.png"
    trajs = save_no_gravity_image_with_background(
        obs, os.path.join(GOLDEN, "texture_card.png"), str(out), boundary_radius=c["r_max"],
        patch_center_theta=np.deg2rad(c["pct"]), patch_center_phi=np.deg2rad(c["pcp"]), patch_size_theta=np.deg2rad(c["pst"]),
        patch_size_phi=np.deg2rad(c["psp"]), flip_theta=c["flt"], flip_phi=c["flp"], return_sampled_trajectories=True,
        n_sampled=4, override_patch_center=True)
    assert out.exists() and (tmp_path / "images" / "scene_full.png").exists()
    assert np.any(np.array(Image.open(out).convert("RGB")) != golden("flat_F2.npz")["img"], axis=2).mean() <= 2e-3
    assert len(trajs) == 4 and all(t.shape == (100, 3) for t in trajs)
    assert all(abs(np.linalg.norm(t[-1]) - c["r_max"]) < 1e-9 for t in trajs)


def test_cuda_flat_raytrace_wrappers(handle, orc):
    """`from simulation.cuda_geodesic import cuda_flat_raytrace[_with_trajectories]` as background.py:77,95 uses them:
    caller-supplied unit directions, flattened (N,3) image, (S,n_points,3) straight-line samples."""
    import schw_b200
    from simulation.cuda_geodesic import cuda_flat_raytrace, cuda_flat_raytrace_with_trajectories
    from helpers import synthetic_sky
    h, w = 120, 200
    tex = synthetic_sky(h, w, 7)
    kw = dict(obs=(25.0, 0, 0), fov=np.radians(75), height=h, width=w, r_max=40.0, patch_center_theta=np.deg2rad(95),
              patch_center_phi=np.deg2rad(170), patch_size_theta=np.deg2rad(150), patch_size_phi=np.deg2rad(300), flip_theta=True)
    # ray directions exactly as background.py:80-91 builds them
    obs = np.array(kw["obs"]); axis = -obs / np.linalg.norm(obs)
    right = np.cross([0.0, 0.0, 1.0], axis); right /= np.linalg.norm(right)
    up = np.cross(axis, right); up /= np.linalg.norm(up)
    pc = obs + axis * 0.2 * np.linalg.norm(obs)
    pw = 2 * 0.2 * np.linalg.norm(obs) * np.tan(kw["fov"] / 2); ph = pw * (h / w)
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    px = pc[None, None, :] + (((jj + 0.5) / w - 0.5) * pw)[..., None] * right + (((ii + 0.5) / h - 0.5) * ph)[..., None] * up
    dirs = px - obs[None, None, :]
    dirs = (dirs / np.linalg.norm(dirs, axis=2, keepdims=True)).reshape(-1, 3)
    patch = (kw["patch_center_theta"], kw["patch_center_phi"], kw["patch_size_theta"], kw["patch_size_phi"], True, False)
    flat = cuda_flat_raytrace(obs, dirs, kw["r_max"], *patch, tex)
    assert flat.shape == (h * w, 3) and flat.dtype == np.uint8 and flat.any()
    img = flat.reshape(h, w, 3)
    for other in (handle.flat_render(schw_b200.make_scene(**kw), tex), orc.flat_render(orc.make_scene(**kw), tex)):
        ndiff, unexplained = flat_mismatch_explained(kw, img, other)
        assert unexplained == 0 and ndiff <= 2e-3 * h * w, (ndiff, unexplained)
    picks = [0, 5 * w + 17, h * w - 1, 5 * w + 17]                      # a duplicate index fills both of its rows
    flat2, trajs = cuda_flat_raytrace_with_trajectories(obs, dirs, kw["r_max"], *patch, tex, picks, n_points=100)
    assert np.array_equal(flat2, flat) and trajs.shape == (4, 100, 3)
    for s, k in enumerate(picks):
        d = dirs[k]; b = 2 * obs @ d; c = obs @ obs - kw["r_max"] ** 2
        hit = obs + (-b + np.sqrt(b * b - 4 * (d @ d) * c)) / (2 * (d @ d)) * d
        want = obs[None, :] + (np.arange(100) / 99.0)[:, None] * (hit - obs)[None, :]
        assert np.allclose(trajs[s], want, rtol=0, atol=1e-12)
    # a ray whose line misses the sphere (observer outside it, looking sideways): black pixel, zero trajectory rows
    miss, tr = cuda_flat_raytrace_with_trajectories(np.array([50.0, 0, 0]), np.array([[0.0, 1.0, 0]]), 10.0, *patch, tex, [0], n_points=5)
    assert not miss.any() and not tr.any()


def test_main_entry_point_end_to_end(tmp_path, monkeypatch, capsys):
    """`python main.py --size 32 ...` (BASELINE config 1 shape, small): files and summary like the reference's main.py."""
    import pandas as pd
    import main as entry
    monkeypatch.chdir(tmp_path)
    entry.main(["--size", "32", "--delta", "0.1", "--background", os.path.join(GOLDEN, "texture_card.png"), "--no-plots"])
    for f in ("images/no_gravity.png", "images/scene_full.png", "images/manual_output.png", "photon_data.csv", "sampled_rays.csv"):
        assert os.path.exists(f), f
    df = pd.read_csv("photon_data.csv")
    assert len(df) == 32 * 32 and set(df["collision"]) == {"bh", "escape_bg"}
    out = capsys.readouterr().out
    assert "Photon summary:" in out and f"Captured by BH: {(df['collision'] == 'bh').sum()}" in out
    sr = pd.read_csv("sampled_rays.csv")
    assert sr["ray_id"].nunique() == 20 and sr.groupby("ray_id").size().max() == 1000


def test_main_entry_point_with_dedup_writes_the_same_frame_and_table(tmp_path, monkeypatch):
    """`main.py --dedup` (additive flag): image and photon_data.csv identical to the default run."""
    import pandas as pd
    from PIL import Image
    import main as entry
    args = ["--size", "40", "--delta", "0.1", "--background", os.path.join(GOLDEN, "texture_card.png"), "--no-plots",
            "--no-flat-trajectories"]
    res = {}
    for tag, extra in (("plain", []), ("dedup", ["--dedup"])):
        d = tmp_path / tag
        d.mkdir()
        monkeypatch.chdir(d)
        entry.main(args + extra)
        res[tag] = (np.array(Image.open("images/manual_output.png").convert("RGB")), pd.read_csv("photon_data.csv"))
    assert np.array_equal(res["plain"][0], res["dedup"][0]) and res["plain"][0].any()
    pd.testing.assert_frame_equal(res["plain"][1], res["dedup"][1], check_exact=True)
