#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

This script only runs in the build container (it needs /root/reference, which
does not exist on the GPU box).  It imports the reference's own modules
(`simulation.cuda_geodesic`, `simulation.utils`, `simulation.raytracing`) with

  * NUMBA_ENABLE_CUDASIM=1, so the reference's Numba-CUDA kernels
    (simulation/cuda_geodesic.py:579-625) execute on the CPU simulator, and
  * `sys.modules` stubs for the two third-party packages that are not
    installed here: `einsteinpy.coordinates.utils` (two 4-line coordinate
    helpers, restated from einsteinpy 0.4.0) and
    `matplotlib` (only `pyplot.imsave` is touched on this path).

Nothing in the reference tree is modified or copied; only its *outputs* on
seeded inputs are recorded as .npz fixtures.  The known-answer data files the
reference ships (tests/einsteinpy-geodesic-test/trajectory.csv and the Plotly
arrays of einsteinpy_ray.ipynb, both genuine EinsteinPy output) are decimated
into .npz as well.

Usage:  python tests/golden/make_golden.py <part> [<part> ...]
Parts:  kat single ic stepper_default stepper_general render_R1 render_R2
        render_R3 render_R4 render_R5 flat_F1 flat_F2 flat_F3   (or "all")
"""
import base64
import json
import os
import random
import sys
import tempfile
import time
import types

os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
import numpy as np  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


# --------------------------------------------------------------------------
# stubs for the un-installed third-party imports of the reference
# --------------------------------------------------------------------------
def install_stubs():
    ep = types.ModuleType("einsteinpy")
    epc = types.ModuleType("einsteinpy.coordinates")
    epu = types.ModuleType("einsteinpy.coordinates.utils")

    def cartesian_to_spherical_fast(t, x, y, z, v_x=None, v_y=None, v_z=None,
                                    velocities_provided=False):
        hxy = np.hypot(x, y)
        r = np.hypot(hxy, z)
        th = np.arctan2(hxy, z)
        p = np.arctan2(y, x)
        return t, r, th, p

    def spherical_to_cartesian_fast(t, r, th, p, v_r=None, v_th=None, v_p=None,
                                    velocities_provided=False):
        x = r * np.cos(p) * np.sin(th)
        y = r * np.sin(p) * np.sin(th)
        z = r * np.cos(th)
        return t, x, y, z

    epu.cartesian_to_spherical_fast = cartesian_to_spherical_fast
    epu.spherical_to_cartesian_fast = spherical_to_cartesian_fast
    ep.coordinates = epc
    epc.utils = epu
    sys.modules["einsteinpy"] = ep
    sys.modules["einsteinpy.coordinates"] = epc
    sys.modules["einsteinpy.coordinates.utils"] = epu

    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    plt.imsave = lambda *a, **k: None
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    if REF not in sys.path:
        sys.path.insert(0, REF)


def camera_pixel(obs_pos, fov, h, w, i, j):
    """Pixel position exactly as the reference builds it (raytracing.py:60-88)."""
    optical_axis = np.array([-1, 0, 0])
    right = np.array([0, 1, 0])
    up_vec = np.array([0, 0, 1])
    plane_dist = 0.2 * np.linalg.norm(obs_pos)
    plane_center = obs_pos + optical_axis * plane_dist
    plane_width = 2 * plane_dist * np.tan(fov / 2)
    plane_height = plane_width * (h / w)
    u = (j + 0.5) / w - 0.5
    v = (i + 0.5) / h - 0.5
    return plane_center + u * plane_width * right + v * plane_height * up_vec


# --------------------------------------------------------------------------
def part_kat():
    import pandas as pd
    df = pd.read_csv(f"{REF}/tests/einsteinpy-geodesic-test/trajectory.csv")
    a = df[["t", "x", "y", "z", "pt", "pr", "pth", "pph"]].to_numpy(np.float64)
    nvalid = 2868                      # rows beyond this are EinsteinPy's r<0 blow-up
    idx = np.unique(np.concatenate([np.arange(0, 64), np.arange(0, nvalid, 4),
                                    np.arange(nvalid - 64, nvalid)]))
    np.savez_compressed(
        f"{OUT}/kat_k1_einsteinpy_trajectory.npz",
        step=(idx + 1).astype(np.int64), vals=a[idx],
        q0=np.array([0.0, 10.0, np.pi / 2, 0.0]),
        p_sph=np.array([1.0, np.pi / 2 - np.deg2rad(60.0), np.pi - np.deg2rad(60.0)]),
        delta=0.05, omega=0.01, mass=1.0,
        note="rows of tests/einsteinpy-geodesic-test/trajectory.csv (genuine EinsteinPy Nulllike output); "
             "row k = state after k+1 steps; cols t,x,y,z,pt,pr,pth,pph")

    nb = json.load(open(f"{REF}/tests/einsteinpy-geodesic-test/einsteinpy_ray.ipynb"))
    out = [o for o in nb["cells"][9]["outputs"] if o["output_type"] == "execute_result"][0]
    tr = [t for t in out["data"]["application/vnd.plotly.v1+json"]["data"] if t.get("type") == "scatter3d"][0]
    xyz = np.stack([np.frombuffer(base64.b64decode(tr[k]["bdata"]), dtype=np.float64) for k in "xyz"], axis=1)
    n = xyz.shape[0]
    idx = np.unique(np.concatenate([np.arange(0, 64), np.arange(0, n, 5), np.arange(n - 64, n)]))
    np.savez_compressed(
        f"{OUT}/kat_k2_einsteinpy_notebook.npz",
        step=(idx + 1).astype(np.int64), xyz=xyz[idx],
        q0=np.array([0.0, 4.0, np.pi / 2, 0.0]),
        p_sph=np.array([1.0, np.pi / 2 - np.deg2rad(70.0), np.pi - np.deg2rad(70.0)]),
        delta=0.001, omega=0.01, mass=1.0,
        note="x,y,z arrays embedded in einsteinpy_ray.ipynb cell 9 (EinsteinPy Nulllike, 10000 steps)")
    print("kat: K1", a.shape, "K2", xyz.shape)


def part_single():
    """single_ray_cuda_test.py:250-321 parameters through the reference kernel."""
    from simulation.cuda_geodesic import CUDASchwarzschildIntegrator
    from simulation.utils import build_null_4momentum_ep_sph
    pos_sph = np.array([35.0, np.pi / 2, 0.0])
    p_dir = np.array([-0.026942690335328513, -0.028502831807219468, 0.06898831276132347])
    # the script's local builder picks the NEGATIVE root with future=True (single_ray_cuda_test.py:219-222);
    # simulation/utils.py's builder gives the same magnitude, negative for future=False.
    p0 = build_null_4momentum_ep_sph(p_dir, pos_sph, mass_bh=1.0, future=False)
    q0 = np.array([0.0, *pos_sph])
    steps, delta, omega, r_max = 200_000, 0.03, 0.01, 50
    integ = CUDASchwarzschildIntegrator(steps=steps, delta=delta, mass=1.0, omega=omega, r_max=r_max)
    t0 = time.time()
    traj = integ.integrate_batch_full(q0[None, :], p0[None, :])[0]
    nrows = int(np.max(np.nonzero(traj[:, 1])[0])) + 1
    print("single: stored rows", nrows, "in", time.time() - t0, "s; last", traj[nrows - 1])
    np.savez_compressed(f"{OUT}/single_ray_k3.npz", q0=q0, p0=p0, traj=traj[:nrows],
                        steps=steps, delta=delta, omega=omega, r_max=float(r_max), mass=1.0,
                        note="reference integrate_batch_full (Numba CUDA simulator) on single_ray_cuda_test.py inputs")


def part_ic():
    from simulation.utils import get_initial_conditions
    cases = []
    rng = np.random.default_rng(1234)
    # (obs_x, fov_deg, h, w, mass, pixel list)
    frames = [
        (30, 80.0, 400, 400, 1.0, [(200, 200), (200, 215), (200, 225), (200, 236), (200, 238), (200, 240), (200, 243),
                                   (200, 250), (200, 275), (200, 320), (200, 399), (0, 0), (399, 399), (0, 399),
                                   (399, 0), (17, 333), (123, 45)]),
        (30, 80.0, 24, 36, 1.0, [(i, j) for i in range(24) for j in range(36)]),
        (30, 80.0, 1080, 1920, 1.0, [(int(rng.integers(0, 1080)), int(rng.integers(0, 1920))) for _ in range(300)]),
        (30, 80.0, 2160, 3840, 1.0, [(int(rng.integers(0, 2160)), int(rng.integers(0, 3840))) for _ in range(300)]),
        (20.0, 90.0, 500, 500, 1.0, [(int(rng.integers(0, 500)), int(rng.integers(0, 500))) for _ in range(100)]),
        (12.5, 35.0, 64, 48, 1.5, [(int(rng.integers(0, 64)), int(rng.integers(0, 48))) for _ in range(100)]),
    ]
    for obs_x, fov_deg, h, w, mass, pix in frames:
        obs = np.array([obs_x, 0, 0])
        fov = np.radians(fov_deg)
        for (i, j) in pix:
            px = camera_pixel(obs, fov, h, w, i, j)
            q0, p0, a0, h_r, h_th, h_ph, beta = get_initial_conditions(obs, px, mass_bh=mass)
            cases.append([obs_x, fov, h, w, mass, i, j, *q0, *p0, a0, h_r, h_th, h_ph, beta, *px])
    arr = np.array(cases, dtype=np.float64)
    np.savez_compressed(f"{OUT}/ic_golden.npz", table=arr,
                        cols="obs_x fov h w mass i j q0_t q0_r q0_th q0_ph p0_t p0_r p0_th p0_ph alpha0 h_r h_theta h_phi beta px py pz",
                        note="reference get_initial_conditions (simulation/utils.py:91-141) on reference camera pixels (raytracing.py:60-88)")
    print("ic:", arr.shape)


def part_stepper_default():
    """Camera rays of the default 400x400 scene at the default step size (config 1)."""
    from simulation.cuda_geodesic import CUDASchwarzschildIntegrator
    from simulation.utils import get_initial_conditions
    obs = np.array([30, 0, 0])
    fov = np.radians(80.0)
    pix = [(200, 200), (200, 215), (200, 225), (200, 236), (200, 238), (200, 240), (200, 243), (200, 250),
           (200, 275), (200, 320), (200, 399), (0, 0), (37, 181), (111, 290), (260, 140), (333, 333),
           (150, 201), (205, 160), (399, 200), (220, 228), (180, 170), (10, 390), (300, 50), (240, 240)]
    q0s, p0s, a0s, betas = [], [], [], []
    for (i, j) in pix:
        q0, p0, a0, _, _, _, beta = get_initial_conditions(obs, camera_pixel(obs, fov, 400, 400, i, j), mass_bh=1.0)
        q0s.append(q0); p0s.append(p0); a0s.append(a0); betas.append(beta)
    q0s, p0s = np.array(q0s), np.array(p0s)
    integ = CUDASchwarzschildIntegrator(steps=200000, delta=0.01, mass=1.0, r_max=31)   # omega defaults to 1.0 (quirk D-1)
    t0 = time.time()
    out, _ = integ.integrate_batch(q0s, p0s)
    print("stepper_default:", len(pix), "rays in", time.time() - t0, "s")
    np.savez_compressed(f"{OUT}/stepper_default.npz", pix=np.array(pix), q0s=q0s, p0s=p0s, alpha0=np.array(a0s),
                        beta=np.array(betas), out_qs=out, steps=200000, delta=0.01, omega=1.0, r_max=31.0, mass=1.0,
                        note="reference integrate_batch (simulator) on default-scene camera rays")


def part_stepper_general():
    """Non-equatorial rays, both omega settings, a step cap that some rays hit."""
    from simulation.cuda_geodesic import CUDASchwarzschildIntegrator
    from simulation.utils import build_null_4momentum_ep_sph
    rng = np.random.default_rng(20261018)
    n = 160
    q0s = np.zeros((n, 4)); p0s = np.zeros((n, 4))
    for k in range(n):
        r0 = rng.uniform(8.0, 40.0)
        th0 = rng.uniform(0.35, np.pi - 0.35)
        ph0 = rng.uniform(-np.pi, np.pi)
        # inward-pointing direction with a random tilt; scaled like the reference's coordinate-basis momenta
        a = np.deg2rad(rng.uniform(0.0, 55.0)); b = rng.uniform(-np.pi, np.pi)
        pr = -np.cos(a) * np.sqrt(1 - 2 / r0)
        pth = np.sin(a) * np.sin(b) * r0
        pph = np.sin(a) * np.cos(b) * r0 * np.sin(th0)
        p0s[k] = build_null_4momentum_ep_sph(np.array([pr, pth, pph]), np.array([r0, th0, ph0]),
                                             mass_bh=1.0, future=bool(k % 2))
        q0s[k] = [0.0, r0, th0, ph0]
    outs = {}
    for tag, omega in (("w1", 1.0), ("w001", 0.01)):
        integ = CUDASchwarzschildIntegrator(steps=2500, delta=0.05, mass=1.0, omega=omega, r_max=45.0)
        t0 = time.time()
        out, _ = integ.integrate_batch(q0s, p0s)
        outs[tag] = out
        print("stepper_general", tag, time.time() - t0, "s")
    np.savez_compressed(f"{OUT}/stepper_general.npz", q0s=q0s, p0s=p0s, out_w1=outs["w1"], out_w001=outs["w001"],
                        steps=2500, delta=0.05, r_max=45.0, mass=1.0,
                        note="reference integrate_batch (simulator), random non-equatorial null rays")


def synth_texture(path, W=96, H=48, seed=7):
    """Small deterministic equirectangular test card (every texel distinct in at least one channel)."""
    from PIL import Image
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W, 3), np.uint8)
    img[..., 0] = (xx * 255 // (W - 1)).astype(np.uint8)
    img[..., 1] = (yy * 255 // (H - 1)).astype(np.uint8)
    img[..., 2] = rng.integers(0, 256, size=(H, W), dtype=np.uint8)
    Image.fromarray(img).save(path)
    return img


RENDERS = {
    # name: dict(h, w, obs_x, fov_deg, mass, steps, delta, omega_cli, r_max, pct, pcp, dth_rel, dph_rel, pst, psp, flt, flp)
    "R1": dict(h=24, w=24, obs_x=30, fov=80.0, mass=1.0, steps=200000, delta=0.05, r_max=31,
               pct=90, pcp=180, rel_t=0, rel_p=0, pst=180, psp=360, flt=False, flp=False),
    "R2": dict(h=20, w=28, obs_x=30, fov=80.0, mass=1.0, steps=200000, delta=0.1, r_max=31,
               pct=90, pcp=180, rel_t=10, rel_p=-25, pst=70, psp=110, flt=True, flp=True),
    "R3": dict(h=16, w=16, obs_x=30, fov=80.0, mass=1.0, steps=200000, delta=0.2, r_max=150,
               pct=90, pcp=180, rel_t=0, rel_p=0, pst=180, psp=360, flt=False, flp=False),
    "R4": dict(h=16, w=20, obs_x=30, fov=80.0, mass=1.0, steps=120, delta=0.1, r_max=31,
               pct=90, pcp=180, rel_t=0, rel_p=0, pst=180, psp=360, flt=False, flp=False),
    "R5": dict(h=18, w=30, obs_x=20.0, fov=90.0, mass=1.0, steps=30000, delta=0.1, r_max=21.0,
               pct=90, pcp=180, rel_t=0, rel_p=40, pst=126, psp=224, flt=False, flp=True),
}


def part_render(name):
    import pandas as pd
    from simulation.blackhole import BlackHole, Observer
    from simulation.raytracing import run_manual_simulation
    from simulation.utils import _apply_relative_offsets
    from PIL import Image
    c = RENDERS[name]
    tex_path = f"{OUT}/texture_card.png"
    if not os.path.exists(tex_path):
        synth_texture(tex_path)
    cwd = os.getcwd()
    tmp = tempfile.mkdtemp(prefix=f"golden_{name}_")
    os.makedirs(f"{tmp}/images", exist_ok=True)
    os.chdir(tmp)
    try:
        bh = BlackHole(mass=c["mass"])
        obs = Observer(position=np.array([c["obs_x"], 0, 0]), fov=np.radians(c["fov"]), image_size=(c["h"], c["w"]))
        pct, pcp = _apply_relative_offsets(c["pct"], c["pcp"], c["rel_t"], c["rel_p"])
        random.seed(0)
        t0 = time.time()
        img, trajs = run_manual_simulation(
            bh, obs, real_time=False, update_every=1, steps=c["steps"], delta=c["delta"], omega=0.01,
            rtol=1e-2, atol=1e-2, order=2, suppress_warnings=False, background_path=tex_path, use_cuda=True,
            boundary_radius=c["r_max"], patch_center_theta=pct, patch_center_phi=pcp,
            patch_size_theta=np.deg2rad(c["pst"]), patch_size_phi=np.deg2rad(c["psp"]),
            flip_theta=c["flt"], flip_phi=c["flp"], n_samples=2)
        print("render", name, "took", time.time() - t0, "s")
        df = pd.read_csv("photon_data.csv")
        sr = pd.read_csv("sampled_rays.csv")
    finally:
        os.chdir(cwd)
    bg = np.array(Image.open(tex_path).convert("RGB").resize((c["w"], c["h"]), Image.LANCZOS))
    classes = ["bh", "numerical error", "escape_bg", "escape_no_patch", "in_domain"]
    cls = df["collision"].map({k: i for i, k in enumerate(classes)}).to_numpy(np.uint8)
    np.savez_compressed(
        f"{OUT}/render_{name}.npz", img=img, bg_array=bg, cls=cls.reshape(c["h"], c["w"]),
        final_r=df["final_r"].to_numpy(), final_th=df["final_th"].to_numpy(), final_ph=df["final_ph"].to_numpy(),
        alpha0=df["alpha0"].to_numpy(), p0=df[["p0_t", "p0_r", "p0_th", "p0_ph"]].to_numpy(),
        h_dir=df[["h_r", "h_theta", "h_phi"]].to_numpy(),
        patch_center=np.array([pct, pcp]),
        sampled_traj0=trajs[0], sampled_traj1=trajs[1],
        sampled_rays=sr.to_numpy(np.float64), sampled_rays_cols=" ".join(sr.columns),
        cfg=json.dumps(c), classes="|".join(classes),
        note="reference run_manual_simulation (raytracing.py:16-314), random.seed(0), n_samples=2, texture_card.png")


FLATS = {
    "F1": dict(h=40, w=56, obs_x=30, fov=80.0, r_max=31, pct=90, pcp=180, pst=180, psp=360, flt=False, flp=False, override=False),
    "F2": dict(h=36, w=48, obs_x=20.0, fov=90.0, r_max=21.0, pct=100, pcp=205, pst=70, psp=110, flt=True, flp=True, override=True),
    "F3": dict(h=32, w=32, obs_x=30, fov=80.0, r_max=31, pct=90, pcp=300, pst=126, psp=224, flt=False, flp=True, override=True),
}


def part_flat(name):
    """Reference flat-space render (simulation/background.py:19-141, CPU loop) on the texture card."""
    from simulation.blackhole import Observer
    from simulation.background import save_no_gravity_image_with_background
    from PIL import Image
    c = FLATS[name]
    tex_path = f"{OUT}/texture_card.png"
    tmp = tempfile.mkdtemp(prefix=f"golden_{name}_")
    obs = Observer(position=np.array([c["obs_x"], 0, 0]), fov=np.radians(c["fov"]), image_size=(c["h"], c["w"]))
    save_no_gravity_image_with_background(
        obs, tex_path, f"{tmp}/images/no_gravity.png", boundary_radius=c["r_max"],
        patch_center_theta=np.deg2rad(c["pct"]), patch_center_phi=np.deg2rad(c["pcp"]),
        patch_size_theta=np.deg2rad(c["pst"]), patch_size_phi=np.deg2rad(c["psp"]),
        flip_theta=c["flt"], flip_phi=c["flp"], return_sampled_trajectories=False, override_patch_center=c["override"])
    img = np.array(Image.open(f"{tmp}/images/no_gravity.png").convert("RGB"))
    bg = np.array(Image.open(tex_path).convert("RGB").resize((c["w"], c["h"]), Image.LANCZOS))
    np.savez_compressed(f"{OUT}/flat_{name}.npz", img=img, bg_array=bg, cfg=json.dumps(c),
                        note="reference save_no_gravity_image_with_background (background.py:19-141), texture_card.png")
    print("flat", name, img.shape, "nonblack", int(np.any(img != 0, axis=2).sum()))


def main():
    install_stubs()
    parts = sys.argv[1:] or ["all"]
    if parts == ["all"]:
        parts = (["kat", "single", "ic", "stepper_default", "stepper_general"] + [f"render_{k}" for k in RENDERS]
                 + [f"flat_{k}" for k in FLATS])
    for p in parts:
        if p.startswith("render_"):
            part_render(p.split("_", 1)[1])
        elif p.startswith("flat_"):
            part_flat(p.split("_", 1)[1])
        else:
            globals()[f"part_{p}"]()


if __name__ == "__main__":
    main()
