"""CPU check of the planar fast path over many parameter sets, with the kernels' OWN source: csrc/schw_device.cuh is
compiled for the host (tests/model/planar_model.cpp, one lane standing for a warp) and every ray is integrated twice
with the same arithmetic -- general flows only vs. step_BCBA_planar driven as persistent_rays drives it.  The two must
agree in every bit of the 16-double state after every step, for the thresholds the library really uses
(schw_debug_step_params: no device needed).  The GPU counterpart is tests/test_gpu_planar.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

MODEL_DIR = os.path.join(ROOT, "tests", "model")
TH_EQ = float(np.pi / 2)


@pytest.fixture(scope="module")
def model():
    import schw_b200
    schw_b200.build()
    lib = schw_b200.load()
    out = os.path.join(MODEL_DIR, "_build", "planar_model.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(MODEL_DIR, "planar_model.cpp")
    hdr = os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200", "csrc", "schw_device.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src])
    m = C.CDLL(out)
    m.model_run_ray.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]

    def params(steps, delta, rs, r_max, omega):
        size = C.c_size_t()
        assert lib.schw_debug_step_params(int(steps), float(delta), float(rs), float(r_max), float(omega), None, 0, C.byref(size)) == 0
        buf = C.create_string_buffer(size.value)
        assert lib.schw_debug_step_params(int(steps), float(delta), float(rs), float(r_max), float(omega), buf, size.value, None) == 0
        return buf

    def run(buf, q0, p0):
        q0 = np.ascontiguousarray(q0, np.float64); p0 = np.ascontiguousarray(p0, np.float64)
        out = np.zeros(4, np.int64); fg = np.zeros(4); fp = np.zeros(4)
        rc = m.model_run_ray(buf, len(buf), q0.ctypes.data, p0.ctypes.data, out.ctypes.data, fg.ctypes.data, fp.ctypes.data)
        assert rc == 0
        return dict(steps=int(out[0]), planar=int(out[1]), first_diff=int(out[2]), same_end=bool(out[3]), fg=fg, fp=fp)
    return params, run


def _camera_ray(alpha, r_obs, mass=1.0, p_theta=-0.0, theta=TH_EQ):
    """Null ray of the reference's camera convention (utils.py:148-194, 39-88) at angle alpha from the optical axis."""
    f = 1.0 - 2.0 * mass / r_obs
    pr = -np.cos(alpha) * np.sqrt(f)
    pph = np.sin(alpha) * r_obs
    Cq = f * pr * pr + p_theta * p_theta / r_obs ** 2 + pph * pph / (r_obs * r_obs * np.sin(theta) ** 2)
    pt = np.sqrt(4.0 * Cq / f) / (2.0 / f)
    return np.array([0.0, r_obs, theta, 0.0]), np.array([pt, pr, p_theta, pph])


# (steps, delta, mass, r_obs, r_max, omega)
SCENES = [
    (200000, 0.01, 1.0, 30.0, 31.0, 1.0),       # default config.py scene
    (200000, 0.005, 1.0, 30.0, 31.0, 1.0),      # sweep step
    (200000, 0.05, 1.0, 45.0, 60.0, 1.0),       # coarse step: small bound T, rays leave the fast path
    (200000, 0.2, 1.0, 30.0, 31.0, 1.0),        # the integrator's default step
    (200000, 0.01, 0.5, 20.0, 25.0, 1.0),       # rs = 1 (literal 2.0 in the metric derivative)
    (200000, 0.01, 2.0, 60.0, 61.0, 1.0),       # rs = 4
    (4000, 0.03, 1.0, 35.0, 50.0, 0.01),        # single_ray_cuda_test coupling, step cap
    (200000, 0.01, 1.0, 30.0, 31.0, 0.3),
    (200000, 0.01, 1.0, 30.0, 31.0, 3.0),       # strong coupling: sin(2 w d) / 2 = 0.03
    (3000, 0.02, 1e-3, 5.0, 8.0, 1.0),          # tiny black hole: r_thr ~ 1.6e-3, rays pass far outside
]


@pytest.mark.parametrize("scene", SCENES, ids=[f"d{s[1]}_m{s[2]}_w{s[5]}" for s in SCENES])
def test_planar_step_equals_general_step_bit_for_bit(model, scene):
    params, run = model
    steps, delta, mass, r_obs, r_max, omega = scene
    P = params(steps, delta, 2.0 * mass, r_max, omega)
    rng = np.random.default_rng(7)
    # angles across the frame, dense around the critical one (where rays wind and p_theta grows), plus exact 0 (radial)
    f = 1.0 - 2.0 * mass / r_obs
    a_crit = np.arcsin(min(1.0, 3 * np.sqrt(3) * mass * np.sqrt(f) / r_obs))      # rough, only to place samples
    alphas = np.concatenate([[0.0, 1e-9, 1e-5], np.deg2rad(rng.uniform(0.0, 50.0, 24)),
                             a_crit * (1 + rng.uniform(-0.08, 0.08, 24)), a_crit * (1 + rng.uniform(-0.004, 0.004, 12))])
    total = planar = 0
    for a in alphas:
        q0, p0 = _camera_ray(a, r_obs, mass)
        r = run(P, q0, p0)
        assert r["first_diff"] == -1 and r["same_end"], (a, r)
        assert np.array_equal(r["fg"].view(np.int64), r["fp"].view(np.int64))
        total += r["steps"]; planar += r["planar"]
    assert total > 0
    if delta <= 0.01 and omega <= 1.0 and 0.5 <= mass <= 1.0:
        assert planar > 0.9 * total, (planar, total)         # camera rays live on the fast path (the sample is dense
                                                             # around the critical angle, where rays wind and leave it)
    else:
        assert planar > 0, (planar, total)


def test_planar_step_leaves_and_rejoins_the_general_flow(model):
    """Rays the fast path must refuse or hand back: polar momentum, tilted start, |p_theta| at the bound, p_phi tiny,
    NaNs.  Same bit-for-bit requirement; the fast-path share is what distinguishes the cases."""
    params, run = model
    P = params(20000, 0.01, 2.0, 31.0, 1.0)
    base_q, base_p = _camera_ray(np.deg2rad(12.0), 30.0)
    cases = {
        "camera": (base_q, base_p, "all"),
        "p_theta=1e-13": (base_q, np.array([base_p[0], base_p[1], 1e-13, base_p[3]]), "none"),
        "p_theta=1e-17": (base_q, np.array([base_p[0], base_p[1], 1e-17, base_p[3]]), "all"),
        "p_theta=2e-15": (base_q, np.array([base_p[0], base_p[1], 2e-15, base_p[3]]), "some"),   # just below T = 2.8e-15
        "theta+1ulp": (np.array([0.0, 30.0, np.nextafter(TH_EQ, 2.0), 0.0]), base_p, "none"),
        "tilted": (np.array([0.0, 30.0, 1.2, 0.0]), np.array([base_p[0], base_p[1], 0.7, base_p[3]]), "none"),
        "p_phi=1e-9": (base_q, np.array([base_p[0], base_p[1], -0.0, 1e-9]), "none"),
        "radial": (base_q, np.array([base_p[0], base_p[1], -0.0, 0.0]), "none"),
        "nan_r": (np.array([0.0, np.nan, TH_EQ, 0.0]), base_p, "any"),
        "nan_p_theta": (base_q, np.array([base_p[0], base_p[1], np.nan, base_p[3]]), "any"),
    }
    for name, (q0, p0, want) in cases.items():
        r = run(P, q0, p0)
        assert r["first_diff"] == -1 and r["same_end"], (name, r)
        assert np.array_equal(r["fg"].view(np.int64), r["fp"].view(np.int64)) or np.isnan(r["fg"]).any(), name
        if want == "all":
            assert r["planar"] == r["steps"] > 0, (name, r)
        elif want == "none":
            assert r["planar"] == 0 and r["steps"] > 0, (name, r)
        elif want == "some":
            assert 0 < r["planar"] < r["steps"], (name, r)


def hi_rcap_of(buf):
    import struct
    return struct.unpack_from("8i", buf, 112)[4]


def hi_rmax_of(buf):
    import struct
    return struct.unpack_from("8i", buf, 112)[5]


def test_thresholds_satisfy_their_derivation(model):
    """(R1)-(R3) of csrc/schw_device.cuh 'Planar steps' from the numbers the library hands to the kernels."""
    import struct
    params, _ = model
    for steps, delta, mass, r_obs, r_max, omega in SCENES + [(10, 0.5, 1.0, 30.0, 31.0, 1.0), (10, 1e-4, 1.0, 30.0, 31.0, 50.0)]:
        buf = params(steps, delta, 2.0 * mass, r_max, omega).raw
        d = struct.unpack_from("14d", buf, 0)
        i = struct.unpack_from("8i", buf, 112)
        h, h_rs, dl, d_rs, mg, mhs, hcB, hcA, rs, zero_hs, rcap, rmax, park_r, _ = d
        # lanes without a ray are parked where the termination pre-test lets them be (strictly between the thresholds' high words)
        hiw = lambda x: struct.unpack("<ii", struct.pack("<d", x))[1]
        assert rcap < park_r < rmax and hi_rcap_of(buf) < hiw(park_r) < hi_rmax_of(buf)
        ph_hi, r_hi, pp_hi, enable, hi_rcap, hi_rmax, nsteps, _ = i
        assert (h, dl, rs, rmax, nsteps) == (0.5 * delta, delta, 2.0 * mass, r_max, steps)
        assert hcA == delta * 6.123233995736766e-17 and hcB == h * 6.123233995736766e-17 and zero_hs == 0.0
        assert 0.0 <= mg <= 1.0
        hi2d = lambda w: struct.unpack("<d", struct.pack("<II", 0, w & 0xffffffff))[0]
        assert hi2d(hi_rcap) <= rcap < hi2d(hi_rcap + 1) and hi2d(hi_rmax) <= rmax < hi2d(hi_rmax + 1)
        if not enable:
            continue
        T, r_thr, pp_thr = hi2d(ph_hi), hi2d(r_hi), hi2d(pp_hi)
        assert 0.0 < r_thr < rcap
        assert T * delta / r_thr ** 2 <= 2.0 ** -54 * (1 + 1e-12)          # (R1), delta >= delta/2
        assert T <= 2.0 ** -28 * pp_thr                                    # (R2)
        assert 2 * T * abs(mhs) <= 2.0 ** -54 * (1 + 1e-12)                # (R3)


def test_random_parameter_sets_and_near_threshold_rays(model):
    """Seeded stress: random (delta, mass, omega, observer) and equatorial rays whose p_theta starts anywhere between 0
    and 100x the bound T, i.e. rays that sit on, cross and re-cross every threshold of the fast path."""
    import struct
    params, run = model
    rng = np.random.default_rng(2026)
    n_rays = n_planar_steps = n_steps = n_mixed = 0
    for _ in range(120):
        delta = 10.0 ** rng.uniform(-3, -0.3)
        mass = 10.0 ** rng.uniform(-2, 0.7)
        omega = rng.choice([0.0, 0.01, 0.3, 1.0, 3.0, 10.0 ** rng.uniform(-2, 1)])
        r_obs = 2.0 * mass * rng.uniform(3.0, 40.0)
        r_max = r_obs * rng.uniform(1.01, 1.5)
        P = params(2500, delta, 2.0 * mass, r_max, omega)
        enable, = struct.unpack_from("i", P.raw, 112 + 12)
        T = struct.unpack("<d", struct.pack("<II", 0, struct.unpack_from("i", P.raw, 112)[0] & 0xffffffff))[0] if enable else 1e-15
        for _ in range(6):
            alpha = np.deg2rad(rng.uniform(0.0, 60.0))
            pth = rng.choice([-0.0, 0.0, 1.0, -1.0]) * T * 10.0 ** rng.uniform(-3, 2)
            q0, p0 = _camera_ray(alpha, r_obs, mass, p_theta=pth)
            r = run(P, q0, p0)
            assert r["first_diff"] == -1 and r["same_end"], (delta, mass, omega, r_obs, alpha, pth, r)
            n_rays += 1; n_steps += r["steps"]; n_planar_steps += r["planar"]
            n_mixed += 0 < r["planar"] < r["steps"]
    assert n_rays == 720 and n_steps > 100000
    assert 0.1 * n_steps < n_planar_steps < 0.95 * n_steps         # both paths are exercised ...
    assert n_mixed > 30                                            # ... also within one ray


def test_device_arithmetic_meets_the_truth_bar_on_the_cpu(model, orc):
    """The kernels' own arithmetic (this header compiled for the host: same FMAs, same reciprocal refinement, same sincos;
    only the 2^-20 hardware reciprocal seed is emulated) against the binary128 truth, with the bar of the GPU parity tests
    (helpers.assert_final_states): a re-derivation of the flows that lost accuracy would fail here, without a GPU."""
    from helpers import assert_final_states, random_rays
    params, run = model
    steps, delta, rs, r_max = 4000, 0.02, 2.0, 45.0
    for omega, n, min_tight in ((1.0, 384, 0.6), (0.01, 192, 0.9)):
        q0s, p0s = random_rays(n, seed=31)
        buf = params(steps, delta, rs, r_max, omega)
        out = np.zeros((n, 4)); ns = np.zeros(n, np.int64)
        for i in range(n):
            r = run(buf, q0s[i], p0s[i])
            out[i] = r["fp"]; ns[i] = r["steps"]
            assert r["first_diff"] == -1 and r["same_end"]
        assert_final_states(orc, out, ns, q0s, p0s, steps, delta, rs, r_max, omega, f"host model, omega {omega}", min_tight)
