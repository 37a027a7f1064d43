"""ctypes front-end of the CPU oracle (oracle/schw_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product (the package directory
`schwarzchild-einsteinpy-raytracing_b200/`) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libschw_oracle.so")

CLASSES = ("bh", "numerical error", "escape_bg", "escape_no_patch", "in_domain")


class OrcScene(C.Structure):
    _fields_ = [
        ("obs", C.c_double * 3), ("fov", C.c_double),
        ("height", C.c_int32), ("width", C.c_int32),
        ("mass", C.c_double), ("r_max", C.c_double),
        ("steps", C.c_int64), ("delta", C.c_double), ("omega", C.c_double),
        ("patch_center_theta", C.c_double), ("patch_center_phi", C.c_double),
        ("patch_size_theta", C.c_double), ("patch_size_phi", C.c_double),
        ("flip_theta", C.c_int32), ("flip_phi", C.c_int32),
        ("has_texture", C.c_int32), ("_pad", C.c_int32),
    ]


def build(force=False):
    src = os.path.join(_HERE, "schw_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_uint8)
        i32p = C.POINTER(C.c_int32)
        L.orc_integrate_batch.argtypes = [dp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                          C.c_double, dp, ip, dp, C.c_int]
        L.orc_integrate_batch_full.argtypes = [dp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                               C.c_double, dp, dp, ip]
        L.orc_initial_conditions.argtypes = [dp, dp, C.c_int64, C.c_double, dp, dp]
        L.orc_camera_pixels.argtypes = [C.POINTER(OrcScene), i32p, C.c_int64, dp]
        L.orc_classify.argtypes = [C.POINTER(OrcScene), dp, dp, dp, C.c_int64, bp, bp, bp, dp, i32p]
        L.orc_render.argtypes = [C.POINTER(OrcScene), bp, C.c_int, C.c_int, C.c_int, bp, bp, dp, dp, dp, i32p, ip,
                                 C.c_int]
        L.orc_flat_render.argtypes = [C.POINTER(OrcScene), bp, bp]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def max_threads():
    return int(lib().orc_max_threads())


def make_scene(obs=(30.0, 0.0, 0.0), fov=np.radians(80.0), height=400, width=400, mass=1.0, r_max=31.0,
               steps=200000, delta=0.01, omega=1.0, patch_center_theta=np.pi / 2, patch_center_phi=np.pi,
               patch_size_theta=np.pi, patch_size_phi=2 * np.pi, flip_theta=False, flip_phi=False,
               has_texture=True):
    s = OrcScene()
    s.obs[0], s.obs[1], s.obs[2] = [float(v) for v in obs]
    s.fov = float(fov); s.height = int(height); s.width = int(width)
    s.mass = float(mass); s.r_max = float(r_max); s.steps = int(steps)
    s.delta = float(delta); s.omega = float(omega)
    s.patch_center_theta = float(patch_center_theta); s.patch_center_phi = float(patch_center_phi)
    s.patch_size_theta = float(patch_size_theta); s.patch_size_phi = float(patch_size_phi)
    s.flip_theta = int(bool(flip_theta)); s.flip_phi = int(bool(flip_phi)); s.has_texture = int(bool(has_texture))
    return s


def integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega, nthreads=0, want_rmin=False):
    """cuda_geodesic.py:639-645.  Returns (out_qs (n,4), nsteps (n,)) [+ rmin (n,) closest approach]."""
    q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
    n = q0s.shape[0]
    out = np.zeros((n, 4)); ns = np.zeros(n, np.int64)
    rmin = np.zeros(n) if want_rmin else None
    lib().orc_integrate_batch(_dp(q0s), _dp(p0s), n, int(steps), float(delta), float(rs), float(r_max),
                              float(omega), _dp(out), _ptr(ns, C.c_int64), _dp(rmin), int(nthreads))
    return (out, ns, rmin) if want_rmin else (out, ns)


def conditioning(q0s, p0s, steps, delta, rs, r_max, omega, eps=2.2e-16):
    """Conditioning probe of the REFERENCE arithmetic: largest relative change of the final state when one
    component of p0 is perturbed by +-eps relative (about 1 ulp).  inf where the step count changes."""
    p0s = np.ascontiguousarray(p0s, np.float64)
    a, na = integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega)
    worst = np.zeros(len(a))
    for sgn in (1.0, -1.0):
        for comp in range(4):
            p = p0s.copy(); p[:, comp] *= (1.0 + sgn * eps)
            b, nb = integrate_batch(q0s, p, steps, delta, rs, r_max, omega)
            with np.errstate(invalid="ignore"):
                d = np.abs(a - b) / np.maximum(np.abs(a), 1.0)
            worst = np.maximum(worst, np.nan_to_num(d.max(1), nan=np.inf))
            worst[na != nb] = np.inf
    return worst


def integrate_batch_full(q0s, p0s, steps, delta, rs, r_max, omega, with_momenta=False):
    """cuda_geodesic.py:647-653.  Returns traj (n,steps,4) [, mom (n,steps,4)], rows (n,)."""
    q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
    n = q0s.shape[0]
    traj = np.zeros((n, int(steps), 4)); rows = np.zeros(n, np.int64)
    mom = np.zeros((n, int(steps), 4)) if with_momenta else None
    lib().orc_integrate_batch_full(_dp(q0s), _dp(p0s), n, int(steps), float(delta), float(rs), float(r_max),
                                   float(omega), _dp(traj), _dp(mom), _ptr(rows, C.c_int64))
    return (traj, mom, rows) if with_momenta else (traj, rows)


def camera_pixels(scene, ij):
    ij = np.ascontiguousarray(ij, np.int32).reshape(-1, 2)
    px = np.zeros((ij.shape[0], 3))
    lib().orc_camera_pixels(C.byref(scene), _ptr(ij, C.c_int32), ij.shape[0], _dp(px))
    return px


def initial_conditions(obs, pixels, mass):
    """utils.py:91-141.  Returns q0s (n,4), ic9 (n,9) = p0[4], alpha0, beta, h_r, h_theta, h_phi; rc."""
    obs = np.ascontiguousarray(obs, np.float64); pixels = np.ascontiguousarray(pixels, np.float64).reshape(-1, 3)
    n = pixels.shape[0]
    q0s = np.zeros((n, 4)); ic = np.zeros((n, 9))
    rc = lib().orc_initial_conditions(_dp(obs), _dp(pixels), n, float(mass), _dp(q0s), _dp(ic))
    return q0s, ic, rc


def classify(scene, final_q, alpha0, beta, tex):
    final_q = np.ascontiguousarray(final_q, np.float64).reshape(-1, 4)
    n = final_q.shape[0]
    alpha0 = np.ascontiguousarray(alpha0, np.float64); beta = np.ascontiguousarray(beta, np.float64)
    tex = np.ascontiguousarray(tex, np.uint8) if tex is not None else None
    rgb = np.zeros((n, 3), np.uint8); cls = np.zeros(n, np.uint8); hit = np.zeros((n, 2)); uv = np.zeros((n, 2), np.int32)
    lib().orc_classify(C.byref(scene), _dp(final_q), _dp(alpha0), _dp(beta), n, _ptr(tex, C.c_uint8),
                       _ptr(rgb, C.c_uint8), _ptr(cls, C.c_uint8), _dp(hit), _ptr(uv, C.c_int32))
    return dict(rgb=rgb, cls=cls, hit=hit, uv=uv)


def render(scene, tex, row_begin=0, row_end=None, row_stride=1, nthreads=0, extras=True):
    h, w = scene.height, scene.width
    row_end = h if row_end is None else row_end
    nrows = len(range(row_begin, row_end, row_stride))
    n = nrows * w
    tex = np.ascontiguousarray(tex, np.uint8) if tex is not None else None
    rgb = np.zeros((nrows, w, 3), np.uint8); cls = np.zeros((nrows, w), np.uint8)
    fq = np.zeros((n, 4)) if extras else None
    ic = np.zeros((n, 9)) if extras else None
    hit = np.zeros((n, 2)) if extras else None
    uv = np.zeros((n, 2), np.int32) if extras else None
    ns = np.zeros(n, np.int64)
    rc = lib().orc_render(C.byref(scene), _ptr(tex, C.c_uint8), row_begin, row_end, row_stride,
                          _ptr(rgb, C.c_uint8), _ptr(cls, C.c_uint8), _dp(fq), _dp(ic), _dp(hit),
                          _ptr(uv, C.c_int32), _ptr(ns, C.c_int64), int(nthreads))
    return dict(rgb=rgb, cls=cls, final_q=fq, ic=ic, hit=hit, uv=uv, nsteps=ns, rc=rc)


def flat_render(scene, tex):
    """background.py:99-132 (patch centre already resolved by the caller, :54-58)."""
    tex = np.ascontiguousarray(tex, np.uint8)
    rgb = np.zeros((scene.height, scene.width, 3), np.uint8)
    lib().orc_flat_render(C.byref(scene), _ptr(tex, C.c_uint8), _ptr(rgb, C.c_uint8))
    return rgb


# ---- binary128 truth (oracle/schw_oracle_q.c) ---------------------------------------------------------
_SOQ = os.path.join(_HERE, "_build", "libschw_oracle_q.so")
_libq = None


def libq():
    global _libq
    if _libq is None:
        src = os.path.join(_HERE, "schw_oracle_q.c")
        if not os.path.exists(_SOQ) or os.path.getmtime(_SOQ) < os.path.getmtime(src):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(_SOQ)
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orcq_integrate_batch.argtypes = [dp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                           ip, dp, dp, ip, C.c_int]
        L.orcq_integrate_batch_full.argtypes = [dp, dp, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_double,
                                                C.c_int, dp, dp, ip, C.c_int]
        _libq = L
    return _libq


def exact_integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega, force_steps=None, nthreads=0, with_momenta=False):
    """The stepper in binary128 on the same double inputs, rounded to double on output.  Returns (out_qs, nsteps)
    [, out_ps].  force_steps: per-ray step counts to take regardless of the termination test."""
    q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
    n = q0s.shape[0]
    out = np.zeros((n, 4)); ns = np.zeros(n, np.int64)
    ps = np.zeros((n, 4)) if with_momenta else None
    fs = None if force_steps is None else np.ascontiguousarray(force_steps, np.int64)
    libq().orcq_integrate_batch(_dp(q0s), _dp(p0s), n, int(steps), float(delta), float(rs), float(r_max), float(omega),
                                _ptr(fs, C.c_int64), _dp(out), _dp(ps), _ptr(ns, C.c_int64), int(nthreads))
    return (out, ns, ps) if with_momenta else (out, ns)


def exact_integrate_batch_full(q0s, p0s, steps, delta, rs, r_max, omega, stop=True, nthreads=0, with_momenta=False):
    """Trajectories (n, steps, 4) in binary128 (rounded to double); stop=False continues past the termination test."""
    q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
    n = q0s.shape[0]
    traj = np.zeros((n, int(steps), 4)); rows = np.zeros(n, np.int64)
    mom = np.zeros((n, int(steps), 4)) if with_momenta else None
    libq().orcq_integrate_batch_full(_dp(q0s), _dp(p0s), n, int(steps), float(delta), float(rs), float(r_max), float(omega),
                                     int(bool(stop)), _dp(traj), _dp(mom), _ptr(rows, C.c_int64), int(nthreads))
    return (traj, mom, rows) if with_momenta else (traj, rows)


def ulp_ensemble(q0s, p0s, steps, delta, rs, r_max, omega, force_steps=None, eps=2.2e-16):
    """FP64 evaluations of the reference arithmetic on the inputs and on +-1-ulp (relative eps) perturbations of each
    component of p0: list of (out_qs, nsteps), unperturbed first.  With force_steps the termination test is replaced
    by the given per-ray step counts (integrate_batch_full semantics are not needed: final states only)."""
    p0s = np.ascontiguousarray(p0s, np.float64)
    runs = [integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega)]
    for sgn in (1.0, -1.0):
        for comp in range(4):
            p = p0s.copy(); p[:, comp] *= (1.0 + sgn * eps)
            runs.append(integrate_batch(q0s, p, steps, delta, rs, r_max, omega))
    return runs
