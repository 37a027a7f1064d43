"""ctypes binding of libschw_b200.so (C ABI: include/schw_b200.h).

This is the only way the Python host side reaches the GPU.  There is no CPU fallback:
if the shared library is missing or no sm_100 device is present the calls raise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# SCHW_B200_LIB selects a tuning variant of the same library (tools/build_variants.sh); default is the in-tree build
LIB_PATH = os.environ.get("SCHW_B200_LIB") or os.path.join(_PKG, "lib", "libschw_b200.so")
CSRC = os.path.join(_PKG, "csrc")

CLASSES = ("bh", "numerical error", "escape_bg", "escape_no_patch", "in_domain")

SCHW_OK, SCHW_E_INVALID, SCHW_E_CUDA, SCHW_E_HORIZON, SCHW_E_NULLCOND, SCHW_E_RAYPLANE, SCHW_E_NOMEM = 0, -1, -2, -3, -4, -5, -6

# every symbol include/schw_b200.h declares (tests/test_abi_symbols.py checks header <-> library <-> this list)
EXPORTS = (
    "schw_abi_version", "schw_last_error", "schw_create", "schw_destroy", "schw_device_info",
    "schw_last_kernel_ms", "schw_integrate_batch", "schw_integrate_batch_dev", "schw_integrate_batch_full", "schw_integrate_batch_rows",
    "schw_initial_conditions", "schw_initial_conditions_px", "schw_classify", "schw_render", "schw_render_dev", "schw_last_total_steps",
    "schw_flat_render", "schw_fp64_peak_probe", "schw_math_probe", "schw_last_counters",
    "schw_host_alloc", "schw_host_free", "schw_set_tex_ready_flag", "schw_debug_step_params",
    "schw_render_dedup", "schw_render_dedup_dev", "schw_last_dedup_stats",
    "schw_render_dev_pitched", "schw_last_status", "schw_flat_raytrace",
    "schw_dedup_plan_dev", "schw_dedup_integrate_dev", "schw_dedup_epilogue_dev", "schw_observer_q0",
    "schw_integrate_batch_full_dev", "schw_step_probe",
)


class SchwScene(C.Structure):
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


class SchwRayOutputs(C.Structure):
    _fields_ = [("final_q", C.c_void_p), ("ic", C.c_void_p), ("hit", C.c_void_p),
                ("texel", C.c_void_p), ("nsteps", C.c_void_p)]


class SchwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libschw_b200 error {code}: {msg}")
        self.code = code


def build(force=False):
    """Compile libschw_b200.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in ("schw_kernels.cu", "schw_device.cuh")]
    srcs.append(os.path.join(os.path.dirname(_PKG), "include", "schw_b200.h"))
    stale = (not os.path.exists(LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def load():
    """dlopen the library and declare prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SchwError(SCHW_E_CUDA, f"{LIB_PATH} not found — run __graft_entry__.build() (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, dp, i64, i32, dbl = C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_double
    scp = C.POINTER(SchwScene)
    xop = C.POINTER(SchwRayOutputs)
    L.schw_abi_version.restype = C.c_int
    L.schw_last_error.restype = C.c_char_p
    L.schw_create.argtypes = [i32, C.POINTER(vp)]
    L.schw_destroy.argtypes = [vp]
    L.schw_device_info.argtypes = [vp, C.c_char_p, i32, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.schw_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]
    L.schw_last_total_steps.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.schw_last_counters.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.schw_integrate_batch.argtypes = [vp, dp, dp, i64, i64, dbl, dbl, dbl, dbl, dp, vp]
    L.schw_integrate_batch_dev.argtypes = [vp, dp, dp, i64, i64, dbl, dbl, dbl, dbl, dp, vp, vp]
    L.schw_integrate_batch_full.argtypes = [vp, dp, dp, i64, i64, dbl, dbl, dbl, dbl, dp, vp]
    L.schw_integrate_batch_full_dev.argtypes = [vp, dp, dp, i64, i64, dbl, dbl, dbl, dbl, dp, vp, i32, vp]
    L.schw_integrate_batch_rows.argtypes = [vp, dp, dp, i64, i64, dbl, dbl, dbl, dbl, vp, i64, dp, vp]
    L.schw_initial_conditions.argtypes = [vp, scp, i32, i32, i32, dp, dp]
    L.schw_initial_conditions_px.argtypes = [vp, dp, dp, i64, dbl, dp, dp]
    L.schw_classify.argtypes = [vp, scp, dp, dp, dp, i64, vp, vp, vp, dp, vp]
    L.schw_render.argtypes = [vp, scp, vp, i32, i32, i32, vp, vp, xop]
    L.schw_render_dev.argtypes = [vp, scp, vp, i32, i32, i32, vp, vp, xop, vp]
    L.schw_render_dev_pitched.argtypes = [vp, scp, vp, i32, i32, i32, vp, i64, vp, i64, xop, vp]
    L.schw_last_status.argtypes = [vp]
    L.schw_flat_raytrace.argtypes = [vp, dp, dp, i64, dbl, dbl, dbl, dbl, dbl, i32, i32, vp, i32, i32, vp, vp, i64, i64, dp]
    L.schw_observer_q0.argtypes = [scp, dp]
    L.schw_step_probe.argtypes = [vp, i32, i32, dbl, C.POINTER(C.c_float), C.POINTER(C.c_uint64), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_double)]
    L.schw_dedup_plan_dev.argtypes = [vp, scp, i32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), vp]
    L.schw_dedup_integrate_dev.argtypes = [vp, scp, i32, vp, vp, vp]
    L.schw_dedup_epilogue_dev.argtypes = [vp, scp, vp, i32, i32, i32, vp, i64, vp, i64, xop, vp, vp, vp]
    L.schw_render_dedup.argtypes = [vp, scp, vp, i32, i32, i32, vp, vp, xop]
    L.schw_render_dedup_dev.argtypes = [vp, scp, vp, i32, i32, i32, vp, vp, xop, vp]
    L.schw_last_dedup_stats.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.schw_flat_render.argtypes = [vp, scp, vp, vp]
    L.schw_math_probe.argtypes = [vp, dp, i64, dp, dp, dp]
    L.schw_fp64_peak_probe.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_float)]
    L.schw_set_tex_ready_flag.argtypes = [vp, vp]
    L.schw_debug_step_params.argtypes = [i64, dbl, dbl, dbl, dbl, vp, C.c_size_t, C.POINTER(C.c_size_t)]
    L.schw_host_alloc.argtypes = [C.c_size_t, C.POINTER(vp)]
    L.schw_host_free.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("schw_last_error",):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _check(rc):
    if rc != 0:
        msg = load().schw_last_error().decode("utf-8", "replace")
        if rc == SCHW_E_HORIZON:
            raise ValueError(msg)                 # simulation/utils.py:67-68
        if rc == SCHW_E_NULLCOND:
            raise RuntimeError(msg)               # simulation/utils.py:80-81
        if rc == SCHW_E_RAYPLANE:
            raise AssertionError(msg)             # simulation/utils.py:113,126
        if rc == SCHW_E_INVALID:
            raise ValueError(msg)
        raise SchwError(rc, msg)


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


# ---- page-locked frame buffers -----------------------------------------------------------------------
# render() returns fresh arrays like the reference's np.zeros (raytracing.py:50), but backed by page-locked memory so
# the D2H of a frame runs at PCIe rate (a 4K frame: 33 MB in ~0.6 ms instead of ~4 ms into pageable memory).  Blocks go
# back to a small pool when the array is garbage-collected.
_PINNED_POOL = {}          # nbytes -> [address, ...]
_PINNED_POOL_MAX = 4       # blocks kept per size


def _host_alloc(nbytes):
    p = C.c_void_p()
    _check(load().schw_host_alloc(nbytes, C.byref(p)))
    return p.value


def _host_free(addr):
    load().schw_host_free(C.c_void_p(addr))


class _PinnedBlock:
    """Owner of one page-locked block: taken from the pool (or allocated) on creation, handed back on collection."""

    def __init__(self, nbytes):
        self.nbytes = nbytes
        free = _PINNED_POOL.get(nbytes)
        self.addr = free.pop() if free else _host_alloc(nbytes)

    def __del__(self):
        try:
            free = _PINNED_POOL.setdefault(self.nbytes, [])
            if len(free) < _PINNED_POOL_MAX:
                free.append(self.addr)
            else:
                _host_free(self.addr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """Uninitialised page-locked ndarray (every element is overwritten by the D2H copy that follows).  A plain
    np.ndarray: its buffer object keeps the block alive, so the block returns to the pool when the last view dies."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes == 0:
        return np.empty(shape, dtype)
    blk = _PinnedBlock(nbytes)
    buf = (C.c_uint8 * nbytes).from_address(blk.addr)
    buf._owner = blk
    return np.frombuffer(buf, dtype=dtype).reshape(shape)


def make_scene(obs=(30.0, 0.0, 0.0), fov=np.radians(80.0), height=400, width=400, mass=1.0, r_max=31.0,
               steps=200000, delta=0.01, omega=1.0, patch_center_theta=np.pi / 2, patch_center_phi=np.pi,
               patch_size_theta=np.pi, patch_size_phi=2 * np.pi, flip_theta=False, flip_phi=False,
               has_texture=True):
    s = SchwScene()
    s.obs[0], s.obs[1], s.obs[2] = [float(v) for v in obs]
    s.fov = float(fov); s.height = int(height); s.width = int(width)
    s.mass = float(mass); s.r_max = float(r_max); s.steps = int(steps)
    s.delta = float(delta); s.omega = float(omega)
    s.patch_center_theta = float(patch_center_theta); s.patch_center_phi = float(patch_center_phi)
    s.patch_size_theta = float(patch_size_theta); s.patch_size_phi = float(patch_size_phi)
    s.flip_theta = int(bool(flip_theta)); s.flip_phi = int(bool(flip_phi)); s.has_texture = int(bool(has_texture))
    return s


def observer_q0(scene):
    """q0 = (0, r_obs, theta_obs, phi_obs) of the scene's camera rays (host arithmetic, no device)."""
    q0 = np.zeros(4)
    _check(load().schw_observer_q0(C.byref(scene), _p(q0)))
    return q0


class Handle:
    """One CUDA stream + scratch on one device (schw_create / schw_destroy)."""

    def __init__(self, device=-1):
        self._L = load()
        self._h = C.c_void_p()
        _check(self._L.schw_create(int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.schw_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- info / measurement -------------------------------------------------------------------
    def device_info(self):
        name = C.create_string_buffer(256)
        sm, khz = C.c_int(), C.c_int()
        _check(self._L.schw_device_info(self._h, name, 256, C.byref(sm), C.byref(khz)))
        return dict(name=name.value.decode(), sm_count=sm.value, clock_khz=khz.value)

    def last_kernel_ms(self):
        ms, n = C.c_float(), C.c_int()
        _check(self._L.schw_last_kernel_ms(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def last_total_steps(self):
        v = C.c_uint64()
        _check(self._L.schw_last_total_steps(self._h, C.byref(v)))
        return int(v.value)

    def last_counters(self):
        """(total_steps, warp_iters, planar_iters) of the last render / integrate_batch call."""
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _check(self._L.schw_last_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return int(a.value), int(b.value), int(c.value)

    def fp64_peak_probe(self):
        t, ms = C.c_double(), C.c_float()
        _check(self._L.schw_fp64_peak_probe(self._h, C.byref(t), C.byref(ms)))
        return t.value, ms.value

    def step_probe(self, variant, iters=20000, delta=0.01):
        """The render kernel's loop body alone: dict(ms, warp_steps, ctas_per_sm, regs, rays_per_thread)."""
        ms, ws, occ, regs, pf = C.c_float(), C.c_uint64(), C.c_int(), C.c_int(), C.c_double()
        _check(self._L.schw_step_probe(self._h, int(variant), int(iters), float(delta), C.byref(ms), C.byref(ws), C.byref(occ),
                                       C.byref(regs), C.byref(pf)))
        return dict(ms=ms.value, warp_steps=int(ws.value), ctas_per_sm=occ.value, regs=regs.value, rays_per_thread=variant // 100,
                    planar_fraction=pf.value)

    def math_probe(self, x):
        x = np.ascontiguousarray(x, np.float64)
        r, s, c = np.zeros_like(x), np.zeros_like(x), np.zeros_like(x)
        _check(self._L.schw_math_probe(self._h, _p(x), x.size, _p(r), _p(s), _p(c)))
        return r, s, c

    # -- integrator -----------------------------------------------------------------------------
    def integrate_batch(self, q0s, p0s, steps, delta, rs, r_max, omega, want_steps=False):
        q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
        if q0s.ndim != 2 or q0s.shape[1] != 4 or p0s.shape != q0s.shape:
            raise ValueError("q0s and p0s must both have shape (n, 4)")
        n = q0s.shape[0]
        out = np.zeros((n, 4), np.float64)
        ns = np.zeros(n, np.int32) if want_steps else None
        _check(self._L.schw_integrate_batch(self._h, _p(q0s), _p(p0s), n, int(steps), float(delta), float(rs),
                                            float(r_max), float(omega), _p(out), _p(ns)))
        return (out, ns) if want_steps else out

    def integrate_batch_dev(self, d_q0s, d_p0s, n, steps, delta, rs, r_max, omega, d_out, d_steps=0, stream=0):
        _check(self._L.schw_integrate_batch_dev(self._h, C.c_void_p(d_q0s), C.c_void_p(d_p0s), int(n), int(steps),
                                                float(delta), float(rs), float(r_max), float(omega),
                                                C.c_void_p(d_out), C.c_void_p(d_steps or None),
                                                C.c_void_p(stream or None)))

    def integrate_batch_full(self, q0s, p0s, steps, delta, rs, r_max, omega, want_rows=False):
        q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
        if q0s.ndim != 2 or q0s.shape[1] != 4 or p0s.shape != q0s.shape:
            raise ValueError("q0s and p0s must both have shape (n, 4)")
        n = q0s.shape[0]
        traj = np.zeros((n, int(steps), 4), np.float64)        # np.zeros like cuda_geodesic.py:649
        rows = np.zeros(n, np.int64)
        _check(self._L.schw_integrate_batch_full(self._h, _p(q0s), _p(p0s), n, int(steps), float(delta), float(rs),
                                                 float(r_max), float(omega), _p(traj), _p(rows)))
        return (traj, rows) if want_rows else traj

    def integrate_batch_full_dev(self, d_q0s, d_p0s, n, steps, delta, rs, r_max, omega, d_traj, d_rows=0, zero_fill=False, stream=0):
        """Device-pointer integrate_batch_full: (n, steps, 4) f64 in HBM; enqueue only."""
        _check(self._L.schw_integrate_batch_full_dev(self._h, C.c_void_p(d_q0s), C.c_void_p(d_p0s), int(n), int(steps),
                                                     float(delta), float(rs), float(r_max), float(omega), C.c_void_p(d_traj),
                                                     C.c_void_p(d_rows or None), int(bool(zero_fill)), C.c_void_p(stream or None)))

    def integrate_batch_rows(self, q0s, p0s, steps, delta, rs, r_max, omega, row_idx, want_rows=False):
        """Rows `row_idx` (strictly ascending step indices) of integrate_batch_full's result, (n, len(row_idx), 4)."""
        q0s = np.ascontiguousarray(q0s, np.float64); p0s = np.ascontiguousarray(p0s, np.float64)
        if q0s.ndim != 2 or q0s.shape[1] != 4 or p0s.shape != q0s.shape:
            raise ValueError("q0s and p0s must both have shape (n, 4)")
        row_idx = np.ascontiguousarray(row_idx, np.int32)
        n = q0s.shape[0]
        out = np.zeros((n, row_idx.size, 4), np.float64)
        rows = np.zeros(n, np.int64)
        _check(self._L.schw_integrate_batch_rows(self._h, _p(q0s), _p(p0s), n, int(steps), float(delta), float(rs),
                                                 float(r_max), float(omega), _p(row_idx), row_idx.size, _p(out), _p(rows)))
        return (out, rows) if want_rows else out

    # -- stages -----------------------------------------------------------------------------------
    def initial_conditions(self, scene, row_begin=0, row_end=None, row_stride=1):
        row_end = scene.height if row_end is None else row_end
        n = len(range(row_begin, row_end, row_stride)) * scene.width
        q0s = np.zeros((n, 4)); ic = np.zeros((n, 9))
        _check(self._L.schw_initial_conditions(self._h, C.byref(scene), row_begin, row_end, row_stride, _p(q0s), _p(ic)))
        return q0s, ic

    def initial_conditions_px(self, obs, pixels, mass):
        obs = np.ascontiguousarray(obs, np.float64); pixels = np.ascontiguousarray(pixels, np.float64).reshape(-1, 3)
        n = pixels.shape[0]
        q0s = np.zeros((n, 4)); ic = np.zeros((n, 9))
        _check(self._L.schw_initial_conditions_px(self._h, _p(obs), _p(pixels), n, float(mass), _p(q0s), _p(ic)))
        return q0s, ic

    def classify(self, scene, final_q, alpha0, beta, tex):
        final_q = np.ascontiguousarray(final_q, np.float64).reshape(-1, 4)
        alpha0 = np.ascontiguousarray(alpha0, np.float64); beta = np.ascontiguousarray(beta, np.float64)
        n = final_q.shape[0]
        tex = None if tex is None else np.ascontiguousarray(tex, np.uint8)
        rgb = np.zeros((n, 3), np.uint8); cls = np.zeros(n, np.uint8)
        hit = np.zeros((n, 2)); uv = np.zeros((n, 2), np.int32)
        _check(self._L.schw_classify(self._h, C.byref(scene), _p(final_q), _p(alpha0), _p(beta), n, _p(tex),
                                     _p(rgb), _p(cls), _p(hit), _p(uv)))
        return dict(rgb=rgb, cls=cls, hit=hit, uv=uv)

    # -- fused render -----------------------------------------------------------------------------
    def last_dedup_stats(self):
        """(distinct rays, steps summed over pixels) of the last render(..., dedup=True)."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._L.schw_last_dedup_stats(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def render(self, scene, tex, row_begin=0, row_end=None, row_stride=1, extras=False, dedup=False):
        """Host-buffer render (H2D of the texture and D2H of the outputs happen inside the call).
        dedup=True (opt-in): pixels with bitwise equal camera angle share one integration; same outputs."""
        row_end = scene.height if row_end is None else row_end
        nrows = len(range(row_begin, row_end, row_stride))
        w = scene.width
        n = nrows * w
        tex = None if tex is None else np.ascontiguousarray(tex, np.uint8)
        if tex is not None and tex.shape != (scene.height, scene.width, 3):
            raise ValueError("texture must be (height, width, 3) uint8 (the sky resized to the frame)")
        rgb = pinned_empty((nrows, w, 3), np.uint8); cls = pinned_empty((nrows, w), np.uint8)
        out = dict(rgb=rgb, cls=cls)
        x = None
        if extras:
            out.update(final_q=np.zeros((n, 4)), ic=np.zeros((n, 9)), hit=np.zeros((n, 2)),
                       uv=np.zeros((n, 2), np.int32), nsteps=np.zeros(n, np.int32))
            x = SchwRayOutputs(out["final_q"].ctypes.data, out["ic"].ctypes.data, out["hit"].ctypes.data,
                               out["uv"].ctypes.data, out["nsteps"].ctypes.data)
        fn = self._L.schw_render_dedup if dedup else self._L.schw_render
        _check(fn(self._h, C.byref(scene), _p(tex), row_begin, row_end, row_stride,
                  _p(rgb), _p(cls), C.byref(x) if x is not None else None))
        return out

    def render_dev(self, scene, d_tex, row_begin, row_end, row_stride, d_rgb, d_cls, stream=0,
                   d_final_q=0, d_ic=0, d_hit=0, d_texel=0, d_nsteps=0, d_tex_ready=0):
        """Device-pointer render: enqueue only (pointers are e.g. torch tensors' data_ptr()).
        d_tex_ready: optional device uint32 the kernel waits on before its first texel lookup (schw_set_tex_ready_flag)."""
        if d_tex_ready:
            _check(self._L.schw_set_tex_ready_flag(self._h, C.c_void_p(d_tex_ready)))
        x = SchwRayOutputs(d_final_q or None, d_ic or None, d_hit or None, d_texel or None, d_nsteps or None)
        _check(self._L.schw_render_dev(self._h, C.byref(scene), C.c_void_p(d_tex or None), row_begin, row_end,
                                       row_stride, C.c_void_p(d_rgb), C.c_void_p(d_cls), C.byref(x),
                                       C.c_void_p(stream or None)))

    def render_dev_pitched(self, scene, d_tex, row_begin, row_end, row_stride, d_rgb, rgb_pitch, d_cls, cls_pitch, stream=0,
                           d_final_q=0, d_ic=0, d_hit=0, d_texel=0, d_nsteps=0, d_tex_ready=0):
        """render_dev with explicit output row pitches in bytes (0 = packed), e.g. rgb and class side by side in the
        (rows, 4w) block of the multi-GPU gather."""
        if d_tex_ready:
            _check(self._L.schw_set_tex_ready_flag(self._h, C.c_void_p(d_tex_ready)))
        x = SchwRayOutputs(d_final_q or None, d_ic or None, d_hit or None, d_texel or None, d_nsteps or None)
        _check(self._L.schw_render_dev_pitched(self._h, C.byref(scene), C.c_void_p(d_tex or None), row_begin, row_end,
                                               row_stride, C.c_void_p(d_rgb), int(rgb_pitch), C.c_void_p(d_cls),
                                               int(cls_pitch), C.byref(x), C.c_void_p(stream or None)))

    def last_status(self):
        """Waits for the last enqueue-only call and raises what its device code flagged (see schw_last_status)."""
        _check(self._L.schw_last_status(self._h))

    def render_dedup_dev(self, scene, d_tex, row_begin, row_end, row_stride, d_rgb, d_cls, stream=0,
                         d_final_q=0, d_ic=0, d_hit=0, d_texel=0, d_nsteps=0):
        """Device-pointer de-duplicated render (opt-in).  Synchronises the stream once unless the
        geodesics of the previous call on this handle can be reused (same camera, rows and integrator parameters)."""
        x = SchwRayOutputs(d_final_q or None, d_ic or None, d_hit or None, d_texel or None, d_nsteps or None)
        _check(self._L.schw_render_dedup_dev(self._h, C.byref(scene), C.c_void_p(d_tex or None), row_begin, row_end,
                                             row_stride, C.c_void_p(d_rgb), C.c_void_p(d_cls), C.byref(x),
                                             C.c_void_p(stream or None)))

    # -- de-duplicated render, stage by stage (the ranks of a multi-GPU job share the integration) ----------
    def dedup_plan_dev(self, scene, world, stream=0):
        """Distinct camera rays of the WHOLE frame; returns (n_unique, chunk).  Synchronises the stream once."""
        a, b = C.c_uint64(), C.c_uint64()
        _check(self._L.schw_dedup_plan_dev(self._h, C.byref(scene), int(world), C.byref(a), C.byref(b),
                                           C.c_void_p(stream or None)))
        return int(a.value), int(b.value)

    def dedup_integrate_dev(self, scene, rank, d_fq_all, d_ns_all, stream=0):
        _check(self._L.schw_dedup_integrate_dev(self._h, C.byref(scene), int(rank), C.c_void_p(d_fq_all),
                                                C.c_void_p(d_ns_all), C.c_void_p(stream or None)))

    def dedup_epilogue_dev(self, scene, d_tex, row_begin, row_end, row_stride, d_rgb, rgb_pitch, d_cls, cls_pitch,
                           d_fq_all, d_ns_all, stream=0, d_final_q=0, d_ic=0, d_hit=0, d_texel=0, d_nsteps=0):
        x = SchwRayOutputs(d_final_q or None, d_ic or None, d_hit or None, d_texel or None, d_nsteps or None)
        _check(self._L.schw_dedup_epilogue_dev(self._h, C.byref(scene), C.c_void_p(d_tex or None), row_begin, row_end,
                                               row_stride, C.c_void_p(d_rgb), int(rgb_pitch), C.c_void_p(d_cls),
                                               int(cls_pitch), C.byref(x), C.c_void_p(d_fq_all), C.c_void_p(d_ns_all),
                                               C.c_void_p(stream or None)))

    def flat_render(self, scene, tex):
        tex = np.ascontiguousarray(tex, np.uint8)
        rgb = np.zeros((scene.height, scene.width, 3), np.uint8)
        _check(self._L.schw_flat_render(self._h, C.byref(scene), _p(tex), _p(rgb)))
        return rgb


    def flat_raytrace(self, obs_pos, ray_dirs, boundary_radius, patch_center_theta, patch_center_phi, patch_size_theta,
                      patch_size_phi, flip_theta, flip_phi, bg_array, sampled_indices=None, n_points=100):
        """cuda_flat_raytrace[_with_trajectories] (simulation/cuda_geodesic.py:214-292) on caller-supplied directions."""
        obs = np.ascontiguousarray(obs_pos, np.float64); dirs = np.ascontiguousarray(ray_dirs, np.float64).reshape(-1, 3)
        bg = np.ascontiguousarray(bg_array, np.uint8)
        n = dirs.shape[0]
        out = np.zeros((n, 3), np.uint8)
        if sampled_indices is None:
            _check(self._L.schw_flat_raytrace(self._h, _p(obs), _p(dirs), n, float(boundary_radius), float(patch_center_theta),
                                              float(patch_center_phi), float(patch_size_theta), float(patch_size_phi),
                                              int(bool(flip_theta)), int(bool(flip_phi)), _p(bg), bg.shape[0], bg.shape[1],
                                              _p(out), None, 0, 0, None))
            return out
        idx = np.ascontiguousarray(sampled_indices, np.int32)
        trajs = np.zeros((idx.size, int(n_points), 3), np.float64)
        _check(self._L.schw_flat_raytrace(self._h, _p(obs), _p(dirs), n, float(boundary_radius), float(patch_center_theta),
                                          float(patch_center_phi), float(patch_size_theta), float(patch_size_phi),
                                          int(bool(flip_theta)), int(bool(flip_phi)), _p(bg), bg.shape[0], bg.shape[1],
                                          _p(out), _p(idx), idx.size, int(n_points), _p(trajs)))
        return out, trajs


_default = {}


def default_handle(device=-1):
    """Process-wide handle per device (created on first use)."""
    if device not in _default:
        _default[device] = Handle(device)
    return _default[device]
