"""Error behaviour and boundary variants of the C ABI on the device: the reference's exceptions that are detected per ray
on the GPU (sticky error word), the upload flag that never arrives, the ray-count limit, pitched outputs."""
import numpy as np
import pytest

from helpers import synthetic_sky

pytestmark = pytest.mark.gpu


def test_ray_plane_assertion_like_the_reference(handle):
    """utils.py:113: `assert abs(ray_xy[2]) < 1e-6` -- fails on NaN, i.e. for a pixel position equal to the observer
    position (the normalisation divides by zero).  The device flags it, the host entry point raises AssertionError."""
    obs = np.array([30.0, 0.0, 0.0])
    good = np.array([[24.0, 0.3, -0.2], [24.0, -1.0, 2.0]])
    q0, ic = handle.initial_conditions_px(obs, good, 1.0)
    assert np.isfinite(ic).all()
    with pytest.raises(AssertionError):
        handle.initial_conditions_px(obs, np.array([[24.0, 0.3, -0.2], [30.0, 0.0, 0.0]]), 1.0)
    # the flag is per call, not sticky across calls
    q0b, icb = handle.initial_conditions_px(obs, good, 1.0)
    assert np.array_equal(icb, ic)
    from simulation.utils import get_initial_conditions
    with pytest.raises(AssertionError):
        get_initial_conditions(obs, obs.copy())


def test_too_many_rays_is_rejected_before_any_work(handle):
    """The kernels index rays with 32-bit integers: a call with more than 2^31-1 rays is refused (ValueError) before any
    buffer is sized or any kernel is launched."""
    import ctypes as C
    import schw_b200
    from schw_b200.native import _check
    sc = schw_b200.make_scene(height=65536, width=32768 + 1, has_texture=False)
    with pytest.raises(ValueError, match="2\\^31"):
        handle.render_dev(sc, 0, 0, sc.height, 1, 8, 8)
    with pytest.raises(ValueError, match="2\\^31"):
        _check(handle._L.schw_render(handle._h, C.byref(sc), None, 0, sc.height, 1, C.c_void_p(8), C.c_void_p(8), None))
    with pytest.raises(ValueError, match="2\\^31"):
        _check(handle._L.schw_render_dedup(handle._h, C.byref(sc), None, 0, sc.height, 1, C.c_void_p(8), C.c_void_p(8), None))


def test_pitched_outputs_equal_packed_outputs(handle):
    """schw_render_dev_pitched: rgb and class side by side in one (rows, 4w) block (the message of the multi-GPU gather)."""
    import torch
    import schw_b200
    h, w = 45, 70
    sc = schw_b200.make_scene(height=h, width=w, delta=0.05)
    tex = synthetic_sky(h, w, 2)
    ref = handle.render(sc, tex, 1, h, 2)
    rows = ref["cls"].shape[0]
    d_tex = torch.from_numpy(tex).cuda()
    blk = torch.full((rows, 4 * w + 8), 77, dtype=torch.uint8, device="cuda")          # pitch > 4w: the padding must stay untouched
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        handle.render_dev_pitched(sc, d_tex.data_ptr(), 1, h, 2, blk.data_ptr(), 4 * w + 8, blk.data_ptr() + 3 * w, 4 * w + 8,
                                  stream=st.cuda_stream)
    handle.last_status()
    got = blk.cpu().numpy()
    assert np.array_equal(got[:, :3 * w].reshape(rows, w, 3), ref["rgb"]) and np.array_equal(got[:, 3 * w:4 * w], ref["cls"])
    assert (got[:, 4 * w:] == 77).all()
    with pytest.raises(ValueError):
        handle.render_dev_pitched(sc, d_tex.data_ptr(), 1, h, 2, blk.data_ptr(), 3 * w - 1, blk.data_ptr() + 3 * w, 4 * w)


def test_texture_flag_that_never_arrives_is_an_error_not_a_wrong_frame(handle):
    """schw_set_tex_ready_flag: a warp waits ~2 s for the upload flag at its first texel lookup; if it never comes the
    call is reported as failed (SCHW_E_CUDA) by schw_last_status instead of returning stale texels with rc 0."""
    import torch
    import schw_b200
    h, w = 8, 64
    sc = schw_b200.make_scene(height=h, width=w, delta=0.2)
    d_tex = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    d_rgb = torch.zeros((h, w, 3), dtype=torch.uint8, device="cuda")
    d_cls = torch.zeros((h, w), dtype=torch.uint8, device="cuda")
    flag = torch.zeros(1, dtype=torch.int32, device="cuda")                             # never raised
    st = torch.cuda.Stream()
    handle.render_dev(sc, d_tex.data_ptr(), 0, h, 1, d_rgb.data_ptr(), d_cls.data_ptr(), stream=st.cuda_stream,
                      d_tex_ready=flag.data_ptr())
    with pytest.raises(schw_b200.SchwError, match="texture upload"):
        handle.last_status()
    # with the flag raised the same call is fine, and the error does not stick to the handle
    flag.fill_(1)
    torch.cuda.synchronize()
    handle.render_dev(sc, d_tex.data_ptr(), 0, h, 1, d_rgb.data_ptr(), d_cls.data_ptr(), stream=st.cuda_stream,
                      d_tex_ready=flag.data_ptr())
    handle.last_status()
    assert (d_cls.cpu().numpy() == handle.render(sc, np.zeros((h, w, 3), np.uint8))["cls"]).all()


def test_integrate_batch_full_dev_equals_the_host_entry_point(handle):
    """schw_integrate_batch_full_dev (trajectories staged in shared memory, flushed as full 128-byte lines) on a ragged
    batch: ray count not a multiple of the CTA, step count not a multiple of the 4-row group, rays ending at every
    position of a group, a NaN ray, rays that end before the first step -- against the host entry point (np.zeros rows
    after termination) and against integrate_batch (the state after k steps = row k)."""
    import torch
    from helpers import random_rays
    q0s, p0s = random_rays(301, seed=17)
    q0s[7, 1] = np.nan                       # never terminates in the reference: NaN rows up to the step cap
    q0s[11, 1] = 2.0                         # starts inside the capture radius: one row
    steps, delta, rs, r_max, omega = 1003, 0.05, 2.0, 41.0, 1.0
    want, rows_want = handle.integrate_batch_full(q0s, p0s, steps, delta, rs, r_max, omega, want_rows=True)
    dq, dp = torch.from_numpy(q0s).cuda(), torch.from_numpy(p0s).cuda()
    traj = torch.full((301, steps, 4), 7.0, dtype=torch.float64, device="cuda")
    rows = torch.zeros(301, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    handle.integrate_batch_full_dev(dq.data_ptr(), dp.data_ptr(), 301, steps, delta, rs, r_max, omega, traj.data_ptr(),
                                    rows.data_ptr(), zero_fill=True, stream=st.cuda_stream)
    handle.last_status()
    assert np.array_equal(rows.cpu().numpy(), rows_want)
    assert np.array_equal(traj.cpu().numpy(), want, equal_nan=True)
    assert rows_want[11] == 1 and rows_want[7] == steps and np.isnan(want[7, 5:]).all()
    assert len(set(int(r) % 4 for r in rows_want)) == 4           # every flush position of a 4-row group occurs
    # without the zero fill the rows past a ray's end are left as they were
    traj.fill_(7.0)
    handle.integrate_batch_full_dev(dq.data_ptr(), dp.data_ptr(), 301, steps, delta, rs, r_max, omega, traj.data_ptr(),
                                    rows.data_ptr(), stream=st.cuda_stream)
    handle.last_status()
    got = traj.cpu().numpy()
    for i in (0, 11, 100, 300):
        assert np.array_equal(got[i, :rows_want[i]], want[i, :rows_want[i]], equal_nan=True) and (got[i, rows_want[i]:] == 7.0).all()
    k = 57
    fin = handle.integrate_batch(q0s, p0s, k, delta, rs, r_max, omega)
    live = rows_want > k
    assert np.array_equal(fin[live], want[live, k], equal_nan=True)


def test_thresholds_a_few_ulps_apart(handle, orc):
    """r_max within the same high word as the capture radius: no radius passes the integer pre-test of the termination
    test, so lanes without a ray cannot be parked 'inside' -- every step then takes the exact FP64 test, which must
    ignore them.  Rays start inside the sliver, outside it, and far away; a mixed warp with empty lanes."""
    rs = 2.0
    rcap = 1.1 * rs
    r_max = rcap * (1 + 3e-9)
    q0s = np.zeros((45, 4)); p0s = np.zeros((45, 4))
    q0s[:, 1] = np.where(np.arange(45) % 3 == 0, rcap * (1 + 1e-9), np.where(np.arange(45) % 3 == 1, rcap * (1 - 1e-9), 30.0))
    q0s[:, 2] = np.pi / 2
    f = 1 - rs / q0s[:, 1]
    p0s[:, 1] = np.where(np.arange(45) % 2 == 0, 1e-9, -1e-9) * f; p0s[:, 3] = 1e-9 * q0s[:, 1]
    p0s[:, 0] = np.sqrt(np.abs(f * p0s[:, 1] ** 2 + p0s[:, 3] ** 2 / q0s[:, 1] ** 2) / np.abs(f)) * np.sign(f)
    for steps in (0, 1, 7, 400):
        out, ns = handle.integrate_batch(q0s, p0s, steps, 0.01, rs, r_max, 1.0, want_steps=True)
        ref, ns_ref = orc.integrate_batch(q0s, p0s, steps, 0.01, rs, r_max, 1.0)
        assert np.array_equal(ns, ns_ref), steps
        assert np.allclose(out, ref, rtol=1e-12, atol=1e-12, equal_nan=True), steps
    assert set(np.unique(ns)) >= {0, 400}


def test_trajectory_kernels_agree_bit_for_bit_on_camera_rays(handle, monkeypatch):
    """integrate_batch_full on planar (camera) rays: the two-rays-per-thread kernel (large batches), the one-ray kernel with the
    fast path (small batches), the default choice and the general kernel (SCHW_PLANAR=0) must store the same rows, bit for bit -- ragged batch (not a multiple
    of 2 x 128 rays), rays ending at every position of a 4-row group, step cap hit by some."""
    import schw_b200
    sc = schw_b200.make_scene(height=19, width=37, delta=0.05)
    q0s, ic = handle.initial_conditions(sc)
    p0s = ic[:, :4].copy()
    steps = 1101
    outs = []
    for env in (dict(SCHW_FULL_RAYS="2"), dict(SCHW_FULL_RAYS="1"), dict(), dict(SCHW_PLANAR="0")):
        for k in ("SCHW_FULL_RAYS", "SCHW_PLANAR"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        traj, rows = handle.integrate_batch_full(q0s, p0s, steps, 0.05, 2.0, 31.0, 1.0, want_rows=True)
        outs.append((traj, rows, handle.last_counters()))
    for traj, rows, ctr in outs[1:]:
        assert np.array_equal(rows, outs[0][1]) and np.array_equal(traj, outs[0][0], equal_nan=True)
    rows = outs[0][1]
    assert rows.min() < steps and rows.max() == steps and len(set(int(r) % 4 for r in rows)) == 4
    # fast path taken (partly: at this coarse step the bound on |p_theta| is small and rays are handed to the general flow) / never
    assert 0 < outs[0][2][2] <= outs[0][2][1] and outs[3][2][2] == 0
    fin = handle.integrate_batch(q0s, p0s, steps, 0.05, 2.0, 31.0, 1.0)
    ended = rows < steps
    assert np.array_equal(fin[ended], outs[0][0][ended, rows[ended] - 1])


def test_trajectory_kernel_choice_by_batch_size_changes_no_bit(handle, monkeypatch):
    """integrate_batch_full picks the two-rays-per-thread kernel once a planar batch fills its resident grid (sm_count x 4 CTAs
    x 256 rays) and the one-ray kernel below that; a batch just above the threshold must give the rows of the forced one-ray
    kernel, bit for bit, and the device counters must show which kernel ran (warp iterations are counted per ray slot)."""
    import schw_b200
    n = handle.device_info()["sm_count"] * 4 * 256 + 77            # ragged, just above the threshold
    sc = schw_b200.make_scene(height=64, width=96)
    q0s, ic = handle.initial_conditions(sc)
    pick = np.arange(n) % q0s.shape[0]
    q0 = np.ascontiguousarray(q0s[pick]); p0 = np.ascontiguousarray(ic[pick, :4])
    steps = 6
    outs = {}
    for mode in ("", "1", "2"):
        monkeypatch.delenv("SCHW_FULL_RAYS", raising=False)
        if mode:
            monkeypatch.setenv("SCHW_FULL_RAYS", mode)
        traj, rows = handle.integrate_batch_full(q0, p0, steps, sc.delta, 2.0 * sc.mass, sc.r_max, 1.0, want_rows=True)
        outs[mode] = (traj, rows, handle.last_counters())
    for mode in ("1", "2"):
        assert np.array_equal(outs[mode][0], outs[""][0]) and np.array_equal(outs[mode][1], outs[""][1])
    assert (outs[""][1] == steps).all()
    # the default took the same kernel as the forced two-ray one, not the one-ray one (padding of the last CTA differs)
    assert outs[""][2][1] == outs["2"][2][1] != outs["1"][2][1]
