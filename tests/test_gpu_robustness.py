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
