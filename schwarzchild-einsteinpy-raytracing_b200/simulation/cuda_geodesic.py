"""GPU integrator with the reference's class name and signatures (simulation/cuda_geodesic.py:629-653).

`CUDASchwarzschildIntegrator` is the drop-in boundary of the hot path: host NumPy (n,4) float64 in,
fresh host NumPy arrays out.  The work is done by the sm_100a kernels in csrc/ through the C ABI
(include/schw_b200.h); there is no Numba and no CPU fallback.
"""
import numpy as np

from schw_b200 import default_handle


class CUDASchwarzschildIntegrator:
    """FANTASY order-2 symplectic integrator for null geodesics in Schwarzschild (cuda_geodesic.py:629-637)."""

    def __init__(self, steps=500, delta=0.2, mass=1.0, omega=1.0, r_max=1e6, device=-1):
        self.steps = steps
        self.delta = delta
        self.rs = 2.0 * mass
        self.omega = omega
        self.r_max = r_max
        self._device = device

    def integrate_batch(self, q0s, p0s):
        """Final copy-1 positions (n,4) of every ray; second return value is None (cuda_geodesic.py:639-645)."""
        out = default_handle(self._device).integrate_batch(q0s, p0s, self.steps, self.delta, self.rs, self.r_max,
                                                           self.omega)
        return out, None

    def integrate_batch_full(self, q0s, p0s):
        """(n, steps, 4) trajectory of copy-1 positions, zero rows after termination (cuda_geodesic.py:647-653)."""
        return default_handle(self._device).integrate_batch_full(q0s, p0s, self.steps, self.delta, self.rs,
                                                                 self.r_max, self.omega)


def _flat_scene(obs_pos, h, w, fov, boundary_radius, pct, pcp, pst, psp, flip_theta, flip_phi):
    from schw_b200 import make_scene
    return make_scene(obs=obs_pos, fov=fov, height=h, width=w, r_max=boundary_radius, patch_center_theta=pct,
                      patch_center_phi=pcp, patch_size_theta=pst, patch_size_phi=psp, flip_theta=flip_theta,
                      flip_phi=flip_phi)
