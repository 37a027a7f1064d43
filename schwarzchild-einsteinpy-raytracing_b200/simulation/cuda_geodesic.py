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


def cuda_flat_raytrace(obs_pos, ray_dirs, boundary_radius,
                       patch_center_theta, patch_center_phi,
                       patch_size_theta, patch_size_phi,
                       flip_theta, flip_phi,
                       bg_array):
    """Flat-space ray tracing of caller-supplied unit directions; returns the flattened (N,3) uint8 RGB array
    (cuda_geodesic.py:214-250, as imported by simulation/background.py:95)."""
    return default_handle().flat_raytrace(obs_pos, ray_dirs, boundary_radius, patch_center_theta, patch_center_phi,
                                          patch_size_theta, patch_size_phi, flip_theta, flip_phi, bg_array)


def cuda_flat_raytrace_with_trajectories(obs_pos, ray_dirs, boundary_radius,
                                         patch_center_theta, patch_center_phi,
                                         patch_size_theta, patch_size_phi,
                                         flip_theta, flip_phi,
                                         bg_array,
                                         sampled_indices, n_points=100):
    """Same, plus `n_points` samples along each ray listed in `sampled_indices` (flat indices): returns
    (image (N,3) uint8, trajectories (len(sampled_indices), n_points, 3) float64)  (cuda_geodesic.py:252-292)."""
    return default_handle().flat_raytrace(obs_pos, ray_dirs, boundary_radius, patch_center_theta, patch_center_phi,
                                          patch_size_theta, patch_size_phi, flip_theta, flip_phi, bg_array,
                                          sampled_indices=list(sampled_indices), n_points=n_points)
