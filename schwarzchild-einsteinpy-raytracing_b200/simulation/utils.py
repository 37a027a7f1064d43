"""Host-side helpers with the reference's names and meaning (reference: simulation/utils.py).

The per-pixel hot path does not go through these (run_manual_simulation fuses camera rays and
initial conditions into the render kernel); they exist so that scripts written against the
reference (single_ray_cuda_test.py, test-band-axis.py, tests/test_initial_conditions.py) keep
working.  `get_initial_conditions` evaluates on the GPU through the C ABI; the closed-form scalar
helpers below it are plain host arithmetic.
"""
import math

import numpy as np


# -- einsteinpy.coordinates.utils (un-vendored third-party helpers the reference re-exports through
#    `from simulation.utils import spherical_to_cartesian_fast, cartesian_to_spherical_fast`,
#    raytracing.py:13).  Restated from einsteinpy 0.4.0; see SURVEY.md App. B-0. -------------------
def cartesian_to_spherical_fast(t, x, y, z, v_x=None, v_y=None, v_z=None, velocities_provided=False):
    hxy = np.hypot(x, y)
    return t, np.hypot(hxy, z), np.arctan2(hxy, z), np.arctan2(y, x)


def spherical_to_cartesian_fast(t, r, th, p, v_r=None, v_th=None, v_p=None, velocities_provided=False):
    return t, r * np.cos(p) * np.sin(th), r * np.sin(p) * np.sin(th), r * np.cos(th)


def _apply_relative_offsets(theta_base_deg, phi_base_deg, dtheta_deg=0.0, dphi_deg=0.0):
    """Patch centre after observer-relative offsets, degrees in, radians out (utils.py:27-36)."""
    theta = np.clip(np.deg2rad(theta_base_deg) + np.deg2rad(dtheta_deg), 0.0, np.pi)
    phi = (np.deg2rad(phi_base_deg) + np.deg2rad(dphi_deg)) % (2 * np.pi)
    return theta, phi


def build_null_4momentum_ep_sph(p_sph, pos_sph, *, mass_bh=1.0, future=False):
    """Null 4-momentum (p_t, p_r, p_th, p_ph) from spatial components at `pos_sph` (utils.py:39-88).

    Raises ValueError when the position is inside the horizon and RuntimeError on a negative
    discriminant, exactly like the reference."""
    pr, pth, pph = p_sph
    r, th, _ = pos_sph
    f = 1.0 - 2.0 * mass_bh / r
    if f <= 0.0:
        raise ValueError("Observer must lie outside the event horizon (r > 2M).")
    A = -1.0 / f
    C = f * pr * pr + (1.0 / (r * r)) * pth * pth + (1.0 / (r * r * math.sin(th) ** 2)) * pph * pph
    disc = -4.0 * A * C
    if disc < 0.0:
        raise RuntimeError("Negative discriminant while enforcing null condition.")
    p_t = math.sqrt(disc) / (2.0 * (-A))
    return np.array([p_t if future else -p_t, pr, pth, pph])


def angles_to_p_sph(alpha, beta, r_obs, *, mass_bh=1.0, normalise=True):
    """Camera angles (alpha right, beta up) -> (p_r, p_th, p_ph) for an observer on +x (utils.py:148-194)."""
    n_r = -math.cos(alpha) * math.cos(beta)
    n_ph = math.sin(alpha) * math.cos(beta)
    n_th = -math.sin(beta)
    if not normalise:
        # the reference leaves p_r/p_th/p_ph undefined in this branch (utils.py:178-194 -> NameError)
        raise NameError("angles_to_p_sph(normalise=False) is undefined in the reference")
    return np.array([n_r * np.sqrt(1.0 - 2.0 * mass_bh / r_obs), n_th * r_obs, n_ph * r_obs])


def get_initial_conditions(observer_pos, pixel_pos, *, mass_bh=1.0):
    """(q0, p0, alpha0, h_r, h_theta, h_phi, beta) for one camera ray (utils.py:91-141), computed on the GPU."""
    from schw_b200 import default_handle
    q0s, ic = default_handle().initial_conditions_px(np.asarray(observer_pos, np.float64),
                                                     np.asarray(pixel_pos, np.float64).reshape(1, 3), mass_bh)
    return q0s[0], ic[0, 0:4].copy(), ic[0, 4], ic[0, 6], ic[0, 7], ic[0, 8], ic[0, 5]
