"""Flat-space (no gravity) comparison render with the reference's entry point
(simulation/background.py:19-141) — SURVEY §8f row N1.  The per-pixel CPU loop (:99-132) is one
sm_100a kernel launch (schw_flat_render); file outputs and return value follow the reference.
"""
import os
import random

import numpy as np
from PIL import Image

from schw_b200 import default_handle, make_scene


def _in_phi_patch(phi, phi0, phi1):
    """True if phi lies in [phi0, phi1] on the circle (background.py:8-17)."""
    phi, phi0, phi1 = phi % (2 * np.pi), phi0 % (2 * np.pi), phi1 % (2 * np.pi)
    return (phi0 <= phi <= phi1) if phi0 <= phi1 else (phi >= phi0 or phi <= phi1)


def save_no_gravity_image_with_background(
    observer, bg_path, out_path, boundary_radius=None,
    patch_center_theta=None, patch_center_phi=None, patch_size_theta=np.deg2rad(10), patch_size_phi=np.deg2rad(10),
    flip_theta=False, flip_phi=False,
    return_sampled_trajectories=False, n_sampled=10,
    override_patch_center=False,
    cuda=False
):
    """Flat-space image of the sky through the same pin-hole camera; saves `out_path` and a copy of the
    sky as scene_full.png next to it; optionally returns `n_sampled` straight-line trajectories (100 points each).
    `cuda` is accepted for compatibility: the GPU kernel is always used."""
    h, w = observer.image_size
    obs_pos = np.array(observer.position, dtype=np.float64)
    if boundary_radius is None:
        boundary_radius = np.linalg.norm(obs_pos) * 2
    bg_img = Image.open(bg_path).convert('RGB')
    bg_array = np.array(bg_img.resize((w, h), Image.LANCZOS))
    if not override_patch_center or patch_center_theta is None or patch_center_phi is None:
        # patch centre forced to the antipode of the observer (background.py:54-58)
        opp = -obs_pos
        patch_center_theta = np.arccos(opp[2] / np.linalg.norm(opp))
        patch_center_phi = np.arctan2(opp[1], opp[0])
    scene = make_scene(obs=obs_pos, fov=float(observer.fov), height=h, width=w, r_max=float(boundary_radius),
                       patch_center_theta=patch_center_theta, patch_center_phi=patch_center_phi,
                       patch_size_theta=patch_size_theta, patch_size_phi=patch_size_phi,
                       flip_theta=flip_theta, flip_phi=flip_phi)
    out_img = default_handle().flat_render(scene, bg_array)

    sampled_trajectories = []
    if return_sampled_trajectories:
        picks = set()
        while len(picks) < n_sampled:
            picks.add((random.randint(0, h - 1), random.randint(0, w - 1)))
        # straight segments observer -> boundary hit for the sampled pixels (background.py:104-117), in scan order
        axis = -obs_pos / np.linalg.norm(obs_pos)
        up_guess = np.array([0.0, 0.0, 1.0])
        if np.allclose(np.cross(axis, up_guess), 0):
            up_guess = np.array([0.0, 1.0, 0.0])
        right = np.cross(up_guess, axis); right /= np.linalg.norm(right)
        up = np.cross(axis, right); up /= np.linalg.norm(up)
        plane_dist = 0.2 * np.linalg.norm(obs_pos)
        pc = obs_pos + axis * plane_dist
        pw = 2 * plane_dist * np.tan(observer.fov / 2); ph = pw * (h / w)
        for (i, j) in sorted(picks):
            d = pc + ((j + 0.5) / w - 0.5) * pw * right + ((i + 0.5) / h - 0.5) * ph * up - obs_pos
            d /= np.linalg.norm(d)
            b = 2 * np.dot(obs_pos, d); c = np.dot(obs_pos, obs_pos) - boundary_radius ** 2
            disc = b * b - 4 * c
            if disc < 0:
                continue
            sampled_trajectories.append(np.linspace(obs_pos, obs_pos + (-b + np.sqrt(disc)) / 2 * d, 100))

    os.makedirs(os.path.dirname(out_path) or '.', exist_ok=True)
    Image.fromarray(out_img).save(out_path)
    print(f"Saved no-gravity image to {out_path}")
    scene_out_path = os.path.join(os.path.dirname(out_path), 'scene_full.png')
    bg_img.save(scene_out_path)
    print(f"Saved full scene image to {scene_out_path}")
    if return_sampled_trajectories:
        return sampled_trajectories
