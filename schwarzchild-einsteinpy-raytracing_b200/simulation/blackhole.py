"""Scene parameter holders (reference: simulation/blackhole.py:4-39).  Geometrized units, G = c = 1."""
import numpy as np


class BlackHole:
    """Schwarzschild black hole of the given mass at `position`; `rs` = 2 M (blackhole.py:10-13)."""

    def __init__(self, mass=1.0, position=np.zeros(3)):
        self.mass = mass
        self.position = np.array(position)
        self.rs = 2 * mass


class Observer:
    """Pin-hole camera: `position` 3-vector, `fov` in radians, `image_size` = (height, width) (blackhole.py:22-25)."""

    def __init__(self, position, fov, image_size):
        self.position = np.array(position)
        self.fov = fov
        self.image_size = image_size


class Photon:
    """Unused by the render path; kept for API compatibility (blackhole.py:27-39)."""

    def __init__(self, position, direction, mesh_idx):
        self.position = np.array(position)
        self.direction = np.array(direction)
        self.mesh_idx = mesh_idx
        self.collision = None
        self.collision_pos = None
