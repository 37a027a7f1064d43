"""Command-line flags of the render entry point — the reference's whole configuration system
(reference: config.py:4-33).  Every flag name, type and default is the reference's; defaults given as
ints there (e.g. --boundary-radius 31, --observer-distance 30) stay ints here, because argparse does
not run `type` on non-string defaults and the reference's arithmetic sees them as such.
New flags are additive and default to the reference's behaviour.
"""
import argparse

# (flag, kwargs) in the reference's order
_FLAGS = [
    ("--size", dict(type=int, default=200, help="Image size (NxN)")),
    ("--fov", dict(type=float, default=80.0, help="Field of view in degrees")),
    ("--background", dict(type=str, default="images/backgrounds/milky-way-equirec.jpg", help="Background image path")),
    ("--steps", dict(type=int, default=200000, help="Step cap of each geodesic")),
    ("--delta", dict(type=float, default=0.01, help="Affine step size")),
    ("--omega", dict(type=float, default=0.01, help="Hamiltonian flow coupling omega (accepted; the render path uses 1.0 like the reference)")),
    ("--rtol", dict(type=float, default=1e-2, help="accepted and ignored, like the reference")),
    ("--atol", dict(type=float, default=1e-2, help="accepted and ignored, like the reference")),
    ("--order", dict(type=int, default=2, choices=[2, 4, 6, 8], help="accepted and ignored, like the reference (order 2 is integrated)")),
    ("--suppress-warnings", dict(action="store_true", help="accepted and ignored, like the reference")),
    ("--cuda", dict(action="store_true", default=True, help="GPU path (always on, like the reference)")),
    ("--bh-mass", dict(type=float, default=1.0, help="Black hole mass")),
    ("--boundary-radius", dict(type=float, default=31, help="Simulation boundary radius")),
    ("--observer-distance", dict(type=float, default=30, help="Observer distance from the black hole")),
    ("--bg-patch-center-theta", dict(type=float, default=90, help="Background patch centre theta (deg)")),
    ("--bg-patch-center-phi", dict(type=float, default=180, help="Background patch centre phi (deg)")),
    ("--bg-patch-center-theta-relobs", dict(type=float, default=0, help="Patch centre theta offset relative to the optical axis (deg)")),
    ("--bg-patch-center-phi-relobs", dict(type=float, default=0, help="Patch centre phi offset relative to the optical axis (deg)")),
    ("--bg-patch-size-theta", dict(type=float, default=180, help="Background patch size theta (deg)")),
    ("--bg-patch-size-phi", dict(type=float, default=360, help="Background patch size phi (deg)")),
    ("--bg-flip-theta", dict(action="store_true", help="Flip theta mapping for the background patch")),
    ("--bg-flip-phi", dict(action="store_true", help="Flip phi mapping for the background patch")),
    ("--no-flat-trajectories", dict(action="store_true", default=False, help="Skip the flat (no-gravity) comparison render")),
]

# additive flags (not in the reference); defaults reproduce the reference's behaviour
_EXTRA_FLAGS = [
    ("--width", dict(type=int, default=None, help="[additive] frame width (default: --size)")),
    ("--height", dict(type=int, default=None, help="[additive] frame height (default: --size)")),
    ("--no-plots", dict(action="store_true", help="[additive] skip the matplotlib diagnostics (also skipped when matplotlib is absent)")),
    ("--dedup", dict(action="store_true", help="[additive] integrate each distinct camera ray once (pixels whose camera angle agrees "
                                               "bitwise share a geodesic); the image and the CSVs are identical, ~8x faster")),
]


def build_parser():
    parser = argparse.ArgumentParser(description="Black Hole Ray Tracing Simulation")
    for flag, kw in _FLAGS + _EXTRA_FLAGS:
        parser.add_argument(flag, **kw)
    return parser


def parse_args(argv=None):
    return build_parser().parse_args(argv)
