"""Host-side Python binding of the B200 Schwarzschild hot path (libschw_b200.so)."""
from .native import (CLASSES, EXPORTS, Handle, SchwError, SchwRayOutputs, SchwScene, build, default_handle, load,
                     make_scene)

__all__ = ["CLASSES", "EXPORTS", "Handle", "SchwError", "SchwRayOutputs", "SchwScene", "build", "default_handle",
           "load", "make_scene"]
