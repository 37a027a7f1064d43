"""Drop-in mirror of the reference's `simulation` package for the per-ray geodesic hot path.

Same module / function / class names and argument meaning as the reference
(simulation/{blackhole,utils,cuda_geodesic,raytracing,background}.py); the arithmetic runs in
hand-written sm_100a CUDA kernels behind the C ABI in include/schw_b200.h.
"""
