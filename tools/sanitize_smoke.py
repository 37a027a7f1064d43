#!/usr/bin/env python
"""Small end-to-end exercise of every kernel, meant to run under compute-sanitizer."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "schwarzchild-einsteinpy-raytracing_b200"))
import schw_b200  # noqa: E402

hd = schw_b200.Handle(0)
rng = np.random.default_rng(0)
h, w = 37, 70                                     # ragged: width not a multiple of 32
tex = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
sc = schw_b200.make_scene(height=h, width=w, delta=0.1)
a = hd.render(sc, tex, extras=True)
b = hd.render(sc, tex, 1, h, 3, extras=True)
assert np.array_equal(b["rgb"], a["rgb"][1::3])
q0s, ic = hd.initial_conditions(sc)
out, ns = hd.integrate_batch(q0s, ic[:, :4].copy(), 5000, 0.1, 2.0, 31.0, 1.0, want_steps=True)
assert np.array_equal(ns, a["nsteps"])
traj, rows = hd.integrate_batch_full(q0s[:5], ic[:5, :4].copy(), 1500, 0.1, 2.0, 31.0, 1.0, want_rows=True)
c = hd.classify(sc, a["final_q"], ic[:, 4].copy(), ic[:, 5].copy(), tex)
assert np.array_equal(c["cls"], a["cls"].reshape(-1))
f = hd.flat_render(schw_b200.make_scene(height=h, width=w, patch_center_phi=3.0, patch_size_phi=2.0), tex)
hd.math_probe(rng.uniform(0.1, 3.0, 1000))
print("sanitize smoke ok", int(ns.sum()), rows.tolist(), int(f.sum()))
