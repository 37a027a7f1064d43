/*
 * schw_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, scalar FP64 restatement of the reference's per-ray geodesic hot
 * path (ryanchen2134/schwarzchild-einsteinpy-raytracing), operation for
 * operation and in the reference's evaluation order.  It is the checker for
 * the CUDA product in tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py.  Nothing in the product path may call,
 * link or import it.
 *
 * Parity pin (tests/test_oracle_golden.py):
 *   - the stepper is checked bit-for-bit against outputs of the reference's own
 *     Numba kernels run under NUMBA_ENABLE_CUDASIM=1 (the .npz files in tests/golden, made
 *     by tests/golden/make_golden.py) and to ~1e-14 against two genuine
 *     EinsteinPy trajectories the reference ships (K1, K2);
 *   - initial conditions / classification are checked against the reference's
 *     get_initial_conditions and run_manual_simulation outputs on the same
 *     inputs (ulp-level only: NumPy's SVML atan2/acos/hypot and OpenBLAS dot
 *     differ from glibc in the last bit);
 *   - the two EinsteinPy coordinate helpers (einsteinpy.coordinates.utils,
 *     un-vendored, no version pinned by the reference; restated from 0.4.0)
 *     are "parity unpinned" except through K1's Cartesian columns.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off matters: the simulator does not fuse multiply-adds.
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference checkout).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_TWO_PI (2.0 * M_PI)

/* ------------------------------------------------------------------------ */
/*  Stepper: simulation/cuda_geodesic.py:432-575                            */
/* ------------------------------------------------------------------------ */

/* _metric_contravariant, cuda_geodesic.py:433-447 */
static void metric_contravariant(const double q[4], double rs, double g[4][4])
{
    double r = q[1], th = q[2];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) g[a][b] = 0.0;
    double inv_fac = 1.0 - rs / r;
    g[0][0] = -1.0 / inv_fac;
    g[1][1] = inv_fac;
    g[2][2] = 1.0 / (r * r);
    double sin_th = sin(th);
    g[3][3] = 1.0 / pow(r * sin_th, 2.0);
}

/* _metric_derivative, cuda_geodesic.py:450-474 (numerators are literal 2.0: M=1 baked in) */
static void metric_derivative(const double q[4], double rs, int wrt, double gp[4][4])
{
    double r = q[1], th = q[2];
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) gp[a][b] = 0.0;
    if (wrt == 1) {
        double denom = r - rs;
        gp[0][0] = 2.0 / (denom * denom);
        gp[1][1] = 2.0 / (r * r);
        gp[2][2] = -2.0 / (r * r * r);
        double sin2 = pow(sin(th), 2.0);
        gp[3][3] = -2.0 / (r * r * r * sin2);
    } else if (wrt == 2) {
        double sin_th = sin(th), cos_th = cos(th);
        gp[3][3] = (-2.0 * cos_th) / ((r * r) * pow(sin_th, 3.0));
    }
}

/* _part_ham_flow, cuda_geodesic.py:477-487 */
static double part_ham_flow(const double q[4], const double p[4], double rs, int wrt)
{
    double gp[4][4];
    metric_derivative(q, rs, wrt, gp);
    double acc = 0.0;
    for (int a = 0; a < 4; ++a)
        for (int b = 0; b < 4; ++b) {
            double val = gp[a][b];
            if (val != 0.0) acc += val * p[a] * p[b];
        }
    return 0.5 * acc;
}

/* _metric_vec_mul, cuda_geodesic.py:490-498 */
static void metric_vec_mul(const double q[4], const double p[4], double rs, double out[4])
{
    double g[4][4];
    metric_contravariant(q, rs, g);
    for (int a = 0; a < 4; ++a) {
        double acc = 0.0;
        for (int b = 0; b < 4; ++b) acc += g[a][b] * p[b];
        out[a] = acc;
    }
}

/* _flow_A_dev, cuda_geodesic.py:501-513: reads (q1,p2), writes p1 and q2 */
static void flow_A(double q1[4], double p1[4], double q2[4], double p2[4], double delta, double rs)
{
    double dH[4], dq[4];
    for (int i = 0; i < 4; ++i) dH[i] = part_ham_flow(q1, p2, rs, i);
    for (int i = 0; i < 4; ++i) p1[i] = p1[i] - delta * dH[i];
    metric_vec_mul(q1, p2, rs, dq);
    for (int i = 0; i < 4; ++i) q2[i] = q2[i] + delta * dq[i];
}

/* _flow_B_dev, cuda_geodesic.py:516-525: reads (q2,p1), writes p2 and q1 */
static void flow_B(double q1[4], double p1[4], double q2[4], double p2[4], double delta, double rs)
{
    double dH[4], dq[4];
    for (int i = 0; i < 4; ++i) dH[i] = part_ham_flow(q2, p1, rs, i);
    for (int i = 0; i < 4; ++i) p2[i] = p2[i] - delta * dH[i];
    metric_vec_mul(q2, p1, rs, dq);
    for (int i = 0; i < 4; ++i) q1[i] = q1[i] + delta * dq[i];
}

/* _flow_mixed_dev, cuda_geodesic.py:528-566 */
static void flow_mixed(double q1[4], double p1[4], double q2[4], double p2[4], double delta, double omega)
{
    double qs[4], qd[4], ps[4], pd[4];
    for (int k = 0; k < 4; ++k) {
        qs[k] = q1[k] + q2[k];
        qd[k] = q1[k] - q2[k];
        ps[k] = p1[k] + p2[k];
        pd[k] = p1[k] - p2[k];
    }
    double c = cos(2.0 * omega * delta);
    double s = sin(2.0 * omega * delta);
    for (int k = 0; k < 4; ++k) {
        q1[k] = 0.5 * (qs[k] + qd[k] * c + pd[k] * s);
        p1[k] = 0.5 * (ps[k] + pd[k] * c - qd[k] * s);
        q2[k] = 0.5 * (qs[k] - qd[k] * c - pd[k] * s);
        p2[k] = 0.5 * (ps[k] - pd[k] * c + qd[k] * s);
    }
}

/* _fantasy_step_ord2, cuda_geodesic.py:569-575 */
static void fantasy_step_ord2(double q1[4], double p1[4], double q2[4], double p2[4],
                              double delta, double rs, double omega)
{
    flow_A(q1, p1, q2, p2, 0.5 * delta, rs);
    flow_B(q1, p1, q2, p2, 0.5 * delta, rs);
    flow_mixed(q1, p1, q2, p2, delta, omega);
    flow_B(q1, p1, q2, p2, 0.5 * delta, rs);
    flow_A(q1, p1, q2, p2, 0.5 * delta, rs);
}

/* One ray of fantasy_integrate_batch_ord2, cuda_geodesic.py:580-602.
 * Returns the number of steps taken.  rmin_out (nullable) receives the smallest r either copy visited. */
static int64_t integrate_one(const double q0[4], const double p0[4], int64_t steps, double delta,
                             double rs, double r_max, double omega, double out_q[4], double *rmin_out)
{
    double q1[4], p1[4], q2[4], p2[4];
    for (int a = 0; a < 4; ++a) { q1[a] = q0[a]; p1[a] = p0[a]; q2[a] = q0[a]; p2[a] = p0[a]; }
    int64_t n = 0;
    double rmin = q1[1];                 /* closest approach of either copy (test bookkeeping for band H) */
    for (int64_t it = 0; it < steps; ++it) {
        if (q1[1] <= 1.1 * rs || q1[1] >= r_max) break;
        fantasy_step_ord2(q1, p1, q2, p2, delta, rs, omega);
        ++n;
        if (q1[1] < rmin) rmin = q1[1];
        if (q2[1] < rmin) rmin = q2[1];
    }
    for (int a = 0; a < 4; ++a) out_q[a] = q1[a];
    if (rmin_out) *rmin_out = rmin;
    return n;
}

/* CUDASchwarzschildIntegrator.integrate_batch, cuda_geodesic.py:639-645 */
int orc_integrate_batch(const double *q0s, const double *p0s, int64_t n, int64_t steps, double delta,
                        double rs, double r_max, double omega, double *out_qs, int64_t *out_steps,
                        double *out_rmin, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t i = 0; i < n; ++i) {
        int64_t k = integrate_one(q0s + 4 * i, p0s + 4 * i, steps, delta, rs, r_max, omega, out_qs + 4 * i,
                                  out_rmin ? out_rmin + i : 0);
        if (out_steps) out_steps[i] = k;
    }
    return 0;
}

/* fantasy_integrate_batch_ord2_full, cuda_geodesic.py:605-625 (+ wrapper :647-653).
 * out_traj is (n, steps, 4), caller-zeroed like np.zeros; out_mom (nullable, same shape)
 * additionally records the copy-1 momenta p1 (not exposed by the reference; used for K1). */
int orc_integrate_batch_full(const double *q0s, const double *p0s, int64_t n, int64_t steps, double delta,
                             double rs, double r_max, double omega, double *out_traj, double *out_mom,
                             int64_t *out_rows)
{
    for (int64_t i = 0; i < n; ++i) {
        double q1[4], p1[4], q2[4], p2[4];
        for (int a = 0; a < 4; ++a) { q1[a] = q0s[4 * i + a]; p1[a] = p0s[4 * i + a]; q2[a] = q1[a]; p2[a] = p1[a]; }
        int64_t rows = 0;
        for (int64_t step = 0; step < steps; ++step) {
            for (int a = 0; a < 4; ++a) out_traj[(i * steps + step) * 4 + a] = q1[a];
            if (out_mom) for (int a = 0; a < 4; ++a) out_mom[(i * steps + step) * 4 + a] = p1[a];
            rows = step + 1;
            if (q1[1] <= 1.1 * rs || q1[1] >= r_max) break;
            fantasy_step_ord2(q1, p1, q2, p2, delta, rs, omega);
        }
        if (out_rows) out_rows[i] = rows;
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/*  EinsteinPy helpers (einsteinpy.coordinates.utils, 0.4.0; un-vendored)   */
/* ------------------------------------------------------------------------ */
static void c2s(double x, double y, double z, double *r, double *th, double *ph)
{
    double hxy = hypot(x, y);
    *r = hypot(hxy, z);
    *th = atan2(hxy, z);
    *ph = atan2(y, x);
}
static void s2c(double r, double th, double ph, double *x, double *y, double *z)
{
    *x = r * cos(ph) * sin(th);
    *y = r * sin(ph) * sin(th);
    *z = r * cos(th);
}

/* Python float modulo (result takes the sign of the divisor) */
static double pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) { if ((m < 0.0) != (b < 0.0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}

/* ------------------------------------------------------------------------ */
/*  Scene                                                                   */
/* ------------------------------------------------------------------------ */
typedef struct orc_scene {
    double obs[3];           /* observer.position (reference scenes use (x,0,0))         */
    double fov;              /* radians                                                  */
    int32_t height, width;   /* observer.image_size = (h, w)                             */
    double mass;             /* bh.mass; rs = 2*mass                                     */
    double r_max;            /* boundary_radius (integrator r_max and escape threshold)  */
    int64_t steps;
    double delta, omega;     /* omega actually used by the integrator (1.0 via run_manual_simulation) */
    double patch_center_theta, patch_center_phi, patch_size_theta, patch_size_phi;
    int32_t flip_theta, flip_phi;
    int32_t has_texture, _pad;
} orc_scene;

/* Camera pixel: simulation/raytracing.py:60-88 */
static void camera_pixel(const orc_scene *sc, int i, int j, double px[3])
{
    int h = sc->height, w = sc->width;
    double nrm = sqrt(sc->obs[0] * sc->obs[0] + sc->obs[1] * sc->obs[1] + sc->obs[2] * sc->obs[2]);
    double plane_dist = 0.2 * nrm;
    double pc[3] = { sc->obs[0] + (-1.0) * plane_dist, sc->obs[1] + 0.0 * plane_dist, sc->obs[2] + 0.0 * plane_dist };
    double plane_width = 2 * plane_dist * tan(sc->fov / 2);
    double plane_height = plane_width * ((double)h / (double)w);
    double u = (j + 0.5) / w - 0.5;
    double v = (i + 0.5) / h - 0.5;
    double uw = u * plane_width, vh = v * plane_height;
    px[0] = (pc[0] + uw * 0.0) + vh * 0.0;
    px[1] = (pc[1] + uw * 1.0) + vh * 0.0;
    px[2] = (pc[2] + uw * 0.0) + vh * 1.0;
}

/* get_initial_conditions (simulation/utils.py:91-141) with angles_to_p_sph (:148-194)
 * and build_null_4momentum_ep_sph (:39-88, future=True).
 * ic9 = p0[4], alpha0, beta, h_r, h_theta, h_phi.  Returns 0, or a negative code mirroring
 * the reference's exceptions: -1 ValueError (f<=0), -2 RuntimeError (disc<0), -3 AssertionError. */
static int initial_conditions(const double obs[3], const double px[3], double mass, double q0[4], double ic9[9])
{
    double d[3] = { px[0] - obs[0], px[1] - obs[1], px[2] - obs[2] };
    double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);       /* np.linalg.norm */
    d[0] /= nrm; d[1] /= nrm; d[2] /= nrm;
    double beta = atan2(d[2], d[1]);
    double c = cos(-beta), s = sin(-beta);
    /* R_x @ ray_dir with R_x = [[1,0,0],[0,c,-s],[0,s,c]] */
    double rx = (1.0 * d[0] + 0.0 * d[1]) + 0.0 * d[2];
    double ry = (0.0 * d[0] + c * d[1]) + (-s) * d[2];
    double rz = (0.0 * d[0] + s * d[1]) + c * d[2];
    if (!(fabs(rz) < 1e-6)) return -3;
    double r_obs, th_obs, ph_obs;
    c2s(obs[0], obs[1], obs[2], &r_obs, &th_obs, &ph_obs);
    double h_r, h_theta, h_phi;
    c2s(rx, ry, rz, &h_r, &h_theta, &h_phi);
    if (!(fabs(h_theta - M_PI / 2) < 1e-6)) return -3;
    /* angles_to_p_sph(pi - h_phi, 0.0, r_obs) */
    double alpha = M_PI - h_phi, bet0 = 0.0;
    double n_rhat = -cos(alpha) * cos(bet0);
    double n_phhat = sin(alpha) * cos(bet0);
    double n_thhat = -sin(bet0);
    double f_r = sqrt(1.0 - 2.0 * mass / r_obs);
    double pr = n_rhat * f_r, pth = n_thhat * r_obs, pph = n_phhat * r_obs;
    /* build_null_4momentum_ep_sph */
    double f = 1.0 - 2.0 * mass / r_obs;
    if (f <= 0.0) return -1;
    double gtt = -1.0 / f, grr = f, gthth = 1.0 / (r_obs * r_obs);
    double gphph = 1.0 / (r_obs * r_obs * pow(sin(th_obs), 2.0));
    double A = gtt;
    double C = grr * pr * pr + gthth * pth * pth + gphph * pph * pph;
    double disc = -4.0 * A * C;
    if (disc < 0.0) return -2;
    double p_t = sqrt(disc) / (2.0 * (-A));
    q0[0] = 0.0; q0[1] = r_obs; q0[2] = th_obs; q0[3] = ph_obs;
    ic9[0] = p_t; ic9[1] = pr; ic9[2] = pth; ic9[3] = pph;
    ic9[4] = acos(-pr / sqrt(1.0 - 2.0 * mass / r_obs));   /* alpha0 */
    ic9[5] = beta;
    c2s(d[0], d[1], d[2], &ic9[6], &ic9[7], &ic9[8]);       /* diagnostic heading of the un-rotated ray */
    return 0;
}

/* public: initial conditions for arbitrary pixel positions (utils.py:91-141) */
int orc_initial_conditions(const double *obs, const double *pixels, int64_t n, double mass,
                           double *q0s, double *ic9s)
{
    int rc = 0;
    for (int64_t k = 0; k < n; ++k) {
        int e = initial_conditions(obs, pixels + 3 * k, mass, q0s + 4 * k, ic9s + 9 * k);
        if (e && !rc) rc = e;
    }
    return rc;
}

/* public: camera pixel positions for listed (i,j) (raytracing.py:60-88) */
int orc_camera_pixels(const orc_scene *sc, const int32_t *ij, int64_t n, double *pixels)
{
    for (int64_t k = 0; k < n; ++k) camera_pixel(sc, ij[2 * k], ij[2 * k + 1], pixels + 3 * k);
    return 0;
}

enum { ORC_BH = 0, ORC_NUMERR = 1, ORC_ESCAPE_BG = 2, ORC_ESCAPE_NO_PATCH = 3, ORC_IN_DOMAIN = 4 };

/* Classification + texel lookup for one ray: simulation/raytracing.py:166-174,181-273.
 * hit2 = (th_hit, ph_hit) as they end up in photon_data.csv; uv = texel or (-1,-1). */
static int classify_one(const orc_scene *sc, const double qf[4], double alpha0, double beta,
                        const uint8_t *tex, uint8_t rgb[3], double hit2[2], int32_t uv[2])
{
    int h = sc->height, w = sc->width;
    double rs = 2 * sc->mass;
    double theta0 = sc->patch_center_theta - sc->patch_size_theta / 2;
    double theta1 = sc->patch_center_theta + sc->patch_size_theta / 2;
    double phi0 = sc->patch_center_phi - sc->patch_size_phi / 2;
    double phi_span = sc->patch_size_phi;
    double r_bh = qf[1], th_hit = qf[2], ph_hit = qf[3];
    double x, y, z;
    s2c(r_bh, th_hit, ph_hit, &x, &y, &z);
    double c = cos(beta), s = sin(beta);
    double x2 = (1.0 * x + 0.0 * y) + 0.0 * z;
    double y2 = (0.0 * x + c * y) + (-s) * z;
    double z2 = (0.0 * x + s * y) + c * z;
    double rr;
    c2s(x2, y2, z2, &rr, &th_hit, &ph_hit);
    double b_crit = 3 * sqrt(3.0) * rs;
    double bh_angle = asin(b_crit / sc->obs[0]) / 2;
    int cls;
    rgb[0] = rgb[1] = rgb[2] = 0;
    uv[0] = uv[1] = -1;
    if ((r_bh <= rs * 1.2) || (alpha0 <= bh_angle)) {
        cls = ORC_BH;
    } else if (r_bh >= 100) {
        rgb[0] = 255; cls = ORC_NUMERR;
    } else if (r_bh >= sc->r_max) {
        if (sc->has_texture) {
            th_hit = pymod(th_hit, ORC_TWO_PI);
            ph_hit = pymod(ph_hit, ORC_TWO_PI);
            double dtheta = fabs(th_hit - sc->patch_center_theta);
            if (sc->flip_phi) ph_hit = -ph_hit;
            double phi_rel = pymod(ph_hit - phi0, ORC_TWO_PI);
            double dphi = fabs(pymod(ph_hit - sc->patch_center_phi + M_PI, ORC_TWO_PI) - M_PI);
            if ((dtheta <= sc->patch_size_theta / 2) && (dphi <= phi_span / 2)) {
                double theta_map = sc->flip_theta ? (M_PI - th_hit) : th_hit;
                double fu = (theta_map - theta0) / (theta1 - theta0) * (h - 1) + 0.5;
                double fv = phi_rel / phi_span * (w - 1) + 0.5;
                int64_t u = (int64_t)fu, v = (int64_t)fv;           /* int(): truncation toward zero */
                if (u < 0) u = 0; if (u > h - 1) u = h - 1;
                if (v < 0) v = 0; if (v > w - 1) v = w - 1;
                const uint8_t *t = tex + ((int64_t)u * w + v) * 3;
                rgb[0] = t[0]; rgb[1] = t[1]; rgb[2] = t[2];
                uv[0] = (int32_t)u; uv[1] = (int32_t)v;
                cls = ORC_ESCAPE_BG;
            } else cls = ORC_ESCAPE_NO_PATCH;
        } else cls = ORC_ESCAPE_NO_PATCH;
    } else cls = ORC_IN_DOMAIN;
    hit2[0] = th_hit; hit2[1] = ph_hit;
    return cls;
}

/* public: classification of given final states (raytracing.py:181-273) */
int orc_classify(const orc_scene *sc, const double *final_q, const double *alpha0, const double *beta, int64_t n,
                 const uint8_t *tex, uint8_t *rgb, uint8_t *cls, double *hit2, int32_t *uv)
{
    for (int64_t k = 0; k < n; ++k) {
        uint8_t c3[3]; double h2[2]; int32_t t2[2];
        int c = classify_one(sc, final_q + 4 * k, alpha0[k], beta[k], tex, c3, h2, t2);
        if (rgb) memcpy(rgb + 3 * k, c3, 3);
        if (cls) cls[k] = (uint8_t)c;
        if (hit2) { hit2[2 * k] = h2[0]; hit2[2 * k + 1] = h2[1]; }
        if (uv) { uv[2 * k] = t2[0]; uv[2 * k + 1] = t2[1]; }
    }
    return 0;
}

/* public: the whole path for rows row_begin, row_begin+row_stride, ... < row_end of the frame
 * (run_manual_simulation, raytracing.py:16-314, image part).  Outputs are packed by visited row.
 * All output pointers except rgb/cls are nullable.  Returns 0 or the first IC error code. */
int orc_render(const orc_scene *sc, const uint8_t *tex, int row_begin, int row_end, int row_stride,
               uint8_t *rgb, uint8_t *cls, double *final_q, double *ic9, double *hit2, int32_t *uv,
               int64_t *nsteps, int nthreads)
{
    int w = sc->width;
    int nrows = row_end > row_begin ? (row_end - row_begin + row_stride - 1) / row_stride : 0;
    int rc = 0;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < (int64_t)nrows * w; ++k) {
        int i = row_begin + (int)(k / w) * row_stride, j = (int)(k % w);
        double px[3], q0[4], ic[9], qf[4], h2[2];
        uint8_t c3[3]; int32_t t2[2];
        camera_pixel(sc, i, j, px);
        int e = initial_conditions(sc->obs, px, sc->mass, q0, ic);
        if (e) {
#pragma omp critical
            if (!rc) rc = e;
            continue;
        }
        int64_t ns = integrate_one(q0, ic, sc->steps, sc->delta, 2.0 * sc->mass, sc->r_max, sc->omega, qf, 0);
        int c = classify_one(sc, qf, ic[4], ic[5], tex, c3, h2, t2);
        memcpy(rgb + 3 * k, c3, 3);
        cls[k] = (uint8_t)c;
        if (final_q) memcpy(final_q + 4 * k, qf, 32);
        if (ic9) memcpy(ic9 + 9 * k, ic, 72);
        if (hit2) { hit2[2 * k] = h2[0]; hit2[2 * k + 1] = h2[1]; }
        if (uv) { uv[2 * k] = t2[0]; uv[2 * k + 1] = t2[1]; }
        if (nsteps) nsteps[k] = ns;
    }
    return rc;
}

/* ------------------------------------------------------------------------ */
/*  Flat-space comparison render: save_no_gravity_image_with_background,     */
/*  simulation/background.py:30-66 (set-up) and :99-132 (CPU pixel loop)     */
/* ------------------------------------------------------------------------ */
static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static int in_phi_patch(double phi, double phi0, double phi1)     /* background.py:8-17 */
{
    phi = pymod(phi, ORC_TWO_PI); phi0 = pymod(phi0, ORC_TWO_PI); phi1 = pymod(phi1, ORC_TWO_PI);
    if (phi0 <= phi1) return phi0 <= phi && phi <= phi1;
    return phi >= phi0 || phi <= phi1;
}

/* Uses sc->obs, fov, height, width, r_max (boundary_radius), patch_*, flip_*.  The caller passes the patch
 * centre already resolved (antipode of the observer unless override_patch_center, background.py:54-58). */
int orc_flat_render(const orc_scene *sc, const uint8_t *tex, uint8_t *rgb)
{
    int h = sc->height, w = sc->width;
    const double *obs = sc->obs;
    double nrm = sqrt(obs[0] * obs[0] + obs[1] * obs[1] + obs[2] * obs[2]);
    double axis[3] = { (0.0 - obs[0]) / nrm, (0.0 - obs[1]) / nrm, (0.0 - obs[2]) / nrm };
    double up_guess[3] = { 0, 0, 1 }, cr[3], right[3], up[3];
    cross3(axis, up_guess, cr);
    if (fabs(cr[0]) <= 1e-8 && fabs(cr[1]) <= 1e-8 && fabs(cr[2]) <= 1e-8) { up_guess[1] = 1; up_guess[2] = 0; }
    cross3(up_guess, axis, right);
    double rn = sqrt(right[0] * right[0] + right[1] * right[1] + right[2] * right[2]);
    for (int a = 0; a < 3; ++a) right[a] /= rn;
    cross3(axis, right, up);
    double un = sqrt(up[0] * up[0] + up[1] * up[1] + up[2] * up[2]);
    for (int a = 0; a < 3; ++a) up[a] /= un;
    double plane_dist = 0.2 * nrm, pc[3];
    for (int a = 0; a < 3; ++a) pc[a] = obs[a] + axis[a] * plane_dist;
    double plane_width = 2 * plane_dist * tan(sc->fov / 2);
    double plane_height = plane_width * ((double)h / (double)w);
    double theta0 = sc->patch_center_theta - sc->patch_size_theta / 2;
    double theta1 = sc->patch_center_theta + sc->patch_size_theta / 2;
    double phi0 = sc->patch_center_phi - sc->patch_size_phi / 2;
    double phi1 = sc->patch_center_phi + sc->patch_size_phi / 2;
    double phi_span = pymod(phi1 - phi0, ORC_TWO_PI);
    if (phi_span == 0) phi_span = ORC_TWO_PI;
    double cc = (obs[0] * obs[0] + obs[1] * obs[1] + obs[2] * obs[2]) - sc->r_max * sc->r_max;
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            uint8_t *o = rgb + ((size_t)i * w + j) * 3;
            o[0] = o[1] = o[2] = 0;
            double u = (j + 0.5) / w - 0.5, v = (i + 0.5) / h - 0.5, d[3];
            for (int a = 0; a < 3; ++a) d[a] = ((pc[a] + (u * plane_width) * right[a]) + (v * plane_height) * up[a]) - obs[a];
            double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
            for (int a = 0; a < 3; ++a) d[a] /= dn;
            double qa = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            double qb = 2 * (obs[0] * d[0] + obs[1] * d[1] + obs[2] * d[2]);
            double disc = qb * qb - 4 * qa * cc;
            if (disc < 0) continue;
            double t = (-qb + sqrt(disc)) / (2 * qa);
            double hp[3] = { obs[0] + t * d[0], obs[1] + t * d[1], obs[2] + t * d[2] };
            double r = sqrt(hp[0] * hp[0] + hp[1] * hp[1] + hp[2] * hp[2]);
            double theta = acos(hp[2] / r), phi = atan2(hp[1], hp[0]);
            if (theta >= theta0 && theta <= theta1 && in_phi_patch(phi, phi0, phi1)) {
                double theta_map = sc->flip_theta ? (M_PI - theta) : theta;
                double phi_map = sc->flip_phi ? -phi : phi;
                int64_t ub = (int64_t)((theta_map - theta0) / (theta1 - theta0) * (h - 1));
                double phi_mod = pymod(phi_map - phi0, ORC_TWO_PI);
                int64_t vb = (int64_t)(phi_mod / phi_span * (w - 1));
                if (ub < 0) ub = 0; if (ub > h - 1) ub = h - 1;
                if (vb < 0) vb = 0; if (vb > w - 1) vb = w - 1;
                const uint8_t *tx = tex + ((size_t)ub * w + vb) * 3;
                o[0] = tx[0]; o[1] = tx[1]; o[2] = tx[2];
            }
        }
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
