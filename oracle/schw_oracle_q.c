/*
 * schw_oracle_q.c — QUAD-PRECISION TRUTH for the stepper.  TEST INFRASTRUCTURE ONLY.
 *
 * The same FANTASY order-2 map as oracle/schw_oracle.c (the reference's simulation/cuda_geodesic.py:432-625,
 * cited per function), with every arithmetic operation carried out in IEEE binary128 (__float128, libquadmath:
 * 113-bit significand, unit round-off 1e-34) on the SAME double-precision inputs.  Over the <= 2e5 steps of a ray
 * its own rounding error stays ~1e-29 times the error amplification of the trajectory, i.e. it is the exact result
 * of the discrete map as far as any FP64 implementation is concerned.  It exists so that the parity tests can
 * state their tolerances against a measured truth:
 *
 *     |gpu - exact|  <=  max( 1e-9 * |exact| ,  4 * max_k |ref_k - exact| )
 *
 * where ref_k are FP64 evaluations of the reference's arithmetic (oracle/schw_oracle.c) on the inputs and on
 * 1-ulp perturbations of them: wherever the reference itself is accurate to 1e-9 the north-star bar applies, and
 * where the reference's own FP64 result is further than that from the truth (exponentially unstable omega = 0.01
 * trajectories, rays grazing the photon sphere) the GPU may be at most 4x as far.
 *
 * What is "exact": q0, p0, delta, rs, r_max, omega are taken as the given doubles; h = delta/2 and the mixing
 * angle 2*omega*delta are formed in binary128 (exactly); cos/sin of it with cosq/sinq; the termination thresholds
 * 1.1*rs and r_max are the reference's FP64 values (they are parameters of the algorithm, cuda_geodesic.py:598).
 * Results are rounded to double on output.  `force_steps` (nullable) makes ray i take exactly force_steps[i] steps,
 * ignoring the termination test: the exact state after the number of steps some FP64 run took.
 *
 * Nothing in the product path may call, link or import this file.  Build: oracle/Makefile (-lquadmath).
 */
#include <quadmath.h>
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef __float128 Q;

/* One Hamiltonian flow, _flow_A_dev / _flow_B_dev (cuda_geodesic.py:501-525) with _part_ham_flow (:477-487),
 * _metric_derivative (:450-474, numerators literal 2.0) and _metric_vec_mul / _metric_contravariant (:433-447,
 * :490-498): position q of one copy, momenta p of the other; updates momenta P and positions X of "the other two". */
static void flow(const Q q[4], const Q p[4], Q P[4], Q X[4], Q h, Q rs)
{
    const Q r = q[1], th = q[2];
    Q s, c;
    sincosq(th, &s, &c);
    const Q d = r - rs;
    /* dH/dr = 1/2 ( 2/(r-rs)^2 pt^2 + 2/r^2 pr^2 - 2/r^3 pth^2 - 2/(r^3 sin^2) pph^2 ) */
    const Q dHr = 0.5Q * ((2.0Q / (d * d)) * p[0] * p[0] + (2.0Q / (r * r)) * p[1] * p[1] +
                          (-2.0Q / (r * r * r)) * p[2] * p[2] + (-2.0Q / (r * r * r * s * s)) * p[3] * p[3]);
    /* dH/dth = 1/2 ( -2 cos / (r^2 sin^3) pph^2 ) */
    const Q dHth = 0.5Q * (((-2.0Q * c) / ((r * r) * (s * s * s))) * p[3] * p[3]);
    P[1] -= h * dHr;
    P[2] -= h * dHth;
    const Q f = 1.0Q - rs / r;
    X[0] += h * ((-1.0Q / f) * p[0]);
    X[1] += h * (f * p[1]);
    X[2] += h * ((1.0Q / (r * r)) * p[2]);
    X[3] += h * ((1.0Q / ((r * s) * (r * s))) * p[3]);
}

/* _flow_mixed_dev, cuda_geodesic.py:528-566 */
static void flow_mixed(Q q1[4], Q p1[4], Q q2[4], Q p2[4], Q c, Q s)
{
    for (int k = 0; k < 4; ++k) {
        const Q qs = q1[k] + q2[k], qd = q1[k] - q2[k], ps = p1[k] + p2[k], pd = p1[k] - p2[k];
        q1[k] = 0.5Q * (qs + qd * c + pd * s);
        p1[k] = 0.5Q * (ps + pd * c - qd * s);
        q2[k] = 0.5Q * (qs - qd * c - pd * s);
        p2[k] = 0.5Q * (ps - pd * c + qd * s);
    }
}

typedef struct { Q q1[4], p1[4], q2[4], p2[4], h, rs, c, s; } Ray;

static void ray_init(Ray *R, const double q0[4], const double p0[4], double delta, double rs, double omega)
{
    for (int a = 0; a < 4; ++a) { R->q1[a] = R->q2[a] = q0[a]; R->p1[a] = R->p2[a] = p0[a]; }
    R->h = 0.5Q * (Q)delta;
    R->rs = rs;
    const Q ang = 2.0Q * (Q)omega * (Q)delta;
    sincosq(ang, &R->s, &R->c);
}

/* _fantasy_step_ord2, cuda_geodesic.py:569-575: A(d/2) B(d/2) C(d) B(d/2) A(d/2) */
static void ray_step(Ray *R)
{
    flow(R->q1, R->p2, R->p1, R->q2, R->h, R->rs);
    flow(R->q2, R->p1, R->p2, R->q1, R->h, R->rs);
    flow_mixed(R->q1, R->p1, R->q2, R->p2, R->c, R->s);
    flow(R->q2, R->p1, R->p2, R->q1, R->h, R->rs);
    flow(R->q1, R->p2, R->p1, R->q2, R->h, R->rs);
}

/* fantasy_integrate_batch_ord2, cuda_geodesic.py:579-602 */
int orcq_integrate_batch(const double *q0s, const double *p0s, int64_t n, int64_t steps, double delta, double rs,
                         double r_max, double omega, const int64_t *force_steps, double *out_qs, double *out_ps,
                         int64_t *out_steps, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const double rcap = 1.1 * rs;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        Ray R;
        ray_init(&R, q0s + 4 * i, p0s + 4 * i, delta, rs, omega);
        int64_t k = 0;
        if (force_steps) {
            for (; k < force_steps[i]; ++k) ray_step(&R);
        } else {
            for (; k < steps; ++k) {
                if (R.q1[1] <= (Q)rcap || R.q1[1] >= (Q)r_max) break;
                ray_step(&R);
            }
        }
        for (int a = 0; a < 4; ++a) out_qs[4 * i + a] = (double)R.q1[a];
        if (out_ps) for (int a = 0; a < 4; ++a) out_ps[4 * i + a] = (double)R.p1[a];
        if (out_steps) out_steps[i] = k;
    }
    return 0;
}

/* fantasy_integrate_batch_ord2_full, cuda_geodesic.py:604-625: row `step` = copy-1 position before the test.
 * out_traj (n, steps, 4) caller-zeroed; out_mom (nullable) the copy-1 momenta.  stop = 0 ignores the termination test
 * (the exact continuation past the point where an FP64 run stopped). */
int orcq_integrate_batch_full(const double *q0s, const double *p0s, int64_t n, int64_t steps, double delta, double rs,
                              double r_max, double omega, int stop, double *out_traj, double *out_mom,
                              int64_t *out_rows, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
    const double rcap = 1.1 * rs;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t i = 0; i < n; ++i) {
        Ray R;
        ray_init(&R, q0s + 4 * i, p0s + 4 * i, delta, rs, omega);
        int64_t rows = 0;
        for (int64_t step = 0; step < steps; ++step) {
            for (int a = 0; a < 4; ++a) out_traj[(i * steps + step) * 4 + a] = (double)R.q1[a];
            if (out_mom) for (int a = 0; a < 4; ++a) out_mom[(i * steps + step) * 4 + a] = (double)R.p1[a];
            rows = step + 1;
            if (stop && (R.q1[1] <= (Q)rcap || R.q1[1] >= (Q)r_max)) break;
            ray_step(&R);
        }
        if (out_rows) out_rows[i] = rows;
    }
    return 0;
}
