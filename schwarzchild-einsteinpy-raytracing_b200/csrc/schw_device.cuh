// schw_device.cuh — device-side arithmetic of the Schwarzschild hot path for sm_100a.
//
// Everything here is FP64, one ray per thread, state in registers.  The formulas are the
// reference's (file:line cited per function, relative to the reference checkout) but the
// evaluation is re-derived for the B200 FP64 pipe (64 DFMA lanes / SM):
//   * two reciprocals per Hamiltonian flow instead of nine divisions: R = 1/(r (r-rs)) gives
//     1/r = R (r-rs) and 1/(r-rs) = R r; 1/sin(th) belongs to the memoised trig factors;
//   * one sincos per flow instead of 3 sin + 1 cos;
//   * the trailing A(d/2) of step k and the leading A(d/2) of step k+1 read the same (q1,p2)
//     (A only writes p1,q2), so they are evaluated once and applied as A(d);
//   * the mixed flow is applied as a small correction q -= w, p -= z with (1-cos(2 w d))/2 and
//     sin(2 w d)/2 computed once on the host (10 instead of 24 flops per component);
//   * a branch-free sincos with constant-bank coefficients (no slow path, no UMOV traffic).
// These change rounding at the ulp level only (the reference's own GPU build already differs
// from its CPU simulator by FMA contraction); parity is asserted at 1e-9 relative on final
// states and bit-exact on class / texel (tests/).
#pragma once
#include <cstdint>
#ifndef SCHW_HOST_MODEL
#include <cuda_runtime.h>
#endif
// SCHW_HOST_MODEL: tests/model/planar_model.cpp compiles this header with g++ (one lane = one "warp", shims for the
// few intrinsics) to check the planar fast path against the general flow on the CPU over many parameter sets.

namespace schw {

// ---------------------------------------------------------------------------------------------
// Parameters (passed by value as __grid_constant__ kernel arguments => constant bank operands)
// ---------------------------------------------------------------------------------------------
struct alignas(16) StepParams {
    // (pairs that one flow reads together sit in one aligned 16-byte slot so that a single LDCU.128 can fetch them)
    double h, h_rs;          // 0.5*delta, h*rs                   (cuda_geodesic.py:571-575)
    double delta, d_rs;      // full step (merged A(d)), delta*rs
    double mg, mhs;          // (1-cos(2 w d))/2, sin(2 w d)/2   (cuda_geodesic.py:545-546; cos/sin from host libm)
    // equatorial fast path (see "Planar steps" below); thresholds are compared on the high 32 bits of a double
    double pl_hcB, pl_hcA;   // (delta/2) * cot(th_eq), delta * cot(th_eq), th_eq = the double nearest pi/2
    double rs;               // 2*mass                             (cuda_geodesic.py:635)
    double pl_zero_hs;       // (+0) * mhs: the signed zero the mixed flow adds to p_theta when q1_theta == q2_theta
    double rcap;             // 1.1*rs evaluated in FP64           (cuda_geodesic.py:598)
    double r_max;
    double park_r;           // radius at which lanes without a ray sit on the fixed point of the flows (make_step_params)
    double _pad8;
    int32_t pl_ph_hi;        // |p_theta| <  T      <=  hi(|p_theta|) <  pl_ph_hi
    int32_t pl_r_hi;         // r         >= r_thr  <=  hi(r)         >= pl_r_hi   (signed compare)
    int32_t pl_pp_hi;        // |p_phi|   >  2^28 T <=  hi(|p_phi|)   >= pl_pp_hi
    int32_t pl_enable;
    // termination pre-test on the high word of r (integer pipe): hi(r) > hi_rcap && hi(r) < hi_rmax  =>  rcap < r < r_max
    int32_t hi_rcap, hi_rmax;
    int32_t steps;
    int32_t _pad;
};

struct SceneParams {
    // camera (raytracing.py:60-88), precomputed on the host in FP64 exactly as the reference does
    double obs[3];
    double pc[3];                 // plane_center
    double plane_width, plane_height;
    // observer in spherical coordinates and the null-momentum constants (utils.py:39-88,118,140,179)
    double r_obs, th_obs, ph_obs;
    double f_r;                   // sqrt(1-2M/r_obs)
    double grr, gthth, gphph;     // f, 1/r^2, 1/(r^2 sin^2 th)
    double m4A;                   // -4*A      (A = -1/f)
    double twomA;                 // 2*(-A)
    // classification (raytracing.py:166-174,205-216)
    double rs12;                  // rs*1.2
    double bh_angle;              // asin(3*sqrt(3)*rs/obs_x)/2
    double r_max;
    double theta0, theta1, phi0, phi_span;
    double pct, pcp, half_pst, half_psp;
    int32_t height, width;
    int32_t flip_theta, flip_phi;
    int32_t has_texture;
    int32_t row_begin, row_stride, nrows;
};

struct RayState {
    double q1t, q1r, q1h, q1p;
    double p1t, p1r, p1h, p1p;
    double q2t, q2r, q2h, q2p;
    double p2t, p2r, p2h, p2p;
};

// ---------------------------------------------------------------------------------------------
// Reciprocal for the flow: MUFU.RCP64H seed (~2^-20) + one cubic Newton step (error e^3 ~ 2^-60,
// i.e. below half an ulp before the final rounding).  r (r-rs) and sin th are normal, non-zero
// numbers on every ray the reference handles (it divides by the same factors), so no special-case
// path is needed: inf/NaN/0 propagate as NaN and the ray ends as 'in_domain' exactly like a NaN in
// the reference.  Accuracy is checked on the device in tests/test_gpu_math.py.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double fast_rcp(double x)
{
    double y;
#ifdef SCHW_HOST_MODEL
    y = __hiloint2double(__double2hiint(1.0 / x), 0);     // stand-in for the ~2^-20 hardware seed
#else
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#endif
    double e = fma(-x, y, 1.0);
    e = fma(e, e, e);
    return fma(y, e, y);
}

// ---------------------------------------------------------------------------------------------
// sincos for the flow.  Branch-free, no slow path: Cody-Waite reduction by pi/2 with FMA (valid to
// 1 ulp of the reduced argument for |x| < 2^20; theta of a geodesic stays O(1)), fdlibm minimax
// kernels on [-pi/4, pi/4] (|err| < 2^-58), quadrant fix-up with integer selects.  Coefficients live
// in the constant bank so every DFMA takes them as a c[][] operand (no UMOV traffic on the issue
// port, which this kernel shares almost 1:1 with the FP64 pipe).  19 FP64 instructions.
// ---------------------------------------------------------------------------------------------
__constant__ double kTrig[18] = {
    0.63661977236758138,        //  0  2/pi
    6755399441055744.0,         //  1  1.5 * 2^52 (round-to-nearest-integer magic)
    1.5707963267948966,         //  2  pi/2 hi
    6.123233995736766e-17,      //  3  pi/2 lo
    -1.66666666666666324348e-01,  //  4  S1
    8.33333333332248946124e-03,   //  5  S2
    -1.98412698298579493134e-04,  //  6  S3
    2.75573137070700676789e-06,   //  7  S4
    -2.50507602534068634195e-08,  //  8  S5
    1.58969099521155010221e-10,   //  9  S6
    4.16666666666666019037e-02,   // 10  C1
    -1.38888888888741095749e-03,  // 11  C2
    2.48015872894767294178e-05,   // 12  C3
    -2.75573143513906633035e-07,  // 13  C4
    2.08757232129817482790e-09,   // 14  C5
    -1.13596475577881948265e-11,  // 15  C6
    0.0, 0.0};

__device__ __forceinline__ void sincos_flow(double x, double &s, double &c)
{
    const double t = fma(x, kTrig[0], kTrig[1]);
    const int q = __double2loint(t);
    const double fq = t - kTrig[1];
    double r = fma(-fq, kTrig[2], x);
    r = fma(-fq, kTrig[3], r);
    const double z = r * r;
    double ps = fma(z, kTrig[9], kTrig[8]);
    double pc = fma(z, kTrig[15], kTrig[14]);
    ps = fma(z, ps, kTrig[7]);
    pc = fma(z, pc, kTrig[13]);
    ps = fma(z, ps, kTrig[6]);
    pc = fma(z, pc, kTrig[12]);
    ps = fma(z, ps, kTrig[5]);
    pc = fma(z, pc, kTrig[11]);
    ps = fma(z, ps, kTrig[4]);
    pc = fma(z, pc, kTrig[10]);
    const double sv = fma(z * r, ps, r);                 // sin(r)
    const double cv = fma(z, fma(z, pc, -0.5), 1.0);     // cos(r)
    // quadrant: sin(x) = {sv, cv, -sv, -cv}[q&3], cos(x) = {cv, -sv, -cv, sv}[q&3]
    const bool odd = q & 1;
    int shi = __double2hiint(odd ? cv : sv), slo = __double2loint(odd ? cv : sv);
    int chi = __double2hiint(odd ? sv : cv), clo = __double2loint(odd ? sv : cv);
    shi ^= (q & 2) << 30;
    chi ^= ((q + 1) & 2) << 30;
    s = __hiloint2double(shi, slo);
    c = __hiloint2double(chi, clo);
}

// ---------------------------------------------------------------------------------------------
// One Hamiltonian flow (A or B):  given the position (r,th) of one phase-space copy and the
// momenta (pt,pr,ph,pp) of the other, advance the *other* momenta / positions by hh.
//   _flow_A_dev / _flow_B_dev           cuda_geodesic.py:501-525
//   _part_ham_flow, _metric_derivative  cuda_geodesic.py:450-487  (numerators literal 2.0)
//   _metric_vec_mul, _metric_contravariant  cuda_geodesic.py:433-447,490-498
// With ir=1/r, id=1/(r-rs), is=1/sin th:
//   dH/dr  = id^2 pt^2 + ir^2 pr^2 - ir^3 pth^2 - ir^3 is^2 pph^2
//   dH/dth = -cos(th) ir^2 is^3 pph^2
//   g.p    = ( -r id pt, (1-rs ir) pr, ir^2 pth, ir^2 is^2 pph )
// ---------------------------------------------------------------------------------------------
// The trig factors of the polar angle of one phase-space copy (1/sin, cos/sin), remembered together
// with the angle they were computed for.  The trig of a flow is re-evaluated only when the angle's BITS changed since the
// last evaluation in some lane of the warp (warp-uniform decision, so no divergence).  This is an
// exact memoisation, not an approximation: results are bit-identical with it on or off.  It pays
// because every camera ray of the reference starts in the equatorial plane with p_theta = -0.0
// (utils.py:129,172) and its theta never leaves the double nearest pi/2 (the per-flow increments
// are ~1e-19, far below half an ulp), while rays with genuine polar motion just pay one compare +
// vote per flow.
struct TrigMemo {
    double th;     // the angle the entry was computed for (compared bitwise)
    double is;     // 1 / sin(th)
    double hcis;   // hh * cos(th) / sin(th), hh = the (fixed) step of the flow this entry serves
};

// The two trig-derived factors a flow needs: 1/sin(th) and cos(th)/sin(th).
__device__ __forceinline__ void trig_factors(double th, double &is, double &cis)
{
    double s, c;
    sincos_flow(th, s, c);
    is = fast_rcp(s);
    cis = c * is;
}

// The part of a flow that does not involve theta or the angular momenta: reciprocals, dt/dl, dr/dl, and the (t,r)
// terms of dH/dr.  Shared by the general and the planar tail below (the same operations in the same order as one
// monolithic flow, so splitting changes no result bit); it also gives the warp votes that choose the tail ~40 cycles
// of independent FP64 work to resolve behind.
struct FlowHead {
    double ir;     // 1 / r
    double dH0;    // (pt/(r-rs))^2 + (pr/r)^2
};
__device__ __forceinline__ FlowHead flow_head(double r, double pt, double pr, double &Qt, double &Qr,
                                              double hh, double hrs, double rs)
{
    const double d   = r - rs;
    const double R   = fast_rcp(r * d);          // 1 / (r (r - rs))
    FlowHead f;
    f.ir = R * d;                                // 1 / r
    const double id  = R * r;                    // 1 / (r - rs)
    const double A1  = id * pt;
    const double A2  = f.ir * pr;
    f.dH0 = fma(A2, A2, A1 * A1);
    Qt = fma(-hh, r * A1, Qt);
    Qr = fma(-hrs, A2, fma(hh, pr, Qr));         // + hh (1 - rs/r) pr
    return f;
}

// Warp-uniform refresh of a memo entry (must be called convergently).
__device__ __forceinline__ void memo_refresh(double th, double hh, TrigMemo &m)
{
    const bool same = __double_as_longlong(th) == __double_as_longlong(m.th);
    if (!__all_sync(0xffffffffu, same)) {
        double cis;
        trig_factors(th, m.is, cis);
        m.hcis = hh * cis;
        m.th = th;
    }
}

__device__ __forceinline__ void flow_tail(const FlowHead &f, double ph, double pp, double is, double hcis,
                                          double &Pr, double &Ph, double &Qh, double &Qp, double hh)
{
    const double A3  = f.ir * ph;
    const double w   = f.ir * is;                // 1 / (r sin th)
    const double A4  = w * pp;
    const double k   = A4 * A4;
    const double e   = fma(A3, A3, k);
    Pr = fma(-hh, fma(-f.ir, e, f.dH0), Pr);
    Ph = fma(hcis, k, Ph);                       // - hh dH/dth,  dH/dth = -cos/sin^3 pph^2 / r^2
    Qh = fma(hh, f.ir * A3, Qh);
    Qp = fma(hh, w * A4, Qp);
}

template <bool kMemo>
__device__ __forceinline__ void flow(double r, double th,
                                     double pt, double pr, double ph, double pp,
                                     double &Pr, double &Ph,                       // momenta updated
                                     double &Qt, double &Qr, double &Qh, double &Qp,  // positions updated
                                     double hh, double hrs, double rs, TrigMemo &m)   // hrs = hh*rs
{
    double is, hcis;
    if (kMemo) {
        // NB: a memo entry serves ONE flow with ONE step size (m1: A(delta), m2: B(delta/2)); full-mask vote =>
        // only call with kMemo from convergent code
        memo_refresh(th, hh, m);
        is = m.is; hcis = m.hcis;
    } else {
        double cis;
        trig_factors(th, is, cis);
        hcis = hh * cis;
    }
    const FlowHead f = flow_head(r, pt, pr, Qt, Qr, hh, hrs, rs);
    flow_tail(f, ph, pp, is, hcis, Pr, Ph, Qh, Qp, hh);
}

// _flow_mixed_dev, cuda_geodesic.py:528-566, for one component.  With g = (1-cos)/2, hs = sin/2:
//   q1' = (q1+q2)/2 + (q1-q2) cos/2 + (p1-p2) sin/2 = q1 - w,   q2' = q2 + w,   w = qd g - pd hs
//   p1' = p1 - z,  p2' = p2 + z,  z = pd g + qd hs
// (1 - cos(2wd) is exact in FP64 by Sterbenz, so g carries exactly the reference's rounded cosine.)
__device__ __forceinline__ void mix(double &q1, double &p1, double &q2, double &p2, double g, double hs)
{
    const double qd = q1 - q2, pd = p1 - p2;
    const double w = fma(qd, g, -(pd * hs));
    const double z = fma(pd, g, qd * hs);
    q1 -= w; q2 += w;
    p1 -= z; p2 += z;
}

// ---------------------------------------------------------------------------------------------
// Planar steps.  Every camera ray of the reference starts at theta = th_eq (the double nearest pi/2)
// with p_theta = -0.0 (utils.py:129,172).  cot(th_eq) = 6.1e-17 then feeds p_theta a few 1e-19 per
// flow, which never moves theta off th_eq: the theta sector only *receives*.  While the following
// holds for every lane of the warp, the theta terms of a flow are provably no-ops in FP64 and are
// skipped (7 of 27 FP64 instructions per flow, 6 of 40 in the mixed flow), bit for bit:
//   (R0) both copies have theta == th_eq bitwise   => 1/sin = 1.0 exactly, cot = 6.123233995736766e-17
//   (R1) hh |p_th| / r^2 <= 2^-54                   => theta + hh g^thth p_th rounds back to theta (ulp(th_eq)/2 = 2^-53)
//   (R2) |p_th| <= 2^-28 |p_ph|                     => fma(A3, A3, k) == k in dH/dr    (A3 = p_th/r, k = (p_ph/r)^2)
//   (R3) 2 T |hs| <= 2^-54 with |p_th| < T          => the mixed flow's w = -(p1_th - p2_th) hs leaves both thetas alone
// (R1)-(R3) follow from |p_th| < T, r >= r_thr, |p_ph| >= 2^28 T with T, r_thr chosen on the host
// (make_step_params); they are tested with integer compares on the high words, which do not occupy
// the FP64 pipe.  p_theta itself is still integrated exactly, so a warp can leave the fast path at any
// sub-step (a high-winding ray whose p_theta outgrows T, a ray loaded with a different theta, ...) and
// continue with the general flow.  B#2 needs no p_theta test: the mixed flow with q1_th == q2_th is a
// convex combination of p1_th and p2_th (0 <= g <= 1).  Results are bit-identical with SCHW_PLANAR=0.
// Every flow is split into a theta-independent head and a general / planar tail (flow_head, flow_tail*).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int hi_abs(double x) { return __double2hiint(x) & 0x7fffffff; }
__device__ __forceinline__ bool is_th_eq(double th) { return __double_as_longlong(th) == 0x3FF921FB54442D18LL; }

__device__ __forceinline__ void flow_tail_planar(const FlowHead &f, double pp, double hc,
                                                 double &Pr, double &Ph, double &Qp, double hh)
{
    const double A4  = f.ir * pp;                // w = ir * (1/sin) = ir * 1.0
    const double k   = A4 * A4;
    Pr = fma(-hh, fma(-f.ir, k, f.dH0), Pr);     // e = fma(A3, A3, k) == k by (R2)
    Ph = fma(hc, k, Ph);
    Qp = fma(hh, f.ir * A4, Qp);
}

// theta component of the mixed flow when q1_th == q2_th and (R3): only the momenta move.
__device__ __forceinline__ void mix_planar(double &p1, double &p2, double g, double zero_hs)
{
    const double pd = p1 - p2;
    const double z = fma(pd, g, zero_hs);
    p1 -= z; p2 += z;
}

struct Memo2 {
    TrigMemo m1, m2;      // for theta of copy 1 (flow A) and copy 2 (flow B)
    // start from a consistent entry (theta = the double nearest pi/2, where parked lanes and every
    // camera ray live); the memo is keyed by the angle's bits, so it can never go stale
    __device__ __forceinline__ void init(double hA, double hB)
    {
        double cis;
        m1.th = m2.th = 1.5707963267948966;
        trig_factors(m1.th, m1.is, cis);
        m2.is = m1.is;
        m1.hcis = hA * cis; m2.hcis = hB * cis;
    }
};

template <bool kMemo>
__device__ __forceinline__ void flow_A(RayState &S, double hh, double hrs, double rs, Memo2 &M)
{
    flow<kMemo>(S.q1r, S.q1h, S.p2t, S.p2r, S.p2h, S.p2p, S.p1r, S.p1h, S.q2t, S.q2r, S.q2h, S.q2p, hh, hrs, rs, M.m1);
}
template <bool kMemo>
__device__ __forceinline__ void flow_B(RayState &S, double hh, double hrs, double rs, Memo2 &M)
{
    flow<kMemo>(S.q2r, S.q2h, S.p1t, S.p1r, S.p1h, S.p1p, S.p2r, S.p2h, S.q1t, S.q1r, S.q1h, S.q1p, hh, hrs, rs, M.m2);
}
__device__ __forceinline__ void flow_C(RayState &S, double g, double hs)
{
    mix(S.q1t, S.p1t, S.q2t, S.p2t, g, hs);
    mix(S.q1r, S.p1r, S.q2r, S.p2r, g, hs);
    mix(S.q1h, S.p1h, S.q2h, S.p2h, g, hs);
    mix(S.q1p, S.p1p, S.q2p, S.p2p, g, hs);
}

// One full loop body B(d/2) C(d) B(d/2) A(d) with the planar fast path.
//   tok    (warp-uniform, carried by the caller, 0/1): (R0) and the tests of the first B hold for every lane, so
//          the step may start on the fast path; the caller clears it whenever lanes were given new rays.
//   pp_hi  this lane's |p_phi| threshold: P.pl_pp_hi while it integrates a ray, 0 while it is parked (parked lanes sit
//          on the fixed point p = 0, theta = th_eq, where both variants of every flow are the identity and every other
//          test passes by itself).
//   next   (out, per lane) this lane passes the tests of the next step's first B; the caller votes on it together
//          with its own end-of-step test and stores the result in tok.
//   n_general (warp-uniform counter) incremented when the step did not run entirely on the fast path.
// Each vote is issued as soon as its inputs exist and consumed only after the next flow_head, so its latency hides
// behind independent FP64 work.  Must be called convergently.
__device__ __forceinline__ void step_BCBA_planar(RayState &S, const StepParams &P, Memo2 &M, int pp_hi, int &tok,
                                                 bool &next, int &n_general)
{
    constexpr unsigned kFullMask = 0xffffffffu;
    const int T = P.pl_ph_hi, RT = P.pl_r_hi;
    if (!tok)       // after a refill, or while some lane carries a ray with polar motion
        tok = __all_sync(kFullMask, is_th_eq(S.q1h) & is_th_eq(S.q2h) &
                                    (hi_abs(S.p1h) < T) & (__double2hiint(S.q2r) >= RT) & (hi_abs(S.p1p) >= pp_hi));
    // (general tails first: ptxas lays the first branch of an if/else out as the fall-through and the second one right
    //  before the join, so the fast path costs one taken branch per flow instead of a not-taken plus a jump)
    // ---- B(d/2): reads (q2, p1), writes p2 and q1 ----
    const bool v1 = tok != 0;
    FlowHead f = flow_head(S.q2r, S.p1t, S.p1r, S.q1t, S.q1r, P.h, P.h_rs, P.rs);
    if (!v1) {
        memo_refresh(S.q2h, P.h, M.m2);
        flow_tail(f, S.p1h, S.p1p, M.m2.is, M.m2.hcis, S.p2r, S.p2h, S.q1h, S.q1p, P.h);
    } else {
        flow_tail_planar(f, S.p1p, P.pl_hcB, S.p2r, S.p2h, S.q1p, P.h);
    }
    // ---- C(d), and the tests of the second B (its r and p_phi are final once their components are mixed) ----
    mix(S.q1r, S.p1r, S.q2r, S.p2r, P.mg, P.mhs);
    mix(S.q1p, S.p1p, S.q2p, S.p2p, P.mg, P.mhs);
    const bool v2 = __all_sync(kFullMask, v1 & (hi_abs(S.p2h) < T) & (__double2hiint(S.q2r) >= RT) & (hi_abs(S.p1p) >= pp_hi));
    mix(S.q1t, S.p1t, S.q2t, S.p2t, P.mg, P.mhs);
    if (!v2) mix(S.q1h, S.p1h, S.q2h, S.p2h, P.mg, P.mhs);
    else mix_planar(S.p1h, S.p2h, P.mg, P.pl_zero_hs);
    // ---- B(d/2) ----
    f = flow_head(S.q2r, S.p1t, S.p1r, S.q1t, S.q1r, P.h, P.h_rs, P.rs);
    if (!v2) {
        memo_refresh(S.q2h, P.h, M.m2);
        flow_tail(f, S.p1h, S.p1p, M.m2.is, M.m2.hcis, S.p2r, S.p2h, S.q1h, S.q1p, P.h);
    } else {
        flow_tail_planar(f, S.p1p, P.pl_hcB, S.p2r, S.p2h, S.q1p, P.h);
    }
    // ---- A(d): reads (q1, p2), writes p1 and q2 ----
    const bool v4 = __all_sync(kFullMask, v2 & (hi_abs(S.p2h) < T) & (__double2hiint(S.q1r) >= RT) & (hi_abs(S.p2p) >= pp_hi));
    f = flow_head(S.q1r, S.p2t, S.p2r, S.q2t, S.q2r, P.delta, P.d_rs, P.rs);
    if (!v4) {
        ++n_general;
        memo_refresh(S.q1h, P.delta, M.m1);
        flow_tail(f, S.p2h, S.p2p, M.m1.is, M.m1.hcis, S.p1r, S.p1h, S.q2h, S.q2p, P.delta);
    } else {
        flow_tail_planar(f, S.p2p, P.pl_hcA, S.p1r, S.p1h, S.q2p, P.delta);
    }
    // the next step's first B: its inputs (p1_th, q2_r, p1_ph) are final now
    next = v4 & (hi_abs(S.p1h) < T) & (__double2hiint(S.q2r) >= RT) & (hi_abs(S.p1p) >= pp_hi);
}

// The same loop body for R rays per thread (R independent geodesics interleaved in one instruction stream): every vote
// covers the R rays of all 32 lanes, so the loop control, the votes, the branches and the constant loads are paid once
// per R ray-steps, and the R dependency chains hide each other's FP64 latency.  R = 1 is step_BCBA_planar itself.
template <int R>
__device__ __forceinline__ void step_BCBA_planar_n(RayState (&S)[R], const StepParams &P, Memo2 *M /*[R], stride mstride*/,
                                                   int mstride, const int (&pp_hi)[R], int &tok, bool &next, int &n_general)
{
    constexpr unsigned kFullMask = 0xffffffffu;
    const int T = P.pl_ph_hi, RT = P.pl_r_hi;
    if (!tok) {
        bool c = true;
#pragma unroll
        for (int r = 0; r < R; ++r)
            c &= is_th_eq(S[r].q1h) & is_th_eq(S[r].q2h) & (hi_abs(S[r].p1h) < T) & (__double2hiint(S[r].q2r) >= RT) &
                 (hi_abs(S[r].p1p) >= pp_hi[r]);
        tok = __all_sync(kFullMask, c);
    }
    const bool v1 = tok != 0;
    FlowHead f[R];
    // ---- B(d/2) ----
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = flow_head(S[r].q2r, S[r].p1t, S[r].p1r, S[r].q1t, S[r].q1r, P.h, P.h_rs, P.rs);
    if (!v1) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Memo2 &m = M[r * mstride];
            memo_refresh(S[r].q2h, P.h, m.m2);
            flow_tail(f[r], S[r].p1h, S[r].p1p, m.m2.is, m.m2.hcis, S[r].p2r, S[r].p2h, S[r].q1h, S[r].q1p, P.h);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) flow_tail_planar(f[r], S[r].p1p, P.pl_hcB, S[r].p2r, S[r].p2h, S[r].q1p, P.h);
    }
    // ---- C(d) ----
    bool c2 = v1;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        mix(S[r].q1r, S[r].p1r, S[r].q2r, S[r].p2r, P.mg, P.mhs);
        mix(S[r].q1p, S[r].p1p, S[r].q2p, S[r].p2p, P.mg, P.mhs);
        c2 &= (hi_abs(S[r].p2h) < T) & (__double2hiint(S[r].q2r) >= RT) & (hi_abs(S[r].p1p) >= pp_hi[r]);
    }
    const bool v2 = __all_sync(kFullMask, c2);
#pragma unroll
    for (int r = 0; r < R; ++r) mix(S[r].q1t, S[r].p1t, S[r].q2t, S[r].p2t, P.mg, P.mhs);
    if (!v2) {
#pragma unroll
        for (int r = 0; r < R; ++r) mix(S[r].q1h, S[r].p1h, S[r].q2h, S[r].p2h, P.mg, P.mhs);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) mix_planar(S[r].p1h, S[r].p2h, P.mg, P.pl_zero_hs);
    }
    // ---- B(d/2) ----
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = flow_head(S[r].q2r, S[r].p1t, S[r].p1r, S[r].q1t, S[r].q1r, P.h, P.h_rs, P.rs);
    if (!v2) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Memo2 &m = M[r * mstride];
            memo_refresh(S[r].q2h, P.h, m.m2);
            flow_tail(f[r], S[r].p1h, S[r].p1p, m.m2.is, m.m2.hcis, S[r].p2r, S[r].p2h, S[r].q1h, S[r].q1p, P.h);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) flow_tail_planar(f[r], S[r].p1p, P.pl_hcB, S[r].p2r, S[r].p2h, S[r].q1p, P.h);
    }
    // ---- A(d) ----
    bool c4 = v2;
#pragma unroll
    for (int r = 0; r < R; ++r) c4 &= (hi_abs(S[r].p2h) < T) & (__double2hiint(S[r].q1r) >= RT) & (hi_abs(S[r].p2p) >= pp_hi[r]);
    const bool v4 = __all_sync(kFullMask, c4);
#pragma unroll
    for (int r = 0; r < R; ++r) f[r] = flow_head(S[r].q1r, S[r].p2t, S[r].p2r, S[r].q2t, S[r].q2r, P.delta, P.d_rs, P.rs);
    if (!v4) {
        ++n_general;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            Memo2 &m = M[r * mstride];
            memo_refresh(S[r].q1h, P.delta, m.m1);
            flow_tail(f[r], S[r].p2h, S[r].p2p, m.m1.is, m.m1.hcis, S[r].p1r, S[r].p1h, S[r].q2h, S[r].q2p, P.delta);
        }
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r) flow_tail_planar(f[r], S[r].p2p, P.pl_hcA, S[r].p1r, S[r].p1h, S[r].q2p, P.delta);
    }
    bool nx = v4;
#pragma unroll
    for (int r = 0; r < R; ++r) nx &= (hi_abs(S[r].p1h) < T) & (__double2hiint(S[r].q2r) >= RT) & (hi_abs(S[r].p1p) >= pp_hi[r]);
    next = nx;
}

// B(d/2) C(d) B(d/2): the part of _fantasy_step_ord2 (cuda_geodesic.py:569-575) between the two
// A half-flows.  After it q1 is final for this step (the trailing A does not touch q1), so the
// termination test may equally be made after the merged A(d) that follows.
template <bool kMemo>
__device__ __forceinline__ void step_BCB(RayState &S, const StepParams &P, Memo2 &M)
{
    flow_B<kMemo>(S, P.h, P.h_rs, P.rs, M);
    flow_C(S, P.mg, P.mhs);
    flow_B<kMemo>(S, P.h, P.h_rs, P.rs, M);
}

// Termination test of cuda_geodesic.py:598 / :623 on copy 1, plus NaN (a NaN r never terminates
// in the reference and ends as an all-NaN state at the step cap; we stop at once, see D-9).
__device__ __forceinline__ bool terminated(double r, const StepParams &P)
{
    return (r <= P.rcap) || (r >= P.r_max);
}
// Cheap necessary condition for the test below, on the high 32 bits of r only (two integer compares instead of two
// FP64 compares per step): when it is false the ray is certainly still strictly inside (rcap, r_max).  The host sets
// hi_rcap / hi_rmax so that this holds for every r, inf and NaN included (make_step_params).
__device__ __forceinline__ bool maybe_finished(double r, const StepParams &P)
{
    const int hr = __double2hiint(r);
    return !((hr > P.hi_rcap) & (hr < P.hi_rmax));
}
// "not strictly inside (rcap, r_max)": the same test with NaN counted as finished (two compares).
__device__ __forceinline__ bool finished_or_nan(double r, const StepParams &P)
{
    return !((r > P.rcap) && (r < P.r_max));
}

__device__ __forceinline__ void init_state(RayState &S, const double q0[4], const double p0[4])
{
    S.q1t = S.q2t = q0[0]; S.q1r = S.q2r = q0[1]; S.q1h = S.q2h = q0[2]; S.q1p = S.q2p = q0[3];
    S.p1t = S.p2t = p0[0]; S.p1r = S.p2r = p0[1]; S.p1h = S.p2h = p0[2]; S.p1p = S.p2p = p0[3];
}

// ---------------------------------------------------------------------------------------------
// EinsteinPy coordinate helpers (einsteinpy.coordinates.utils, restated)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void c2s(double x, double y, double z, double &r, double &th, double &ph)
{
    const double hxy = hypot(x, y);
    r = hypot(hxy, z);
    th = atan2(hxy, z);
    ph = atan2(y, x);
}

// Python float modulo: result takes the sign of the divisor.
__device__ __forceinline__ double pymod(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) { if ((m < 0.0) != (b < 0.0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}

// ---------------------------------------------------------------------------------------------
// Stages 1+2: pixel (i,j) -> q0, p0, alpha0, beta (+ diagnostic heading)
//   camera plane            raytracing.py:74-88
//   get_initial_conditions  utils.py:91-141;  angles_to_p_sph :148-194;
//   build_null_4momentum_ep_sph :39-88 (future=True)
// ---------------------------------------------------------------------------------------------
// error bits raised by the device code and reported by the host entry points (Counters::err in schw_kernels.cu)
constexpr unsigned kErrTexWait  = 1u;   // the texture-ready flag never arrived (wait_flag_warp timed out)
constexpr unsigned kErrRayPlane = 2u;   // utils.py:113,126: the rotated ray does not lie in the equatorial plane
constexpr unsigned kErrNullCond = 4u;   // utils.py:80-81: negative discriminant of the null condition

struct RayIC {
    unsigned err;      // kErr* bits of this ray (0 for every finite camera ray)
    double p0[4];
    double alpha;      // camera angle pi - h_phi: p0 is a function of it alone (key of the geodesic de-duplication)
    double alpha0, beta;
    double dir[3];     // normalised un-rotated ray (for the diagnostic h_r,h_theta,h_phi)
};

__device__ __forceinline__ void initial_conditions_px(const SceneParams &sc, double px, double py, double pz, RayIC &ic)
{
    double d0 = px - sc.obs[0], d1 = py - sc.obs[1], d2 = pz - sc.obs[2];
    const double nrm = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    d0 /= nrm; d1 /= nrm; d2 /= nrm;
    const double beta = atan2(d2, d1);
    double s, c;
    sincos(-beta, &s, &c);
    const double rx = d0;
    const double ry = c * d1 - s * d2;
    // utils.py:113 `assert abs(ray_xy[2]) < 1e-6` and :126 `assert abs(h_theta - pi/2) < 1e-6`: for a unit vector the
    // second (|atan2(rz, hypot(rx, ry))| < 1e-6) holds whenever the first does, up to the last digit of the 1e-6; both
    // fail on NaN (pixel == observer).  Reported as SCHW_E_RAYPLANE by the host entry points.
    const double rz = s * d1 + c * d2;
    ic.err = (fabs(rz) < 1e-6) ? 0u : kErrRayPlane;
    // h_phi = atan2(y', x') of the rotated ray (cartesian_to_spherical_fast, utils.py:124)
    const double h_phi = atan2(ry, rx);
    const double alpha = 3.141592653589793 - h_phi;
    double sa, ca;
    sincos(alpha, &sa, &ca);
    const double pr  = (-ca) * sc.f_r;
    const double pth = -0.0 * sc.r_obs;          // n_thhat = -sin(0.0)
    const double pph = sa * sc.r_obs;
    const double C = sc.grr * pr * pr + sc.gthth * pth * pth + sc.gphph * pph * pph;
    const double disc = sc.m4A * C;
    if (disc < 0.0) ic.err |= kErrNullCond;      // utils.py:80-81 (unreachable for f > 0; NaN passes, as in the reference)
    ic.p0[0] = sqrt(disc) / sc.twomA;
    ic.p0[1] = pr; ic.p0[2] = pth; ic.p0[3] = pph;
    ic.alpha = alpha;
    ic.alpha0 = acos(-pr / sc.f_r);
    ic.beta = beta;
    ic.dir[0] = d0; ic.dir[1] = d1; ic.dir[2] = d2;
}

__device__ __forceinline__ void initial_conditions(const SceneParams &sc, int i, int j, RayIC &ic)
{
    const double u = (j + 0.5) / sc.width - 0.5;
    const double v = (i + 0.5) / sc.height - 0.5;
    initial_conditions_px(sc, sc.pc[0], sc.pc[1] + u * sc.plane_width, sc.pc[2] + v * sc.plane_height, ic);
}

// ---------------------------------------------------------------------------------------------
// Stage 5: rotate back by +beta, classify, texel lookup   (raytracing.py:166-174,181-273)
// ---------------------------------------------------------------------------------------------
struct RayHit {
    double th_hit, ph_hit;   // as written to photon_data.csv
    int32_t u, v;            // texel or -1
    uint8_t cls, r, g, b;
};

__device__ __forceinline__ void classify(const SceneParams &sc, const uint8_t *__restrict__ tex,
                                         double r_bh, double th, double ph, double alpha0, double beta,
                                         RayHit &o)
{
    double sph, cph, sth, cth, sb, cb;
    sincos(ph, &sph, &cph);
    sincos(th, &sth, &cth);
    const double x = r_bh * cph * sth;
    const double y = r_bh * sph * sth;
    const double z = r_bh * cth;
    sincos(beta, &sb, &cb);
    const double y2 = cb * y - sb * z;
    const double z2 = sb * y + cb * z;
    double rr, th_hit, ph_hit;
    c2s(x, y2, z2, rr, th_hit, ph_hit);
    o.u = o.v = -1;
    o.r = o.g = o.b = 0;
    if ((r_bh <= sc.rs12) || (alpha0 <= sc.bh_angle)) {
        o.cls = 0;
    } else if (r_bh >= 100.0) {
        o.cls = 1; o.r = 255;
    } else if (r_bh >= sc.r_max) {
        o.cls = 3;
        if (sc.has_texture) {
            const double two_pi = 6.283185307179586;
            th_hit = pymod(th_hit, two_pi);
            ph_hit = pymod(ph_hit, two_pi);
            const double dtheta = fabs(th_hit - sc.pct);
            if (sc.flip_phi) ph_hit = -ph_hit;
            const double phi_rel = pymod(ph_hit - sc.phi0, two_pi);
            const double dphi = fabs(pymod(ph_hit - sc.pcp + 3.141592653589793, two_pi) - 3.141592653589793);
            if ((dtheta <= sc.half_pst) && (dphi <= sc.half_psp)) {
                const double theta_map = sc.flip_theta ? (3.141592653589793 - th_hit) : th_hit;
                const double fu = (theta_map - sc.theta0) / (sc.theta1 - sc.theta0) * (sc.height - 1) + 0.5;
                const double fv = phi_rel / sc.phi_span * (sc.width - 1) + 0.5;
                long long uu = (long long)fu, vv = (long long)fv;       // int(): truncation toward zero
                uu = uu < 0 ? 0 : (uu > sc.height - 1 ? sc.height - 1 : uu);
                vv = vv < 0 ? 0 : (vv > sc.width - 1 ? sc.width - 1 : vv);
                const uint8_t *t = tex + ((size_t)uu * sc.width + vv) * 3;
                o.r = __ldg(t); o.g = __ldg(t + 1); o.b = __ldg(t + 2);
                o.u = (int32_t)uu; o.v = (int32_t)vv;
                o.cls = 2;
            }
        }
    } else {
        o.cls = 4;
    }
    o.th_hit = th_hit; o.ph_hit = ph_hit;
}

}  // namespace schw
