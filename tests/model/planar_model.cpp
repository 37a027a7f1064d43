// Host-compiled model of the step loop (test infrastructure): csrc/schw_device.cuh built with g++, one lane standing
// for a warp.  It integrates one ray twice with the SAME arithmetic -- once through the general flows only, once
// through step_BCBA_planar driven exactly like persistent_rays drives it -- and reports the first step at which the
// two states differ in any bit (-1 if none), so the "the fast path is a no-op elimination" claim can be checked on the
// CPU over many parameter sets.  fast_rcp uses a stand-in seed (the hardware's is not available), which is irrelevant
// for this comparison: both integrations share it.
#define SCHW_HOST_MODEL 1
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __forceinline__ inline
#define __constant__ static const
#define __restrict__
#define __ldg(p) (*(p))
static inline int __double2hiint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo)
{
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; memcpy(&x, &u, 8); return x;
}
static inline long long __double_as_longlong(double x) { long long u; memcpy(&u, &x, 8); return u; }
static inline double __longlong_as_double(long long u) { double x; memcpy(&x, &u, 8); return x; }
static inline bool __all_sync(unsigned, bool p) { return p; }
using std::fma;

#include "../../schwarzchild-einsteinpy-raytracing_b200/csrc/schw_device.cuh"

using namespace schw;

// bitwise equality of the 16 doubles; two NaNs count as equal whatever their payload (the kernel emits one canonical
// NaN for a ray whose r went NaN)
static bool same_bits(const RayState &a, const RayState &b)
{
    const double *x = &a.q1t, *y = &b.q1t;
    for (int k = 0; k < 16; ++k)
        if (memcmp(x + k, y + k, 8) != 0 && !(x[k] != x[k] && y[k] != y[k])) return false;
    return true;
}

// P: the bytes of schw::StepParams from schw_debug_step_params.  out: [0] steps taken, [1] steps entirely on the fast
// path, [2] first differing step or -1, [3] 1 if both integrations ended by the same test at the same step.
// fin_general / fin_planar: (t, r, theta, phi) of copy 1 at the end.
extern "C" int model_run_ray(const void *params, size_t params_size, const double *q0, const double *p0,
                             long long *out, double *fin_general, double *fin_planar)
{
    if (params_size != sizeof(StepParams)) return -1;
    StepParams P;
    memcpy(&P, params, sizeof(P));
    RayState G, F;
    init_state(G, q0, p0);
    init_state(F, q0, p0);
    Memo2 MG, MF;
    MG.init(P.delta, P.h);
    MF.init(P.delta, P.h);
    long long n = 0, planar_steps = 0, first_diff = -1;
    bool same_end = true;
    if (!(P.steps <= 0 || terminated(G.q1r, P))) {
        flow_A<false>(G, P.h, P.h_rs, P.rs, MG);          // leading A(d/2), done in the refill by the kernel
        flow_A<false>(F, P.h, P.h_rs, P.rs, MF);
        int tok = 0, n_general = 0;
        const int pp_hi = P.pl_pp_hi;
        while (true) {
            step_BCB<true>(G, P, MG);
            flow_A<true>(G, P.delta, P.d_rs, P.rs, MG);
            bool next = false;
            if (P.pl_enable) {
                const int before = n_general;
                step_BCBA_planar(F, P, MF, pp_hi, tok, next, n_general);
                planar_steps += (n_general == before);
            } else {
                step_BCB<true>(F, P, MF);
                flow_A<true>(F, P.delta, P.d_rs, P.rs, MF);
            }
            ++n;
            if (first_diff < 0 && !same_bits(G, F)) first_diff = n;
            const bool maybe = (n >= P.steps) | maybe_finished(F.q1r, P);
            tok = P.pl_enable ? (next & !maybe) : 0;
            const bool done_f = maybe && (n >= P.steps || finished_or_nan(F.q1r, P));
            const bool done_g = n >= P.steps || finished_or_nan(G.q1r, P);
            if (done_f != done_g) same_end = false;
            if (done_f || done_g) break;
        }
    }
    out[0] = n; out[1] = planar_steps; out[2] = first_diff; out[3] = same_end ? 1 : 0;
    fin_general[0] = G.q1t; fin_general[1] = G.q1r; fin_general[2] = G.q1h; fin_general[3] = G.q1p;
    fin_planar[0] = F.q1t; fin_planar[1] = F.q1r; fin_planar[2] = F.q1h; fin_planar[3] = F.q1p;
    return 0;
}
