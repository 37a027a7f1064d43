// schw_kernels.cu — sm_100a kernels and the C ABI (include/schw_b200.h) of libschw_b200.so.
//
// Kernel inventory:
//   persistent_tiles<RenderJob,R>   camera IC -> FANTASY stepper -> terminate -> classify -> texel: THE hot path (whole
//                                   32-pixel tiles per warp, R = 2 rays per thread on large launches)
//   persistent_tiles<BatchJob,R>    the distinct rays of a de-duplicated frame (coherent: sorted by camera angle)
//   persistent_rays<Job,memo,planar>  per-lane refill: integrate_batch on caller-supplied (q0,p0); render with an
//                                   observer off the equatorial plane; A/B switch SCHW_TILE_MODE
//   integrate_full_planar_kernel<2>, integrate_full_kernel<planar>, integrate_rows_kernel
//                                   integrate_batch_full (per-step trajectory store, full 128-byte lines) and its decimated form
//   dedup_{key,heads,scatter,epilogue}_kernel   geodesic de-duplication (sort/scan by cub)
//   ic_kernel / ic_px_kernel / classify_kernel  stages 1+2 and 5 alone (boundary + parity testing)
//   flat_kernel / flat_dirs_kernel / flat_traj_kernel   flat-space comparison render
//   dfma_probe_kernel, step_probe_kernel<R,CTAs>, math_probe_kernel   measurement and accuracy probes
//
// Scheduling: FP64-pipe bound, no reuse, ~4 B of output per ~1.5 M flops => no shared-memory tiling or TMA staging is
// useful.  What matters is keeping all 64 FP64 lanes of every SM busy although rays need 3 k .. 200 k steps: the kernels
// are persistent (grid = SMs x resident CTAs) over a global queue of 32-pixel tiles in longest-first order; a warp takes
// whole tiles (neighbouring pixels need almost the same number of steps) or, for incoherent batches, every lane refills
// itself the moment its ray ends (warp-ballot compaction + one aggregated atomic per warp).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <new>
#include <vector>
#include <algorithm>
#include <utility>
#include <type_traits>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "schw_device.cuh"
#include "../../include/schw_b200.h"

namespace schw {

#ifndef SCHW_BLOCK
#define SCHW_BLOCK 128
#endif
#ifndef SCHW_MINBLOCKS
#define SCHW_MINBLOCKS 6   // 7 (72 registers, spills) measures the same, 5 is 1.7 % slower
#endif
#ifndef SCHW_MINBLOCKS_GENERAL
#define SCHW_MINBLOCKS_GENERAL 5     // kernels without the planar fast path keep the trig memo in registers (94 of them)
#endif
#ifndef SCHW_MEMO_SMEM
#define SCHW_MEMO_SMEM 1
#endif
constexpr int kBlock = SCHW_BLOCK;   // threads per CTA (tuning variants: make VARIANT_FLAGS=-DSCHW_BLOCK=.. OUT=..)
constexpr unsigned kFull = 0xffffffffu;

struct Counters {
    unsigned long long next;         // ray queue head
    unsigned long long total_steps;  // integrator steps executed
    unsigned long long warp_iters;   // loop bodies executed, summed over warps (32 x this = lane-steps issued)
    unsigned long long planar_iters; // of those, completed entirely on the planar fast path
    int batch_planar;                // verdict of batch_probe_kernel
    unsigned err;                    // kErr* bits raised by the device code of this call (schw_device.cuh)
    unsigned long long pixel_steps;  // de-duplicated render: steps summed over PIXELS (what one ray per pixel would take)
};

// ---------------------------------------------------------------------------------------------
// Jobs: where a ray comes from and where its result goes.
// ---------------------------------------------------------------------------------------------
struct BatchJob {
    const double *__restrict__ q0s;
    const double *__restrict__ p0s;
    double *__restrict__ out_qs;
    int32_t *__restrict__ out_steps;
    long long n;
    // Kernel selection on the device: when `select` is set, this launch runs only if *select == want (see
    // batch_probe_kernel / integrate_batch_impl); otherwise the whole grid exits at once.
    const int *__restrict__ select;
    int want;

    __device__ __forceinline__ bool skip() const { return select && *select != want; }
    __device__ __forceinline__ const unsigned *ready_flag() const { return nullptr; }
    __device__ __forceinline__ unsigned *err_word() const { return nullptr; }
    __device__ __forceinline__ bool whole_tiles() const { return false; }
    __device__ __forceinline__ bool ends_at_once(const StepParams &) const { return false; }
    __device__ __forceinline__ long long queue_len() const { return n; }
    __device__ __forceinline__ long long resolve(long long q) const { return q; }
    // load(): initial (q0, p0) of ray k into this thread's shared-memory staging area st[0..3*kBlock]
    // (st[0] = (q0_t,q0_r), st[kBlock] = (q0_th,q0_ph), st[2 kBlock] = (p0_t,p0_r), st[3 kBlock] = (p0_th,p0_ph)).
    __device__ __forceinline__ void load(long long k, double2 *st, double2 &) const
    {
        const double2 *q = reinterpret_cast<const double2 *>(q0s + 4 * k);
        const double2 *p = reinterpret_cast<const double2 *>(p0s + 4 * k);
        st[0] = __ldg(q); st[kBlock] = __ldg(q + 1); st[2 * kBlock] = __ldg(p); st[3 * kBlock] = __ldg(p + 1);
    }
    // cost of one divergent service pass (load + store) in lane-steps, for the batching rule of the kernel
    static constexpr int kServiceCost = 32;
    // caller-supplied rays may have any polar motion: keep the trig memo in registers
    static constexpr bool kMemoSmem = false;
    __device__ __forceinline__ void store(long long k, double t, double r, double th, double ph, int nsteps, bool nan_exit,
                                          const double2 &) const
    {
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        double2 *o = reinterpret_cast<double2 *>(out_qs + 4 * k);
        o[0] = nan_exit ? make_double2(nanv, nanv) : make_double2(t, r);
        o[1] = nan_exit ? make_double2(nanv, nanv) : make_double2(th, ph);
        if (out_steps) out_steps[k] = nsteps;
    }
};

// Warp-uniform bounded spin (about 2 s) until the word at `flag` is non-zero.  The flag is written by a copy-engine
// memcpy that is stream-ordered after the texture upload, so a frame's integration starts while its texture is still
// crossing PCIe; the first rays end milliseconds later, normally long after the flag.  If the flag never arrives
// (a stalled pageable upload, a time-sliced GPU) the texels that follow may be stale: kErrTexWait is raised in *err and
// every host entry point / schw_last_status turns it into SCHW_E_CUDA -- never a silently wrong frame.
// The exit test is a vote, so the loop is convergent by construction (call from convergent code only).
__device__ __forceinline__ void wait_flag_warp(const unsigned *flag, unsigned *err)
{
    const volatile unsigned *f = flag;
    bool seen = false;
    for (int spin = 0; spin < (1 << 22) && !(seen = __any_sync(0xffffffffu, *f != 0u)); ++spin) __nanosleep(512);
    if (!seen) atomicOr(err, kErrTexWait);
    __threadfence();
}

struct RenderJob {
    SceneParams sc;
    const uint8_t *__restrict__ tex;
    const unsigned *tex_ready;          // NULL, or the flag that says the texture upload has landed (see wait_flag)
    unsigned *err;                      // Counters::err of the call
    uint8_t *__restrict__ rgb;
    uint8_t *__restrict__ cls;
    long long rgb_pitch, cls_pitch;     // bytes between visited rows (3w and w when packed; 4w for both inside a (rows, 4w)
                                        // rgb ++ cls gather block, see schw_render_dev_pitched)
    double *__restrict__ final_q;
    double *__restrict__ ic;
    double *__restrict__ hit;
    int32_t *__restrict__ texel;
    int32_t *__restrict__ nsteps;
    long long n;
    // Work order: the queue hands out 32-pixel row segments ("tiles") in longest-processing-time-first order
    // (tile_order, built on the host from the camera geometry, see build_tile_order).  Slot q of the queue is
    // pixel (q & 31) of tile tile_order[q >> 5]; slots past the end of a row are empty.
    const int32_t *__restrict__ tile_order;
    int32_t tiles_per_row;
    // Refill policy.  0: every lane takes its next ray the moment its own ray ends (best when a lane handles many rays: the
    // lanes of a warp drift apart, but the queue keeps them all busy).  1: a warp takes a whole 32-pixel tile only when all
    // its rays have ended -- neighbouring pixels need almost the same number of steps, so the warp stays full until it
    // retires and the frame does not end with every warp carrying a few stragglers (small frames / shards: few rays per lane).
    int32_t tile_mode;
    long long qlen;

    __device__ __forceinline__ bool skip() const { return false; }
    // the flag a warp must see before its first epilogue (texel lookup), or NULL
    __device__ __forceinline__ const unsigned *ready_flag() const { return tex_ready; }
    __device__ __forceinline__ unsigned *err_word() const { return err; }
    __device__ __forceinline__ bool whole_tiles() const { return tile_mode != 0; }
    // every ray of the frame starts at r_obs: do the rays end before their first step (cuda_geodesic.py:598)?
    __device__ __forceinline__ bool ends_at_once(const StepParams &P) const { return P.steps <= 0 || terminated(sc.r_obs, P); }
    __device__ __forceinline__ long long queue_len() const { return qlen; }
    __device__ __forceinline__ long long resolve(long long q) const
    {
        if (!tile_order) return q;
        const int tile = __ldg(tile_order + (q >> 5));
        const int row = tile / tiles_per_row, seg = tile - row * tiles_per_row;
        const int j = seg * 32 + (int)(q & 31);
        return j < sc.width ? (long long)row * sc.width + j : -1;
    }
    // slot = this thread's shared-memory (alpha0, beta): kept out of registers across the step loop.
    __device__ __forceinline__ void load(long long k, double2 *st, double2 &slot) const
    {
        const int row = (int)(k / sc.width), j = (int)(k - (long long)row * sc.width);
        const int i = sc.row_begin + row * sc.row_stride;
        RayIC r;
        initial_conditions(sc, i, j, r);
        st[0] = make_double2(0.0, sc.r_obs); st[kBlock] = make_double2(sc.th_obs, sc.ph_obs);
        st[2 * kBlock] = make_double2(r.p0[0], r.p0[1]); st[3 * kBlock] = make_double2(r.p0[2], r.p0[3]);
        slot = make_double2(r.alpha0, r.beta);
        if (r.err) atomicOr(err, r.err);
        if (ic) {
            double *o = ic + 9 * k;
            o[0] = r.p0[0]; o[1] = r.p0[1]; o[2] = r.p0[2]; o[3] = r.p0[3];
            o[4] = r.alpha0; o[5] = r.beta;
            c2s(r.dir[0], r.dir[1], r.dir[2], o[6], o[7], o[8]);
        }
    }
    // one service pass = classification epilogue + camera/initial-condition prologue, ~1500 instructions executed for
    // however few lanes need it = ~9 warp-steps = ~288 lane-steps
    static constexpr int kServiceCost = 288;
    // the planar kernel is only launched for an observer in the equatorial plane (schw_render_dev): every ray is planar
    static constexpr bool kMemoSmem = true;
    __device__ __forceinline__ void store(long long k, double t_, double r_, double th_, double ph_, int ns, bool nan_exit,
                                          const double2 &slot) const
    {
        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
        const double t = nan_exit ? nanv : t_, r = nan_exit ? nanv : r_;
        const double th = nan_exit ? nanv : th_, ph = nan_exit ? nanv : ph_;
        RayHit o;
        classify(sc, tex, r, th, ph, slot.x, slot.y, o);
        const int row = (int)(k / sc.width), j = (int)(k - (long long)row * sc.width);
        uint8_t *c3 = rgb + row * rgb_pitch + 3 * j;
        c3[0] = o.r; c3[1] = o.g; c3[2] = o.b;
        cls[row * cls_pitch + j] = o.cls;
        if (final_q) {
            double2 *q = reinterpret_cast<double2 *>(final_q + 4 * k);
            q[0] = make_double2(t, r); q[1] = make_double2(th, ph);
        }
        if (hit) *reinterpret_cast<double2 *>(hit + 2 * k) = make_double2(o.th_hit, o.ph_hit);
        if (texel) *reinterpret_cast<int2 *>(texel + 2 * k) = make_int2(o.u, o.v);
        if (nsteps) nsteps[k] = ns;
    }
};

// ---------------------------------------------------------------------------------------------
// The persistent ray kernel.
//
// Loop structure per lane (reference: fantasy_integrate_batch_ord2, cuda_geodesic.py:579-602):
//     fetch ray -> [test] A(d/2) | B(d/2) C(d) B(d/2) A(d) [count, test] | B C B A(d) [test] | ...
// A(d) = the trailing A(d/2) of one step merged with the leading A(d/2) of the next (A does not touch q1, so the
// test on q1 may follow it; the A applied after a ray's last step is discarded with the rest of its state).
// All 32 lanes execute the same B,C,B,A sequence; only the service (epilogue + refill) is divergent.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void park(RayState &S, const StepParams &P)
{
    const double q0[4] = {0.0, P.park_r, 1.5707963267948966, 0.0}, p0[4] = {0.0, 0.0, 0.0, 0.0};
    init_state(S, q0, p0);
}

template <class Job, bool kMemo, bool kPlanar>
__global__ void __launch_bounds__(kBlock, kPlanar ? SCHW_MINBLOCKS : SCHW_MINBLOCKS_GENERAL)
persistent_rays(const __grid_constant__ Job job, const __grid_constant__ StepParams P, Counters *ctr)
{
    if (job.skip()) return;                   // (grid-uniform)
    __shared__ double2 s_slot[kBlock];        // (alpha0, beta) of the ray in flight, for its epilogue
    __shared__ double2 s_fin[4 * kBlock];     // staging: final (t,r),(theta,phi) of a finished ray awaiting service;
                                              // (q0, p0) of a fetched ray between load() and the state registers
    __shared__ int s_nsteps[kBlock];
    __shared__ unsigned long long s_total;    // steps integrated by this CTA (metric numerator)
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();
    double2 &slot = s_slot[threadIdx.x];
    const unsigned lane = threadIdx.x & 31u;
    RayState S;
    // On the planar fast path the trig memo is not touched, so for jobs known to live on it (Job::kMemoSmem) it sits in
    // shared memory (12 registers less, which buys the 6th resident CTA).
    constexpr bool kSmemMemo = SCHW_MEMO_SMEM && kPlanar && Job::kMemoSmem;
    __shared__ Memo2 s_memo[kSmemMemo ? kBlock : 1];
    Memo2 M_reg;
    Memo2 &M = kSmemMemo ? s_memo[threadIdx.x] : M_reg;
    M.init(P.delta, P.h);        // m1 serves the merged A(delta) of the step loop, m2 the two B(delta/2)
    int k = -1;                  // ray in flight, or finished ray awaiting service (n < 2^31 is checked on the host)
    int it_end = 0x7fffffff;     // value of `it` at which the ray in flight reaches the step cap (never, while parked); the
                                 // steps it has taken are it - s_nsteps[threadIdx.x] (= `it` when it was fetched)
    bool active = false;         // integrating a ray
    bool waiting = false;        // ray finished: result parked in shared memory, epilogue + refill pending
    bool drained = false;        // queue is empty
    int it = 0;                  // warp-uniform iteration counter
    int debt = 0;                // lane-steps spent idle by waiting lanes since the last service (warp-uniform)
    int tok = 0;                 // planar fast path: the next step may start on it (warp-uniform, see step_BCBA_planar)
    int n_general = 0;           // iterations that left the planar fast path (warp-uniform)
    int pp_hi = 0;               // planar fast path: this lane's |p_phi| threshold (0 while parked)

    // Lanes without a ray sit on a fixed point of all three flows (p = 0 => dq = dH = 0, mixing of
    // equal copies is the identity), so they execute the same instructions without ever producing
    // inf/NaN operands (which would send the trig/reciprocal code down slow paths for the whole warp).
    park(S, P);

    bool inputs_ready = job.ready_flag() == nullptr;     // (warp-uniform) the texture upload of this frame has landed
    while (true) {
        // ---- service (divergent, so batched): epilogue of the finished rays, then refill ----
        if (!inputs_ready && (job.ends_at_once(P) || __any_sync(kFull, waiting))) {
            wait_flag_warp(job.ready_flag(), job.err_word());                      // the first epilogue of this warp: its texel lookup
            inputs_ready = true;
        }
        if (waiting) {
            const double2 a = s_fin[threadIdx.x], b = s_fin[kBlock + threadIdx.x];
            job.store(k, a.x, a.y, b.x, b.y, s_nsteps[threadIdx.x], !(a.y == a.y), slot);
            waiting = false; k = -1;
        }
        // (whole tiles only: nobody is refilled while a ray of the warp is still running, see RenderJob::tile_mode; voted
        //  once, outside the refill loop -- a second vote inside it costs the compiler its proof that the step loop's votes
        //  are reached convergently, and with it 13 % of the frame time: tests/test_abi_symbols.py::test_step_loop_sass_budget)
        const bool hold = job.whole_tiles() && __any_sync(kFull, active);
        while (true) {                                             // loop: a ray may end at step 0
            const unsigned need = __ballot_sync(kFull, !active && !drained && !hold);
            if (need == 0u) break;
            long long base = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = (long long)atomicAdd(&ctr->next, (unsigned long long)__popc(need));
            base = __shfl_sync(kFull, base, leader);
            if (!active && !drained) {
                const long long mine = base + __popc(need & ((1u << lane) - 1u));
                if (mine < job.queue_len()) {
                    const int kk = (int)job.resolve(mine);
                    if (kk >= 0) {                                 // (< 0: empty slot at the ragged end of a row)
                        job.load(kk, &s_fin[threadIdx.x], slot);
                        {
                            const double2 qa = s_fin[threadIdx.x], qb = s_fin[kBlock + threadIdx.x];
                            const double2 pa = s_fin[2 * kBlock + threadIdx.x], pb = s_fin[3 * kBlock + threadIdx.x];
                            const double q0[4] = {qa.x, qa.y, qb.x, qb.y}, p0[4] = {pa.x, pa.y, pb.x, pb.y};
                            init_state(S, q0, p0);
                        }
                        if (P.steps <= 0 || terminated(S.q1r, P)) {
                            // cuda_geodesic.py:598: the test precedes the first step
                            job.store(kk, S.q1t, S.q1r, S.q1h, S.q1p, 0, false, slot);
                            park(S, P);
                        } else {
                            // leading A(d/2) of the ray's first step (cuda_geodesic.py:571); every later A is the merged
                            // trailing+leading A(d) at the end of the loop body.  Divergent code: no memo vote here.
                            flow_A<false>(S, P.h, P.h_rs, P.rs, M);
                            k = kk; active = true;
                            s_nsteps[threadIdx.x] = it;
                            it_end = P.steps <= 0x7fffffff - it ? it + P.steps : 0x7fffffff;
                        }
                    }
                } else {
                    drained = true;
                }
            }
        }
        if (__ballot_sync(kFull, active) == 0u) break;             // every lane is out of work
        debt = 0;
        tok = 0;                                                   // new rays on board: re-establish (R0)
        pp_hi = active ? P.pl_pp_hi : 0;

        // ---- integrate; finished lanes park their result and idle until servicing them pays ----
        while (true) {
            bool next = false;
            if (kPlanar) {
                step_BCBA_planar(S, P, M, pp_hi, tok, next, n_general);
            } else {
                step_BCB<kMemo>(S, P, M);                           // q1 is now final for this step
                flow_A<kMemo>(S, P.delta, P.d_rs, P.rs, M);         // trailing A(d/2) + next step's leading A(d/2)
            }
            ++it;
            // termination (cuda_geodesic.py:598): a two-compare integer pre-test on the high word of r every step, the exact
            // FP64 test only in the (rare) steps where some lane's pre-test cannot rule termination out
            // (no `active &`: a parked lane has it_end = INT_MAX and sits at P.park_r, which passes the pre-test whenever a radius
            //  strictly between the thresholds exists -- make_step_params; otherwise it only sends the warp to the exact test)
            const bool maybe = (it >= it_end) | maybe_finished(S.q1r, P);
            const bool tick = (it & 15) == 0;
            // one vote per step on the fast path: "nobody may have finished and everybody may start the next step planar"
            if (kPlanar) tok = __all_sync(kFull, next & !maybe);
            if ((kPlanar && tok) ? tick : (__any_sync(kFull, maybe) || tick)) {
                const bool done = active && maybe && (it >= it_end || finished_or_nan(S.q1r, P));
                const bool any_done = __any_sync(kFull, done);
                // idle cost so far: every waiting lane wastes one lane-step per iteration (booked every 16 iterations)
                const unsigned wm = __ballot_sync(kFull, waiting);
                if (tick) debt += __popc(wm) * 16;
                if (done) {
                    s_fin[threadIdx.x] = make_double2(S.q1t, S.q1r);
                    s_fin[kBlock + threadIdx.x] = make_double2(S.q1h, S.q1p);
                    const int n = it - s_nsteps[threadIdx.x];
                    s_nsteps[threadIdx.x] = n;
                    atomicAdd(&s_total, (unsigned long long)n);
                    it_end = 0x7fffffff;
                    active = false; waiting = true; pp_hi = 0;
                    park(S, P);
                }
                // Service now if nobody is left integrating, or once the accumulated idle cost reaches the cost of
                // one (divergent) service pass -- the ski-rental rule: waiting never costs more than 2x the optimum.
                if ((wm != 0u || any_done) &&
                    ((debt >= Job::kServiceCost && !job.whole_tiles()) || __ballot_sync(kFull, active) == 0u)) break;
            }
        }
    }

    // ---- total step count (metric numerator) and loop statistics ----
    if (lane == 0 && it) {
        atomicAdd(&ctr->warp_iters, (unsigned long long)it);
        if (kPlanar) atomicAdd(&ctr->planar_iters, (unsigned long long)(it - n_general));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(&ctr->total_steps, s_total);
}

// ---------------------------------------------------------------------------------------------
// The render kernel for whole tiles, R rays per thread (the shipped render path when the observer lies in the
// equatorial plane, i.e. for the reference's only camera).
//
// A warp takes R consecutive 32-pixel tiles of the (longest-first) queue at a time and integrates their 32 R rays in
// lock step -- lane l carries pixel l of each of the R tiles -- until the last of them has ended, then classifies them
// all and takes the next R tiles.  Neighbouring pixels, and neighbouring tiles of the sorted queue, need almost the same
// number of steps, so the warp stays full until it retires.  Compared with persistent_rays (per-lane refill):
//   * all rays of a warp start together: ONE step counter per warp, no per-lane bookkeeping in the loop, the service
//     (epilogue + prologue, ~1500 instructions per ray) runs with all 32 lanes instead of a few;
//   * R independent geodesics per thread share the loop control, the four votes, the branches and the constant loads of
//     a loop body (paid once per R ray-steps) and hide each other's FP64 dependency latency.
// Same arithmetic per ray (step_BCBA_planar_n<R> issues exactly the instructions of step_BCBA_planar for each ray), so
// the outputs are bit-identical to persistent_rays' (tests/test_gpu_planar.py, tools/ab_tile.py).
// ---------------------------------------------------------------------------------------------
#ifndef SCHW_TILE_RAYS
#define SCHW_TILE_RAYS 2
#endif
#ifndef SCHW_TILE_MINBLOCKS
#define SCHW_TILE_MINBLOCKS 4
#endif
template <class Job, int R>
__global__ void __launch_bounds__(kBlock, R == 1 ? SCHW_MINBLOCKS : SCHW_TILE_MINBLOCKS)
persistent_tiles(const __grid_constant__ Job job, const __grid_constant__ StepParams P, Counters *ctr)
{
    if (job.skip()) return;                       // (grid-uniform)
    __shared__ double2 s_slot[R * kBlock];        // (alpha0, beta) per ray, for its epilogue
    __shared__ double2 s_fin[R * 2 * kBlock];     // final (t,r),(theta,phi) of an ended ray
    __shared__ double2 s_stage[4 * kBlock];       // (q0, p0) between load() and the state registers
    __shared__ int s_n[R * kBlock];               // steps taken by an ended ray
    __shared__ Memo2 s_memo[R * kBlock];
    __shared__ unsigned long long s_total;
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    RayState S[R];
    int kk[R], pp_hi[R];
    bool active[R];
#pragma unroll
    for (int r = 0; r < R; ++r) s_memo[r * kBlock + threadIdx.x].init(P.delta, P.h);
    int it = 0, n_general = 0;
    unsigned long long mine = 0ull;               // steps integrated by this thread's rays
    bool inputs_ready = job.ready_flag() == nullptr;
#pragma unroll
    for (int r = 0; r < R; ++r) kk[r] = -1;
    while (true) {
        // ---- epilogue of the tiles integrated in the previous iteration: rotate back, classify, texel (all lanes together).
        //      (At the TOP of the loop, like persistent_rays' service: with the divergent per-ray code right in front of the
        //       loop's back edge the compiler no longer proves that the votes of the step loop are reached convergently.)
        if (!inputs_ready && __any_sync(kFull, kk[0] >= 0)) {       // (tile 0 of a group is never empty when the group exists)
            wait_flag_warp(job.ready_flag(), job.err_word());         // the first texel lookup of this warp
            inputs_ready = true;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (kk[r] >= 0) {
                const double2 a = s_fin[(2 * r) * kBlock + threadIdx.x], b = s_fin[(2 * r + 1) * kBlock + threadIdx.x];
                const int ns = s_n[r * kBlock + threadIdx.x];
                job.store(kk[r], a.x, a.y, b.x, b.y, ns, !(a.y == a.y), s_slot[r * kBlock + threadIdx.x]);
                mine += (unsigned long long)ns;
            }
        }
        // ---- R tiles for the warp ----
        long long base = 0;
        if (lane == 0) base = (long long)atomicAdd(&ctr->next, (unsigned long long)(32 * R));
        base = __shfl_sync(kFull, base, 0);
        // (the exit test is a vote so that the compiler can see it is warp-uniform: every *_sync of the step loop must be
        //  reached provably convergently, tests/test_abi_symbols.py::test_step_loop_sass_budget)
        if (__all_sync(kFull, base >= job.queue_len())) break;
        // ---- prologue: camera ray, initial conditions, leading A(d/2) ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const long long q = base + 32 * r + lane;
            kk[r] = q < job.queue_len() ? (int)job.resolve(q) : -1;
            active[r] = false;
            if (kk[r] >= 0) {
                job.load(kk[r], &s_stage[threadIdx.x], s_slot[r * kBlock + threadIdx.x]);
                const double2 qa = s_stage[threadIdx.x], qb = s_stage[kBlock + threadIdx.x];
                const double2 pa = s_stage[2 * kBlock + threadIdx.x], pb = s_stage[3 * kBlock + threadIdx.x];
                const double q0[4] = {qa.x, qa.y, qb.x, qb.y}, p0[4] = {pa.x, pa.y, pb.x, pb.y};
                init_state(S[r], q0, p0);
                if (P.steps <= 0 || terminated(S[r].q1r, P)) {           // cuda_geodesic.py:598: the test precedes the first step
                    s_fin[(2 * r) * kBlock + threadIdx.x] = make_double2(S[r].q1t, S[r].q1r);
                    s_fin[(2 * r + 1) * kBlock + threadIdx.x] = make_double2(S[r].q1h, S[r].q1p);
                    s_n[r * kBlock + threadIdx.x] = 0;
                    park(S[r], P);
                } else {
                    flow_A<false>(S[r], P.h, P.h_rs, P.rs, s_memo[r * kBlock + threadIdx.x]);
                    active[r] = true;
                }
            } else {
                park(S[r], P);
            }
            pp_hi[r] = active[r] ? P.pl_pp_hi : 0;
        }
        // ---- integrate the 32 R rays in lock step: n = steps taken so far by every ray still running ----
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) any |= active[r];
        int n = 0, tok = 0;
        if (__any_sync(kFull, any)) {
            while (true) {
                bool next = false;
                step_BCBA_planar_n<R>(S, P, &s_memo[threadIdx.x], kBlock, pp_hi, tok, next, n_general);
                ++n; ++it;
                const bool cap = n >= P.steps;                                   // (warp-uniform)
                bool maybe = false;
#pragma unroll
                for (int r = 0; r < R; ++r) maybe |= maybe_finished(S[r].q1r, P);   // (parked lanes pass the pre-test: P.park_r)
                tok = __all_sync(kFull, next & !maybe) && !cap;
                if (!tok && (cap || __any_sync(kFull, maybe))) {
                    bool left = false;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        if (active[r] && (cap || finished_or_nan(S[r].q1r, P))) {
                            s_fin[(2 * r) * kBlock + threadIdx.x] = make_double2(S[r].q1t, S[r].q1r);
                            s_fin[(2 * r + 1) * kBlock + threadIdx.x] = make_double2(S[r].q1h, S[r].q1p);
                            s_n[r * kBlock + threadIdx.x] = n;
                            active[r] = false; pp_hi[r] = 0;
                            park(S[r], P);
                        }
                        left |= active[r];
                    }
                    if (!__any_sync(kFull, left)) break;
                }
            }
        }
    }
    // (one atomic per thread, after the loop: a lane-conditional atomic in front of the loop's back edge would cost the
    //  compiler its proof that the votes of the next iteration are reached convergently)
    atomicAdd(&s_total, mine);
    if (lane == 0 && it) {
        atomicAdd(&ctr->warp_iters, (unsigned long long)it * R);      // loop bodies per RAY slot, as persistent_rays counts them
        atomicAdd(&ctr->planar_iters, (unsigned long long)(it - n_general) * R);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(&ctr->total_steps, s_total);
}

// ---------------------------------------------------------------------------------------------
// integrate_batch_full: fantasy_integrate_batch_ord2_full, cuda_geodesic.py:604-625.
// Row `step` of a ray holds q1 *before* the termination test of that step (the terminal state is stored, the state
// after the last of `steps` steps is not); rows after a ray's last stored row are left untouched (the host entry point
// copies only the stored rows into the caller's np.zeros buffer, the _dev one zero-fills on request).
//
// The path is bound by the FP64 pipe like the render (the same B C B A(d) loop body: 1.6e11 steps/s at best) and writes
// 32 B per step per ray, i.e. up to 5 TB/s of the 6.5 TB/s HBM write bandwidth when the GPU is full -- so the stores
// must not cost issue slots or sector efficiency: one ray per thread, rows staged in shared memory four steps at a
// time (2 STS.128 per step, XOR-swizzled: conflict-free) and flushed by the warp as FULL 128-BYTE LINES (8 lanes x 16 B
// per ray and group, 8 coalesced STG.128 per thread every 4 steps, streaming).  Lanes whose ray has ended sit on the
// fixed point of the flows (park) until the warp's last ray ends; the planar fast path is used when the rays allow it.
// ---------------------------------------------------------------------------------------------
template <bool kPlanar>
__global__ void __launch_bounds__(kBlock, kPlanar ? SCHW_MINBLOCKS : SCHW_MINBLOCKS_GENERAL)
integrate_full_kernel(const double *__restrict__ q0s, const double *__restrict__ p0s, long long n,
                      const __grid_constant__ StepParams P, double *__restrict__ out_traj,
                      long long *__restrict__ out_rows, Counters *ctr, const int *__restrict__ select, int want)
{
    if (select && *select != want) return;                 // (grid-uniform) see batch_probe_kernel
    constexpr int kWarps = kBlock / 32;
    constexpr int kPitch = 9;                              // double2 per ray in the staging area: 4 rows x 2 + 1 of padding,
                                                           // so that lane-strided 16-byte accesses fall into distinct banks
    __shared__ double2 s_stage[kWarps][32 * kPitch];       // per warp: 32 rays x (4 rows x 32 B)
    __shared__ int s_rows[kBlock];                         // rows stored by the ray of each lane (INT_MAX while it runs)
    __shared__ unsigned long long s_total;
    __shared__ Memo2 s_memo[kPlanar ? kBlock : 1];
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const long long i = (long long)blockIdx.x * kBlock + threadIdx.x;
    const long long ray0 = i - lane;                       // first ray of this warp
    Memo2 M_reg;
    Memo2 &M = kPlanar ? s_memo[threadIdx.x] : M_reg;
    M.init(P.delta, P.h);
    RayState S;
    bool active = i < n && P.steps > 0;
    if (active) {
        const double2 *q = reinterpret_cast<const double2 *>(q0s + 4 * i), *p = reinterpret_cast<const double2 *>(p0s + 4 * i);
        const double2 qa = __ldg(q), qb = __ldg(q + 1), pa = __ldg(p), pb = __ldg(p + 1);
        const double q0[4] = {qa.x, qa.y, qb.x, qb.y}, p0[4] = {pa.x, pa.y, pb.x, pb.y};
        init_state(S, q0, p0);
    } else {
        park(S, P);
    }
    s_rows[threadIdx.x] = active ? 0x7fffffff : 0;
    int pp_hi = active ? P.pl_pp_hi : 0, tok = 0, n_general = 0, it = 0;
    double2 *const stage = s_stage[warp];
    double2 *const mine = stage + lane * kPitch;           // this lane's four staged rows
    const int *rows_w = s_rows + (threadIdx.x & ~31);
    // flush geometry of this lane: in store j (0..7) it writes 16-byte piece `piece` of the 128-byte group of the warp's ray
    // L = 4 j + l3
    const int l3 = (int)(lane >> 3), piece = (int)(lane & 7);
    double2 *const dst0 = reinterpret_cast<double2 *>(out_traj + (size_t)(ray0 + l3) * (size_t)P.steps * 4) + piece;
    const size_t dstride = (size_t)4 * (size_t)P.steps * 2;       // double2 units between the rays of store j and store j+1
    const double2 *const src0 = stage + l3 * kPitch + piece;
    bool all_running = __all_sync(kFull, active);          // (warp-uniform) no ray of the warp has ended yet: full-line flushes
    bool any_active = __any_sync(kFull, active);           // (warp-uniform)
    bool quiet = false;                                    // (warp-uniform) no ray can end at the row about to be stored
    bool first = true;
    // (unrolling the four steps of a row group, so that staging offsets and the flush test become compile-time constants,
    //  measured equal: 0.567 vs 0.562 of the HBM figure -- the rolled loop is kept)
    for (int step = 0; any_active; ++step) {               // (warp-uniform trip count)
        const int slot = step & 3;
        // row `step`: q1 before the test (parked lanes stage their fixed point: never flushed)
        mine[2 * slot] = make_double2(S.q1t, S.q1r);
        mine[2 * slot + 1] = make_double2(S.q1h, S.q1p);
        if (!quiet || step == P.steps - 1) {
            // someone may end here: the exact FP64 test, per lane (cuda_geodesic.py:623), then the warp's bookkeeping
            if (active) {
                if (terminated(S.q1r, P) || step == P.steps - 1) {
                    active = false; pp_hi = 0;
                    s_rows[threadIdx.x] = step + 1;
                    atomicAdd(&s_total, (unsigned long long)step);
                    park(S, P);
                } else if (!(S.q1r == S.q1r)) {
                    // a NaN r never terminates in the reference: every later row is NaN
                    const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                    S.q1t = S.q1h = S.q1p = S.q2t = S.q2r = S.q2h = S.q2p = nanv;
                }
            }
            const unsigned am = __ballot_sync(kFull, active);
            any_active = am != 0u;
            all_running = all_running && am == kFull;
        }
        if (slot == 3 || !any_active) {                    // flush rows [step - slot, step] of the warp's 32 rays
            __syncwarp();
            const int base = step - slot;
            double2 *dst = dst0 + (size_t)base * 2;
            if (all_running && slot == 3) {                // all four rows of all 32 rays are valid: eight full-line stores
#pragma unroll
                for (int j = 0; j < 8; ++j) __stcs(dst + j * dstride, src0[4 * j * kPitch]);
            } else {
#pragma unroll 1
                for (int j = 0; j < 8; ++j) {
                    const int L = 4 * j + l3;
                    if ((piece >> 1) <= slot && base + (piece >> 1) < rows_w[L]) __stcs(dst + j * dstride, src0[4 * j * kPitch]);
                }
            }
            __syncwarp();
        }
        if (!any_active) break;
        // one step: A(d/2) B C B A(d/2) = [leading A(d/2) once] then B(d/2) C(d) B(d/2) A(d) with the merged A (A never
        // touches q1, so q1 after B C B is the row of the next step whatever A follows)
        if (first) { flow_A<false>(S, P.h, P.h_rs, P.rs, M); first = false; }
        if (kPlanar) {
            bool next = false;
            step_BCBA_planar(S, P, M, pp_hi, tok, next, n_general);
            // one vote per step: "everybody may start the next step planar and nobody can end at the next row" (integer
            // pre-test on the high word of r; parked lanes pass it, StepParams::park_r)
            const bool maybe = maybe_finished(S.q1r, P);
            tok = __all_sync(kFull, next & !maybe);
            quiet = tok ? true : (bool)__all_sync(kFull, !maybe);
        } else {                                           // tilted rays: the trig memo would never hit
            step_BCB<false>(S, P, M);
            flow_A<false>(S, P.delta, P.d_rs, P.rs, M);
            quiet = __all_sync(kFull, !maybe_finished(S.q1r, P));
        }
        ++it;
    }
    if (i < n && out_rows) out_rows[i] = s_rows[threadIdx.x];
    if (lane == 0 && it) {
        atomicAdd(&ctr->warp_iters, (unsigned long long)it);
        if (kPlanar) atomicAdd(&ctr->planar_iters, (unsigned long long)(it - n_general));
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(&ctr->total_steps, s_total);
}

// The same kernel for planar rays with R rays per thread (R = 2: the loop control, the votes and the constant loads of a
// loop body are shared by two geodesics, and at 4 CTAs/SM the compiler has the registers to keep staging pointers, store
// addresses and constants resident instead of re-materialising them every step).  A warp owns R groups of 32 consecutive
// rays; lane l carries ray l of each group.  Rows are staged and flushed exactly as above, group by group.
template <int R>
__global__ void __launch_bounds__(kBlock, R == 1 ? SCHW_MINBLOCKS : SCHW_TILE_MINBLOCKS)
integrate_full_planar_kernel(const double *__restrict__ q0s, const double *__restrict__ p0s, long long n,
                             const __grid_constant__ StepParams P, double *__restrict__ out_traj,
                             long long *__restrict__ out_rows, Counters *ctr, const int *__restrict__ select, int want)
{
    if (select && *select != want) return;                 // (grid-uniform) see batch_probe_kernel
    constexpr int kWarps = kBlock / 32;
    // staging: 8 double2 (4 rows x 32 B) per ray, 16-byte pieces XOR-swizzled with the ray's lane so that both the
    // lane-strided staging stores and the ray-contiguous flush loads are bank-conflict free (the padded layout of the
    // one-ray kernel would exceed the 48 KB of static shared memory here)
    __shared__ double2 s_stage[R][kWarps][32 * 8];
    __shared__ int s_rows[R][kBlock];                      // rows stored by each ray (INT_MAX while it runs)
    __shared__ unsigned long long s_total;
    __shared__ Memo2 s_memo[R * kBlock];
    if (threadIdx.x == 0) s_total = 0ull;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const int l3 = (int)(lane >> 3), piece = (int)(lane & 7);
    const size_t dstride = (size_t)4 * (size_t)P.steps * 2;       // double2 units between the rays of store j and store j+1
    const int sw = (int)(lane & 7);
    // in store j (0..7) this lane writes piece `piece` of ray L = 4 j + l3, staged at L*8 + (piece ^ (L & 7)), (L & 7) = 4 (j & 1) + l3
    const int sidx0 = l3 * 8 + (piece ^ l3), sidx1 = (4 + l3) * 8 + (piece ^ (4 + l3));
    RayState S[R];
    bool active[R];
    int pp_hi[R];
    double2 *mine[R];
    double2 *dst0[R];
    const double2 *src0[R];
    bool all_running = true, any_active = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long i = ((long long)blockIdx.x * R + r) * kBlock + threadIdx.x;
        s_memo[r * kBlock + threadIdx.x].init(P.delta, P.h);
        active[r] = i < n && P.steps > 0;
        if (active[r]) {
            const double2 *q = reinterpret_cast<const double2 *>(q0s + 4 * i), *p = reinterpret_cast<const double2 *>(p0s + 4 * i);
            const double2 qa = __ldg(q), qb = __ldg(q + 1), pa = __ldg(p), pb = __ldg(p + 1);
            const double q0[4] = {qa.x, qa.y, qb.x, qb.y}, p0[4] = {pa.x, pa.y, pb.x, pb.y};
            init_state(S[r], q0, p0);
        } else {
            park(S[r], P);
        }
        s_rows[r][threadIdx.x] = active[r] ? 0x7fffffff : 0;
        pp_hi[r] = active[r] ? P.pl_pp_hi : 0;
        mine[r] = s_stage[r][warp] + lane * 8;
        src0[r] = s_stage[r][warp];
        dst0[r] = reinterpret_cast<double2 *>(out_traj + (size_t)(i - lane + l3) * (size_t)P.steps * 4) + piece;
        all_running = all_running && __all_sync(kFull, active[r]);
        any_active = any_active || __any_sync(kFull, active[r]);
    }
    int tok = 0, n_general = 0, it = 0;
    bool quiet = false;                                    // (warp-uniform) no ray can end at the row about to be stored
    bool first = true;
    for (int step = 0; any_active; ++step) {               // (warp-uniform trip count)
        const int slot = step & 3;
#pragma unroll
        for (int r = 0; r < R; ++r) {                      // row `step`: q1 before the test
            mine[r][(2 * slot) ^ sw] = make_double2(S[r].q1t, S[r].q1r);
            mine[r][(2 * slot + 1) ^ sw] = make_double2(S[r].q1h, S[r].q1p);
        }
        if (!quiet || step == P.steps - 1) {
            unsigned am_all = kFull, am_any = 0u;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (active[r]) {
                    if (terminated(S[r].q1r, P) || step == P.steps - 1) {
                        active[r] = false; pp_hi[r] = 0;
                        s_rows[r][threadIdx.x] = step + 1;
                        atomicAdd(&s_total, (unsigned long long)step);
                        park(S[r], P);
                    } else if (!(S[r].q1r == S[r].q1r)) {
                        // a NaN r never terminates in the reference: every later row is NaN
                        const double nanv = __longlong_as_double(0x7ff8000000000000LL);
                        S[r].q1t = S[r].q1h = S[r].q1p = S[r].q2t = S[r].q2r = S[r].q2h = S[r].q2p = nanv;
                    }
                }
                const unsigned am = __ballot_sync(kFull, active[r]);
                am_all &= am; am_any |= am;
            }
            any_active = am_any != 0u;
            all_running = all_running && am_all == kFull;
        }
        if (slot == 3 || !any_active) {                    // flush rows [step - slot, step] of the warp's 32 R rays
            __syncwarp();
            const int base = step - slot;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                double2 *dst = dst0[r] + (size_t)base * 2;
                if (all_running && slot == 3) {            // all four rows of all rays are valid: eight full-line stores per group
#pragma unroll
                    for (int j = 0; j < 8; ++j) __stcs(dst + j * dstride, src0[r][32 * j + ((j & 1) ? sidx1 - 32 : sidx0)]);
                } else {
                    const int *rows_w = s_rows[r] + (threadIdx.x & ~31);
#pragma unroll 1
                    for (int j = 0; j < 8; ++j) {
                        const int L = 4 * j + l3;
                        if ((piece >> 1) <= slot && base + (piece >> 1) < rows_w[L])
                            __stcs(dst + j * dstride, src0[r][L * 8 + (piece ^ (L & 7))]);
                    }
                }
            }
            __syncwarp();
        }
        if (!any_active) break;
        if (first) {
#pragma unroll
            for (int r = 0; r < R; ++r) flow_A<false>(S[r], P.h, P.h_rs, P.rs, s_memo[r * kBlock + threadIdx.x]);
            first = false;
        }
        bool next = false;
        step_BCBA_planar_n<R>(S, P, &s_memo[threadIdx.x], kBlock, pp_hi, tok, next, n_general);
        bool maybe = false;
#pragma unroll
        for (int r = 0; r < R; ++r) maybe |= maybe_finished(S[r].q1r, P);
        tok = __all_sync(kFull, next & !maybe);
        quiet = tok ? true : (bool)__all_sync(kFull, !maybe);
        ++it;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const long long i = ((long long)blockIdx.x * R + r) * kBlock + threadIdx.x;
        if (i < n && out_rows) out_rows[i] = s_rows[r][threadIdx.x];
    }
    if (lane == 0 && it) {
        atomicAdd(&ctr->warp_iters, (unsigned long long)it * R);
        atomicAdd(&ctr->planar_iters, (unsigned long long)(it - n_general) * R);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_total) atomicAdd(&ctr->total_steps, s_total);
}

// ---------------------------------------------------------------------------------------------
// Decimated trajectories: the same integration as integrate_full_kernel, but only the rows
// whose step index appears in the ascending list row_idx[0..nidx) are written, to out (n, nidx, 4).  This is what
// the reference's consumer keeps anyway (np.linspace over the allocated rows, raytracing.py:139-143), without
// the (n, steps, 4) buffer (6.4 MB per ray at the default step cap).  Rows the ray never reaches stay zero.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
integrate_rows_kernel(const double *__restrict__ q0s, const double *__restrict__ p0s, long long n,
                      const __grid_constant__ StepParams P, const int32_t *__restrict__ row_idx, int nidx,
                      double *__restrict__ out, long long *__restrict__ out_rows, Counters *ctr)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayState S;
    {
        const double q0[4] = {q0s[4 * i], q0s[4 * i + 1], q0s[4 * i + 2], q0s[4 * i + 3]};
        const double p0[4] = {p0s[4 * i], p0s[4 * i + 1], p0s[4 * i + 2], p0s[4 * i + 3]};
        init_state(S, q0, p0);
    }
    Memo2 M;      // unused (kMemo = false): lanes of this kernel leave the loop at different times
    double2 *row = reinterpret_cast<double2 *>(out + (size_t)i * (size_t)nidx * 4);
    int cur = 0;
    int next = nidx > 0 ? row_idx[0] : -1;
    long long rows = 0;
    bool fresh = true;
    for (int step = 0; step < P.steps; ++step) {
        const bool nan_r = !(S.q1r == S.q1r);
        if (step == next) {
            row[2 * (size_t)cur] = make_double2(S.q1t, S.q1r);
            row[2 * (size_t)cur + 1] = make_double2(S.q1h, S.q1p);
            ++cur;
            next = cur < nidx ? row_idx[cur] : -1;
        }
        rows = step + 1;
        if (terminated(S.q1r, P)) break;
        if (nan_r) {
            // NaN never terminates in the reference: every later row is NaN
            const double nanv = __longlong_as_double(0x7ff8000000000000LL);
            for (; cur < nidx; ++cur) {
                row[2 * (size_t)cur] = make_double2(nanv, nanv);
                row[2 * (size_t)cur + 1] = make_double2(nanv, nanv);
            }
            rows = P.steps;
            break;
        }
        flow_A<false>(S, fresh ? P.h : P.delta, fresh ? P.h_rs : P.d_rs, P.rs, M);
        fresh = false;
        step_BCB<false>(S, P, M);
    }
    if (out_rows) out_rows[i] = rows;
    atomicAdd(&ctr->total_steps, (unsigned long long)(rows > 0 ? rows - 1 : 0));
}

// ---------------------------------------------------------------------------------------------
// Is a batch of caller-supplied rays worth the kernel with the planar fast path?  Only if (nearly) all of them start
// in the equatorial plane without polar momentum, like the reference's camera rays (utils.py:129,172): a warp leaves
// the fast path as soon as one of its lanes carries a tilted ray, and the fast-path kernel is slower than the plain one
// on such warps.  One CTA looks at up to 4096 evenly spaced rays and leaves the verdict in *flag; both kernels are then
// enqueued and the one that was not chosen exits immediately (BatchJob::skip) -- no host round trip.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
batch_probe_kernel(const double *__restrict__ q0s, const double *__restrict__ p0s, long long n, int *__restrict__ flag)
{
    __shared__ int hits;
    if (threadIdx.x == 0) hits = 0;
    __syncthreads();
    const long long m = n < 4096 ? n : 4096;
    int mine = 0;
    for (long long k = threadIdx.x; k < m; k += blockDim.x) {
        const long long i = (k * n) / m;
        mine += is_th_eq(q0s[4 * i + 2]) && fabs(p0s[4 * i + 2]) < 1e-18;
    }
    if (mine) atomicAdd(&hits, mine);
    __syncthreads();
    if (threadIdx.x == 0) *flag = (long long)hits * 10 >= m * 9;
}

// ---------------------------------------------------------------------------------------------
// Geodesic de-duplication (opt-in: schw_render_dedup*).  Every camera ray starts at the same q0 with
// p0 = F(alpha), alpha = its angle to the optical axis (utils.py:148-194), so pixels whose alpha agrees BITWISE -- the
// mirror images of a pixel, and more -- own the same geodesic: 8.3 M pixels of a 4K frame hold 0.95 M distinct alphas.
// The frame is rendered exactly as without it (same bits in every output), integrating each distinct ray once:
//   keys (alpha bits per pixel) -> radix sort -> run heads -> one integrate_batch over the distinct rays
//   -> per-pixel epilogue (own beta: rotate back, classify, texel) reading its geodesic's final state.
// ---------------------------------------------------------------------------------------------
__global__ void dedup_key_kernel(const __grid_constant__ SceneParams sc, long long n, unsigned long long *__restrict__ keys,
                                 unsigned *__restrict__ pix)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int row = (int)(k / sc.width), j = (int)(k - (long long)row * sc.width);
    RayIC r;
    initial_conditions(sc, sc.row_begin + row * sc.row_stride, j, r);
    keys[k] = (unsigned long long)__double_as_longlong(r.alpha);
    pix[k] = (unsigned)k;
}

__global__ void dedup_heads_kernel(const unsigned long long *__restrict__ keys_sorted, long long n, int *__restrict__ head)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (i == 0 || keys_sorted[i] != keys_sorted[i - 1]) ? 1 : 0;
}

// uid1[i] = inclusive sum of head[] = 1-based index (in order of alpha) of the distinct ray of sorted position i.
// Distinct ray u lives in SLOT (u % world) * chunk + u / world: with `world` ranks sharing the integration, rank r owns the
// contiguous slots [r*chunk, r*chunk + count_r) and every world-th distinct ray in alpha order -- neighbouring alphas cost
// the same, so the shares are balanced; world = 1 gives slot = u.
__global__ void dedup_scatter_kernel(const __grid_constant__ SceneParams sc, long long n, const unsigned *__restrict__ pix_sorted,
                                     const int *__restrict__ head, const int *__restrict__ uid1, int *__restrict__ ray_of_pixel,
                                     double *__restrict__ uq0, double *__restrict__ up0, int world, int chunk)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned k = pix_sorted[i];
    const int u0 = uid1[i] - 1;
    const int u = (u0 % world) * chunk + u0 / world;
    ray_of_pixel[k] = u;
    if (head[i]) {
        const int row = (int)(k / sc.width), j = (int)(k - (long long)row * sc.width);
        RayIC r;
        initial_conditions(sc, sc.row_begin + row * sc.row_stride, j, r);
        double2 *q = reinterpret_cast<double2 *>(uq0 + 4 * (size_t)u), *p = reinterpret_cast<double2 *>(up0 + 4 * (size_t)u);
        q[0] = make_double2(0.0, sc.r_obs); q[1] = make_double2(sc.th_obs, sc.ph_obs);
        p[0] = make_double2(r.p0[0], r.p0[1]); p[1] = make_double2(r.p0[2], r.p0[3]);
    }
}

// per pixel: its own (alpha0, beta) again, the final state of its geodesic, then the same epilogue as the fused kernel
__global__ void __launch_bounds__(kBlock)
dedup_epilogue_kernel(const __grid_constant__ RenderJob job, const int *__restrict__ ray_of_pixel, int plan_rb, int plan_rs,
                      const double *__restrict__ ufq, const int32_t *__restrict__ unsteps, Counters *ctr)
{
    __shared__ double2 s_stage[4 * kBlock];
    __shared__ unsigned long long s_steps;
    if (threadIdx.x == 0) s_steps = 0ull;
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < job.n) {
        double2 slot;
        job.load(k, &s_stage[threadIdx.x], slot);           // writes ic[k] if requested; slot = (alpha0, beta)
        // the plan may cover more rows than this call renders (multi-GPU: the whole frame): packed ray k of this call is
        // pixel (row_begin + row*row_stride, j) = packed pixel ((i - plan_rb) / plan_rs, j) of the plan
        const int row = (int)(k / job.sc.width), j = (int)(k - (long long)row * job.sc.width);
        const long long kp = (long long)((job.sc.row_begin + row * job.sc.row_stride - plan_rb) / plan_rs) * job.sc.width + j;
        const int u = ray_of_pixel[kp];
        const double2 a = *reinterpret_cast<const double2 *>(ufq + 4 * (size_t)u);
        const double2 b = *reinterpret_cast<const double2 *>(ufq + 4 * (size_t)u + 2);
        const int ns = unsteps[u];
        job.store(k, a.x, a.y, b.x, b.y, ns, !(a.y == a.y), slot);
        atomicAdd(&s_steps, (unsigned long long)ns);
    }
    __syncthreads();
    if (threadIdx.x == 0 && s_steps) atomicAdd(&ctr->pixel_steps, s_steps);
}

// ---------------------------------------------------------------------------------------------
// Stage kernels (boundary / parity testing): one thread per ray, static mapping.
// ---------------------------------------------------------------------------------------------
__global__ void ic_kernel(const __grid_constant__ SceneParams sc, long long n, double *__restrict__ q0s,
                          double *__restrict__ ic, unsigned *err)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int row = (int)(k / sc.width), j = (int)(k - (long long)row * sc.width);
    RayIC r;
    initial_conditions(sc, sc.row_begin + row * sc.row_stride, j, r);
    if (r.err) atomicOr(err, r.err);
    q0s[4 * k] = 0.0; q0s[4 * k + 1] = sc.r_obs; q0s[4 * k + 2] = sc.th_obs; q0s[4 * k + 3] = sc.ph_obs;
    double *o = ic + 9 * k;
    o[0] = r.p0[0]; o[1] = r.p0[1]; o[2] = r.p0[2]; o[3] = r.p0[3];
    o[4] = r.alpha0; o[5] = r.beta;
    c2s(r.dir[0], r.dir[1], r.dir[2], o[6], o[7], o[8]);
}

__global__ void ic_px_kernel(const __grid_constant__ SceneParams sc, const double *__restrict__ pixels, long long n,
                             double *__restrict__ q0s, double *__restrict__ ic, unsigned *err)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    RayIC r;
    initial_conditions_px(sc, pixels[3 * k], pixels[3 * k + 1], pixels[3 * k + 2], r);
    if (r.err) atomicOr(err, r.err);
    q0s[4 * k] = 0.0; q0s[4 * k + 1] = sc.r_obs; q0s[4 * k + 2] = sc.th_obs; q0s[4 * k + 3] = sc.ph_obs;
    double *o = ic + 9 * k;
    o[0] = r.p0[0]; o[1] = r.p0[1]; o[2] = r.p0[2]; o[3] = r.p0[3];
    o[4] = r.alpha0; o[5] = r.beta;
    c2s(r.dir[0], r.dir[1], r.dir[2], o[6], o[7], o[8]);
}

__global__ void classify_kernel(const __grid_constant__ SceneParams sc, const double *__restrict__ final_q,
                                const double *__restrict__ alpha0, const double *__restrict__ beta, long long n,
                                const uint8_t *__restrict__ tex, uint8_t *__restrict__ rgb,
                                uint8_t *__restrict__ cls, double *__restrict__ hit, int32_t *__restrict__ texel)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    RayHit o;
    classify(sc, tex, final_q[4 * k + 1], final_q[4 * k + 2], final_q[4 * k + 3], alpha0[k], beta[k], o);
    rgb[3 * k] = o.r; rgb[3 * k + 1] = o.g; rgb[3 * k + 2] = o.b;
    cls[k] = o.cls;
    if (hit) { hit[2 * k] = o.th_hit; hit[2 * k + 1] = o.ph_hit; }
    if (texel) { texel[2 * k] = o.u; texel[2 * k + 1] = o.v; }
}

// ---------------------------------------------------------------------------------------------
// Flat-space comparison render: save_no_gravity_image_with_background CPU loop,
// simulation/background.py:99-132 (+ camera basis :37-52, patch :54-66).  Texel map TRUNCATES
// (background.py:125-127) unlike the curved path.
// ---------------------------------------------------------------------------------------------
struct FlatParams {
    double obs[3], pc[3], right[3], up[3];
    double plane_width, plane_height;
    double r2;                      // boundary_radius**2
    double obs_dot_obs;
    double theta0, theta1, phi0, phi1, phi0m, phi1m, phi_span;
    int32_t height, width, flip_theta, flip_phi;
};

// Shared by both flat kernels: unit direction d from the observer -> hit on the boundary sphere -> patch test ->
// truncating texel map (background.py:106-132; the same arithmetic as the reference's Numba variant,
// cuda_geodesic.py:45-113).  Returns false (black pixel) when the ray misses the sphere; hit[] receives the hit point.
__device__ __forceinline__ bool flat_shade(const FlatParams &fp, const double d[3], const uint8_t *__restrict__ tex,
                                           uint8_t &r8, uint8_t &g8, uint8_t &b8, double hit[3])
{
    r8 = g8 = b8 = 0;
    const double a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
    const double b = 2 * (fp.obs[0] * d[0] + fp.obs[1] * d[1] + fp.obs[2] * d[2]);
    const double c = fp.obs_dot_obs - fp.r2;
    const double disc = b * b - 4 * a * c;
    if (disc < 0) return false;
    const double t = (-b + sqrt(disc)) / (2 * a);
    const double hx = fp.obs[0] + t * d[0], hy = fp.obs[1] + t * d[1], hz = fp.obs[2] + t * d[2];
    hit[0] = hx; hit[1] = hy; hit[2] = hz;
    const double r = sqrt(hx * hx + hy * hy + hz * hz);
    const double theta = acos(hz / r);
    const double phi = atan2(hy, hx);
    const double two_pi = 6.283185307179586;
    // _in_phi_patch, background.py:8-17
    const double pm = pymod(phi, two_pi);
    const bool in_phi = (fp.phi0m <= fp.phi1m) ? (fp.phi0m <= pm && pm <= fp.phi1m) : (pm >= fp.phi0m || pm <= fp.phi1m);
    if (theta >= fp.theta0 && theta <= fp.theta1 && in_phi) {
        const double theta_map = fp.flip_theta ? (3.141592653589793 - theta) : theta;
        const double phi_map = fp.flip_phi ? -phi : phi;
        long long ub = (long long)((theta_map - fp.theta0) / (fp.theta1 - fp.theta0) * (fp.height - 1));
        const double phi_mod = pymod(phi_map - fp.phi0, two_pi);
        long long vb = (long long)(phi_mod / fp.phi_span * (fp.width - 1));
        ub = ub < 0 ? 0 : (ub > fp.height - 1 ? fp.height - 1 : ub);
        vb = vb < 0 ? 0 : (vb > fp.width - 1 ? fp.width - 1 : vb);
        const uint8_t *tx = tex + ((size_t)ub * fp.width + vb) * 3;
        r8 = __ldg(tx); g8 = __ldg(tx + 1); b8 = __ldg(tx + 2);
    }
    return true;
}

__global__ void flat_kernel(const __grid_constant__ FlatParams fp, const uint8_t *__restrict__ tex,
                            uint8_t *__restrict__ rgb)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n = (long long)fp.height * fp.width;
    if (k >= n) return;
    const int i = (int)(k / fp.width), j = (int)(k - (long long)i * fp.width);
    const double u = (j + 0.5) / fp.width - 0.5;
    const double v = (i + 0.5) / fp.height - 0.5;
    double d[3], hit[3];
    for (int a = 0; a < 3; ++a) {
        const double px = (fp.pc[a] + (u * fp.plane_width) * fp.right[a]) + (v * fp.plane_height) * fp.up[a];
        d[a] = px - fp.obs[a];
    }
    const double nrm = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    d[0] /= nrm; d[1] /= nrm; d[2] /= nrm;
    uint8_t r8, g8, b8;
    flat_shade(fp, d, tex, r8, g8, b8, hit);
    rgb[3 * k] = r8; rgb[3 * k + 1] = g8; rgb[3 * k + 2] = b8;
}

// cuda_flat_raytrace (cuda_geodesic.py:214-250, kernel :18-113): caller-supplied unit directions, one ray per thread;
// fp.height / fp.width are the TEXTURE's dimensions (the reference reads them from bg_array.shape, :223).
__global__ void flat_dirs_kernel(const __grid_constant__ FlatParams fp, const double *__restrict__ dirs, long long n,
                                 const uint8_t *__restrict__ tex, uint8_t *__restrict__ rgb)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double d[3] = {dirs[3 * k], dirs[3 * k + 1], dirs[3 * k + 2]};
    double hit[3];
    uint8_t r8, g8, b8;
    flat_shade(fp, d, tex, r8, g8, b8, hit);
    rgb[3 * k] = r8; rgb[3 * k + 1] = g8; rgb[3 * k + 2] = b8;
}

// The trajectory half of cuda_flat_raytrace_with_trajectories (kernel :150-163): n_points equally spaced points from the
// observer to the hit point of ray sampled[s]; one thread per (s, point).  Rows of rays that miss the sphere stay zero.
__global__ void flat_traj_kernel(const __grid_constant__ FlatParams fp, const double *__restrict__ dirs, long long n,
                                 const int32_t *__restrict__ sampled, int n_sampled, int n_points, double *__restrict__ out)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_sampled * n_points) return;
    const int s = (int)(t / n_points), p = (int)(t - (long long)s * n_points);
    const long long k = sampled[s];
    if (k < 0 || k >= n) return;
    const double dx = dirs[3 * k], dy = dirs[3 * k + 1], dz = dirs[3 * k + 2];
    const double a = dx * dx + dy * dy + dz * dz;
    const double b = 2.0 * (fp.obs[0] * dx + fp.obs[1] * dy + fp.obs[2] * dz);
    const double disc = b * b - 4.0 * a * (fp.obs_dot_obs - fp.r2);
    if (disc < 0.0) return;
    const double tt = (-b + sqrt(disc)) / (2.0 * a);
    const double hx = fp.obs[0] + tt * dx, hy = fp.obs[1] + tt * dy, hz = fp.obs[2] + tt * dz;
    const double alpha = p / (n_points - 1.0);
    double *o = out + 3 * t;
    o[0] = fp.obs[0] + alpha * (hx - fp.obs[0]);
    o[1] = fp.obs[1] + alpha * (hy - fp.obs[1]);
    o[2] = fp.obs[2] + alpha * (hz - fp.obs[2]);
}

// Device math primitives of the flow, exposed for accuracy tests (tests/test_gpu_math.py).
__global__ void math_probe_kernel(const double *__restrict__ x, long long n, double *__restrict__ rcp,
                                  double *__restrict__ sn, double *__restrict__ cs)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    rcp[k] = fast_rcp(x[k]);
    sincos_flow(x[k], sn[k], cs[k]);
}

// ---------------------------------------------------------------------------------------------
// FP64 peak probe: 8 independent DFMA chains per thread, register resident.
// ---------------------------------------------------------------------------------------------
template <int kChains>
__global__ void __launch_bounds__(256)
dfma_probe_kernel(double *out, int iters, double b, double c)
{
    double a[kChains];
#pragma unroll
    for (int u = 0; u < kChains; ++u) a[u] = threadIdx.x * 1e-3 + u;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int u = 0; u < kChains; ++u) a[u] = fma(a[u], b, c);
        }
    }
    double sum = 0.0;
#pragma unroll
    for (int u = 0; u < kChains; ++u) sum += a[u];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = sum;
}

// ---------------------------------------------------------------------------------------------
// Step-loop probe (measurement only): the loop body of the render kernel on R rays per thread, without queue, refill or
// epilogue -- how many SM cycles a warp-step costs as a function of rays per thread and resident CTAs (DESIGN.md §4.1:
// is the step loop limited by the FP64 pipe, by issue slots or by dependent-issue latency?).
// ---------------------------------------------------------------------------------------------
template <int R, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
step_probe_kernel(const __grid_constant__ StepParams P, int rounds, int period, double *__restrict__ out, Counters *ctr)
{
    __shared__ Memo2 s_memo[R * kBlock];
    RayState S[R];
    int pp_hi[R];
    int n_general = 0, it = 0;
    double acc = 0.0;
    for (int round = 0; round < rounds; ++round) {
        // camera-like rays of the default scene (utils.py:148-194) at angles 15..45 degrees, different per thread and slot;
        // re-started every `period` steps so that they stay on the planar fast path like the rays of a frame (which end
        // after 3-7 k steps), instead of winding on until p_theta outgrows its bound
#pragma unroll
        for (int r = 0; r < R; ++r) {
            s_memo[r * kBlock + threadIdx.x].init(P.delta, P.h);
            const double al = 0.26 + 0.52 * ((threadIdx.x * 7 + r * 131 + blockIdx.x * 13 + round * 17) % 1024) / 1024.0;
            const double r0 = 30.0, f = 1.0 - P.rs / r0;
            const double pr = -cos(al) * sqrt(f), pph = sin(al) * r0;
            const double Cq = f * pr * pr + pph * pph / (r0 * r0);
            const double q0[4] = {0.0, r0, 1.5707963267948966, 0.0}, p0[4] = {sqrt(4.0 * Cq / f) / (2.0 / f), pr, -0.0, pph};
            init_state(S[r], q0, p0);
            pp_hi[r] = P.pl_pp_hi;
            flow_A<false>(S[r], P.h, P.h_rs, P.rs, s_memo[r * kBlock + threadIdx.x]);
        }
        int tok = 0;
        for (int k = 0; k < period; ++k) {
            bool next = false;
            step_BCBA_planar_n<R>(S, P, &s_memo[threadIdx.x], kBlock, pp_hi, tok, next, n_general);
            bool maybe = false;
#pragma unroll
            for (int r = 0; r < R; ++r) maybe |= maybe_finished(S[r].q1r, P);
            tok = __all_sync(kFull, next & !maybe);
            ++it;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) acc += S[r].q1t + S[r].q1r + S[r].q1p + S[r].p1t + S[r].p1r + S[r].p1h;
    }
    out[(size_t)blockIdx.x * kBlock + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&ctr->warp_iters, (unsigned long long)it);
        atomicAdd(&ctr->planar_iters, (unsigned long long)(it - n_general));
    }
}

}  // namespace schw

// =================================================================================================
//                                         C ABI
// =================================================================================================
using namespace schw;

static thread_local std::string g_err;

static int fail(int code, const std::string &msg)
{
    g_err = msg;
    return code;
}
#define CU(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(SCHW_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                 \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
};

struct schw_handle {
    int device = 0;
    int sm_count = 0;
    int clock_khz = 0;
    char name[256] = {0};
    cudaStream_t stream = nullptr;
    cudaStream_t last_stream = nullptr;   // stream of the last timed call (for result read-back ordering)
    cudaStream_t copy_stream = nullptr;   // texture uploads of schw_render, overlapped with the start of the kernel
    cudaEvent_t ev_copy0 = nullptr, ev_copy1 = nullptr;
    cudaEvent_t ev_dd = nullptr;          // the cached geodesics of the de-duplicated render are complete
    unsigned *tex_flag = nullptr;         // device word: "texture of the current frame has landed"
    unsigned *one_pinned = nullptr;       // page-locked {1, err}: [0] the source of the copy that raises tex_flag,
                                          // [1] receives Counters::err at the end of a host-pointer call
    const unsigned *next_tex_ready = nullptr;   // one-shot flag for the next schw_render_dev (schw_set_tex_ready_flag)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    int launches = 0;
    Counters *ctr = nullptr;
    int occ_render[3] = {0, 0, 0}, occ_batch[3] = {0, 0, 0};   // filled on first use, see launch_persistent
    int occ_tiles[2][3] = {{0, 0, 0}, {0, 0, 0}};               // persistent_tiles<RenderJob | BatchJob, R>, index [job][R]
    // grow-only scratch
    DevBuf in_a, in_b, in_c, tex, rgb, cls, fq, ic, hit, texel, nsteps, traj, rows, order;
    DevBuf dd_keys, dd_keys2, dd_pix, dd_pix2, dd_head, dd_uid, dd_map, dd_tmp, dd_q0, dd_p0, dd_fq, dd_ns;   // de-duplication
    bool in_pipeline = false;             // a multi-kernel entry point owns the timing events and the counters
    long long last_unique = 0;            // distinct rays of the last de-duplicated render
    bool dd_valid = false;                // dd_fq / dd_ns hold the geodesics of the frame described by dd_key
    bool dd_plan_valid = false;           // dd_map / dd_q0 / dd_p0 hold the plan described by dd_plan_key
    struct { double v[16]; } dd_plan_key = {{0}};
    int dd_world = 1, dd_chunk = 0;       // slot layout of the plan: distinct ray u -> (u % world) * chunk + u / world
    int dd_plan_rb = 0, dd_plan_rs = 1, dd_plan_rows = 0;   // rows the plan covers
    cudaEvent_t ev_plan = nullptr;
    double dd_key[16] = {0};
    double order_key[12] = {0};   // camera geometry the cached tile order was built for
    long long order_tiles = 0;
};

static int ensure(DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return SCHW_OK;
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.cap = 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) return fail(SCHW_E_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    b.cap = bytes;
    return SCHW_OK;
}
#define ENSURE(buf, bytes)                           \
    do {                                             \
        int rc_ = ensure(buf, bytes);                \
        if (rc_) return rc_;                         \
    } while (0)

// The error bits the device code raised during a call (Counters::err), as the reference's exceptions:
// AssertionError (utils.py:113,126) before RuntimeError (utils.py:80-81), in the order get_initial_conditions raises them.
static int device_error_code(unsigned e)
{
    if (e & kErrRayPlane)
        return fail(SCHW_E_RAYPLANE, "a ray does not lie in the equatorial plane after the rotation by -beta (|z'| >= 1e-6 or NaN: "
                                     "pixel position equal to the observer position?)");
    if (e & kErrNullCond) return fail(SCHW_E_NULLCOND, "Negative discriminant while enforcing null condition.");
    if (e & kErrTexWait)
        return fail(SCHW_E_CUDA, "the texture upload did not arrive within ~2 s of the first texel lookup (tex_ready flag never "
                                 "raised); the frame is invalid");
    return SCHW_OK;
}
// Last step of a host-pointer entry point: wait for the stream, then report what the device code flagged.
static int sync_and_check(schw_handle *h, cudaStream_t st)
{
    CU(cudaMemcpyAsync(&h->one_pinned[1], &h->ctr->err, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return device_error_code(h->one_pinned[1]);
}

extern "C" int schw_abi_version(void) { return SCHW_ABI_VERSION; }
extern "C" const char *schw_last_error(void) { return g_err.c_str(); }

extern "C" int schw_create(int device, schw_handle **out)
{
    if (!out) return fail(SCHW_E_INVALID, "schw_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SCHW_E_CUDA, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                     " (libschw_b200 has no CPU fallback)");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= count) return fail(SCHW_E_INVALID, "schw_create: device index out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(SCHW_E_CUDA, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major * 10 + prop.minor) +
                                     "; libschw_b200 is built for sm_100a only");
    schw_handle *h = new (std::nothrow) schw_handle();
    if (!h) return fail(SCHW_E_NOMEM, "schw_create: out of host memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    CU(cudaDeviceGetAttribute(&h->clock_khz, cudaDevAttrClockRate, device));
    snprintf(h->name, sizeof(h->name), "%s", prop.name);
    CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_copy0, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_copy1, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_dd, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&h->ev_plan, cudaEventDisableTiming));
    CU(cudaMalloc(&h->tex_flag, sizeof(unsigned)));
    CU(cudaHostAlloc(&h->one_pinned, 2 * sizeof(unsigned), cudaHostAllocDefault));
    h->one_pinned[0] = 1u; h->one_pinned[1] = 0u;
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaMalloc(&h->ctr, sizeof(Counters)));
    CU(cudaMemset(h->ctr, 0, sizeof(Counters)));
    *out = h;
    return SCHW_OK;
}

extern "C" int schw_destroy(schw_handle *h)
{
    if (!h) return SCHW_OK;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf *bufs[] = {&h->in_a, &h->in_b, &h->in_c, &h->tex, &h->rgb, &h->cls, &h->fq, &h->ic, &h->hit, &h->texel, &h->nsteps, &h->traj, &h->rows, &h->order,
                      &h->dd_keys, &h->dd_keys2, &h->dd_pix, &h->dd_pix2, &h->dd_head, &h->dd_uid, &h->dd_map, &h->dd_tmp,
                      &h->dd_q0, &h->dd_p0, &h->dd_fq, &h->dd_ns};
    for (DevBuf *b : bufs) if (b->p) cudaFree(b->p);
    if (h->ctr) cudaFree(h->ctr);
    if (h->tex_flag) cudaFree(h->tex_flag);
    if (h->one_pinned) cudaFreeHost(h->one_pinned);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->ev_copy0) cudaEventDestroy(h->ev_copy0);
    if (h->ev_copy1) cudaEventDestroy(h->ev_copy1);
    if (h->ev_dd) cudaEventDestroy(h->ev_dd);
    if (h->ev_plan) cudaEventDestroy(h->ev_plan);
    cudaEventDestroy(h->ev0);
    cudaEventDestroy(h->ev1);
    cudaStreamDestroy(h->stream);
    delete h;
    return SCHW_OK;
}

extern "C" int schw_host_alloc(size_t bytes, void **out)
{
    if (!out) return fail(SCHW_E_INVALID, "schw_host_alloc: out is NULL");
    *out = nullptr;
    if (bytes == 0) return SCHW_OK;
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        *out = nullptr;
        return fail(SCHW_E_NOMEM, std::string("cudaHostAlloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e));
    }
    return SCHW_OK;
}

extern "C" int schw_host_free(void *p)
{
    if (!p) return SCHW_OK;
    CU(cudaFreeHost(p));
    return SCHW_OK;
}

extern "C" int schw_device_info(schw_handle *h, char *name, int name_cap, int *sm_count, int *clock_khz)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (name && name_cap > 0) { strncpy(name, h->name, name_cap - 1); name[name_cap - 1] = 0; }
    if (sm_count) *sm_count = h->sm_count;
    if (clock_khz) *clock_khz = h->clock_khz;
    return SCHW_OK;
}

extern "C" int schw_last_kernel_ms(schw_handle *h, float *ms, int *launches)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (!h->timed) return fail(SCHW_E_INVALID, "no timed call on this handle yet");
    CU(cudaSetDevice(h->device));
    CU(cudaEventSynchronize(h->ev1));
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, h->ev0, h->ev1));
    if (ms) *ms = t;
    if (launches) *launches = h->launches;
    return SCHW_OK;
}

extern "C" int schw_last_total_steps(schw_handle *h, uint64_t *total_steps)
{
    if (!h || !total_steps) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    Counters c;
    CU(cudaStreamSynchronize(h->last_stream ? h->last_stream : h->stream));
    CU(cudaMemcpy(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    *total_steps = c.total_steps;
    return SCHW_OK;
}

// For the enqueue-only (_dev) entry points: waits for the last call on this handle and returns what its device code flagged.
extern "C" int schw_last_status(schw_handle *h)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    Counters c;
    CU(cudaStreamSynchronize(h->last_stream ? h->last_stream : h->stream));
    CU(cudaMemcpy(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    return device_error_code(c.err);
}

extern "C" int schw_last_counters(schw_handle *h, uint64_t *total_steps, uint64_t *warp_iters, uint64_t *planar_iters)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    Counters c;
    CU(cudaStreamSynchronize(h->last_stream ? h->last_stream : h->stream));
    CU(cudaMemcpy(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    if (total_steps) *total_steps = c.total_steps;
    if (warp_iters) *warp_iters = c.warp_iters;
    if (planar_iters) *planar_iters = c.planar_iters;
    return SCHW_OK;
}

// ---- parameter preparation (host FP64, glibc — the same libm NumPy uses for sin/cos) ------------
// SCHW_PLANAR=0 disables the equatorial fast path (A/B measurements; results are bit-identical).
// (The A/B switches are read on every call, so a test can flip them inside one process.)
static bool env_switch_on(const char *name)
{
    const char *e = getenv(name);
    return !(e && e[0] == '0');
}
static bool planar_enabled() { return env_switch_on("SCHW_PLANAR"); }
// SCHW_PLANAR=2: use the fast-path kernel for every integrate_batch call, whatever the rays look like (tests of the
// per-warp switching between the two paths on mixed batches; by default the device picks, see batch_probe_kernel)
static bool planar_forced()
{
    const char *e = getenv("SCHW_PLANAR");
    return e && e[0] == '2';
}
static int32_t hi_word(double x)
{
    uint64_t u;
    memcpy(&u, &x, 8);
    return (int32_t)(u >> 32);
}
static double from_hi_word(int32_t hi)
{
    const uint64_t u = (uint64_t)(uint32_t)hi << 32;
    double x;
    memcpy(&x, &u, 8);
    return x;
}
// Thresholds of the planar fast path (schw_device.cuh, "Planar steps"): pick r_thr below the capture radius, then
// the largest |p_theta| bound T that makes (R1) and (R3) hold, then the |p_phi| bound that makes (R2) hold.
static void make_planar_params(StepParams &P)
{
    const double cot_eq = 6.123233995736766e-17;   // cos/sin of the double nearest pi/2 (sin is exactly 1.0)
    P.pl_hcA = P.delta * cot_eq;
    P.pl_hcB = P.h * cot_eq;
    P.pl_zero_hs = 0.0 * P.mhs;
    P.pl_enable = 0; P.pl_ph_hi = 0; P.pl_r_hi = 0x7ff00000; P.pl_pp_hi = 0x7ff00000;
    if (!planar_enabled()) return;
    if (!(P.delta > 0.0) || !(P.delta < 1e6) || !(P.rcap > 0.0) || !(P.rcap < 1e300) || !(fabs(P.mhs) <= 0.5)) return;
    const int32_t r_hi = hi_word(0.75 * P.rcap) + 1;               // hi(r) >= r_hi  =>  r >= r_thr > 0.75 rcap
    const double r_thr = from_hi_word(r_hi);
    double T = ldexp(1.0, -54) * r_thr * r_thr / P.delta;          // (R1), delta >= delta/2 covers both flows
    if (P.mhs != 0.0) T = fmin(T, ldexp(1.0, -55) / fabs(P.mhs));  // (R3)
    T = fmin(T, ldexp(1.0, -30));
    if (!(T >= ldexp(1.0, -900))) return;
    P.pl_ph_hi = hi_word(T);                                       // hi(|p|) < hi(T)  =>  |p| < T
    P.pl_pp_hi = hi_word(ldexp(T, 28)) + 1;                        // hi(|p|) >= this  =>  |p| > 2^28 T
    P.pl_r_hi = r_hi;
    P.pl_enable = 1;
}

static int make_step_params(int64_t steps, double delta, double rs, double r_max, double omega, StepParams &P)
{
    if (steps < 0 || steps > 0x7fffffffLL) return fail(SCHW_E_INVALID, "steps must be in [0, 2^31-1]");
    P.h = 0.5 * delta;
    P.delta = delta;
    P.h_rs = P.h * rs;
    P.d_rs = delta * rs;
    P.rs = rs;
    P.rcap = 1.1 * rs;
    P.r_max = r_max;
    const double ang = 2.0 * omega * delta;
    P.mg = 0.5 * (1.0 - cos(ang));      // 1 - cos is exact (Sterbenz): g carries the reference's rounded cosine
    P.mhs = 0.5 * sin(ang);
    P.steps = (int32_t)steps;
    P._pad = 0;
    // high-word pre-test of the termination test (maybe_finished): for 0 <= rcap and 0 < r_max the signed high words
    // order like the doubles; otherwise (negative or NaN thresholds) the pre-test never rules termination out
    P.hi_rcap = (P.rcap >= 0.0) ? hi_word(P.rcap) : 0x7fffffff;
    P.hi_rmax = (P.r_max > 0.0) ? hi_word(P.r_max) : (int32_t)0x80000000;
    // where lanes without a ray sit: a radius whose high word lies strictly between the two thresholds, so that the
    // termination pre-test is false for them; if there is none (r_max within a few ulps of the capture radius, or
    // degenerate thresholds) any finite radius does -- the pre-test then always sends the warp to the exact test, which
    // looks at active lanes only
    P.park_r = 4.0 * rs + 8.0;
    if (P.hi_rcap >= 0 && P.hi_rmax > 0 && (int64_t)P.hi_rmax - (int64_t)P.hi_rcap >= 2 && P.hi_rmax < 0x7ff00000) {
        const double mid = from_hi_word((int32_t)(((int64_t)P.hi_rcap + (int64_t)P.hi_rmax) / 2));
        if (!(hi_word(P.park_r) > P.hi_rcap && hi_word(P.park_r) < P.hi_rmax)) P.park_r = mid;
    }
    make_planar_params(P);
    return SCHW_OK;
}

static int make_scene_params(const schw_scene *s, int row_begin, int row_end, int row_stride, SceneParams &sc)
{
    if (!s) return fail(SCHW_E_INVALID, "scene is NULL");
    if (s->height <= 0 || s->width <= 0) return fail(SCHW_E_INVALID, "image_size must be positive");
    if (row_stride <= 0 || row_begin < 0 || row_end > s->height || row_begin > row_end)
        return fail(SCHW_E_INVALID, "bad row range");
    const int h = s->height, w = s->width;
    // camera plane, raytracing.py:74-81
    const double nrm = sqrt(s->obs[0] * s->obs[0] + s->obs[1] * s->obs[1] + s->obs[2] * s->obs[2]);
    const double plane_dist = 0.2 * nrm;
    for (int a = 0; a < 3; ++a) sc.obs[a] = s->obs[a];
    sc.pc[0] = s->obs[0] + (-1.0) * plane_dist;
    sc.pc[1] = s->obs[1] + 0.0 * plane_dist;
    sc.pc[2] = s->obs[2] + 0.0 * plane_dist;
    sc.plane_width = 2 * plane_dist * tan(s->fov / 2);
    sc.plane_height = sc.plane_width * ((double)h / (double)w);
    // observer in spherical coordinates, utils.py:118
    const double hxy = hypot(s->obs[0], s->obs[1]);
    sc.r_obs = hypot(hxy, s->obs[2]);
    sc.th_obs = atan2(hxy, s->obs[2]);
    sc.ph_obs = atan2(s->obs[1], s->obs[0]);
    // build_null_4momentum_ep_sph constants, utils.py:66-84
    const double f = 1.0 - 2.0 * s->mass / sc.r_obs;
    if (f <= 0.0) return fail(SCHW_E_HORIZON, "Observer must lie outside the event horizon (r > 2M).");
    sc.f_r = sqrt(1.0 - 2.0 * s->mass / sc.r_obs);
    const double A = -1.0 / f;
    sc.grr = f;
    sc.gthth = 1.0 / (sc.r_obs * sc.r_obs);
    const double sth = sin(sc.th_obs);
    sc.gphph = 1.0 / (sc.r_obs * sc.r_obs * (sth * sth));
    sc.m4A = -4.0 * A;
    sc.twomA = 2.0 * (-A);
    // classification constants, raytracing.py:166-174,205-216
    const double rs = 2 * s->mass;
    sc.rs12 = rs * 1.2;
    sc.bh_angle = asin(3 * sqrt(3.0) * rs / s->obs[0]) / 2;
    sc.r_max = s->r_max;
    sc.theta0 = s->patch_center_theta - s->patch_size_theta / 2;
    sc.theta1 = s->patch_center_theta + s->patch_size_theta / 2;
    sc.phi0 = s->patch_center_phi - s->patch_size_phi / 2;
    sc.phi_span = s->patch_size_phi;
    sc.pct = s->patch_center_theta;
    sc.pcp = s->patch_center_phi;
    sc.half_pst = s->patch_size_theta / 2;
    sc.half_psp = s->patch_size_phi / 2;
    sc.height = h; sc.width = w;
    sc.flip_theta = s->flip_theta; sc.flip_phi = s->flip_phi;
    sc.has_texture = s->has_texture;
    sc.row_begin = row_begin; sc.row_stride = row_stride;
    sc.nrows = row_end > row_begin ? (row_end - row_begin + row_stride - 1) / row_stride : 0;
    return SCHW_OK;
}

// q0 = (0, r_obs, theta_obs, phi_obs) of every camera ray of the scene (utils.py:118,140), host arithmetic only.
extern "C" int schw_observer_q0(const schw_scene *scene, double *q0)
{
    if (!scene || !q0) return fail(SCHW_E_INVALID, "null argument");
    SceneParams sc;
    int rc = make_scene_params(scene, 0, scene->height, 1, sc);
    if (rc) return rc;
    q0[0] = 0.0; q0[1] = sc.r_obs; q0[2] = sc.th_obs; q0[3] = sc.ph_obs;
    return SCHW_OK;
}

// Diagnostic (no device needed): the step parameters -- incl. the thresholds of the planar fast path and of the
// termination pre-test -- exactly as the kernels receive them, copied to `out` (sizeof(schw::StepParams) bytes).
extern "C" int schw_debug_step_params(int64_t steps, double delta, double rs, double r_max, double omega,
                                      void *out, size_t cap, size_t *size)
{
    StepParams P;
    int rc = make_step_params(steps, delta, rs, r_max, omega, P);
    if (rc) return rc;
    if (size) *size = sizeof(P);
    if (out) {
        if (cap < sizeof(P)) return fail(SCHW_E_INVALID, "schw_debug_step_params: buffer too small");
        memcpy(out, &P, sizeof(P));
    }
    return SCHW_OK;
}

static void begin_timed(schw_handle *h, cudaStream_t st)
{
    if (h->in_pipeline) return;
    cudaEventRecord(h->ev0, st);
    h->launches = 0;
    h->last_stream = st;
}
static void end_timed(schw_handle *h, cudaStream_t st)
{
    if (h->in_pipeline) return;
    cudaEventRecord(h->ev1, st);
    h->timed = true;
}

// SCHW_TRIG_MEMO=0 disables the exact sin/cos memoisation (A/B measurements; results are bit-identical).
static bool trig_memo_enabled() { return env_switch_on("SCHW_TRIG_MEMO"); }

template <class Job, bool kMemo, bool kPlanar>
static int launch_variant(schw_handle *h, const Job &job, const StepParams &P, int &occ, cudaStream_t st)
{
    auto kern = persistent_rays<Job, kMemo, kPlanar>;
    if (occ == 0) {      // resident CTAs per SM of this variant (they differ in registers per thread)
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, 0));
        if (occ < 1) return fail(SCHW_E_CUDA, "kernel does not fit on an SM");
    }
    const long long want_blocks = (job.n + kBlock - 1) / kBlock;
    long long grid = (long long)h->sm_count * occ;
    if (grid > want_blocks) grid = want_blocks;
    kern<<<(unsigned)grid, kBlock, 0, st>>>(job, P, h->ctr);
    h->launches += 1;
    CU(cudaGetLastError());
    return SCHW_OK;
}

// occ: the handle's three occupancy slots of this job type {no memo, memo, memo + planar}
template <class Job>
static int launch_persistent(schw_handle *h, const Job &job, const StepParams &P, int *occ, cudaStream_t st,
                             bool reset_counters = true)
{
    if (reset_counters) CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
    if (job.n <= 0) return SCHW_OK;
    if (!trig_memo_enabled()) return launch_variant<Job, false, false>(h, job, P, occ[0], st);
    if (!P.pl_enable) return launch_variant<Job, true, false>(h, job, P, occ[1], st);
    return launch_variant<Job, true, true>(h, job, P, occ[2], st);
}

// The whole-tile kernel (persistent_tiles<Job, R>), R rays per thread.  Two rays per thread (4 CTAs/SM) are faster once a
// warp gets a few tiles -- 4K frame 297.9 -> 292.4 ms, 1080p 75.5 -> 74.3 ms, a 1/8 shard of the 4K frame 37.9 -> 37.7 ms,
// a 1/16 shard 19.48 -> 19.39 ms -- while one ray per thread (6 CTAs/SM: smaller units of work) ends a short launch sooner:
// 14 400 tiles 17.38 vs 17.52 ms, 400x400 (5200 tiles) 6.85 vs 7.1 ms (profiles/r02_ab_tile.json, tools/ab_tail.sh).
// SCHW_TILE_RAYS=1|2 forces one (A/B measurements).
static int tile_rays(long long qlen)
{
    const char *e = getenv("SCHW_TILE_RAYS");
    if (e && (e[0] == '1' || e[0] == '2')) return e[0] - '0';
    return qlen / 32 >= 16000 ? SCHW_TILE_RAYS : 1;
}
static int tile_mode_for(const schw_handle *, long long);

template <class Job, int R>
static int launch_tiles_variant(schw_handle *h, const Job &job, long long qlen, const StepParams &P, cudaStream_t st)
{
    auto kern = persistent_tiles<Job, R>;
    int &occ = h->occ_tiles[std::is_same<Job, RenderJob>::value ? 0 : 1][R];
    if (occ == 0) {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBlock, 0));
        if (occ < 1) return fail(SCHW_E_CUDA, "kernel does not fit on an SM");
    }
    const long long want_blocks = (qlen + (long long)kBlock * R - 1) / ((long long)kBlock * R);
    int use = occ;
    if (const char *e = getenv("SCHW_TILE_CTAS")) { const int c = atoi(e); if (c >= 1 && c < occ) use = c; }   // (A/B measurements)
    long long grid = (long long)h->sm_count * use;
    if (grid > want_blocks) grid = want_blocks;
    kern<<<(unsigned)grid, kBlock, 0, st>>>(job, P, h->ctr);
    h->launches += 1;
    CU(cudaGetLastError());
    return SCHW_OK;
}
template <class Job>
static int launch_tiles(schw_handle *h, const Job &job, long long qlen, const StepParams &P, cudaStream_t st, bool reset_counters = true)
{
    if (reset_counters) CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
    if (job.n <= 0) return SCHW_OK;
    return tile_rays(qlen) == 1 ? launch_tiles_variant<Job, 1>(h, job, qlen, P, st) : launch_tiles_variant<Job, 2>(h, job, qlen, P, st);
}

// ---- integrate_batch ----------------------------------------------------------------------------
static int integrate_batch_impl(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n,
                                int64_t steps, double delta, double rs, double r_max, double omega,
                                double *d_out_qs, int32_t *d_out_steps, void *cuda_stream, bool coherent = false)
{
    // coherent: neighbouring rays of the batch need about the same number of steps (the distinct camera rays of a
    // de-duplicated frame, in order of camera angle), so planar batches go to the whole-tile kernel
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (n < 0 || n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "n must be in [0, 2^31-1]");
    if (n > 0 && (!d_q0s || !d_p0s || !d_out_qs)) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    StepParams P;
    int rc = make_step_params(steps, delta, rs, r_max, omega, P);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    BatchJob job{d_q0s, d_p0s, d_out_qs, d_out_steps, (long long)n, nullptr, 0};
    begin_timed(h, st);
    if (n > 0 && P.pl_enable && trig_memo_enabled() && !planar_forced()) {
        // let the device choose between the kernel with the planar fast path and the plain one (batch_probe_kernel)
        CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
        batch_probe_kernel<<<1, 256, 0, st>>>(d_q0s, d_p0s, (long long)n, &h->ctr->batch_planar);
        h->launches += 1;
        CU(cudaGetLastError());
        job.select = &h->ctr->batch_planar;
        job.want = 1;
        if (coherent && tile_mode_for(h, n) == 2) rc = launch_tiles(h, job, (long long)n, P, st, false);
        else rc = launch_persistent(h, job, P, h->occ_batch, st, false);
        if (rc) return rc;
        // not planar: the angle of a tilted ray changes every step, the trig memo never hits and its compare + vote per flow
        // only cost issue slots (4e5 random rays: 44.6 -> 38.8 ms without it)
        job.want = 0;
        P.pl_enable = 0;
        rc = launch_variant<BatchJob, false, false>(h, job, P, h->occ_batch[0], st);
    } else {
        rc = launch_persistent(h, job, P, h->occ_batch, st);
    }
    end_timed(h, st);
    return rc;
}

extern "C" int schw_integrate_batch_dev(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n,
                                        int64_t steps, double delta, double rs, double r_max, double omega,
                                        double *d_out_qs, int32_t *d_out_steps, void *cuda_stream)
{
    return integrate_batch_impl(h, d_q0s, d_p0s, n, steps, delta, rs, r_max, omega, d_out_qs, d_out_steps, cuda_stream);
}

extern "C" int schw_integrate_batch(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                                    int64_t steps, double delta, double rs, double r_max, double omega,
                                    double *out_qs, int32_t *out_steps)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (n < 0 || n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "n must be in [0, 2^31-1]");
    if (n > 0 && (!q0s || !p0s || !out_qs)) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    const size_t b = (size_t)n * 4 * sizeof(double);
    ENSURE(h->in_a, b); ENSURE(h->in_b, b); ENSURE(h->fq, b);
    if (out_steps) ENSURE(h->nsteps, (size_t)n * sizeof(int32_t));
    if (n > 0) {
        CU(cudaMemcpyAsync(h->in_a.p, q0s, b, cudaMemcpyHostToDevice, h->stream));
        CU(cudaMemcpyAsync(h->in_b.p, p0s, b, cudaMemcpyHostToDevice, h->stream));
    }
    int rc = integrate_batch_impl(h, (const double *)h->in_a.p, (const double *)h->in_b.p, n, steps, delta, rs, r_max,
                                  omega, (double *)h->fq.p, out_steps ? (int32_t *)h->nsteps.p : nullptr, nullptr);
    if (rc) return rc;
    if (n > 0) {
        CU(cudaMemcpyAsync(out_qs, h->fq.p, b, cudaMemcpyDeviceToHost, h->stream));
        if (out_steps) CU(cudaMemcpyAsync(out_steps, h->nsteps.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    return SCHW_OK;
}

// ---- integrate_batch_full -----------------------------------------------------------------------
static int integrate_full_impl(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n, const StepParams &P0,
                               double *d_traj, long long *d_rows, cudaStream_t st)
{
    StepParams P = P0;
    const unsigned grid = (unsigned)((n + kBlock - 1) / kBlock);
    const unsigned grid2 = (unsigned)((n + 2 * kBlock - 1) / (2 * kBlock));
    // Planar batches: two rays per thread once the batch fills the resident grid of that kernel (sm_count x 4 CTAs x 256
    // rays = 151 552 on a B200: measured 5 % faster there, 10 % at 454 656 rays), one ray per thread below it, where twice
    // the CTAs spread over more SMs (4096 rays x 8192 rows: 3.7 vs 4.8 ms).  SCHW_FULL_RAYS=1|2 forces one (A/B measurements).
    const char *fr = getenv("SCHW_FULL_RAYS");
    const bool two = (fr && (fr[0] == '1' || fr[0] == '2')) ? fr[0] == '2'
                                                            : n >= (int64_t)h->sm_count * SCHW_TILE_MINBLOCKS * 2 * kBlock;
    CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
    begin_timed(h, st);
    if (P.pl_enable && trig_memo_enabled() && !planar_forced()) {
        // the device chooses between the kernel with the planar fast path and the plain one (batch_probe_kernel)
        batch_probe_kernel<<<1, 256, 0, st>>>(d_q0s, d_p0s, (long long)n, &h->ctr->batch_planar);
        if (two) integrate_full_planar_kernel<2><<<grid2, kBlock, 0, st>>>(d_q0s, d_p0s, (long long)n, P, d_traj, d_rows, h->ctr, &h->ctr->batch_planar, 1);
        else integrate_full_kernel<true><<<grid, kBlock, 0, st>>>(d_q0s, d_p0s, (long long)n, P, d_traj, d_rows, h->ctr, &h->ctr->batch_planar, 1);
        P.pl_enable = 0;
        integrate_full_kernel<false><<<grid, kBlock, 0, st>>>(d_q0s, d_p0s, (long long)n, P, d_traj, d_rows, h->ctr, &h->ctr->batch_planar, 0);
        h->launches += 3;
    } else if (P.pl_enable && trig_memo_enabled()) {
        integrate_full_kernel<true><<<grid, kBlock, 0, st>>>(d_q0s, d_p0s, (long long)n, P, d_traj, d_rows, h->ctr, nullptr, 0);
        h->launches += 1;
    } else {
        P.pl_enable = 0;
        integrate_full_kernel<false><<<grid, kBlock, 0, st>>>(d_q0s, d_p0s, (long long)n, P, d_traj, d_rows, h->ctr, nullptr, 0);
        h->launches += 1;
    }
    CU(cudaGetLastError());
    end_timed(h, st);
    return SCHW_OK;
}

extern "C" int schw_integrate_batch_full_dev(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n,
                                             int64_t steps, double delta, double rs, double r_max, double omega,
                                             double *d_out_traj, int64_t *d_out_rows, int zero_fill, void *cuda_stream)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (n < 0 || n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "n must be in [0, 2^31-1]");
    if (n > 0 && (!d_q0s || !d_p0s || (!d_out_traj && steps > 0))) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    StepParams P;
    int rc = make_step_params(steps, delta, rs, r_max, omega, P);
    if (rc) return rc;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    if (n == 0 || steps == 0) return SCHW_OK;
    if (zero_fill) CU(cudaMemsetAsync(d_out_traj, 0, (size_t)n * (size_t)steps * 32, st));      // np.zeros, cuda_geodesic.py:649
    long long *d_rows = (long long *)d_out_rows;
    if (!d_rows) { ENSURE(h->rows, (size_t)n * sizeof(long long)); d_rows = (long long *)h->rows.p; }
    return integrate_full_impl(h, d_q0s, d_p0s, n, P, d_out_traj, d_rows, st);
}

extern "C" int schw_integrate_batch_full(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                                         int64_t steps, double delta, double rs, double r_max, double omega,
                                         double *out_traj, int64_t *out_rows)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (n < 0 || n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "n must be in [0, 2^31-1]");
    if (n > 0 && (!q0s || !p0s || (!out_traj && steps > 0))) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    StepParams P;
    int rc = make_step_params(steps, delta, rs, r_max, omega, P);
    if (rc) return rc;
    if (n == 0 || steps == 0) return SCHW_OK;
    const size_t b = (size_t)n * 4 * sizeof(double);
    const size_t tb = (size_t)n * (size_t)steps * 4 * sizeof(double);
    ENSURE(h->in_a, b); ENSURE(h->in_b, b); ENSURE(h->traj, tb); ENSURE(h->rows, (size_t)n * sizeof(long long));
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(h->in_a.p, q0s, b, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->in_b.p, p0s, b, cudaMemcpyHostToDevice, st));
    // No device-side zero fill: only the stored rows are copied back into the caller's np.zeros buffer.
    rc = integrate_full_impl(h, (const double *)h->in_a.p, (const double *)h->in_b.p, n, P, (double *)h->traj.p,
                             (long long *)h->rows.p, st);
    if (rc) return rc;
    long long *rows = (long long *)malloc((size_t)n * sizeof(long long));
    if (!rows) return fail(SCHW_E_NOMEM, "host malloc");
    cudaError_t e = cudaMemcpyAsync(rows, h->rows.p, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    for (int64_t i = 0; e == cudaSuccess && i < n; ++i) {
        if (out_rows) out_rows[i] = rows[i];
        if (rows[i] > 0)
            e = cudaMemcpyAsync(out_traj + (size_t)i * steps * 4, (const double *)h->traj.p + (size_t)i * steps * 4,
                                (size_t)rows[i] * 4 * sizeof(double), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(rows);
    if (e != cudaSuccess) return fail(SCHW_E_CUDA, std::string("integrate_batch_full copy-back: ") + cudaGetErrorString(e));
    return SCHW_OK;
}

extern "C" int schw_integrate_batch_rows(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                                         int64_t steps, double delta, double rs, double r_max, double omega,
                                         const int32_t *row_idx, int64_t nidx, double *out, int64_t *out_rows)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    if (n < 0 || n > 0x7fffffffLL || nidx < 0 || nidx > 0x7fffffffLL) return fail(SCHW_E_INVALID, "n, nidx must be in [0, 2^31-1]");
    if (n > 0 && (!q0s || !p0s)) return fail(SCHW_E_INVALID, "null array");
    if (n > 0 && nidx > 0 && (!row_idx || !out)) return fail(SCHW_E_INVALID, "null array");
    for (int64_t k = 0; k < nidx; ++k)
        if (row_idx[k] < 0 || (k > 0 && row_idx[k] <= row_idx[k - 1]))
            return fail(SCHW_E_INVALID, "row_idx must be strictly ascending and non-negative");
    CU(cudaSetDevice(h->device));
    StepParams P;
    int rc = make_step_params(steps, delta, rs, r_max, omega, P);
    if (rc) return rc;
    if (n == 0) return SCHW_OK;
    const size_t b = (size_t)n * 4 * sizeof(double);
    const size_t ob = (size_t)n * (size_t)nidx * 4 * sizeof(double);
    ENSURE(h->in_a, b); ENSURE(h->in_b, b); ENSURE(h->traj, ob + 8); ENSURE(h->rows, (size_t)n * sizeof(long long));
    ENSURE(h->in_c, (size_t)(nidx + 1) * sizeof(int32_t));
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(h->in_a.p, q0s, b, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->in_b.p, p0s, b, cudaMemcpyHostToDevice, st));
    if (nidx > 0) CU(cudaMemcpyAsync(h->in_c.p, row_idx, (size_t)nidx * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
    if (ob) CU(cudaMemsetAsync(h->traj.p, 0, ob, st));
    begin_timed(h, st);
    integrate_rows_kernel<<<(unsigned)((n + 31) / 32), 32, 0, st>>>((const double *)h->in_a.p, (const double *)h->in_b.p,
                                                                      (long long)n, P, (const int32_t *)h->in_c.p, (int)nidx,
                                                                      (double *)h->traj.p, (long long *)h->rows.p, h->ctr);
    h->launches += 1;
    CU(cudaGetLastError());
    end_timed(h, st);
    if (ob) CU(cudaMemcpyAsync(out, h->traj.p, ob, cudaMemcpyDeviceToHost, st));
    if (out_rows) CU(cudaMemcpyAsync(out_rows, h->rows.p, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SCHW_OK;
}

// ---- stages 1+2 ---------------------------------------------------------------------------------
extern "C" int schw_initial_conditions(schw_handle *h, const schw_scene *scene, int row_begin, int row_end,
                                       int row_stride, double *q0s, double *ic)
{
    if (!h || !q0s || !ic) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    SceneParams sc;
    int rc = make_scene_params(scene, row_begin, row_end, row_stride, sc);
    if (rc) return rc;
    const long long n = (long long)sc.nrows * sc.width;
    if (n == 0) return SCHW_OK;
    ENSURE(h->fq, (size_t)n * 4 * sizeof(double));
    ENSURE(h->ic, (size_t)n * 9 * sizeof(double));
    CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), h->stream));
    begin_timed(h, h->stream);
    ic_kernel<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(sc, n, (double *)h->fq.p, (double *)h->ic.p, &h->ctr->err);
    h->launches += 1;
    CU(cudaGetLastError());
    end_timed(h, h->stream);
    CU(cudaMemcpyAsync(q0s, h->fq.p, (size_t)n * 4 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(ic, h->ic.p, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    return sync_and_check(h, h->stream);
}

extern "C" int schw_initial_conditions_px(schw_handle *h, const double *obs, const double *pixels, int64_t n,
                                          double mass, double *q0s, double *ic)
{
    if (!h || !obs || (n > 0 && (!pixels || !q0s || !ic))) return fail(SCHW_E_INVALID, "null argument");
    if (n < 0) return fail(SCHW_E_INVALID, "n < 0");
    CU(cudaSetDevice(h->device));
    schw_scene s;
    memset(&s, 0, sizeof(s));
    s.obs[0] = obs[0]; s.obs[1] = obs[1]; s.obs[2] = obs[2];
    s.fov = 1.0; s.height = 1; s.width = 1; s.mass = mass; s.r_max = 1.0; s.delta = 1.0; s.omega = 1.0;
    SceneParams sc;
    int rc = make_scene_params(&s, 0, 1, 1, sc);
    if (rc) return rc;
    if (n == 0) return SCHW_OK;
    cudaStream_t st = h->stream;
    ENSURE(h->in_a, (size_t)n * 24); ENSURE(h->fq, (size_t)n * 32); ENSURE(h->ic, (size_t)n * 72);
    CU(cudaMemcpyAsync(h->in_a.p, pixels, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
    begin_timed(h, st);
    ic_px_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc, (const double *)h->in_a.p, (long long)n,
                                                               (double *)h->fq.p, (double *)h->ic.p, &h->ctr->err);
    h->launches += 1;
    CU(cudaGetLastError());
    end_timed(h, st);
    CU(cudaMemcpyAsync(q0s, h->fq.p, (size_t)n * 32, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ic, h->ic.p, (size_t)n * 72, cudaMemcpyDeviceToHost, st));
    return sync_and_check(h, st);
}

// ---- stage 5 ------------------------------------------------------------------------------------
extern "C" int schw_classify(schw_handle *h, const schw_scene *scene, const double *final_q, const double *alpha0,
                             const double *beta, int64_t n, const uint8_t *tex, uint8_t *rgb, uint8_t *cls,
                             double *hit, int32_t *texel)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    if (n < 0) return fail(SCHW_E_INVALID, "n < 0");
    if (n > 0 && (!final_q || !alpha0 || !beta || !rgb || !cls)) return fail(SCHW_E_INVALID, "null array");
    if (scene->has_texture && !tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    CU(cudaSetDevice(h->device));
    SceneParams sc;
    int rc = make_scene_params(scene, 0, scene->height, 1, sc);
    if (rc) return rc;
    if (n == 0) return SCHW_OK;
    cudaStream_t st = h->stream;
    const size_t tb = (size_t)sc.height * sc.width * 3;
    ENSURE(h->fq, (size_t)n * 32); ENSURE(h->in_a, (size_t)n * 8); ENSURE(h->in_b, (size_t)n * 8);
    ENSURE(h->rgb, (size_t)n * 3); ENSURE(h->cls, (size_t)n);
    ENSURE(h->hit, (size_t)n * 16); ENSURE(h->texel, (size_t)n * 8);
    if (scene->has_texture) { ENSURE(h->tex, tb); CU(cudaMemcpyAsync(h->tex.p, tex, tb, cudaMemcpyHostToDevice, st)); }
    CU(cudaMemcpyAsync(h->fq.p, final_q, (size_t)n * 32, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->in_a.p, alpha0, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->in_b.p, beta, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    begin_timed(h, st);
    classify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc, (const double *)h->fq.p, (const double *)h->in_a.p,
                                                                  (const double *)h->in_b.p, (long long)n,
                                                                  (const uint8_t *)h->tex.p, (uint8_t *)h->rgb.p,
                                                                  (uint8_t *)h->cls.p, (double *)h->hit.p, (int32_t *)h->texel.p);
    h->launches += 1;
    CU(cudaGetLastError());
    end_timed(h, st);
    CU(cudaMemcpyAsync(rgb, h->rgb.p, (size_t)n * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cls, h->cls.p, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (hit) CU(cudaMemcpyAsync(hit, h->hit.p, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
    if (texel) CU(cudaMemcpyAsync(texel, h->texel.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SCHW_OK;
}

// ---- work order of the render queue -----------------------------------------------------------------
// Longest-processing-time-first over 32-pixel row segments.  The number of steps of a ray depends only on its
// camera angle alpha (distance rho from the image centre): escaped rays need the most steps just outside the
// critical impact parameter b = 3 sqrt(3) M and fewer towards the corners, captured rays the fewest.  Handing
// out the expensive tiles first leaves the cheap captured rays for the end of the launch, which halves the
// end-of-frame tail (it matters when a frame is split over 8 GPUs: ~50 ms of work per GPU against a ~4 ms tail).
// Purely a scheduling hint: every ray's result is independent of the order.
static int build_tile_order(schw_handle *h, const schw_scene *s, const SceneParams &sc, cudaStream_t st,
                            const int32_t **d_order, int *tiles_per_row, long long *qlen)
{
    const int tpr = (sc.width + 31) / 32;
    const long long ntiles = (long long)sc.nrows * tpr;
    *tiles_per_row = tpr;
    *qlen = ntiles * 32;
    *d_order = nullptr;
    if (ntiles == 0) return SCHW_OK;
    const double key[12] = {(double)sc.height, (double)sc.width, (double)sc.row_begin, (double)sc.row_stride, (double)sc.nrows,
                            sc.plane_width, sc.plane_height, sc.obs[0], sc.obs[1], sc.obs[2], s->mass, 1.0};
    if (h->order_tiles == ntiles && memcmp(key, h->order_key, sizeof(key)) == 0) {
        *d_order = (const int32_t *)h->order.p;
        return SCHW_OK;
    }
    // critical camera angle: b(alpha) = r sin(a) / sqrt(f (f^2 cos^2 a + sin^2 a)) = 3 sqrt(3) M  (reference momentum convention)
    const double r = sc.r_obs, f = sc.grr, bc = 3.0 * sqrt(3.0) * s->mass;
    double lo = 0.0, hi = 0.5 * M_PI;
    for (int it = 0; it < 60; ++it) {
        const double a = 0.5 * (lo + hi);
        const double b = r * sin(a) / sqrt(f * (f * f * cos(a) * cos(a) + sin(a) * sin(a)));
        if (b < bc) lo = a; else hi = a;
    }
    const double plane_dist = 0.2 * sqrt(sc.obs[0] * sc.obs[0] + sc.obs[1] * sc.obs[1] + sc.obs[2] * sc.obs[2]);
    const double rho_c = plane_dist * tan(0.5 * (lo + hi));
    const double rho_c2 = rho_c * rho_c;
    std::vector<std::pair<float, int32_t>> tiles((size_t)ntiles);
    for (int row = 0; row < sc.nrows; ++row) {
        const int i = sc.row_begin + row * sc.row_stride;
        const double v = ((i + 0.5) / sc.height - 0.5) * sc.plane_height;
        for (int seg = 0; seg < tpr; ++seg) {
            const int j0 = seg * 32, j1 = (j0 + 31 < sc.width - 1) ? j0 + 31 : sc.width - 1;
            // the pixel of the segment closest to the shadow edge decides (it is the most expensive one)
            const double u0 = ((j0 + 0.5) / sc.width - 0.5) * sc.plane_width, u1 = ((j1 + 0.5) / sc.width - 0.5) * sc.plane_width;
            const double ua = fmin(fabs(u0), fabs(u1)), ub = fmax(fabs(u0), fabs(u1));
            const double umin = (u0 * u1 < 0.0) ? 0.0 : ua;
            const double rmin2 = umin * umin + v * v, rmax2 = ub * ub + v * v;
            float k;
            if (rmax2 >= rho_c2) k = (float)(rmin2 > rho_c2 ? rmin2 - rho_c2 : 0.0);        // contains escaped rays: most expensive first
            else k = (float)(1e30 + (rho_c2 - rmax2));                                       // all captured: last, edge inward
            tiles[(size_t)row * tpr + seg] = std::make_pair(k, (int32_t)(row * tpr + seg));
        }
    }
    std::sort(tiles.begin(), tiles.end());
    std::vector<int32_t> order((size_t)ntiles);
    for (size_t t = 0; t < (size_t)ntiles; ++t) order[t] = tiles[t].second;
    // (Measured and rejected, r02: a "stratified" last round -- every c-th of the c*W shortest tiles instead of the W shortest,
    //  so that warps which become free at different times could end together -- helps one shard size by 0.5 ms and costs
    //  0.2-0.5 ms at others: bin-packing noise, not a rule.)
    ENSURE(h->order, (size_t)ntiles * sizeof(int32_t));
    CU(cudaMemcpyAsync(h->order.p, order.data(), (size_t)ntiles * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaStreamSynchronize(st));       // `order` is a local
    memcpy(h->order_key, key, sizeof(key));
    h->order_tiles = ntiles;
    *d_order = (const int32_t *)h->order.p;
    return SCHW_OK;
}

// SCHW_LPT_ORDER=0: plain row-major queue (A/B measurements)
static bool lpt_order_enabled() { return env_switch_on("SCHW_LPT_ORDER"); }
// Refill policy of the render queue (RenderJob::tile_mode): whole tiles by default -- measured faster at every frame size
// (profiles/r02_ab_tile.json: 400x400 7.12 -> 6.81 ms, a 1/8 shard of the 4K frame 39.26 -> 39.02 ms, the 4K frame
// 308.7 -> 306.1 ms, same output bits); SCHW_TILE_MODE=0 selects the per-lane refill (A/B measurements).
// SCHW_TILE_MODE: 0 = per-lane refill (persistent_rays), 1 = whole tiles inside persistent_rays, 2 (default) = the
// whole-tile kernel persistent_tiles<R>.
static int tile_mode_for(const schw_handle *, long long)
{
    const char *e = getenv("SCHW_TILE_MODE");
    if (e && (e[0] == '0' || e[0] == '1')) return e[0] - '0';
    return 2;
}

// ---- fused render -------------------------------------------------------------------------------
extern "C" int schw_render_dev_pitched(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex, int row_begin,
                                       int row_end, int row_stride, uint8_t *d_rgb, int64_t rgb_row_pitch,
                                       uint8_t *d_cls, int64_t cls_row_pitch, const schw_ray_outputs *x, void *cuda_stream)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    SceneParams sc;
    int rc = make_scene_params(scene, row_begin, row_end, row_stride, sc);
    if (rc) return rc;
    if (rgb_row_pitch == 0) rgb_row_pitch = 3LL * sc.width;
    if (cls_row_pitch == 0) cls_row_pitch = sc.width;
    if (rgb_row_pitch < 3LL * sc.width || cls_row_pitch < sc.width) return fail(SCHW_E_INVALID, "row pitch smaller than a row");
    if ((long long)sc.nrows * sc.width > 0x7fffffffLL)
        return fail(SCHW_E_INVALID, "more than 2^31-1 rays in one call (the kernel indexes rays with 32-bit integers)");
    StepParams P;
    rc = make_step_params(scene->steps, scene->delta, 2.0 * scene->mass, scene->r_max, scene->omega, P);
    if (rc) return rc;
    if (scene->has_texture && !d_tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    const long long n = (long long)sc.nrows * sc.width;
    if (n > 0 && (!d_rgb || !d_cls)) return fail(SCHW_E_INVALID, "null output");
    // Camera rays start at theta_obs with p_theta = -0.0: they are eligible for the planar fast path iff theta_obs is
    // the double nearest pi/2 (observer in the equatorial plane, the reference's only camera: raytracing.py:60-72).
    const bool off_plane = sc.th_obs != 1.5707963267948966;
    if (off_plane) P.pl_enable = 0;
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    RenderJob job;
    job.sc = sc; job.tex = d_tex; job.rgb = d_rgb; job.cls = d_cls;
    job.rgb_pitch = rgb_row_pitch; job.cls_pitch = cls_row_pitch; job.err = &h->ctr->err;
    job.tex_ready = scene->has_texture ? h->next_tex_ready : nullptr;
    h->next_tex_ready = nullptr;
    job.final_q = x ? x->final_q : nullptr;
    job.ic = x ? x->ic : nullptr;
    job.hit = x ? x->hit : nullptr;
    job.texel = x ? x->texel : nullptr;
    job.nsteps = x ? x->nsteps : nullptr;
    job.n = n;
    job.tile_order = nullptr; job.tiles_per_row = 1; job.qlen = n; job.tile_mode = 0;
    if (n > 0 && lpt_order_enabled()) {
        const int32_t *d_order = nullptr;
        rc = build_tile_order(h, scene, sc, st, &d_order, &job.tiles_per_row, &job.qlen);
        if (rc) return rc;
        job.tile_order = d_order;
        job.tile_mode = tile_mode_for(h, n);
    }
    begin_timed(h, st);
    // equatorial observer (every ray planar) + whole tiles: the tile kernel; otherwise the per-lane-refill kernel
    if (job.tile_mode == 2 && P.pl_enable && trig_memo_enabled()) rc = launch_tiles(h, job, job.qlen, P, st);
    else if (off_plane) {                       // tilted camera rays: no fast path, and the trig memo would never hit
        CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
        rc = n > 0 ? launch_variant<RenderJob, false, false>(h, job, P, h->occ_render[0], st) : SCHW_OK;
    } else rc = launch_persistent(h, job, P, h->occ_render, st);
    end_timed(h, st);
    return rc;
}

extern "C" int schw_render_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex, int row_begin,
                               int row_end, int row_stride, uint8_t *d_rgb, uint8_t *d_cls,
                               const schw_ray_outputs *x, void *cuda_stream)
{
    return schw_render_dev_pitched(h, scene, d_tex, row_begin, row_end, row_stride, d_rgb, 0, d_cls, 0, x, cuda_stream);
}

extern "C" int schw_set_tex_ready_flag(schw_handle *h, const uint32_t *d_flag)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    h->next_tex_ready = reinterpret_cast<const unsigned *>(d_flag);
    return SCHW_OK;
}

extern "C" int schw_render(schw_handle *h, const schw_scene *scene, const uint8_t *tex, int row_begin, int row_end,
                           int row_stride, uint8_t *rgb, uint8_t *cls, const schw_ray_outputs *x)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    if (scene->height <= 0 || scene->width <= 0) return fail(SCHW_E_INVALID, "image_size must be positive");
    if (row_stride <= 0 || row_begin < 0 || row_end > scene->height || row_begin > row_end)
        return fail(SCHW_E_INVALID, "bad row range");
    if (scene->has_texture && !tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    CU(cudaSetDevice(h->device));
    const int nrows = row_end > row_begin ? (row_end - row_begin + row_stride - 1) / row_stride : 0;
    const size_t n = (size_t)nrows * scene->width;
    if (n > 0x7fffffffULL) return fail(SCHW_E_INVALID, "more than 2^31-1 rays in one call (the kernel indexes rays with 32-bit integers)");
    if (n > 0 && (!rgb || !cls)) return fail(SCHW_E_INVALID, "null output");
    cudaStream_t st = h->stream;
    const size_t tb = (size_t)scene->height * scene->width * 3;
    if (scene->has_texture) {
        // Upload on the copy stream and raise tex_flag behind it; the kernel is launched at once on `st` and only the
        // epilogue of a ray (its texel lookup) waits for the flag.
        ENSURE(h->tex, tb);
        CU(cudaMemsetAsync(h->tex_flag, 0, sizeof(unsigned), st));
        CU(cudaEventRecord(h->ev_copy0, st));                   // (also orders this upload after the previous frame's kernel)
        CU(cudaStreamWaitEvent(h->copy_stream, h->ev_copy0, 0));
        CU(cudaMemcpyAsync(h->tex.p, tex, tb, cudaMemcpyHostToDevice, h->copy_stream));
        CU(cudaMemcpyAsync(h->tex_flag, h->one_pinned, sizeof(unsigned), cudaMemcpyHostToDevice, h->copy_stream));
        CU(cudaEventRecord(h->ev_copy1, h->copy_stream));
        h->next_tex_ready = h->tex_flag;
    }
    ENSURE(h->rgb, n * 3 + 1); ENSURE(h->cls, n + 1);
    schw_ray_outputs d = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (x) {
        if (x->final_q) { ENSURE(h->fq, n * 32); d.final_q = (double *)h->fq.p; }
        if (x->ic) { ENSURE(h->ic, n * 72); d.ic = (double *)h->ic.p; }
        if (x->hit) { ENSURE(h->hit, n * 16); d.hit = (double *)h->hit.p; }
        if (x->texel) { ENSURE(h->texel, n * 8); d.texel = (int32_t *)h->texel.p; }
        if (x->nsteps) { ENSURE(h->nsteps, n * 4); d.nsteps = (int32_t *)h->nsteps.p; }
    }
    int rc = schw_render_dev(h, scene, scene->has_texture ? (const uint8_t *)h->tex.p : nullptr, row_begin, row_end,
                             row_stride, (uint8_t *)h->rgb.p, (uint8_t *)h->cls.p, &d, nullptr);
    h->next_tex_ready = nullptr;
    if (scene->has_texture) CU(cudaStreamWaitEvent(st, h->ev_copy1, 0));   // the upload is part of this call
    if (rc) return rc;
    if (n > 0) {
        CU(cudaMemcpyAsync(rgb, h->rgb.p, n * 3, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(cls, h->cls.p, n, cudaMemcpyDeviceToHost, st));
        if (d.final_q) CU(cudaMemcpyAsync(x->final_q, d.final_q, n * 32, cudaMemcpyDeviceToHost, st));
        if (d.ic) CU(cudaMemcpyAsync(x->ic, d.ic, n * 72, cudaMemcpyDeviceToHost, st));
        if (d.hit) CU(cudaMemcpyAsync(x->hit, d.hit, n * 16, cudaMemcpyDeviceToHost, st));
        if (d.texel) CU(cudaMemcpyAsync(x->texel, d.texel, n * 8, cudaMemcpyDeviceToHost, st));
        if (d.nsteps) CU(cudaMemcpyAsync(x->nsteps, d.nsteps, n * 4, cudaMemcpyDeviceToHost, st));
    }
    return sync_and_check(h, st);
}

// ---- de-duplicated render (N4, opt-in) ------------------------------------------------------------
// Same contract and same output bits as schw_render_dev, but pixels whose camera angle agrees bitwise share one
// integration (see dedup_key_kernel).  Three stages, also exported one by one so that the ranks of a multi-GPU job can
// share the integration (schw_b200/multigpu.py): plan (keys -> sort -> distinct rays, ONE stream synchronisation: the
// number of distinct rays sizes what follows), integrate (a rank's share of the distinct rays), epilogue (per pixel).
struct DedupPlanKey { double v[16]; };
static DedupPlanKey dedup_plan_key(const schw_scene *scene, int rb, int rs, int nrows, int world)
{
    DedupPlanKey k = {{scene->obs[0], scene->obs[1], scene->obs[2], scene->fov, (double)scene->height, (double)scene->width,
                      (double)rb, (double)rs, (double)nrows, scene->mass, (double)world, 0.0, 0.0, 0.0, 0.0, 1.0}};
    return k;
}

// Stage 1.  Rows (row_begin, row_end, row_stride) of the frame -> h->dd_map (slot of every packed pixel), h->dd_q0/p0
// (initial conditions per slot), h->last_unique, h->dd_chunk.  Cached per camera geometry + rows + world.
static int dedup_plan(schw_handle *h, const schw_scene *scene, int row_begin, int row_end, int row_stride, int world,
                      cudaStream_t st)
{
    SceneParams sc;
    int rc = make_scene_params(scene, row_begin, row_end, row_stride, sc);
    if (rc) return rc;
    if (world < 1) return fail(SCHW_E_INVALID, "world must be >= 1");
    const long long n = (long long)sc.nrows * sc.width;
    if (n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "more than 2^31-1 rays");
    const DedupPlanKey key = dedup_plan_key(scene, sc.row_begin, sc.row_stride, sc.nrows, world);
    if (h->dd_plan_valid && memcmp(&key, &h->dd_plan_key, sizeof(key)) == 0) {
        CU(cudaStreamWaitEvent(st, h->ev_plan, 0));
        return SCHW_OK;
    }
    h->dd_plan_valid = false; h->dd_valid = false;
    h->last_unique = 0; h->dd_chunk = 0; h->dd_world = world;
    h->dd_plan_rb = sc.row_begin; h->dd_plan_rs = sc.row_stride; h->dd_plan_rows = sc.nrows;
    if (n == 0) { memcpy(&h->dd_plan_key, &key, sizeof(key)); h->dd_plan_valid = true; CU(cudaEventRecord(h->ev_plan, st)); return SCHW_OK; }
    ENSURE(h->dd_keys, (size_t)n * 8); ENSURE(h->dd_keys2, (size_t)n * 8);
    ENSURE(h->dd_pix, (size_t)n * 4); ENSURE(h->dd_pix2, (size_t)n * 4);
    ENSURE(h->dd_head, (size_t)n * 4); ENSURE(h->dd_uid, (size_t)n * 4); ENSURE(h->dd_map, (size_t)n * 4);
    auto *keys = (unsigned long long *)h->dd_keys.p, *keys2 = (unsigned long long *)h->dd_keys2.p;
    auto *pix = (unsigned *)h->dd_pix.p, *pix2 = (unsigned *)h->dd_pix2.p;
    int *head = (int *)h->dd_head.p, *uid1 = (int *)h->dd_uid.p;
    size_t tmp_sort = 0, tmp_scan = 0;
    CU(cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys, keys2, pix, pix2, (int)n, 0, 64, st));
    CU(cub::DeviceScan::InclusiveSum(nullptr, tmp_scan, head, uid1, (int)n, st));
    ENSURE(h->dd_tmp, tmp_sort > tmp_scan ? tmp_sort : tmp_scan);
    const unsigned blocks = (unsigned)((n + 255) / 256);
    dedup_key_kernel<<<blocks, 256, 0, st>>>(sc, n, keys, pix);
    CU(cudaGetLastError());
    size_t tmp = h->dd_tmp.cap;
    CU(cub::DeviceRadixSort::SortPairs(h->dd_tmp.p, tmp, keys, keys2, pix, pix2, (int)n, 0, 64, st));
    dedup_heads_kernel<<<blocks, 256, 0, st>>>(keys2, n, head);
    CU(cudaGetLastError());
    tmp = h->dd_tmp.cap;
    CU(cub::DeviceScan::InclusiveSum(h->dd_tmp.p, tmp, head, uid1, (int)n, st));
    int n_unique = 0;
    CU(cudaMemcpyAsync(&n_unique, uid1 + (n - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));          // (the cub sort/scan are library calls on an opt-in path; this is its one host sync)
    if (n_unique <= 0 || n_unique > n) return fail(SCHW_E_CUDA, "de-duplication produced an impossible ray count");
    const int chunk = (n_unique + world - 1) / world;
    ENSURE(h->dd_q0, (size_t)chunk * world * 32); ENSURE(h->dd_p0, (size_t)chunk * world * 32);
    dedup_scatter_kernel<<<blocks, 256, 0, st>>>(sc, n, pix2, head, uid1, (int *)h->dd_map.p, (double *)h->dd_q0.p,
                                                  (double *)h->dd_p0.p, world, chunk);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ev_plan, st));
    h->launches += 5;                                   // key, sort, heads, scan, scatter
    h->last_unique = n_unique; h->dd_chunk = chunk;
    memcpy(&h->dd_plan_key, &key, sizeof(key));
    h->dd_plan_valid = true;
    return SCHW_OK;
}

// Stage 2.  Rank `rank`'s share of the distinct rays of the current plan: slots [rank*chunk, rank*chunk + count) of
// d_fq_all (world*chunk, 4) / d_ns_all (world*chunk).
static int dedup_integrate(schw_handle *h, const schw_scene *scene, int rank, double *d_fq_all, int32_t *d_ns_all, cudaStream_t st)
{
    if (!h->dd_plan_valid) return fail(SCHW_E_INVALID, "no de-duplication plan on this handle");
    const int world = h->dd_world, chunk = h->dd_chunk;
    if (rank < 0 || rank >= world) return fail(SCHW_E_INVALID, "rank out of range");
    const long long U = h->last_unique;
    const long long count = U > rank ? (U - rank + world - 1) / world : 0;       // distinct rays u = rank (mod world)
    const size_t first = (size_t)rank * chunk;
    return integrate_batch_impl(h, (const double *)h->dd_q0.p + 4 * first, (const double *)h->dd_p0.p + 4 * first, count,
                                scene->steps, scene->delta, 2.0 * scene->mass, scene->r_max, scene->omega,
                                d_fq_all + 4 * first, d_ns_all + first, st, /*coherent=*/true);
}

// Stage 3.  Per-pixel epilogue of the rows of THIS call (a subset of the plan's rows) from the distinct rays' final states.
static int dedup_epilogue(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex, int row_begin, int row_end,
                          int row_stride, uint8_t *d_rgb, long long rgb_pitch, uint8_t *d_cls, long long cls_pitch,
                          const schw_ray_outputs *x, const double *d_fq_all, const int32_t *d_ns_all, cudaStream_t st)
{
    if (!h->dd_plan_valid) return fail(SCHW_E_INVALID, "no de-duplication plan on this handle");
    SceneParams sc;
    int rc = make_scene_params(scene, row_begin, row_end, row_stride, sc);
    if (rc) return rc;
    const long long n = (long long)sc.nrows * sc.width;
    if (n == 0) return SCHW_OK;
    if (scene->has_texture && !d_tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    if (!d_rgb || !d_cls) return fail(SCHW_E_INVALID, "null output");
    const int prb = h->dd_plan_rb, prs = h->dd_plan_rs;
    const int last = sc.row_begin + (sc.nrows - 1) * sc.row_stride;
    if ((sc.row_begin - prb) % prs != 0 || (sc.nrows > 1 && sc.row_stride % prs != 0) || sc.row_begin < prb ||
        (last - prb) / prs >= h->dd_plan_rows)
        return fail(SCHW_E_INVALID, "the rows of this call are not covered by the de-duplication plan");
    if (rgb_pitch == 0) rgb_pitch = 3LL * sc.width;
    if (cls_pitch == 0) cls_pitch = sc.width;
    if (rgb_pitch < 3LL * sc.width || cls_pitch < sc.width) return fail(SCHW_E_INVALID, "row pitch smaller than a row");
    RenderJob job;
    job.sc = sc; job.tex = d_tex; job.tex_ready = nullptr; job.rgb = d_rgb; job.cls = d_cls;
    job.rgb_pitch = rgb_pitch; job.cls_pitch = cls_pitch; job.err = &h->ctr->err;
    job.final_q = x ? x->final_q : nullptr;
    job.ic = x ? x->ic : nullptr;
    job.hit = x ? x->hit : nullptr;
    job.texel = x ? x->texel : nullptr;
    job.nsteps = x ? x->nsteps : nullptr;
    job.n = n;
    job.tile_order = nullptr; job.tiles_per_row = 1; job.qlen = n; job.tile_mode = 0;
    CU(cudaMemsetAsync(&h->ctr->pixel_steps, 0, sizeof(unsigned long long), st));
    dedup_epilogue_kernel<<<(unsigned)((n + kBlock - 1) / kBlock), kBlock, 0, st>>>(job, (const int *)h->dd_map.p, prb, prs,
                                                                                  d_fq_all, d_ns_all, h->ctr);
    h->launches += 1;
    CU(cudaGetLastError());
    return SCHW_OK;
}

extern "C" int schw_render_dedup_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex, int row_begin,
                                     int row_end, int row_stride, uint8_t *d_rgb, uint8_t *d_cls,
                                     const schw_ray_outputs *x, void *cuda_stream)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    SceneParams sc;
    int rc = make_scene_params(scene, row_begin, row_end, row_stride, sc);
    if (rc) return rc;
    if (scene->has_texture && !d_tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    const long long n = (long long)sc.nrows * sc.width;
    if (n > 0x7fffffffLL) return fail(SCHW_E_INVALID, "more than 2^31-1 rays");
    if (n > 0 && (!d_rgb || !d_cls)) return fail(SCHW_E_INVALID, "null output");
    {   // validate the step parameters before any work is enqueued
        StepParams P;
        rc = make_step_params(scene->steps, scene->delta, 2.0 * scene->mass, scene->r_max, scene->omega, P);
        if (rc) return rc;
    }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    begin_timed(h, st);
    if (n == 0) { h->last_unique = 0; CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st)); end_timed(h, st); return SCHW_OK; }
    // The geodesics of a frame depend on the camera, the rows and the integrator parameters -- not on the sky patch, the
    // flips or the texture.  A call that differs from the previous one on this handle only in those (an azimuth sweep:
    // SURVEY §8d config 5, "integrate once, 360 lookups") reuses them and runs the per-pixel epilogue alone.
    const double key[16] = {scene->obs[0], scene->obs[1], scene->obs[2], scene->fov, (double)scene->height, (double)scene->width,
                            (double)sc.row_begin, (double)sc.row_stride, (double)sc.nrows, scene->mass, scene->r_max,
                            (double)scene->steps, scene->delta, scene->omega, (double)n, 1.0};
    const bool reuse = env_switch_on("SCHW_DEDUP_REUSE") && h->dd_valid && h->dd_plan_valid && h->dd_world == 1 &&
                       memcmp(key, h->dd_key, sizeof(key)) == 0;
    if (reuse) {
        CU(cudaStreamWaitEvent(st, h->ev_dd, 0));                       // (the cache may have been filled on another stream)
        CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));          // no step is integrated by this call
    } else {
        h->dd_valid = false;
        if (!env_switch_on("SCHW_DEDUP_REUSE")) h->dd_plan_valid = false;
        const bool owns = h->in_pipeline;
        h->in_pipeline = true;                                          // this entry point owns the timing events
        rc = dedup_plan(h, scene, row_begin, row_end, row_stride, 1, st);
        if (!rc) {
            ENSURE(h->dd_fq, (size_t)h->dd_chunk * 32); ENSURE(h->dd_ns, (size_t)h->dd_chunk * 4);
            rc = dedup_integrate(h, scene, 0, (double *)h->dd_fq.p, (int32_t *)h->dd_ns.p, st);
        }
        h->in_pipeline = owns;
        if (rc) return rc;
        CU(cudaEventRecord(h->ev_dd, st));
        memcpy(h->dd_key, key, sizeof(key));
        h->dd_valid = true;
    }
    rc = dedup_epilogue(h, scene, d_tex, row_begin, row_end, row_stride, d_rgb, 0, d_cls, 0, x, (const double *)h->dd_fq.p,
                        (const int32_t *)h->dd_ns.p, st);
    if (rc) return rc;
    end_timed(h, st);
    return SCHW_OK;
}

// The three stages for the ranks of a multi-GPU job (schw_b200/multigpu.py): every rank plans the WHOLE frame (identical
// on all ranks: the kernels and cub's radix sort are deterministic), integrates its share of the distinct rays into its
// slots of the (world*chunk, 4) / (world*chunk) arrays, the caller all-gathers those, and every rank runs the epilogue
// of its own rows.
extern "C" int schw_dedup_plan_dev(schw_handle *h, const schw_scene *scene, int world, uint64_t *n_unique, uint64_t *chunk,
                                   void *cuda_stream)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    begin_timed(h, st);
    h->in_pipeline = true;
    int rc = dedup_plan(h, scene, 0, scene->height, 1, world, st);
    h->in_pipeline = false;
    if (rc) return rc;
    end_timed(h, st);
    if (n_unique) *n_unique = (uint64_t)h->last_unique;
    if (chunk) *chunk = (uint64_t)h->dd_chunk;
    return SCHW_OK;
}

extern "C" int schw_dedup_integrate_dev(schw_handle *h, const schw_scene *scene, int rank, double *d_fq_all,
                                        int32_t *d_ns_all, void *cuda_stream)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    if (!d_fq_all || !d_ns_all) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    return dedup_integrate(h, scene, rank, d_fq_all, d_ns_all, st);
}

extern "C" int schw_dedup_epilogue_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex, int row_begin,
                                       int row_end, int row_stride, uint8_t *d_rgb, int64_t rgb_row_pitch, uint8_t *d_cls,
                                       int64_t cls_row_pitch, const schw_ray_outputs *x, const double *d_fq_all,
                                       const int32_t *d_ns_all, void *cuda_stream)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    if (!d_fq_all || !d_ns_all) return fail(SCHW_E_INVALID, "null array");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    begin_timed(h, st);
    int rc = dedup_epilogue(h, scene, d_tex, row_begin, row_end, row_stride, d_rgb, rgb_row_pitch, d_cls, cls_row_pitch, x,
                            d_fq_all, d_ns_all, st);
    if (rc) return rc;
    end_timed(h, st);
    return SCHW_OK;
}

extern "C" int schw_render_dedup(schw_handle *h, const schw_scene *scene, const uint8_t *tex, int row_begin, int row_end,
                                 int row_stride, uint8_t *rgb, uint8_t *cls, const schw_ray_outputs *x)
{
    if (!h || !scene) return fail(SCHW_E_INVALID, "null argument");
    if (scene->height <= 0 || scene->width <= 0) return fail(SCHW_E_INVALID, "image_size must be positive");
    if (row_stride <= 0 || row_begin < 0 || row_end > scene->height || row_begin > row_end)
        return fail(SCHW_E_INVALID, "bad row range");
    if (scene->has_texture && !tex) return fail(SCHW_E_INVALID, "has_texture set but tex is NULL");
    CU(cudaSetDevice(h->device));
    const int nrows = row_end > row_begin ? (row_end - row_begin + row_stride - 1) / row_stride : 0;
    const size_t n = (size_t)nrows * scene->width;
    if (n > 0x7fffffffULL) return fail(SCHW_E_INVALID, "more than 2^31-1 rays in one call (the kernel indexes rays with 32-bit integers)");
    if (n > 0 && (!rgb || !cls)) return fail(SCHW_E_INVALID, "null output");
    cudaStream_t st = h->stream;
    const size_t tb = (size_t)scene->height * scene->width * 3;
    if (scene->has_texture) { ENSURE(h->tex, tb); CU(cudaMemcpyAsync(h->tex.p, tex, tb, cudaMemcpyHostToDevice, st)); }
    ENSURE(h->rgb, n * 3 + 1); ENSURE(h->cls, n + 1);
    schw_ray_outputs d = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (x) {
        if (x->final_q) { ENSURE(h->fq, n * 32); d.final_q = (double *)h->fq.p; }
        if (x->ic) { ENSURE(h->ic, n * 72); d.ic = (double *)h->ic.p; }
        if (x->hit) { ENSURE(h->hit, n * 16); d.hit = (double *)h->hit.p; }
        if (x->texel) { ENSURE(h->texel, n * 8); d.texel = (int32_t *)h->texel.p; }
        if (x->nsteps) { ENSURE(h->nsteps, n * 4); d.nsteps = (int32_t *)h->nsteps.p; }
    }
    int rc = schw_render_dedup_dev(h, scene, scene->has_texture ? (const uint8_t *)h->tex.p : nullptr, row_begin, row_end,
                                   row_stride, (uint8_t *)h->rgb.p, (uint8_t *)h->cls.p, &d, nullptr);
    if (rc) return rc;
    if (n > 0) {
        CU(cudaMemcpyAsync(rgb, h->rgb.p, n * 3, cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(cls, h->cls.p, n, cudaMemcpyDeviceToHost, st));
        if (d.final_q) CU(cudaMemcpyAsync(x->final_q, d.final_q, n * 32, cudaMemcpyDeviceToHost, st));
        if (d.ic) CU(cudaMemcpyAsync(x->ic, d.ic, n * 72, cudaMemcpyDeviceToHost, st));
        if (d.hit) CU(cudaMemcpyAsync(x->hit, d.hit, n * 16, cudaMemcpyDeviceToHost, st));
        if (d.texel) CU(cudaMemcpyAsync(x->texel, d.texel, n * 8, cudaMemcpyDeviceToHost, st));
        if (d.nsteps) CU(cudaMemcpyAsync(x->nsteps, d.nsteps, n * 4, cudaMemcpyDeviceToHost, st));
    }
    return sync_and_check(h, st);
}

// Distinct rays integrated by the last schw_render_dedup* call, and the steps summed over its pixels (what one ray
// per pixel would have taken; schw_last_total_steps reports the steps actually integrated).
extern "C" int schw_last_dedup_stats(schw_handle *h, uint64_t *unique_rays, uint64_t *pixel_steps)
{
    if (!h) return fail(SCHW_E_INVALID, "null handle");
    CU(cudaSetDevice(h->device));
    Counters c;
    CU(cudaStreamSynchronize(h->last_stream ? h->last_stream : h->stream));
    CU(cudaMemcpy(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    if (unique_rays) *unique_rays = (uint64_t)h->last_unique;
    if (pixel_steps) *pixel_steps = c.pixel_steps;
    return SCHW_OK;
}

// ---- flat render (N1) ---------------------------------------------------------------------------
static void cross3(const double a[3], const double b[3], double o[3])
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static double pymod_host(double a, double b)
{
    double m = fmod(a, b);
    if (m != 0.0) { if ((m < 0.0) != (b < 0.0)) m += b; }
    else m = copysign(0.0, b);
    return m;
}

// sphere + patch + texture part of FlatParams (background.py:60-66 / cuda_geodesic.py:65-78)
static void flat_patch_params(FlatParams &fp, const double obs[3], double boundary_radius, double pct, double pcp, double pst,
                              double psp, int flip_theta, int flip_phi, int tex_h, int tex_w)
{
    for (int a = 0; a < 3; ++a) fp.obs[a] = obs[a];
    fp.r2 = boundary_radius * boundary_radius;
    fp.obs_dot_obs = obs[0] * obs[0] + obs[1] * obs[1] + obs[2] * obs[2];
    fp.theta0 = pct - pst / 2;
    fp.theta1 = pct + pst / 2;
    fp.phi0 = pcp - psp / 2;
    fp.phi1 = pcp + psp / 2;
    const double two_pi = 2 * M_PI;
    fp.phi_span = pymod_host(fp.phi1 - fp.phi0, two_pi);
    if (fp.phi_span == 0) fp.phi_span = two_pi;
    fp.phi0m = pymod_host(fp.phi0, two_pi);
    fp.phi1m = pymod_host(fp.phi1, two_pi);
    fp.height = tex_h; fp.width = tex_w; fp.flip_theta = flip_theta; fp.flip_phi = flip_phi;
}

extern "C" int schw_flat_raytrace(schw_handle *h, const double *obs, const double *ray_dirs, int64_t n, double boundary_radius,
                                  double patch_center_theta, double patch_center_phi, double patch_size_theta,
                                  double patch_size_phi, int flip_theta, int flip_phi, const uint8_t *bg, int tex_h, int tex_w,
                                  uint8_t *out_rgb, const int32_t *sampled_indices, int64_t n_sampled, int64_t n_points,
                                  double *out_trajs)
{
    if (!h || !obs || !bg) return fail(SCHW_E_INVALID, "null argument");
    if (n < 0 || n > 0x7fffffffLL || n_sampled < 0 || n_points < 0 || n_sampled * n_points > 0x7fffffffLL)
        return fail(SCHW_E_INVALID, "bad size");
    if (tex_h <= 0 || tex_w <= 0) return fail(SCHW_E_INVALID, "texture size must be positive");
    if (n > 0 && (!ray_dirs || !out_rgb)) return fail(SCHW_E_INVALID, "null array");
    if (n_sampled > 0 && n_points > 0 && (!sampled_indices || !out_trajs)) return fail(SCHW_E_INVALID, "null trajectory array");
    CU(cudaSetDevice(h->device));
    FlatParams fp;
    memset(&fp, 0, sizeof(fp));
    flat_patch_params(fp, obs, boundary_radius, patch_center_theta, patch_center_phi, patch_size_theta, patch_size_phi,
                      flip_theta, flip_phi, tex_h, tex_w);
    if (n == 0) return SCHW_OK;
    cudaStream_t st = h->stream;
    const size_t tb = (size_t)tex_h * tex_w * 3, nt = (size_t)n_sampled * n_points;
    ENSURE(h->tex, tb); ENSURE(h->rgb, (size_t)n * 3 + 1); ENSURE(h->in_a, (size_t)n * 24);
    CU(cudaMemcpyAsync(h->tex.p, bg, tb, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->in_a.p, ray_dirs, (size_t)n * 24, cudaMemcpyHostToDevice, st));
    begin_timed(h, st);
    flat_dirs_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp, (const double *)h->in_a.p, (long long)n,
                                                                   (const uint8_t *)h->tex.p, (uint8_t *)h->rgb.p);
    h->launches += 1;
    CU(cudaGetLastError());
    if (nt > 0) {
        ENSURE(h->in_c, (size_t)n_sampled * sizeof(int32_t)); ENSURE(h->traj, nt * 24);
        CU(cudaMemcpyAsync(h->in_c.p, sampled_indices, (size_t)n_sampled * sizeof(int32_t), cudaMemcpyHostToDevice, st));
        CU(cudaMemsetAsync(h->traj.p, 0, nt * 24, st));
        flat_traj_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, st>>>(fp, (const double *)h->in_a.p, (long long)n,
                                                                        (const int32_t *)h->in_c.p, (int)n_sampled, (int)n_points,
                                                                        (double *)h->traj.p);
        h->launches += 1;
        CU(cudaGetLastError());
    }
    end_timed(h, st);
    CU(cudaMemcpyAsync(out_rgb, h->rgb.p, (size_t)n * 3, cudaMemcpyDeviceToHost, st));
    if (nt > 0) CU(cudaMemcpyAsync(out_trajs, h->traj.p, nt * 24, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SCHW_OK;
}

extern "C" int schw_flat_render(schw_handle *h, const schw_scene *s, const uint8_t *tex, uint8_t *rgb)
{
    if (!h || !s || !tex || !rgb) return fail(SCHW_E_INVALID, "null argument");
    if (s->height <= 0 || s->width <= 0) return fail(SCHW_E_INVALID, "image_size must be positive");
    CU(cudaSetDevice(h->device));
    FlatParams fp;
    // camera basis, background.py:37-52
    const double nrm = sqrt(s->obs[0] * s->obs[0] + s->obs[1] * s->obs[1] + s->obs[2] * s->obs[2]);
    if (!(nrm > 0.0)) return fail(SCHW_E_INVALID, "observer at the origin");
    double axis[3] = {(0.0 - s->obs[0]) / nrm, (0.0 - s->obs[1]) / nrm, (0.0 - s->obs[2]) / nrm};
    double up_guess[3] = {0, 0, 1}, cr[3];
    cross3(axis, up_guess, cr);
    if (fabs(cr[0]) <= 1e-8 && fabs(cr[1]) <= 1e-8 && fabs(cr[2]) <= 1e-8) { up_guess[1] = 1; up_guess[2] = 0; }   // np.allclose(.., 0)
    double right[3], up[3];
    cross3(up_guess, axis, right);
    double rn = sqrt(right[0] * right[0] + right[1] * right[1] + right[2] * right[2]);
    for (int a = 0; a < 3; ++a) right[a] /= rn;
    cross3(axis, right, up);
    double un = sqrt(up[0] * up[0] + up[1] * up[1] + up[2] * up[2]);
    for (int a = 0; a < 3; ++a) up[a] /= un;
    const double plane_dist = 0.2 * nrm;
    for (int a = 0; a < 3; ++a) { fp.pc[a] = s->obs[a] + axis[a] * plane_dist; fp.right[a] = right[a]; fp.up[a] = up[a]; }
    fp.plane_width = 2 * plane_dist * tan(s->fov / 2);
    fp.plane_height = fp.plane_width * ((double)s->height / (double)s->width);
    flat_patch_params(fp, s->obs, s->r_max, s->patch_center_theta, s->patch_center_phi, s->patch_size_theta, s->patch_size_phi,
                      s->flip_theta, s->flip_phi, s->height, s->width);
    const size_t n = (size_t)s->height * s->width;
    cudaStream_t st = h->stream;
    ENSURE(h->tex, n * 3); ENSURE(h->rgb, n * 3 + 1);
    CU(cudaMemcpyAsync(h->tex.p, tex, n * 3, cudaMemcpyHostToDevice, st));
    begin_timed(h, st);
    flat_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp, (const uint8_t *)h->tex.p, (uint8_t *)h->rgb.p);
    h->launches += 1;
    CU(cudaGetLastError());
    end_timed(h, st);
    CU(cudaMemcpyAsync(rgb, h->rgb.p, n * 3, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SCHW_OK;
}

// ---- device math probe (diagnostic) -------------------------------------------------------------
extern "C" int schw_math_probe(schw_handle *h, const double *x, int64_t n, double *rcp, double *sn, double *cs)
{
    if (!h || n < 0 || (n > 0 && (!x || !rcp || !sn || !cs))) return fail(SCHW_E_INVALID, "bad argument");
    CU(cudaSetDevice(h->device));
    if (n == 0) return SCHW_OK;
    cudaStream_t st = h->stream;
    const size_t b = (size_t)n * 8;
    ENSURE(h->in_a, b); ENSURE(h->in_b, b); ENSURE(h->in_c, b); ENSURE(h->fq, b);
    CU(cudaMemcpyAsync(h->in_a.p, x, b, cudaMemcpyHostToDevice, st));
    math_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const double *)h->in_a.p, (long long)n, (double *)h->in_b.p,
                                                                    (double *)h->in_c.p, (double *)h->fq.p);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(rcp, h->in_b.p, b, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sn, h->in_c.p, b, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(cs, h->fq.p, b, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SCHW_OK;
}

// ---- FP64 peak probe ----------------------------------------------------------------------------
extern "C" int schw_fp64_peak_probe(schw_handle *h, double *tflops, float *ms_out)
{
    if (!h || !tflops) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    // Several shapes (chains per thread x threads per block x blocks per SM); the best one is the peak.
    struct Shape { int chains, threads, blocks_per_sm; };
    const Shape shapes[] = {{8, 256, 8}, {4, 256, 8}, {16, 128, 8}, {8, 128, 4}, {12, 256, 4}};
    const int iters = 1 << 13;
    ENSURE(h->in_c, (size_t)h->sm_count * 8 * 256 * sizeof(double));
    cudaStream_t st = h->stream;
    double best_tflops = 0.0;
    float best = 0.f;
    for (const Shape &sh : shapes) {
        const int blocks = h->sm_count * sh.blocks_per_sm;
        for (int rep = 0; rep < 3; ++rep) {
            CU(cudaEventRecord(h->ev0, st));
            switch (sh.chains) {
            case 4: dfma_probe_kernel<4><<<blocks, sh.threads, 0, st>>>((double *)h->in_c.p, iters, 1.0000001, 1e-9); break;
            case 8: dfma_probe_kernel<8><<<blocks, sh.threads, 0, st>>>((double *)h->in_c.p, iters, 1.0000001, 1e-9); break;
            case 12: dfma_probe_kernel<12><<<blocks, sh.threads, 0, st>>>((double *)h->in_c.p, iters, 1.0000001, 1e-9); break;
            default: dfma_probe_kernel<16><<<blocks, sh.threads, 0, st>>>((double *)h->in_c.p, iters, 1.0000001, 1e-9); break;
            }
            CU(cudaGetLastError());
            CU(cudaEventRecord(h->ev1, st));
            CU(cudaEventSynchronize(h->ev1));
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
            const double tf = 2.0 * 8.0 * sh.chains * (double)iters * (double)blocks * sh.threads / (ms * 1e-3) / 1e12;
            if (rep > 0 && tf > best_tflops) { best_tflops = tf; best = ms; }
        }
    }
    h->timed = true; h->launches = 15;
    *tflops = best_tflops;
    if (ms_out) *ms_out = best;
    return SCHW_OK;
}

// ---- step-loop probe (diagnostic) -----------------------------------------------------------------
// variant: rays per thread * 100 + CTAs per SM requested (106, 105, 203, 204).  Returns kernel ms and the warp-steps run.
extern "C" int schw_step_probe(schw_handle *h, int variant, int iters, double delta, float *ms, uint64_t *warp_steps,
                               int *ctas_per_sm, int *regs, double *planar_fraction)
{
    if (!h || !ms || !warp_steps) return fail(SCHW_E_INVALID, "null argument");
    CU(cudaSetDevice(h->device));
    StepParams P;
    int rc = make_step_params(1 << 30, delta, 2.0, 1e6, 1.0, P);
    if (rc) return rc;
    ENSURE(h->in_c, (size_t)h->sm_count * 8 * kBlock * sizeof(double));
    cudaStream_t st = h->stream;
    const void *fn = nullptr;
    switch (variant) {
    case 106: fn = (const void *)step_probe_kernel<1, 6>; break;
    case 105: fn = (const void *)step_probe_kernel<1, 5>; break;
    case 104: fn = (const void *)step_probe_kernel<1, 4>; break;
    case 103: fn = (const void *)step_probe_kernel<1, 3>; break;
    case 102: fn = (const void *)step_probe_kernel<1, 2>; break;
    case 101: fn = (const void *)step_probe_kernel<1, 1>; break;
    case 201: fn = (const void *)step_probe_kernel<2, 1>; break;
    case 203: fn = (const void *)step_probe_kernel<2, 3>; break;
    case 204: fn = (const void *)step_probe_kernel<2, 4>; break;
    case 202: fn = (const void *)step_probe_kernel<2, 2>; break;
    default: return fail(SCHW_E_INVALID, "unknown probe variant");
    }
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, kBlock, 0));
    cudaFuncAttributes fa;
    CU(cudaFuncGetAttributes(&fa, fn));
    const int want = variant % 100;
    if (occ > want) occ = want;
    const int blocks = h->sm_count * occ;
    double *out = (double *)h->in_c.p;
    int period = 2000, rounds = (iters + period - 1) / period;
    Counters *ctr = h->ctr;
    void *args[] = {(void *)&P, (void *)&rounds, (void *)&period, (void *)&out, (void *)&ctr};
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaMemsetAsync(h->ctr, 0, sizeof(Counters), st));
        CU(cudaEventRecord(h->ev0, st));
        CU(cudaLaunchKernel(fn, dim3(blocks), dim3(kBlock), args, 0, st));
        CU(cudaEventRecord(h->ev1, st));
        CU(cudaEventSynchronize(h->ev1));
        float t = 0.f;
        CU(cudaEventElapsedTime(&t, h->ev0, h->ev1));
        if (t < best) best = t;
    }
    *ms = best;
    *warp_steps = (uint64_t)blocks * (kBlock / 32) * (uint64_t)rounds * (uint64_t)period;
    Counters c;
    CU(cudaMemcpy(&c, h->ctr, sizeof(c), cudaMemcpyDeviceToHost));
    if (planar_fraction) *planar_fraction = c.warp_iters ? (double)c.planar_iters / (double)c.warp_iters : 0.0;
    if (ctas_per_sm) *ctas_per_sm = occ;
    if (regs) *regs = fa.numRegs;
    return SCHW_OK;
}
