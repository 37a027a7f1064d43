/*
 * schw_b200.h — C ABI of libschw_b200.so, the B200 (sm_100a) replacement for the per-ray
 * geodesic hot path of ryanchen2134/schwarzchild-einsteinpy-raytracing.
 *
 * Plain C: opaque handle, pointers and sizes only; no torch / numpy / C++ types.
 * Every entry point names the reference interface it replaces (file:line relative to the
 * reference checkout).  All functions return 0 on success or a negative SCHW_E_* code;
 * schw_last_error() returns a thread-local human-readable message for the last failure.
 *
 * Threading: a handle owns one CUDA stream and its scratch buffers; calls on one handle must
 * not overlap (thread-compatible, not thread-safe).  Different handles are independent.
 * There is no CPU fallback: every compute entry point fails with SCHW_E_CUDA when no
 * sm_100-class device is usable.
 *
 * Pointer residency: entry points without a suffix take HOST pointers and do the H2D/D2H
 * copies themselves on the handle's stream (synchronous on return, like the reference's
 * implicit Numba transfers at cuda_geodesic.py:644,652).  `_dev` entry points take DEVICE
 * pointers (e.g. torch tensors' data_ptr()) and only enqueue work on the given stream.
 */
#ifndef SCHW_B200_H
#define SCHW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCHW_ABI_VERSION 2

/* error codes */
#define SCHW_OK            0
#define SCHW_E_INVALID    -1   /* bad argument (null pointer, negative size, ...)                      */
#define SCHW_E_CUDA       -2   /* CUDA runtime error or no usable device                               */
#define SCHW_E_HORIZON    -3   /* observer inside the horizon: reference raises ValueError   (simulation/utils.py:67-68)  */
#define SCHW_E_NULLCOND   -4   /* negative discriminant:        reference raises RuntimeError (simulation/utils.py:80-81)  */
#define SCHW_E_RAYPLANE   -5   /* ray not in the equatorial plane after rotation: AssertionError (simulation/utils.py:113,126) */
/* The last two are detected per ray ON THE DEVICE (a sticky error word, read back before a host-pointer entry point
 * returns; the outputs are then undefined, as after the reference's exception).  Enqueue-only `_dev` entry points
 * cannot report them at call time: schw_last_status() does, after the work has finished.  The same word carries a
 * texture upload that never arrived (see schw_set_tex_ready_flag), reported as SCHW_E_CUDA. */
#define SCHW_E_NOMEM      -6

/* ray classes, in the order of the reference's `collision` strings (simulation/raytracing.py:210-272) */
#define SCHW_CLS_BH               0   /* 'bh'               */
#define SCHW_CLS_NUMERICAL_ERROR  1   /* 'numerical error'  */
#define SCHW_CLS_ESCAPE_BG        2   /* 'escape_bg'        */
#define SCHW_CLS_ESCAPE_NO_PATCH  3   /* 'escape_no_patch'  */
#define SCHW_CLS_IN_DOMAIN        4   /* 'in_domain'        */

typedef struct schw_handle schw_handle;

/* Scene = the arguments of run_manual_simulation (simulation/raytracing.py:16-24) plus the
 * Observer / BlackHole fields it reads (simulation/blackhole.py:4-25). */
typedef struct schw_scene {
    double  obs[3];              /* observer.position; the reference's camera axes are hard-wired to an
                                    observer on the +x axis (raytracing.py:60-72)                          */
    double  fov;                 /* observer.fov, radians                                                  */
    int32_t height, width;       /* observer.image_size = (h, w)                                           */
    double  mass;                /* bh.mass (rs = 2*mass, blackhole.py:13)                                 */
    double  r_max;               /* boundary_radius: integrator r_max (raytracing.py:118) and escape test (:216) */
    int64_t steps;               /* step cap (cuda_geodesic.py:596)                                        */
    double  delta;               /* affine step                                                            */
    double  omega;               /* FANTASY coupling actually used; run_manual_simulation always uses 1.0
                                    (raytracing.py:118 drops its omega argument, cuda_geodesic.py:632)     */
    double  patch_center_theta, patch_center_phi, patch_size_theta, patch_size_phi;   /* radians          */
    int32_t flip_theta, flip_phi;
    int32_t has_texture;         /* 0 <=> background_path is None (raytracing.py:217,265-268)              */
    int32_t _pad;
} schw_scene;

/* Optional per-ray outputs of schw_render*; any member may be NULL.  All arrays are indexed by the
 * packed ray number k = (visited_row_index * width + j).  */
typedef struct schw_ray_outputs {
    double  *final_q;     /* (n,4)  q1 = (t,r,theta,phi) at termination            (raytracing.py:181-183)   */
    double  *ic;          /* (n,9)  p0_t,p0_r,p0_th,p0_ph, alpha0, beta, h_r,h_theta,h_phi (utils.py:141)     */
    double  *hit;         /* (n,2)  th_hit, ph_hit as written to photon_data.csv    (raytracing.py:276)       */
    int32_t *texel;       /* (n,2)  (u,v) texture indices, (-1,-1) when no lookup   (raytracing.py:241-245)   */
    int32_t *nsteps;      /* (n,)   integrator steps taken                                                     */
} schw_ray_outputs;

/* ---- lifetime ------------------------------------------------------------------------------- */
int  schw_abi_version(void);
const char *schw_last_error(void);
/* device < 0 selects the current device.  Replaces the implicit Numba context creation. */
int  schw_create(int device, schw_handle **out);
int  schw_destroy(schw_handle *h);
/* name (<=255 chars), SM count, SM clock in kHz of the handle's device */
int  schw_device_info(schw_handle *h, char *name, int name_cap, int *sm_count, int *clock_khz);
/* CUDA-event duration (ms) of the kernels enqueued by the last compute call on this handle,
 * measured on the stream they were launched on; and how many kernels that call launched. */
int  schw_last_kernel_ms(schw_handle *h, float *ms, int *launches);

/* Page-locked host memory for the caller's frame buffers (optional): host-pointer entry points copy at full PCIe rate
 * into / out of such buffers, while ordinary pageable memory is staged by the driver at a fraction of it.  The reference
 * has no counterpart (Numba's implicit transfers at cuda_geodesic.py:644,652 use pageable NumPy arrays). */
int  schw_host_alloc(size_t bytes, void **out);
int  schw_host_free(void *p);

/* ---- stage 3+4: the integrator --------------------------------------------------------------
 * Replaces CUDASchwarzschildIntegrator.integrate_batch (simulation/cuda_geodesic.py:639-645,
 * kernel :579-602).  q0s,p0s: (n,4) f64 C-contiguous.  out_qs: (n,4).  out_steps: (n,) or NULL.
 * rs = 2*mass as the class computes it (:635). */
int  schw_integrate_batch(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                          int64_t steps, double delta, double rs, double r_max, double omega,
                          double *out_qs, int32_t *out_steps);
int  schw_integrate_batch_dev(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n,
                              int64_t steps, double delta, double rs, double r_max, double omega,
                              double *d_out_qs, int32_t *d_out_steps, void *cuda_stream);

/* Replaces CUDASchwarzschildIntegrator.integrate_batch_full (cuda_geodesic.py:647-653, kernel
 * :604-625).  out_traj: (n,steps,4) f64; rows after a ray's termination are zero (np.zeros, :649).
 * out_rows: (n,) number of stored rows per ray, or NULL.  The host variant requires out_traj to be
 * ZERO-FILLED by the caller (np.zeros) and only copies the stored rows back. */
int  schw_integrate_batch_full(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                               int64_t steps, double delta, double rs, double r_max, double omega,
                               double *out_traj, int64_t *out_rows);

/* Device-resident variant for trajectory dumps at scale: (n,steps,4) f64 in HBM, 32 B per step per ray,
 * written as full 128-byte lines.  zero_fill != 0 clears the buffer first (the reference's np.zeros, :649); otherwise rows
 * after a ray's last stored row are left untouched.  d_out_rows: (n,) i64 or NULL.  Enqueue only. */
int  schw_integrate_batch_full_dev(schw_handle *h, const double *d_q0s, const double *d_p0s, int64_t n,
                                   int64_t steps, double delta, double rs, double r_max, double omega,
                                   double *d_out_traj, int64_t *d_out_rows, int zero_fill, void *cuda_stream);

/* Decimated variant: same integration, but only the rows whose step index is listed in the strictly
 * ascending row_idx[0..nidx) are produced, out: (n,nidx,4) f64 (zero where the ray ended earlier).  With
 * row_idx = np.linspace(0, steps-1, min(1000, steps), dtype=int32) this is exactly what the reference keeps of
 * integrate_batch_full's (n,steps,4) result (simulation/raytracing.py:139-150), without the 6.4 MB/ray buffer. */
int  schw_integrate_batch_rows(schw_handle *h, const double *q0s, const double *p0s, int64_t n,
                               int64_t steps, double delta, double rs, double r_max, double omega,
                               const int32_t *row_idx, int64_t nidx, double *out, int64_t *out_rows);

/* ---- stages 1+2: camera rays -> initial conditions -------------------------------------------
 * Replaces the pixel loop of run_manual_simulation (simulation/raytracing.py:60-88,98-116) and
 * get_initial_conditions / angles_to_p_sph / build_null_4momentum_ep_sph (simulation/utils.py:
 * 91-141,148-194,39-88) for rows row_begin, row_begin+row_stride, ... < row_end.
 * q0s,(n,4); ic (n,9) as in schw_ray_outputs.  */
int  schw_initial_conditions(schw_handle *h, const schw_scene *scene, int row_begin, int row_end,
                             int row_stride, double *q0s, double *ic);

/* q0 = (0, r_obs, theta_obs, phi_obs), the initial position shared by every camera ray of the scene (simulation/utils.py:
 * 118,140: cartesian_to_spherical_fast of the observer); host arithmetic, needs no device. */
int  schw_observer_q0(const schw_scene *scene, double *q0);

/* Same for caller-supplied pixel positions: get_initial_conditions(observer_pos, pixel_pos, mass_bh=)
 * (simulation/utils.py:91-141) as called per ray by the reference's scripts.  obs: (3,), pixels: (n,3). */
int  schw_initial_conditions_px(schw_handle *h, const double *obs, const double *pixels, int64_t n,
                                double mass, double *q0s, double *ic);

/* ---- stage 5: rotate back, classify, texel lookup --------------------------------------------
 * Replaces simulation/raytracing.py:166-174,181-273 on caller-supplied final states.
 * tex: (height,width,3) u8 (the sky already resized to the frame, raytracing.py:36-37) or NULL. */
int  schw_classify(schw_handle *h, const schw_scene *scene, const double *final_q,
                   const double *alpha0, const double *beta, int64_t n, const uint8_t *tex,
                   uint8_t *rgb, uint8_t *cls, double *hit, int32_t *texel);

/* ---- the fused path ---------------------------------------------------------------------------
 * Replaces run_manual_simulation's image production (simulation/raytracing.py:50-120,162-273):
 * camera rays -> initial conditions -> FANTASY stepper with capture/escape/max-step termination
 * -> rotate back -> classify -> equirectangular texel lookup, one launch, for rows
 * row_begin, row_begin+row_stride, ... < row_end.  rgb: (nrows,width,3) u8; cls: (nrows,width) u8. */
int  schw_render(schw_handle *h, const schw_scene *scene, const uint8_t *tex,
                 int row_begin, int row_end, int row_stride,
                 uint8_t *rgb, uint8_t *cls, const schw_ray_outputs *extra);
/* Device-resident variant: d_tex is the (height,width,3) u8 texture already in HBM; outputs are
 * device pointers; work is enqueued on cuda_stream (a cudaStream_t; NULL = the handle's stream). */
int  schw_render_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex,
                     int row_begin, int row_end, int row_stride,
                     uint8_t *d_rgb, uint8_t *d_cls, const schw_ray_outputs *d_extra,
                     void *cuda_stream);
/* Same with explicit row pitches (bytes between consecutive visited rows; 0 = packed: 3*width and width), so the kernel can
 * write straight into a caller-defined block -- e.g. rgb and class side by side in the (rows, 4*width) u8 message of the
 * multi-GPU gather: d_cls = d_rgb + 3*width, both pitches 4*width. */
int  schw_render_dev_pitched(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex,
                             int row_begin, int row_end, int row_stride,
                             uint8_t *d_rgb, int64_t rgb_row_pitch, uint8_t *d_cls, int64_t cls_row_pitch,
                             const schw_ray_outputs *d_extra, void *cuda_stream);
/* Waits for the last call on this handle and returns SCHW_OK or the error its device code flagged (SCHW_E_RAYPLANE,
 * SCHW_E_NULLCOND, SCHW_E_CUDA for a texture that never arrived).  Host-pointer entry points check this themselves. */
int  schw_last_status(schw_handle *h);
/* ---- next rows (opt-in): geodesic de-duplication -------------------------------
 * Same arguments, same outputs -- bit for bit -- as schw_render / schw_render_dev.  Every camera ray of the reference
 * starts at the same q0 with a 4-momentum that depends only on its angle alpha to the optical axis (simulation/utils.py:
 * 148-194, 39-88), so pixels whose alpha agrees bitwise (mirror images and more: 8.3 M pixels of a 4K frame hold 0.95 M
 * distinct values) own the same geodesic; each distinct one is integrated once and every pixel applies its own rotation
 * by beta, classification and texel lookup (simulation/raytracing.py:181-273) to the shared final state.  The reference
 * integrates every pixel; this entry point is therefore opt-in and reports its own counters: schw_last_total_steps =
 * steps actually integrated, schw_last_dedup_stats = distinct rays and the steps summed over pixels.
 * The _dev variant synchronises its stream once (the distinct-ray count sizes the second half of the pipeline). */
int  schw_render_dedup(schw_handle *h, const schw_scene *scene, const uint8_t *tex,
                       int row_begin, int row_end, int row_stride,
                       uint8_t *rgb, uint8_t *cls, const schw_ray_outputs *extra);
int  schw_render_dedup_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex,
                           int row_begin, int row_end, int row_stride,
                           uint8_t *d_rgb, uint8_t *d_cls, const schw_ray_outputs *d_extra,
                           void *cuda_stream);
int  schw_last_dedup_stats(schw_handle *h, uint64_t *unique_rays, uint64_t *pixel_steps);
/* The same pipeline stage by stage, for the ranks of a multi-GPU job that SHARE the integration of the distinct rays
 * (schw_b200/multigpu.py; the reference integrates every pixel in one process, simulation/raytracing.py:98-120):
 *   plan       every rank: keys of the WHOLE frame -> sort -> distinct rays (identical on all ranks; synchronises the
 *              stream once).  Distinct ray u (in order of camera angle) gets slot (u % world)*chunk + u / world,
 *              chunk = ceil(n_unique / world): rank r owns slots [r*chunk, r*chunk + count_r) = every world-th ray.
 *   integrate  rank r integrates its slots into d_fq_all (world*chunk, 4) f64 / d_ns_all (world*chunk) i32.
 *              -- the caller all-gathers the two arrays (chunk rows per rank) --
 *   epilogue   every rank: rotate back, classify, texel lookup for its own rows from the gathered final states;
 *              outputs and pitches as schw_render_dev_pitched. */
int  schw_dedup_plan_dev(schw_handle *h, const schw_scene *scene, int world, uint64_t *n_unique, uint64_t *chunk,
                         void *cuda_stream);
int  schw_dedup_integrate_dev(schw_handle *h, const schw_scene *scene, int rank, double *d_fq_all, int32_t *d_ns_all,
                              void *cuda_stream);
int  schw_dedup_epilogue_dev(schw_handle *h, const schw_scene *scene, const uint8_t *d_tex,
                             int row_begin, int row_end, int row_stride,
                             uint8_t *d_rgb, int64_t rgb_row_pitch, uint8_t *d_cls, int64_t cls_row_pitch,
                             const schw_ray_outputs *d_extra, const double *d_fq_all, const int32_t *d_ns_all,
                             void *cuda_stream);

/* Overlap of the texture upload with the start of a frame (optional, one-shot): the next schw_render_dev call on this
 * handle may be enqueued BEFORE its texture has arrived in d_tex; the kernel waits -- per ray, only at the texel lookup
 * that ends it -- until the 32-bit word at d_flag is non-zero.  The caller clears the word, uploads the texture on
 * another stream and sets the word with a copy ordered after that upload.  schw_render does this internally. */
int  schw_set_tex_ready_flag(schw_handle *h, const uint32_t *d_flag);
/* Sum of integrator steps over the rays of the last schw_render* / schw_integrate_batch* call on
 * this handle (device-side counter; used for the steps/s metric). */
int  schw_last_total_steps(schw_handle *h, uint64_t *total_steps);
/* Loop statistics of the same call (any pointer may be NULL): total_steps as above; warp_iters = step-loop bodies
 * executed summed over warps (total_steps / (32 * warp_iters) = fraction of issued lane-steps that advanced a ray,
 * the rest being lanes that idle between a ray's end and the next refill); planar_iters = how many of those bodies
 * ran entirely on the equatorial fast path (csrc/schw_device.cuh "Planar steps").  Zero for the trajectory kernels. */
int  schw_last_counters(schw_handle *h, uint64_t *total_steps, uint64_t *warp_iters, uint64_t *planar_iters);

/* ---- next rows: flat-space comparison render ----------------------------------
 * Replaces the pixel loop of save_no_gravity_image_with_background (simulation/background.py:
 * 99-132): ray-sphere intersection at boundary radius r_max, truncating texel map.
 * Uses scene->obs, fov, height, width, r_max, patch_*, flip_*.  rgb: (height,width,3) u8. */
int  schw_flat_render(schw_handle *h, const schw_scene *scene, const uint8_t *tex, uint8_t *rgb);
/* Replaces cuda_flat_raytrace and cuda_flat_raytrace_with_trajectories (simulation/cuda_geodesic.py:214-292, kernels
 * :18-208; imported by simulation/background.py:77,95): caller-supplied unit ray directions (n,3) from obs (3,),
 * bg (tex_h,tex_w,3) u8, out_rgb (n,3) u8.  With n_sampled > 0 also the n_points equally spaced points from the observer
 * to the hit point of every ray listed in sampled_indices (flat ray numbers): out_trajs (n_sampled,n_points,3) f64, zero
 * rows for rays that miss the boundary sphere. */
int  schw_flat_raytrace(schw_handle *h, const double *obs, const double *ray_dirs, int64_t n, double boundary_radius,
                        double patch_center_theta, double patch_center_phi, double patch_size_theta,
                        double patch_size_phi, int flip_theta, int flip_phi, const uint8_t *bg, int tex_h, int tex_w,
                        uint8_t *out_rgb, const int32_t *sampled_indices, int64_t n_sampled, int64_t n_points,
                        double *out_trajs);

/* ---- measurement helper ------------------------------------------------------------------------
 * Register-resident DFMA chains on every SM: measured FP64 FMA peak of the device in TFLOP/s
 * (FMA = 2 flops), the denominator of the FP64 roofline (not in MEASURED_PEAKS.json). */
int  schw_fp64_peak_probe(schw_handle *h, double *tflops, float *ms);

/* Diagnostic: evaluates the kernels' own reciprocal and sincos primitives (csrc/schw_device.cuh:
 * fast_rcp, sincos_flow — the replacements for the reference's `/` and np.sin/np.cos inside
 * simulation/cuda_geodesic.py:433-474) on n host values, for accuracy tests. */
int  schw_math_probe(schw_handle *h, const double *x, int64_t n, double *rcp, double *sn, double *cs);

/* Diagnostic / measurement: the loop body of the render kernel alone (no queue, refill or epilogue) on camera-like rays,
 * variant = 100 * rays per thread + resident CTAs per SM requested (106, 105, 104, 204, 203, 202): kernel time and
 * warp-steps executed, from which SM cycles per warp-step follow. */
int  schw_step_probe(schw_handle *h, int variant, int iters, double delta, float *ms, uint64_t *warp_steps,
                     int *ctas_per_sm, int *regs, double *planar_fraction);

/* Diagnostic, needs no device: the step parameters exactly as the kernels receive them (csrc/schw_device.cuh:
 * StepParams, incl. the thresholds of the planar fast path), for the host-compiled model of the step loop in
 * tests/model/.  out may be NULL to query the size. */
int  schw_debug_step_params(int64_t steps, double delta, double rs, double r_max, double omega,
                            void *out, size_t cap, size_t *size);

#ifdef __cplusplus
}
#endif
#endif /* SCHW_B200_H */
