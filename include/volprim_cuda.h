/*
 * volprim_cuda.h -- C ABI of libvolprim_cuda.so, the B200 (sm_100a) implementation of volprim's
 * per-ray volumetric-primitive integration (the loop shared by the `volprim_rf` and
 * `volprim_tomography` Mitsuba integrators of gitmon/volprim-balance).
 *
 * The reference has no FFI: its boundary is Mitsuba's Python integrator-plugin API.  Each entry
 * point below names the reference interface it stands in for (paths relative to the reference
 * repository root).  The Python host side (volprim_balance_b200/) binds these with ctypes; see
 * INTEGRATION.md for the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative VP_E_* code otherwise; the text of the last
 *     error of a context is available from vp_last_error().  No C++ exception crosses the ABI.
 *   - all array arguments are DEVICE pointers owned by the caller (torch allocations on the
 *     context's device) unless named `host_*`; NULL is allowed where stated.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - a vp_ctx is not thread-safe; distinct contexts (one per GPU / process) are independent.
 *   - layouts are the reference's: `data10` [N*10] = center3, scale3, quaternion (i,j,k,r)
 *     (volprim/integrators/common.py:55-74), `attr` [N] = `opacities` (rf) or `sigma_t` (tomography),
 *     `sh` [N*C] coefficient-major / channel-minor with C = 3 (D+1)^2 (volprim_rf.py:88-95).
 */
#ifndef VOLPRIM_CUDA_H
#define VOLPRIM_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VP_VERSION 200 /* 0.2.0 */

#if defined(__GNUC__)
#define VP_API __attribute__((visibility("default")))
#else
#define VP_API
#endif

enum { VP_OK = 0, VP_E_INVALID = -1, VP_E_CUDA = -2, VP_E_STATE = -3, VP_E_OOM = -4 };

enum { VP_INTEGRATOR_RF = 0, VP_INTEGRATOR_TOMO = 1 };   /* plugin names volprim_rf / volprim_tomography */
enum { VP_KERNEL_GAUSSIAN = 0, VP_KERNEL_EPANECHNIKOV = 1 }; /* Kernel.factory, common.py:96-105 */

typedef struct vp_ctx vp_ctx;

/* Integrator state: the plugin parameters of volprim_rf.py:23-46 / volprim_tomography.py:24-35. */
typedef struct vp_params {
    int32_t integrator;      /* VP_INTEGRATOR_*                                              */
    int32_t kernel;          /* VP_KERNEL_*            `kernel_type`                         */
    uint32_t max_depth;      /* `max_depth`; 0xFFFFFFFF = unlimited (-1 in the plugin)       */
    int32_t srgb_primitives; /* `srgb_primitives` (rf only, default 1)   volprim_rf.py:41,189 */
    int32_t hide_emitters;   /* `hide_emitters` (tomography)     volprim_tomography.py:106   */
    float t_cutoff;          /* 0.01: transmittance cut-off               volprim_rf.py:173  */
    float eps_advance;       /* 1e-4: origin advance after a hit          volprim_rf.py:149  */
    float env[3];            /* constant environment radiance    volprim_tomography.py:107   */
    int32_t image_width;     /* >0: rays are W x H pixel grids; threads walk them in 8x4 tiles */
    int32_t image_height;
    int32_t use_rr;          /* Russian roulette active (rr_depth >= 0 and (rr_depth < max_depth or max_depth == -1),
                                volprim_rf.py:39,177-183); primal pass only, like the reference                 */
    uint32_t rr_depth;       /* `rr_depth`                                                                      */
    uint32_t rr_seed;        /* seed of the `independent` sampler (PCG32 per ray, stream = TEA(seed, ray index)) */
    uint32_t rr_skip;        /* 1-D samples every ray has drawn before sample() is entered (2 under mi.render)  */
} vp_params;

/* Perspective sensor, Mitsuba `perspective` plugin semantics (volprim/cameras.py:114-137). */
typedef struct vp_camera {
    float to_world[12]; /* rows of the 3x4 camera-to-world matrix (Mitsuba look_at convention) */
    float fov_x_deg;    /* `fov` with fov_axis = 'x'                                           */
    float near_clip, far_clip;
    float cx, cy;       /* principal_point_offset_{x,y}                                        */
    int32_t width, height;
} vp_camera;

/* Where the rays of a vp_render_* call come from: an explicit batch (the calling convention of
 * scripts/radiosity/radiance_cache.py:252-266), or a perspective sensor whose rays are generated INSIDE the trace
 * kernels (Sensor.sample_ray fused into the integrator launch: no ray buffers in HBM, 28 B per ray less traffic). */
typedef struct vp_ray_source {
    const float *ray_o;      /* device [R*3], or NULL when `camera` is set                                   */
    const float *ray_d;      /* device [R*3], or NULL when `camera` is set                                   */
    const float *ray_maxt;   /* device [R] or NULL (= infinity); unused with `camera` (far - near along the ray) */
    const vp_camera *camera; /* HOST pointer or NULL.  Rays are pixel-major, then sample (like vp_raygen_perspective) */
    const float *jitter;     /* device [R*2] sub-pixel offsets in [0,1) or NULL = pixel centres (camera only) */
    int32_t spp;             /* samples per pixel (camera only)                                              */
    int32_t row_begin;       /* first film row of this call; R = width * row_count * spp (camera only)        */
    int32_t row_count;       /* 0 = all rows from row_begin to the bottom of the film                         */
    int32_t reserved;
} vp_ray_source;

/* Ordered hit lists of a primal pass in compressed-row form (what the adjoint replays instead of walking the BVH a
 * second time).  All arrays are caller-owned DEVICE memory.  Bytes kept per view: 4 (20 with `state`) per recorded
 * hit + 8 per ray; or a dense block (see `dense`). */
typedef struct vp_hit_record {
    int64_t *ray_offsets;   /* rows:  [n_rays + 1]   list of ray r = ids[ray_offsets[r] .. ray_offsets[r + 1])      */
    int32_t *ids;           /* rows:  [capacity] primitive ids (numbering of vp_set_primitives), front to back
                               dense: [id_cap * n_rays], hit k of ray r at ids[k * n_rays + r]                      */
    float *state;           /* rows:  [capacity * 4] or NULL; dense: [id_cap * n_rays * 4], same indexing:
                               (colour r, g, b, transmittance) of every recorded hit (volprim_rf).  With it the
                               adjoint's ray-major pass is the PRB recurrence alone -- no primitive is loaded or
                               shaded a second time; 20 instead of 4 bytes per hit                                  */
    uint32_t *counts;       /* dense: [n_rays] recorded hits per ray (written by vp_render_forward); rows: unused   */
    int64_t *total;         /* [2]            entries all lists hold; rays whose list was cut at `id_cap`          */
    int64_t capacity;       /* rows: entries `ids` holds; dense: the bound on the entries the gather adjoint sizes its
                               buckets with (< 2^32).  A record is usable iff total[0] <= capacity and total[1] == 0;
                               vp_render_adjoint does nothing otherwise (the caller re-traces)                     */
    int32_t id_cap;         /* most hits recorded per ray                                                          */
    int32_t dense;          /* 0: compressed rows (4 or 20 bytes per hit, compacted from a transient band scratch);
                               1: dense hit-major block in the caller's buffers with per-hit state -- no compaction
                               pass, coalesced replay: 20 B * id_cap * n_rays of memory for the fastest step        */
} vp_hit_record;

/* Reconstruction filters of the film (Mitsuba rfilter plugins `box`, `tent`, `gaussian`; volprim/cameras.py:117). */
enum { VP_RFILTER_BOX = 0, VP_RFILTER_TENT = 1, VP_RFILTER_GAUSSIAN = 2 };

/* Work counters of the last trace call (device-accumulated, read back on request). */
typedef struct vp_stats {
    uint64_t rays;
    uint64_t hits;        /* primitive evaluations (accepted hits)             */
    uint64_t candidates;  /* exact ray/ellipsoid tests during traversal        */
    uint64_t node_visits; /* BVH internal nodes fetched                        */
    uint64_t passes;      /* intervals walked (hit-list refills), summed over rays */
    uint64_t stack_overflows;  /* traversal-stack overflows of the per-ray walker: must be 0, a subtree was skipped */
    uint64_t interval_retries; /* intervals walked again with a shorter width because a list did not fit (per ray;
                                  performance statistic only) */
} vp_stats;

VP_API int vp_version(void);

/* Per-GPU context: owns the LBVH, the Morton-sorted SoA copy of the primitives and scratch.
 * Stands in for the scene-side acceleration structure Mitsuba builds in mi.load_dict()
 * (examples/render_3dg_asset.py:59-66). */
VP_API int vp_create(int device, vp_ctx **out);
VP_API int vp_destroy(vp_ctx *ctx);
VP_API const char *vp_last_error(const vp_ctx *ctx); /* ctx may be NULL: error of the last failed vp_create */

/* Upload the primitive set in the reference layouts (replaces params['primitives.data'] = ...,
 * params['primitives.opacities'|'sigma_t'], params['primitives.sh_coeffs'];
 * examples/refine_3dg_dataset.py:131-159).  `sh` may be NULL (tomography). */
VP_API int vp_set_primitives(vp_ctx *ctx, int64_t n, const float *data10, const float *attr, const float *sh,
                      int32_t sh_floats, float extent, void *stream);

/* params.update() -> acceleration-structure (re)build (examples/refine_3dg_dataset.py:159).
 * vp_build: Morton codes -> radix sort -> Karras hierarchy -> bottom-up AABB fit.
 * vp_refit: keep order and topology of the last build, refresh SoA and boxes (same N required). */
VP_API int vp_build(vp_ctx *ctx, void *stream);
VP_API int vp_refit(vp_ctx *ctx, void *stream);

/* Integrator.sample(mode=Primal, ...) for a batch of explicit rays
 * (volprim_rf.py:103-192, volprim_tomography.py:47-127; call shape of
 * scripts/radiosity/radiance_cache.py:252-266).
 *   ray_o, ray_d [R*3]; ray_maxt [R] or NULL (= infinity)
 *   out_rgb [R*3]; out_T [R] final throughput beta (NULL ok); out_nhits [R] (NULL ok)
 *   out_hit_ids: NULL, or the ordered primitive-ID list of every ray, element (ray r, hit k) at
 *                out_hit_ids[r * id_ray_stride + k * id_hit_stride], k < min(out_nhits[r], id_cap).  Entries beyond
 *                a ray's hit count are NOT written (pre-fill the buffer if padding is wanted). */
VP_API int vp_trace_forward(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, float *out_rgb, float *out_T, uint32_t *out_nhits, int32_t *out_hit_ids,
                     int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride, void *stream);

/* Integrator.sample(mode=Backward, ...): PRB adjoint (volprim_rf.py:106-165,
 * volprim_tomography.py:57-101).  d_L [R*3] = delta L, state_in [R*3] = the primal call's state_out.
 * Gradients are ADDED into the caller-zeroed reference-layout buffers g_data10 [N*10],
 * g_attr [N], g_sh [N*C] (NULL ok for tomography).
 * hit_ids: NULL -> the hit sequence is re-traced exactly like the primal; otherwise the lists a
 * previous vp_trace_forward recorded for the same rays (same strides), which are replayed without
 * touching the BVH (hit_counts [R] required then; a ray with hit_counts[r] > id_cap replays only its first id_cap
 * hits -- record with id_cap >= max_depth, or check the counts).
 * Alignment: g_data10 8 bytes, g_sh 16 bytes when C % 4 == 0 (vector reductions); VP_E_INVALID otherwise. */
VP_API int vp_trace_adjoint(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, const float *d_L, const float *state_in, const int32_t *hit_ids,
                     const uint32_t *hit_counts, int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride,
                     float *g_data10, float *g_attr, float *g_sh, void *stream);

/* Sensor.sample_ray for a `perspective` sensor (Mitsuba plugin; parameters as produced by
 * CameraSpecs.to_dict, volprim/cameras.py:114-137).  Writes W*H*spp rays, pixel-major then sample.
 * `jitter` NULL -> pixel centres; else [W*H*spp*2] sub-pixel offsets in [0,1). */
VP_API int vp_raygen_perspective(vp_ctx *ctx, const vp_camera *cam, int32_t spp, const float *jitter, float *ray_o,
                          float *ray_d, float *ray_maxt, void *stream);

/* ---- sensor-fused / record-replay entry points (what volprim_balance_b200.render() uses) ------------------------
 * vp_render_forward: Integrator.sample(Primal) over `rays` (explicit batch, or rays generated in-kernel from a
 * perspective sensor).  out_T / out_nhits may be NULL.  With `record` != NULL the ordered hit lists are kept in
 * compressed-row form: the trace kernel writes them hit-major into a transient scratch of the context (bounded by
 * splitting the call into row bands), a compaction pass copies them to record->ids at the exclusive-scan offsets of
 * the hit counts.  (The trace kernel itself contains no atomics: see csrc/vp_trace.cu.) */
VP_API int vp_render_forward(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays,
                             float *out_rgb, float *out_T, uint32_t *out_nhits, const vp_hit_record *record,
                             void *stream);

/* vp_render_adjoint: Integrator.sample(Backward) replaying `record` (volprim_rf.py:106-165).  volprim_rf uses the
 * GATHER formulation: a flat counting pass over the record sizes one bucket per primitive, a ray-major pass replays
 * every list and writes a 32-byte entry per hit (one full-sector store) into the bucket of the hit primitive, then one warp per primitive (and
 * per further 256 entries of a big bucket) accumulates its bucket in registers and adds the 10 + 1 + C gradient floats
 * to the caller's buffers -- no per-hit global reductions (the scatter formulation of vp_trace_adjoint saturates the L2
 * reduction units).  volprim_tomography replays with vector reductions like vp_trace_adjoint.
 * Gradients are ADDED to g_data10 [N*10], g_attr [N], g_sh [N*C] (8-, 4- and 8-byte aligned); the buffers must not
 * be written by anything else while the call runs.  Equivalent to vp_adjoint_begin + vp_adjoint_finish(0, N). */
VP_API int vp_render_adjoint(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays,
                             const float *d_L, const float *state_in, const vp_hit_record *record, float *g_data10,
                             float *g_attr, float *g_sh, void *stream);

/* The two halves of vp_render_adjoint, so that a caller can overlap the gradient all-reduce of one primitive range
 * with the accumulation of the next (examples/refine_3dg_dataset.py on several GPUs): vp_adjoint_begin runs the
 * ray-major pass into the context's scratch; vp_adjoint_finish accumulates primitives [prim_begin, prim_end). */
VP_API int vp_adjoint_begin(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays,
                            const float *d_L, const float *state_in, const vp_hit_record *record, float *g_data10,
                            float *g_attr, float *g_sh, void *stream);
VP_API int vp_adjoint_finish(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays,
                             const vp_hit_record *record, int64_t prim_begin, int64_t prim_end, float *g_data10,
                             float *g_attr, float *g_sh, void *stream);

/* Film (Mitsuba hdrfilm + reconstruction filter; batch film layout of examples/refine_3dg_dataset.py:96-107).
 * Samples are pixel-major then sample, at pixel + jitter (NULL = centres).  vp_film_splat ADDS weight * radiance
 * and the weights into accum [H*W*4] (caller-zeroed); vp_film_develop writes image [H*W*3] = rgb / weight;
 * vp_film_adjoint gathers d_L [S*3] = sum over the pixels a sample touches of d_image * weight / pixel weight. */
VP_API int vp_film_splat(int32_t width, int32_t height, int32_t spp, int32_t rfilter, const float *jitter,
                         const float *radiance, float *accum, void *stream);
VP_API int vp_film_develop(int32_t width, int32_t height, const float *accum, float *image, int64_t image_row_stride,
                           void *stream);
VP_API int vp_film_adjoint(int32_t width, int32_t height, int32_t spp, int32_t rfilter, const float *jitter,
                           const float *accum, const float *d_image, int64_t d_image_row_stride, float *d_L,
                           void *stream);

/* Tunables of a context: "record_scratch_bytes" (transient dense hit-list scratch of vp_render_forward, default
 * 6 GiB: a 1080p view with per-hit state at a cap of 128 hits fits one band; every band is a kernel launch with its own
 * tail).  Returns VP_E_INVALID for an unknown name. */
VP_API int vp_set_option(vp_ctx *ctx, const char *name, int64_t value);

/* Copy the work counters of the most recent trace call to the host (synchronises `stream`). */
VP_API int vp_get_stats(vp_ctx *ctx, vp_stats *host_out, void *stream);

/* BoundedAdam.step for one parameter tensor, fused into one pass (volprim/optimizers.py:72-146; non-uniform,
 * unmasked variant -- the one the reference's examples use).  Host scalars are doubles like the reference's Python
 * floats (1 - beta is formed in double before rounding to fp32).  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t) is
 * computed by the caller.  param / m / v are updated in place; all pointers 16-byte aligned.  NaN gradients count
 * as zero; an element whose step would cross a bound moves half-way to it and has its moments reset. */
VP_API int vp_bounded_adam_step(int64_t n, float *param, const float *grad, float *m, float *v, double lr_t, double beta1,
                                double beta2, double eps, int has_lower, float lower, int has_upper, float upper,
                                void *stream);

/* l1(reference, image) (volprim/optimizers.py:170-174), the seed dr.backward(loss) hands the render op (d loss / d image
 * = sign(image - reference) / n_total) and the squared error of psnr() (:180-186) in one pass over n floats.  ADDS
 * sum |diff| / n_total to sums[0] and sum diff^2 / n_total to sums[1] (device, caller-zeroed); n_total = element count of
 * the whole batch film, so that the views of a batch can be processed one at a time. */
VP_API int vp_l1_loss_grad(int64_t n, const float *image, const float *reference, double n_total, float *d_image, float *sums,
                           void *stream);

/* Introspection for tests: copies the BVH node array (16 floats per internal node, layout in
 * csrc/vp_build.cu) and the sorted->original index into caller device buffers (either may be NULL);
 * *n_internal receives the internal-node count (N-1). */
VP_API int vp_debug_bvh(vp_ctx *ctx, float *out_nodes, int32_t *out_perm, int64_t *n_internal, void *stream);

/* Device self-test for tests.  which = 0: the correctly rounded shared-reciprocal division used by the exact
 * ray/ellipsoid intersection (the arithmetic that must equal the reference formulas of common.py:346-367 bit for bit
 * in fp32) against __fdiv_rn on n pseudo-random operand pairs; *mismatches receives the number of differing results.
 * Synchronous, current device. */
VP_API int vp_debug_selftest(int32_t which, int64_t n, uint64_t seed, int64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* VOLPRIM_CUDA_H */
