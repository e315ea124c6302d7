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

#define VP_VERSION 100 /* 0.1.0 */

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
} vp_params;

/* Perspective sensor, Mitsuba `perspective` plugin semantics (volprim/cameras.py:114-137). */
typedef struct vp_camera {
    float to_world[12]; /* rows of the 3x4 camera-to-world matrix (Mitsuba look_at convention) */
    float fov_x_deg;    /* `fov` with fov_axis = 'x'                                           */
    float near_clip, far_clip;
    float cx, cy;       /* principal_point_offset_{x,y}                                        */
    int32_t width, height;
} vp_camera;

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
 *                out_hit_ids[r * id_ray_stride + k * id_hit_stride], k < id_cap, -1 padded. */
VP_API int vp_trace_forward(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, float *out_rgb, float *out_T, uint32_t *out_nhits, int32_t *out_hit_ids,
                     int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride, void *stream);

/* Integrator.sample(mode=Backward, ...): PRB adjoint (volprim_rf.py:106-165,
 * volprim_tomography.py:57-101).  d_L [R*3] = delta L, state_in [R*3] = the primal call's state_out.
 * Gradients are ADDED into the caller-zeroed reference-layout buffers g_data10 [N*10],
 * g_attr [N], g_sh [N*C] (NULL ok for tomography).
 * hit_ids: NULL -> the hit sequence is re-traced exactly like the primal; otherwise the lists a
 * previous vp_trace_forward recorded for the same rays (same strides), which are replayed without
 * touching the BVH (hit_counts [R] required then). */
VP_API int vp_trace_adjoint(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, const float *d_L, const float *state_in, const int32_t *hit_ids,
                     const uint32_t *hit_counts, int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride,
                     float *g_data10, float *g_attr, float *g_sh, void *stream);

/* Sensor.sample_ray for a `perspective` sensor (Mitsuba plugin; parameters as produced by
 * CameraSpecs.to_dict, volprim/cameras.py:114-137).  Writes W*H*spp rays, pixel-major then sample.
 * `jitter` NULL -> pixel centres; else [W*H*spp*2] sub-pixel offsets in [0,1). */
VP_API int vp_raygen_perspective(vp_ctx *ctx, const vp_camera *cam, int32_t spp, const float *jitter, float *ray_o,
                          float *ray_d, float *ray_maxt, void *stream);

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

/* Introspection for tests: copies the BVH node array (16 floats per internal node, layout in
 * csrc/vp_build.cu) and the sorted->original index into caller device buffers (either may be NULL);
 * *n_internal receives the internal-node count (N-1). */
VP_API int vp_debug_bvh(vp_ctx *ctx, float *out_nodes, int32_t *out_perm, int64_t *n_internal, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* VOLPRIM_CUDA_H */
