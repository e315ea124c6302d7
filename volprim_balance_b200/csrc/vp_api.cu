// vp_api.cu -- extern "C" surface of libvolprim_cuda.so (declared in include/volprim_cuda.h).
#include "vp_internal.cuh"

#include <cmath>
#include <cstring>
#include <new>

namespace {
std::string g_create_error;

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

void free_buf(DevBuffer &b)
{
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
}
}  // namespace

int vp_fail(vp_ctx *ctx, int code, const std::string &msg)
{
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
    return code;
}

int vp_ensure(vp_ctx *ctx, DevBuffer &b, size_t bytes)
{
    if (bytes == 0) bytes = 16;
    if (b.cap >= bytes) return VP_OK;
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr;
    b.cap = 0;
    size_t want = bytes + bytes / 8;  // slack so that slowly growing clouds do not reallocate every step
    cudaError_t e = cudaMalloc(&b.ptr, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        e = cudaMalloc(&b.ptr, bytes);
        want = bytes;
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        b.ptr = nullptr;
        return vp_fail(ctx, VP_E_OOM, std::string("cudaMalloc of ") + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
    }
    b.cap = want;
    return VP_OK;
}

DevScene vp_dev_scene(const vp_ctx *ctx)
{
    DevScene S;
    S.n = (int32_t)ctx->n;
    S.sh_floats = ctx->sh_floats;
    // sh_degree = int(sqrt(C // 3 - 1))   (reference volprim_rf.py:89, reproduced literally)
    S.sh_degree = ctx->sh_floats > 0 ? (int)std::sqrt((double)(ctx->sh_floats / 3 - 1)) : -1;
    if (ctx->sh_floats > 0 && 3 * (S.sh_degree + 1) * (S.sh_degree + 1) != ctx->sh_floats) S.sh_degree = 99;  // rejected by dispatch
    S.sh_stride4 = (ctx->sh_floats + 3) / 4;
    S.extent = ctx->extent;
    S.root = ctx->root;
    S.geo0 = (const float4 *)ctx->geo0.ptr;
    S.geo1 = (const float4 *)ctx->geo1.ptr;
    S.geo2 = (const float4 *)ctx->geo2.ptr;
    S.sh4 = (const float4 *)ctx->sh4.ptr;
    S.xf = (const float4 *)ctx->xf.ptr;
    S.info = (const float *)ctx->info.ptr;
    S.nodes = (const float4 *)ctx->nodes.ptr;
    S.perm = (const int32_t *)ctx->perm.ptr;
    S.inv_perm = (const int32_t *)ctx->inv_perm.ptr;
    return S;
}

// vp_debug_selftest(0): the shared-reciprocal division of exact_isect (vp_div_rn) against __fdiv_rn on pseudo-random
// operands of the magnitudes that occur there (numerators 2^-40 .. 2^20, denominators 2^-24 .. 2^10, both signs)
__global__ void k_selftest_div(int64_t n, uint64_t seed, unsigned long long *mismatches)
{
    unsigned long long bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint64_t z = seed + 0x9e3779b97f4a7c15ull * (uint64_t)(i + 1);
        z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
        z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
        z ^= z >> 31;
        const uint32_t lo = (uint32_t)z, hi = (uint32_t)(z >> 32);
        const uint32_t ea = 127u - 40u + (lo >> 23) % 61u, eb = 127u - 24u + (hi >> 23) % 35u;
        const float a = __uint_as_float((lo & 0x807fffffu) | (ea << 23));
        const float b = __uint_as_float((hi & 0x807fffffu) | (eb << 23));
        const float want = __fdiv_rn(a, b);
        const float got = vp_div_rn(a, vp_divisor(b));
        bad += __float_as_uint(want) != __float_as_uint(got);
    }
    if (bad) atomicAdd(mismatches, bad);
}

extern "C" {

int vp_version(void) { return VP_VERSION; }

int vp_create(int device, vp_ctx **out)
{
    if (!out) return vp_fail(nullptr, VP_E_INVALID, "vp_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return vp_fail(nullptr, VP_E_CUDA, std::string("vp_create: no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= count) return vp_fail(nullptr, VP_E_INVALID, "vp_create: device index out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return vp_fail(nullptr, VP_E_CUDA, std::string("vp_create: ") + cudaGetErrorString(e));
    if (prop.major != 10)
        return vp_fail(nullptr, VP_E_CUDA, std::string("vp_create: libvolprim_cuda.so is built for sm_100a only; device is sm_") +
                                               std::to_string(prop.major) + std::to_string(prop.minor));
    vp_ctx *ctx = new (std::nothrow) vp_ctx();
    if (!ctx) return vp_fail(nullptr, VP_E_OOM, "vp_create: out of host memory");
    ctx->device = device;
    *out = ctx;
    return VP_OK;
}

int vp_destroy(vp_ctx *ctx)
{
    if (!ctx) return VP_OK;
    DeviceGuard g(ctx->device);
    DevBuffer *all[] = { &ctx->raw_data, &ctx->raw_attr, &ctx->raw_sh, &ctx->geo0, &ctx->geo1, &ctx->geo2, &ctx->sh4, &ctx->xf, &ctx->info,
                         &ctx->nodes, &ctx->perm, &ctx->inv_perm, &ctx->leaf_lo, &ctx->leaf_hi, &ctx->keys[0],
                         &ctx->keys[1], &ctx->vals[0], &ctx->vals[1], &ctx->hist, &ctx->parent, &ctx->counters,
                         &ctx->bounds, &ctx->stats, &ctx->scan_tmp, &ctx->rec_dense, &ctx->rec_dense_state, &ctx->rec_counts, &ctx->rec_nhits,
                         &ctx->adj_offsets, &ctx->adj_rank, &ctx->adj_state, &ctx->adj_extra, &ctx->adj_items };
    for (DevBuffer *b : all) free_buf(*b);
    delete ctx;
    return VP_OK;
}

const char *vp_last_error(const vp_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int vp_set_primitives(vp_ctx *ctx, int64_t n, const float *data10, const float *attr, const float *sh,
                      int32_t sh_floats, float extent, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0) return vp_fail(ctx, VP_E_INVALID, "vp_set_primitives: negative primitive count");
    if (n > 0 && !data10) return vp_fail(ctx, VP_E_INVALID, "vp_set_primitives: data10 is NULL");
    if (sh && sh_floats <= 0) return vp_fail(ctx, VP_E_INVALID, "vp_set_primitives: sh given but sh_floats <= 0");
    if (!(extent > 0.f)) return vp_fail(ctx, VP_E_INVALID, "vp_set_primitives: extent must be > 0");
    if (!sh) sh_floats = 0;
    int rc;
    if ((rc = vp_ensure(ctx, ctx->raw_data, sizeof(float) * 10 * (size_t)n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->raw_attr, sizeof(float) * (size_t)n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->raw_sh, sizeof(float) * (size_t)n * (size_t)sh_floats))) return rc;
    if (n > 0) {
        VP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->raw_data.ptr, data10, sizeof(float) * 10 * (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (attr) VP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->raw_attr.ptr, attr, sizeof(float) * (size_t)n, cudaMemcpyDeviceToDevice, st));
        if (sh_floats)
            VP_CUDA_CHECK(ctx, cudaMemcpyAsync(ctx->raw_sh.ptr, sh, sizeof(float) * (size_t)n * sh_floats, cudaMemcpyDeviceToDevice, st));
    }
    ctx->have_attr = attr != nullptr;
    if (ctx->n != n || ctx->sh_floats != sh_floats) ctx->built = false;
    ctx->n = n;
    ctx->sh_floats = sh_floats;
    ctx->extent = extent;
    ctx->have_prims = true;
    return VP_OK;
}

int vp_build(vp_ctx *ctx, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_build_impl(ctx, false, (cudaStream_t)stream);
}

int vp_refit(vp_ctx *ctx, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_build_impl(ctx, true, (cudaStream_t)stream);
}

int vp_trace_forward(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, float *out_rgb, float *out_T, uint32_t *out_nhits, int32_t *out_hit_ids,
                     int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_trace_forward_impl(ctx, params, n_rays, ray_o, ray_d, ray_maxt, out_rgb, out_T, out_nhits, out_hit_ids,
                                 id_cap, id_ray_stride, id_hit_stride, (cudaStream_t)stream);
}

int vp_trace_adjoint(vp_ctx *ctx, const vp_params *params, int64_t n_rays, const float *ray_o, const float *ray_d,
                     const float *ray_maxt, const float *d_L, const float *state_in, const int32_t *hit_ids,
                     const uint32_t *hit_counts, int32_t id_cap, int64_t id_ray_stride, int64_t id_hit_stride,
                     float *g_data10, float *g_attr, float *g_sh, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_trace_adjoint_impl(ctx, params, n_rays, ray_o, ray_d, ray_maxt, d_L, state_in, hit_ids, hit_counts, id_cap,
                                 id_ray_stride, id_hit_stride, g_data10, g_attr, g_sh, (cudaStream_t)stream);
}

int vp_raygen_perspective(vp_ctx *ctx, const vp_camera *cam, int32_t spp, const float *jitter, float *ray_o,
                          float *ray_d, float *ray_maxt, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_raygen_impl(ctx, cam, spp, jitter, ray_o, ray_d, ray_maxt, (cudaStream_t)stream);
}

int vp_render_forward(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays, float *out_rgb,
                      float *out_T, uint32_t *out_nhits, const vp_hit_record *record, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_render_forward_impl(ctx, params, rays, n_rays, out_rgb, out_T, out_nhits, record, (cudaStream_t)stream);
}

int vp_adjoint_begin(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays, const float *d_L,
                     const float *state_in, const vp_hit_record *record, float *g_data10, float *g_attr, float *g_sh,
                     void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_adjoint_begin_impl(ctx, params, rays, n_rays, d_L, state_in, record, g_data10, g_attr, g_sh, (cudaStream_t)stream);
}

int vp_adjoint_finish(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays,
                      const vp_hit_record *record, int64_t prim_begin, int64_t prim_end, float *g_data10, float *g_attr,
                      float *g_sh, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    return vp_adjoint_finish_impl(ctx, params, rays, n_rays, record, prim_begin, prim_end, g_data10, g_attr, g_sh,
                                  (cudaStream_t)stream);
}

int vp_render_adjoint(vp_ctx *ctx, const vp_params *params, const vp_ray_source *rays, int64_t n_rays, const float *d_L,
                      const float *state_in, const vp_hit_record *record, float *g_data10, float *g_attr, float *g_sh,
                      void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    int rc = vp_adjoint_begin_impl(ctx, params, rays, n_rays, d_L, state_in, record, g_data10, g_attr, g_sh, (cudaStream_t)stream);
    if (rc) return rc;
    return vp_adjoint_finish_impl(ctx, params, rays, n_rays, record, 0, ctx->n, g_data10, g_attr, g_sh, (cudaStream_t)stream);
}

int vp_film_splat(int32_t width, int32_t height, int32_t spp, int32_t rfilter, const float *jitter, const float *radiance,
                  float *accum, void *stream)
{
    return vp_film_splat_impl(width, height, spp, rfilter, jitter, radiance, accum, (cudaStream_t)stream);
}

int vp_film_develop(int32_t width, int32_t height, const float *accum, float *image, int64_t image_row_stride, void *stream)
{
    return vp_film_develop_impl(width, height, accum, image, image_row_stride, (cudaStream_t)stream);
}

int vp_film_adjoint(int32_t width, int32_t height, int32_t spp, int32_t rfilter, const float *jitter, const float *accum,
                    const float *d_image, int64_t d_image_row_stride, float *d_L, void *stream)
{
    return vp_film_adjoint_impl(width, height, spp, rfilter, jitter, accum, d_image, d_image_row_stride, d_L, (cudaStream_t)stream);
}

int vp_set_option(vp_ctx *ctx, const char *name, int64_t value)
{
    if (!ctx || !name) return VP_E_INVALID;
    if (!std::strcmp(name, "record_scratch_bytes")) {
        if (value < (1 << 20)) return vp_fail(ctx, VP_E_INVALID, "vp_set_option: record_scratch_bytes must be at least 1 MiB");
        ctx->record_scratch_bytes = value;
        return VP_OK;
    }
    return vp_fail(ctx, VP_E_INVALID, std::string("vp_set_option: unknown option '") + name + "'");
}

int vp_get_stats(vp_ctx *ctx, vp_stats *host_out, void *stream)
{
    if (!ctx || !host_out) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    std::memset(host_out, 0, sizeof *host_out);
    if (!ctx->stats.ptr) return VP_OK;
    VP_CUDA_CHECK(ctx, cudaMemcpyAsync(host_out, ctx->stats.ptr, sizeof *host_out, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    VP_CUDA_CHECK(ctx, cudaStreamSynchronize((cudaStream_t)stream));
    return VP_OK;
}

int vp_debug_bvh(vp_ctx *ctx, float *out_nodes, int32_t *out_perm, int64_t *n_internal, void *stream)
{
    if (!ctx) return VP_E_INVALID;
    DeviceGuard g(ctx->device);
    if (!ctx->built) return vp_fail(ctx, VP_E_STATE, "vp_debug_bvh: not built");
    int64_t ni = ctx->n > 1 ? ctx->n - 1 : 0;
    if (n_internal) *n_internal = ni;
    cudaStream_t st = (cudaStream_t)stream;
    if (out_nodes && ni > 0)
        VP_CUDA_CHECK(ctx, cudaMemcpyAsync(out_nodes, ctx->nodes.ptr, sizeof(float) * 16 * (size_t)ni, cudaMemcpyDeviceToDevice, st));
    if (out_perm && ctx->n > 0)
        VP_CUDA_CHECK(ctx, cudaMemcpyAsync(out_perm, ctx->perm.ptr, sizeof(int32_t) * (size_t)ctx->n, cudaMemcpyDeviceToDevice, st));
    return VP_OK;
}

int vp_debug_selftest(int32_t which, int64_t n, uint64_t seed, int64_t *mismatches)
{
    if (which != 0 || n < 0 || !mismatches) return VP_E_INVALID;
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, sizeof *d) != cudaSuccess) return VP_E_CUDA;
    cudaMemset(d, 0, sizeof *d);
    k_selftest_div<<<148 * 8, 256>>>(n, seed, d);
    unsigned long long h = 0;
    const cudaError_t e = cudaMemcpy(&h, d, sizeof h, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return VP_E_CUDA;
    *mismatches = (int64_t)h;
    return VP_OK;
}

}  // extern "C"
