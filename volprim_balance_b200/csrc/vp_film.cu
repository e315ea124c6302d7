// vp_film.cu -- film reconstruction: the step after the integrator in mi.render().
//
// Replaces Mitsuba's hdrfilm + reconstruction filter for the sensors volprim configures (volprim/cameras.py:114-137:
// `hdrfilm` with a `tent` rfilter; batch film of examples/refine_3dg_dataset.py:96-107 through a row stride).
// Restated semantics (third-party, unpinned): a sample at film position (x, y) contributes to every pixel whose
// centre lies within the filter radius, with the separable weight f(dx) f(dy); a pixel's value is the weighted
// sum divided by the sum of the weights.
//   box       radius 0.5   f = 1
//   tent      radius 1     f = max(0, 1 - |d|)
//   gaussian  radius 2     f = max(0, exp(-d^2 / (2 * 0.5^2)) - exp(-2^2 / (2 * 0.5^2)))      (stddev 0.5, 4 stddev)
// Three HBM-bound kernels: splat (one thread per sample, 128-bit vector reductions into an RGBW accumulator),
// develop (one thread per pixel), adjoint (one thread per sample gathers d image * weight / pixel weight).
#include "vp_internal.cuh"

namespace {

template <int F>
__device__ __forceinline__ float filter_eval(float d)
{
    d = fabsf(d);
    if (F == VP_RFILTER_BOX) return d <= 0.5f ? 1.f : 0.f;
    if (F == VP_RFILTER_TENT) return fmaxf(0.f, 1.f - d);
    return fmaxf(0.f, expf(-2.f * d * d) - 3.3546262790251185e-4f);   // exp(-8)
}

template <int F>
struct Window {
    static constexpr int K = F == VP_RFILTER_BOX ? 1 : (F == VP_RFILTER_TENT ? 2 : 4);
    int x0, y0;
    float wx[K], wy[K];
};

// pixels [x0, x0 + K) x [y0, y0 + K) a sample can touch and the separable weights
template <int F>
__device__ __forceinline__ Window<F> sample_window(int width, int spp, const float *__restrict__ jitter, int64_t i)
{
    Window<F> w;
    const int64_t pix = i / spp;
    const int x = (int)(pix % width), y = (int)(pix / width);
    const float jx = jitter ? jitter[2 * i] : 0.5f, jy = jitter ? jitter[2 * i + 1] : 0.5f;
    constexpr int K = Window<F>::K;
    if (F == VP_RFILTER_BOX) {
        w.x0 = x; w.y0 = y;
        w.wx[0] = w.wy[0] = 1.f;
        return w;
    }
    // first pixel whose centre can lie within the radius: floor(s - r + 0.5), s = pixel + jitter
    constexpr float rad = F == VP_RFILTER_TENT ? 1.f : 2.f;
    const int ox = (int)floorf(jx - rad + 0.5f), oy = (int)floorf(jy - rad + 0.5f);
    w.x0 = x + ox; w.y0 = y + oy;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        w.wx[k] = filter_eval<F>(jx - ((float)(ox + k) + 0.5f));
        w.wy[k] = filter_eval<F>(jy - ((float)(oy + k) + 0.5f));
    }
    return w;
}

template <int F>
__global__ void k_film_splat(int width, int height, int spp, const float *__restrict__ jitter,
                             const float *__restrict__ radiance, float4 *__restrict__ accum)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)width * height * spp) return;
    const Window<F> w = sample_window<F>(width, spp, jitter, i);
    const float r = radiance[3 * i], g = radiance[3 * i + 1], b = radiance[3 * i + 2];
    constexpr int K = Window<F>::K;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int py = w.y0 + ky;
        if (py < 0 || py >= height) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int px = w.x0 + kx;
            const float wt = w.wx[kx] * w.wy[ky];
            if (px < 0 || px >= width || !(wt > 0.f)) continue;
            atomicAdd(accum + (int64_t)py * width + px, make_float4(wt * r, wt * g, wt * b, wt));
        }
    }
}

__global__ void k_film_develop(int width, int height, const float4 *__restrict__ accum, float *__restrict__ image,
                               int64_t row_stride)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (int64_t)width * height) return;
    const int x = (int)(p % width), y = (int)(p / width);
    const float4 a = accum[p];
    const float inv = a.w > 0.f ? 1.f / a.w : 0.f;
    float *dst = image + (int64_t)y * row_stride + 3 * x;
    dst[0] = a.x * inv; dst[1] = a.y * inv; dst[2] = a.z * inv;
}

template <int F>
__global__ void k_film_adjoint(int width, int height, int spp, const float *__restrict__ jitter,
                               const float4 *__restrict__ accum, const float *__restrict__ d_image, int64_t row_stride,
                               float *__restrict__ d_L)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)width * height * spp) return;
    const Window<F> w = sample_window<F>(width, spp, jitter, i);
    float r = 0.f, g = 0.f, b = 0.f;
    constexpr int K = Window<F>::K;
#pragma unroll
    for (int ky = 0; ky < K; ++ky) {
        const int py = w.y0 + ky;
        if (py < 0 || py >= height) continue;
#pragma unroll
        for (int kx = 0; kx < K; ++kx) {
            const int px = w.x0 + kx;
            const float wt = w.wx[kx] * w.wy[ky];
            if (px < 0 || px >= width || !(wt > 0.f)) continue;
            const float pw = accum[(int64_t)py * width + px].w;
            if (!(pw > 0.f)) continue;
            const float *src = d_image + (int64_t)py * row_stride + 3 * px;
            const float k = wt / pw;
            r = fmaf(src[0], k, r); g = fmaf(src[1], k, g); b = fmaf(src[2], k, b);
        }
    }
    d_L[3 * i] = r; d_L[3 * i + 1] = g; d_L[3 * i + 2] = b;
}

bool bad_film(int32_t w, int32_t h, int32_t spp, int32_t f) { return w <= 0 || h <= 0 || spp <= 0 || f < 0 || f > VP_RFILTER_GAUSSIAN; }

}  // namespace

int vp_film_splat_impl(int32_t w, int32_t h, int32_t spp, int32_t rfilter, const float *jitter, const float *radiance,
                       float *accum, cudaStream_t st)
{
    if (bad_film(w, h, spp, rfilter) || !radiance || !accum || (uintptr_t)accum % 16) return VP_E_INVALID;
    const int64_t n = (int64_t)w * h * spp;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    float4 *acc = reinterpret_cast<float4 *>(accum);
    if (rfilter == VP_RFILTER_BOX) k_film_splat<VP_RFILTER_BOX><<<blocks, 256, 0, st>>>(w, h, spp, jitter, radiance, acc);
    else if (rfilter == VP_RFILTER_TENT) k_film_splat<VP_RFILTER_TENT><<<blocks, 256, 0, st>>>(w, h, spp, jitter, radiance, acc);
    else k_film_splat<VP_RFILTER_GAUSSIAN><<<blocks, 256, 0, st>>>(w, h, spp, jitter, radiance, acc);
    return cudaGetLastError() == cudaSuccess ? VP_OK : VP_E_CUDA;
}

int vp_film_develop_impl(int32_t w, int32_t h, const float *accum, float *image, int64_t row_stride, cudaStream_t st)
{
    if (w <= 0 || h <= 0 || !accum || !image || row_stride < 3ll * w || (uintptr_t)accum % 16) return VP_E_INVALID;
    const int64_t n = (int64_t)w * h;
    k_film_develop<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(w, h, reinterpret_cast<const float4 *>(accum), image, row_stride);
    return cudaGetLastError() == cudaSuccess ? VP_OK : VP_E_CUDA;
}

int vp_film_adjoint_impl(int32_t w, int32_t h, int32_t spp, int32_t rfilter, const float *jitter, const float *accum,
                         const float *d_image, int64_t row_stride, float *d_L, cudaStream_t st)
{
    if (bad_film(w, h, spp, rfilter) || !accum || !d_image || !d_L || row_stride < 3ll * w || (uintptr_t)accum % 16) return VP_E_INVALID;
    const int64_t n = (int64_t)w * h * spp;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    const float4 *acc = reinterpret_cast<const float4 *>(accum);
    if (rfilter == VP_RFILTER_BOX) k_film_adjoint<VP_RFILTER_BOX><<<blocks, 256, 0, st>>>(w, h, spp, jitter, acc, d_image, row_stride, d_L);
    else if (rfilter == VP_RFILTER_TENT) k_film_adjoint<VP_RFILTER_TENT><<<blocks, 256, 0, st>>>(w, h, spp, jitter, acc, d_image, row_stride, d_L);
    else k_film_adjoint<VP_RFILTER_GAUSSIAN><<<blocks, 256, 0, st>>>(w, h, spp, jitter, acc, d_image, row_stride, d_L);
    return cudaGetLastError() == cudaSuccess ? VP_OK : VP_E_CUDA;
}
