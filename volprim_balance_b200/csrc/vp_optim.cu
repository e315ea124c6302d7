// vp_optim.cu -- fused bounds-aware Adam step (consumer of the adjoint's gradients).
//
// Replaces the chain of Dr.Jit element-wise kernels of BoundedAdam.step (reference volprim/optimizers.py:72-146) by ONE
// pass over (param, grad, m, v): 16 B read + 12 B written per element, HBM-bound.  Semantics reproduced literally:
//   g = isnan(g) ? 0 : g                                                    (:88)
//   m = b1 m + (1 - b1) g ;  v = b2 v + (1 - b2) g^2                         (:97-98)
//   u = p - lr_t m / (sqrt(v) + eps),  lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) (:83-85, :112)
//   upper:  over = u >= upper ; p' = (over && p >= upper) ? upper : p ; u = over ? p' + (upper - p') / 2 : u   (:124-127)
//   lower:  over = u <= lower ; p' = (over && p' <= lower) ? lower : p' ; u = over ? p' - (p' - lower) / 2 : u  (:128-131)
//   moments reset where `over` -- and, as in the reference, the LOWER test overwrites the upper mask (:129, :134-138)
#include "vp_internal.cuh"

namespace {

__global__ void __launch_bounds__(256) k_bounded_adam(int64_t n, float *__restrict__ p, const float *__restrict__ g,
                                                      float *__restrict__ m, float *__restrict__ v, float lr_t, float b1,
                                                      float omb1, float b2, float omb2, float eps, int has_lower, float lower, int has_upper,
                                                      float upper)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
    for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 < n; i0 += stride) {
        float pv[4], gv[4], mv[4], vv[4];
        const bool full = i0 + 4 <= n;
        if (full) {
            float4 a = *reinterpret_cast<const float4 *>(p + i0), b = __ldg(reinterpret_cast<const float4 *>(g + i0));
            float4 c = *reinterpret_cast<const float4 *>(m + i0), d = *reinterpret_cast<const float4 *>(v + i0);
            pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w; gv[0] = b.x; gv[1] = b.y; gv[2] = b.z; gv[3] = b.w;
            mv[0] = c.x; mv[1] = c.y; mv[2] = c.z; mv[3] = c.w; vv[0] = d.x; vv[1] = d.y; vv[2] = d.z; vv[3] = d.w;
        } else {
            for (int k = 0; k < 4; ++k) {
                const bool ok = i0 + k < n;
                pv[k] = ok ? p[i0 + k] : 0.f; gv[k] = ok ? g[i0 + k] : 0.f;
                mv[k] = ok ? m[i0 + k] : 0.f; vv[k] = ok ? v[i0 + k] : 0.f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gg = isnan(gv[k]) ? 0.f : gv[k];
            float mt = b1 * mv[k] + omb1 * gg;
            float vt = b2 * vv[k] + omb2 * gg * gg;
            float val = pv[k];
            float u = val - lr_t * mt / (sqrtf(vt) + eps);
            bool over = false;
            if (has_upper) {
                over = u >= upper;
                if (over && val >= upper) val = upper;
                if (over) u = val + 0.5f * (upper - val);
            }
            if (has_lower) {
                over = u <= lower;
                if (over && val <= lower) val = lower;
                if (over) u = val - 0.5f * (val - lower);
            }
            if (over) { mt = 0.f; vt = 0.f; }
            pv[k] = u; mv[k] = mt; vv[k] = vt;
        }
        if (full) {
            *reinterpret_cast<float4 *>(p + i0) = make_float4(pv[0], pv[1], pv[2], pv[3]);
            *reinterpret_cast<float4 *>(m + i0) = make_float4(mv[0], mv[1], mv[2], mv[3]);
            *reinterpret_cast<float4 *>(v + i0) = make_float4(vv[0], vv[1], vv[2], vv[3]);
        } else {
            for (int k = 0; k < 4 && i0 + k < n; ++k) { p[i0 + k] = pv[k]; m[i0 + k] = mv[k]; v[i0 + k] = vv[k]; }
        }
    }
}

// l1(reference, image) of volprim/optimizers.py:170-174 together with what dr.backward(loss) seeds the render op with
// (d loss / d image = sign(image - reference) / n_total) and the squared error psnr() needs (:180-186): one pass, 8 B read
// + 4 B written per element, block sums reduced into sums[0] (sum |diff| / n_total) and sums[1] (sum diff^2 / n_total).
__global__ void __launch_bounds__(256) k_l1_loss_grad(int64_t n, const float *__restrict__ image, const float *__restrict__ reference,
                                                      float inv_n_total, float *__restrict__ d_image, float *__restrict__ sums)
{
    float a = 0.f, q = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float diff = image[i] - __ldg(reference + i);
        a += fabsf(diff);
        q = fmaf(diff, diff, q);
        d_image[i] = diff > 0.f ? inv_n_total : (diff < 0.f ? -inv_n_total : 0.f);
    }
    for (int off = 16; off; off >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    __shared__ float sa[8], sq[8];
    if ((threadIdx.x & 31) == 0) { sa[threadIdx.x >> 5] = a; sq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float ta = 0.f, tq = 0.f;
        for (int k = 0; k < 8; ++k) { ta += sa[k]; tq += sq[k]; }
        atomicAdd(sums, ta * inv_n_total);
        atomicAdd(sums + 1, tq * inv_n_total);
    }
}

}  // namespace

extern "C" int vp_l1_loss_grad(int64_t n, const float *image, const float *reference, double n_total, float *d_image,
                               float *sums, void *stream)
{
    if (n < 0 || !(n_total > 0) || (n > 0 && (!image || !reference || !d_image || !sums))) return VP_E_INVALID;
    if (n == 0) return VP_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k_l1_loss_grad<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(n, image, reference, (float)(1.0 / n_total), d_image, sums);
    return cudaGetLastError() == cudaSuccess ? VP_OK : VP_E_CUDA;
}

extern "C" int vp_bounded_adam_step(int64_t n, float *param, const float *grad, float *m, float *v, double lr_t, double beta1,
                                    double beta2, double eps, int has_lower, float lower, int has_upper, float upper,
                                    void *stream)
{
    if (n < 0 || (n > 0 && (!param || !grad || !m || !v))) return VP_E_INVALID;
    if (n == 0) return VP_OK;
    if (((uintptr_t)param | (uintptr_t)grad | (uintptr_t)m | (uintptr_t)v) & 15) return VP_E_INVALID;  // 128-bit accesses
    int64_t blocks = (n / 4 + 255) / 256;
    const int64_t cap = 148 * 16;   // a few waves of the 148 SMs; grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_bounded_adam<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        n, param, grad, m, v, (float)lr_t, (float)beta1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
        has_lower, lower, has_upper, upper);  // 1 - beta in double first: 1.f - 0.999f is off by 1.3e-5 relative
    return cudaGetLastError() == cudaSuccess ? VP_OK : VP_E_CUDA;
}
