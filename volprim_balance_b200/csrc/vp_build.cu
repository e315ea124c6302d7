// vp_build.cu -- GPU construction of the ellipsoid-primitive acceleration structure.
//
// Replaces the acceleration-structure build that Mitsuba performs inside mi.load_dict() and again on
// every params.update() (reference: examples/refine_3dg_dataset.py:155-159, SURVEY.md section 2.1 "Accel
// rebuild").  Design: LBVH (Karras 2012) over the bounding boxes of the primitives' bounding ellipsoids
// { x : |diag(1/(extent s)) R^T (x - c)| = 1 }  (reference common.py:346-352).
//
//   1. scene bounds of the centres            (block reduce + ordered-uint atomics)
//   2. 63-bit Morton code per primitive
//   3. LSD radix sort of (code, index), 8 bits per pass, hand-written (tile histogram / multi-block scan / ranked scatter)
//   4. gather the reference AoS records into the Morton-ordered 128-bit SoA + leaf boxes
//   5. Karras hierarchy over the sorted codes
//   6. bottom-up box fit with per-node arrival counters
//
// vp_refit re-runs 4 and 6 only (topology and order of the last build are kept).
//
// Node layout, 16 floats (4 x float4) per internal node i:
//   [0..2] left.min   [3..5] left.max   [6..8] right.min   [9..11] right.max
//   [12] left link    [13] right link   [14] parent        [15] unused
// A link >= 0 is an internal node, a link < 0 is the leaf ~link (= sorted primitive position).
#include "vp_internal.cuh"
#include "vp_scan.cuh"

#include <cfloat>

namespace {

constexpr int RADIX_BITS = 8;
constexpr int RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__device__ __forceinline__ uint32_t ordered_from_float(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float float_from_ordered(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_init_bounds(uint32_t *bounds)
{
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

// 1. bounds of the primitive centres (reference layout: data10[10*j + 0..2])
__global__ void k_center_bounds(const float *__restrict__ data10, int n, uint32_t *bounds)
{
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float c = data10[10ll * j + a];
            if (isfinite(c)) { lo[a] = fminf(lo[a], c); hi[a] = fmaxf(hi[a], c); }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        for (int off = 16; off; off >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], off));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], off));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bounds[a], ordered_from_float(lo[a]));
            atomicMax(&bounds[3 + a], ordered_from_float(hi[a]));
        }
    }
}

__device__ __forceinline__ uint64_t spread21(uint64_t x)
{
    x &= 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

// 2. 63-bit Morton codes
__global__ void k_morton(const float *__restrict__ data10, int n, const uint32_t *__restrict__ bounds,
                         uint64_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint64_t code = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = float_from_ordered(bounds[a]), hi = float_from_ordered(bounds[3 + a]);
        float c = data10[10ll * j + a];
        float ext = hi - lo;
        float u = (ext > 0.f && isfinite(c)) ? (c - lo) / ext : 0.f;
        u = fminf(fmaxf(u, 0.f), 1.f);
        uint32_t q = (uint32_t)fminf(u * 2097152.f, 2097151.f);
        code |= spread21(q) << (2 - a);
    }
    keys[j] = code;
    vals[j] = (uint32_t)j;
}

// 3a. per-tile digit histogram, stored digit-major: hist[digit * n_tiles + tile]
__global__ void __launch_bounds__(SORT_THREADS) k_radix_hist(const uint64_t *__restrict__ keys, int n, int shift,
                                                              int n_tiles, uint32_t *__restrict__ hist)
{
    __shared__ uint32_t sh[RADIX];
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) sh[i] = 0;
    __syncthreads();
    int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        int i = base + k * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&sh[(uint32_t)(keys[i] >> shift) & (RADIX - 1)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RADIX; i += SORT_THREADS) hist[(size_t)i * n_tiles + blockIdx.x] = sh[i];
}

// 3b. exclusive scan of the digit-major histogram [256 * n_tiles]: vpscan::exclusive_scan (multi-block, vp_scan.cuh)

// 3c. stable ranked scatter.  Each warp owns a contiguous 512-key slice of the tile and walks it in
// 32-key rounds; __match_any_sync groups equal digits, the group's lowest lane owns the counter.
__global__ void __launch_bounds__(SORT_THREADS) k_radix_scatter(const uint64_t *__restrict__ keys_in,
                                                                 const uint32_t *__restrict__ vals_in, int n, int shift,
                                                                 int n_tiles, const uint32_t *__restrict__ offsets,
                                                                 uint64_t *__restrict__ keys_out,
                                                                 uint32_t *__restrict__ vals_out)
{
    constexpr int WARPS = SORT_THREADS / 32;
    __shared__ uint32_t wcount[WARPS][RADIX];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < WARPS * RADIX; i += SORT_THREADS) (&wcount[0][0])[i] = 0;
    __syncthreads();

    const int wbase = blockIdx.x * SORT_TILE + wid * (32 * SORT_ITEMS);
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        int i = wbase + k * 32 + lane;
        bool ok = i < n;
        key[k] = ok ? keys_in[i] : ~0ull;
        val[k] = ok ? vals_in[i] : 0u;
        uint32_t dgt = (uint32_t)(key[k] >> shift) & (RADIX - 1);
        uint32_t peers = __match_any_sync(0xffffffffu, ok ? dgt : 0xffffffffu);
        if (ok && (peers & ((1u << lane) - 1)) == 0) wcount[wid][dgt] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps of this tile + global offset of (digit, tile)
    for (int dgt = threadIdx.x; dgt < RADIX; dgt += SORT_THREADS) {
        uint32_t run = offsets[(size_t)dgt * n_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            uint32_t c = wcount[w][dgt];
            wcount[w][dgt] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        int i = wbase + k * 32 + lane;
        bool ok = i < n;
        uint32_t dgt = (uint32_t)(key[k] >> shift) & (RADIX - 1);
        uint32_t peers = __match_any_sync(0xffffffffu, ok ? dgt : 0xffffffffu);
        uint32_t below = peers & ((1u << lane) - 1);
        uint32_t dst = 0;
        if (ok) dst = wcount[wid][dgt] + __popc(below);
        __syncwarp();
        if (ok && below == 0) wcount[wid][dgt] += __popc(peers);
        __syncwarp();
        if (ok) { keys_out[dst] = key[k]; vals_out[dst] = val[k]; }
    }
}

// 4. Morton-ordered SoA + padded leaf boxes
__global__ void k_gather_soa(const float *__restrict__ data10, const float *__restrict__ attr,
                             const float *__restrict__ sh, int n, int sh_floats, int sh_stride4, float extent,
                             const uint32_t *__restrict__ order, float4 *__restrict__ geo0, float4 *__restrict__ geo1,
                             float4 *__restrict__ geo2, float4 *__restrict__ sh4, float4 *__restrict__ xf, float *__restrict__ info,
                             int32_t *__restrict__ perm,
                             int32_t *__restrict__ inv_perm, float4 *__restrict__ leaf_lo, float4 *__restrict__ leaf_hi)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    float area = 0.f, half = 0.f;
    if (p < n) {
    int j = (int)order[p];
    const float *rec = data10 + 10ll * j;
    float c[3] = { rec[0], rec[1], rec[2] }, s[3] = { rec[3], rec[4], rec[5] };
    float4 q = make_float4(rec[6], rec[7], rec[8], rec[9]);
    geo0[p] = make_float4(c[0], c[1], c[2], attr ? attr[j] : 1.f);
    geo1[p] = make_float4(s[0], s[1], s[2], __int_as_float(j));
    geo2[p] = q;
    perm[p] = j;
    inv_perm[j] = p;
    if (sh_floats > 0) {
        const float *f = sh + (size_t)j * sh_floats;
        for (int i = 0; i < sh_stride4; ++i) {
            float4 v;
            v.x = (4 * i + 0 < sh_floats) ? f[4 * i + 0] : 0.f;
            v.y = (4 * i + 1 < sh_floats) ? f[4 * i + 1] : 0.f;
            v.z = (4 * i + 2 < sh_floats) ? f[4 * i + 2] : 0.f;
            v.w = (4 * i + 3 < sh_floats) ? f[4 * i + 3] : 0.f;
            sh4[(size_t)p * sh_stride4 + i] = v;
        }
    }
    // world box of { x : |diag(1/(extent s)) R^T (x - c)| = 1 } = c + R^{-T} diag(extent s) u, |u| = 1;
    // R^{-1} by cofactors so that un-normalised quaternions (reference quirk Q6) stay bounded correctly.
    double x = q.x, y = q.y, z = q.z, w = q.w;
    double R[3][3] = { { 1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w) },
                       { 2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w) },
                       { 2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y) } };
    double det = R[0][0] * (R[1][1] * R[2][2] - R[1][2] * R[2][1]) - R[0][1] * (R[1][0] * R[2][2] - R[1][2] * R[2][0])
               + R[0][2] * (R[1][0] * R[2][1] - R[1][1] * R[2][0]);
    double inv[3][3];
    inv[0][0] = (R[1][1] * R[2][2] - R[1][2] * R[2][1]) / det;
    inv[0][1] = (R[0][2] * R[2][1] - R[0][1] * R[2][2]) / det;
    inv[0][2] = (R[0][1] * R[1][2] - R[0][2] * R[1][1]) / det;
    inv[1][0] = (R[1][2] * R[2][0] - R[1][0] * R[2][2]) / det;
    inv[1][1] = (R[0][0] * R[2][2] - R[0][2] * R[2][0]) / det;
    inv[1][2] = (R[0][2] * R[1][0] - R[0][0] * R[1][2]) / det;
    inv[2][0] = (R[1][0] * R[2][1] - R[1][1] * R[2][0]) / det;
    inv[2][1] = (R[0][1] * R[2][0] - R[0][0] * R[2][1]) / det;
    inv[2][2] = (R[0][0] * R[1][1] - R[0][1] * R[1][0]) / det;
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double h2 = 0;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double b = inv[k][a] * (double)s[k] * (double)extent;
            h2 += b * b;
        }
        double h = sqrt(h2);
        if (!isfinite(h)) h = 1e30;
        double pad = 1e-4 * h + 1e-6 * (fabs((double)c[a]) + 1.0);
        lo[a] = __double2float_rd((double)c[a] - h - pad);
        hi[a] = __double2float_ru((double)c[a] + h + pad);
        if (!isfinite(c[a])) { lo[a] = FLT_MAX; hi[a] = -FLT_MAX; }  // unhittable
    }
    leaf_lo[p] = make_float4(lo[0], lo[1], lo[2], 0.f);
    leaf_hi[p] = make_float4(hi[0], hi[1], hi[2], 0.f);
    // unit-sphere transform for the ordering test: M = diag(1/(extent s)) R^T (fp32 from the double matrix)
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double se = (double)s[i] * (double)extent;
        xf[3ll * p + i] = make_float4((float)(R[0][i] / se), (float)(R[1][i] / se), (float)(R[2][i] / se), c[i]);
    }
    // mean projected area of the leaf box = surface area / 4 (feeds the initial interval width)
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    if (isfinite(ex) && isfinite(ey) && isfinite(ez) && ex > 0.f && ey > 0.f && ez > 0.f) {
        area = 0.5f * (ex * ey + ey * ez + ez * ex);
        half = (ex + ey + ez) * (1.f / 6.f);
    }
    }
    // The two sums are accumulated as 64-bit fixed-point integers (2^-32 units): integer addition is associative, so the
    // build is bit-reproducible whatever order the warps arrive in.  (A float atomicAdd here made delta0 differ in its
    // last bits from run to run, and with it the interval partition -- and the hit order of fragile rays -- between the
    // replicas of a multi-GPU job.)
    for (int off = 16; off; off >>= 1) area += __shfl_xor_sync(0xffffffffu, area, off);
    for (int off = 16; off; off >>= 1) half += __shfl_xor_sync(0xffffffffu, half, off);
    unsigned long long *acc = reinterpret_cast<unsigned long long *>(info + 12);
    if ((threadIdx.x & 31) == 0 && area > 0.f) atomicAdd(acc, (unsigned long long)fmin((double)area * 4294967296.0, 1.8e19));
    if ((threadIdx.x & 31) == 0 && half > 0.f) atomicAdd(acc + 1, (unsigned long long)fmin((double)half * 4294967296.0, 1.8e19));
}

// scene box (union of the root's child boxes, or the single leaf) and the initial interval width
// delta0 = 2 TARGET / (box crossings per unit length), crossings = sum(mean projected leaf area) / volume
__global__ void k_scene_info(int n, const float *__restrict__ nodes, const float4 *__restrict__ leaf_lo,
                             const float4 *__restrict__ leaf_hi, float *__restrict__ info)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    float lo[3], hi[3];
    if (n == 1) {
        lo[0] = leaf_lo[0].x; lo[1] = leaf_lo[0].y; lo[2] = leaf_lo[0].z;
        hi[0] = leaf_hi[0].x; hi[1] = leaf_hi[0].y; hi[2] = leaf_hi[0].z;
    } else {
        for (int a = 0; a < 3; ++a) {
            lo[a] = fminf(nodes[a], nodes[6 + a]);
            hi[a] = fmaxf(nodes[3 + a], nodes[9 + a]);
        }
    }
    float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
    float diag = sqrtf(ex * ex + ey * ey + ez * ez);
    float vol = ex * ey * ez;
    const unsigned long long *acc = reinterpret_cast<const unsigned long long *>(info + 12);
    float area = (float)((double)acc[0] * (1.0 / 4294967296.0));
    info[7] = area;
    info[8] = (float)((double)acc[1] * (1.0 / 4294967296.0));
    float delta0 = diag * (1.f / 64.f);
    if (vol > 0.f && area > 0.f && isfinite(vol) && isfinite(area)) delta0 = 2.f * 12.f * vol / area;
    if (!(delta0 > 0.f) || !isfinite(delta0)) delta0 = 1.f;
    delta0 = fminf(delta0, fmaxf(diag, 1e-20f));
    for (int a = 0; a < 3; ++a) { info[a] = lo[a]; info[3 + a] = hi[a]; }
    info[6] = delta0;
    info[8] = info[8] / (float)n;   // mean half-extent of a leaf box (tile-vs-per-ray walker heuristic)
}

// 5. Karras 2012
__device__ __forceinline__ int lcp(const uint64_t *__restrict__ keys, int n, int i, uint64_t ki, int j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t kj = keys[j];
    if (ki == kj) return 64 + __clz(i ^ j);
    return __clzll((long long)(ki ^ kj));
}

__global__ void k_hierarchy(const uint64_t *__restrict__ keys, int n, float *__restrict__ nodes,
                            int32_t *__restrict__ leaf_parent, int32_t *__restrict__ counters)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint64_t ki = keys[i];
    int d = (lcp(keys, n, i, ki, i + 1) - lcp(keys, n, i, ki, i - 1)) >= 0 ? 1 : -1;
    int dmin = lcp(keys, n, i, ki, i - d);
    int lmax = 2;
    while (lcp(keys, n, i, ki, i + lmax * d) > dmin) {
        lmax <<= 1;
        if (lmax > (1 << 30)) break;
    }
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) {
        long long jj = (long long)i + (long long)(l + t) * d;
        if (jj >= 0 && jj < n && lcp(keys, n, i, ki, (int)jj) > dmin) l += t;
    }
    int j = i + l * d;
    int dnode = lcp(keys, n, i, ki, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lcp(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    float *nd = nodes + 16ll * i;
    nd[12] = __int_as_float(left);
    nd[13] = __int_as_float(right);
    if (i == 0) nd[14] = __int_as_float(-1);
    nd[15] = 0.f;
    if (left < 0) leaf_parent[~left] = i; else nodes[16ll * left + 14] = __int_as_float(i);
    if (right < 0) leaf_parent[~right] = i; else nodes[16ll * right + 14] = __int_as_float(i);
    counters[i] = 0;
}

// 6. bottom-up fit: the second thread to arrive at a node owns it and continues upward.
__global__ void k_refit(int n, float *__restrict__ nodes, const int32_t *__restrict__ leaf_parent,
                        const float4 *__restrict__ leaf_lo, const float4 *__restrict__ leaf_hi,
                        int32_t *__restrict__ counters)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    float4 l4 = leaf_lo[p], h4 = leaf_hi[p];
    float lo[3] = { l4.x, l4.y, l4.z }, hi[3] = { h4.x, h4.y, h4.z };
    int child = ~p;
    int node = leaf_parent[p];
    while (node >= 0) {
        float *nd = nodes + 16ll * node;
        int left = __float_as_int(__ldcg(nd + 12));
        float *slot = nd + (left == child ? 0 : 6);
        slot[0] = lo[0]; slot[1] = lo[1]; slot[2] = lo[2];
        slot[3] = hi[0]; slot[4] = hi[1]; slot[5] = hi[2];
        __threadfence();
        if (atomicAdd(&counters[node], 1) == 0) return;
        __threadfence();
        const float *other = nd + (left == child ? 6 : 0);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = fminf(lo[a], __ldcg(other + a));
            hi[a] = fmaxf(hi[a], __ldcg(other + 3 + a));
        }
        child = node;
        node = __float_as_int(__ldcg(nd + 14));
    }
}

__global__ void k_zero_i32(int32_t *p, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = 0;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace

int vp_build_impl(vp_ctx *ctx, bool refit_only, cudaStream_t st)
{
    if (!ctx->have_prims) return vp_fail(ctx, VP_E_STATE, "vp_build/vp_refit: no primitives set (call vp_set_primitives)");
    const int64_t n64 = ctx->n;
    if (n64 > 0x3fffffff) return vp_fail(ctx, VP_E_INVALID, "vp_build: more than 2^30 primitives are not supported");
    const int n = (int)n64;
    if (refit_only && (!ctx->built || ctx->built_n != n64))
        return vp_fail(ctx, VP_E_STATE, "vp_refit: needs a previous vp_build with the same primitive count");
    const int sh_stride4 = (ctx->sh_floats + 3) / 4;
    if (n == 0) {
        ctx->built = true;
        ctx->built_n = 0;
        ctx->root = 0;
        return VP_OK;
    }
    int rc;
    if ((rc = vp_ensure(ctx, ctx->geo0, sizeof(float4) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->geo1, sizeof(float4) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->geo2, sizeof(float4) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->sh4, sizeof(float4) * (size_t)n * (sh_stride4 > 0 ? sh_stride4 : 1)))) return rc;
    if ((rc = vp_ensure(ctx, ctx->xf, sizeof(float4) * 3 * (size_t)n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->info, sizeof(float) * 16))) return rc;
    if ((rc = vp_ensure(ctx, ctx->nodes, sizeof(float) * 16 * (size_t)(n > 1 ? n - 1 : 1)))) return rc;
    if ((rc = vp_ensure(ctx, ctx->perm, sizeof(int32_t) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->inv_perm, sizeof(int32_t) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->leaf_lo, sizeof(float4) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->leaf_hi, sizeof(float4) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->parent, sizeof(int32_t) * n))) return rc;
    if ((rc = vp_ensure(ctx, ctx->counters, sizeof(int32_t) * n))) return rc;
    for (int k = 0; k < 2; ++k) {
        if ((rc = vp_ensure(ctx, ctx->keys[k], sizeof(uint64_t) * n))) return rc;
        if ((rc = vp_ensure(ctx, ctx->vals[k], sizeof(uint32_t) * n))) return rc;
    }
    const int n_tiles = cdiv(n, SORT_TILE);
    if ((rc = vp_ensure(ctx, ctx->hist, sizeof(uint32_t) * (size_t)RADIX * n_tiles))) return rc;
    if ((rc = vp_ensure(ctx, ctx->bounds, sizeof(uint32_t) * 8))) return rc;
    if ((rc = vp_ensure(ctx, ctx->scan_tmp, sizeof(uint64_t) * (size_t)(vpscan::n_tiles((int64_t)RADIX * n_tiles) + 1)))) return rc;

    const float *data10 = (const float *)ctx->raw_data.ptr;
    const float *attr = ctx->have_attr ? (const float *)ctx->raw_attr.ptr : nullptr;
    const float *sh = ctx->sh_floats ? (const float *)ctx->raw_sh.ptr : nullptr;
    const int B = 256;
    const uint32_t *order;
    const uint64_t *sorted_keys = nullptr;

    if (!refit_only) {
        uint32_t *bounds = (uint32_t *)ctx->bounds.ptr;
        k_init_bounds<<<1, 32, 0, st>>>(bounds);
        k_center_bounds<<<min(cdiv(n, B), 148 * 8), B, 0, st>>>(data10, n, bounds);
        int cur = 0;
        k_morton<<<cdiv(n, B), B, 0, st>>>(data10, n, bounds, (uint64_t *)ctx->keys[0].ptr, (uint32_t *)ctx->vals[0].ptr);
        for (int shift = 0; shift < 63; shift += RADIX_BITS) {
            uint64_t *kin = (uint64_t *)ctx->keys[cur].ptr, *kout = (uint64_t *)ctx->keys[cur ^ 1].ptr;
            uint32_t *vin = (uint32_t *)ctx->vals[cur].ptr, *vout = (uint32_t *)ctx->vals[cur ^ 1].ptr;
            uint32_t *hist = (uint32_t *)ctx->hist.ptr;
            k_radix_hist<<<n_tiles, SORT_THREADS, 0, st>>>(kin, n, shift, n_tiles, hist);
            vpscan::exclusive_scan<uint32_t>(hist, (int64_t)RADIX * n_tiles, hist, (uint32_t *)ctx->scan_tmp.ptr, nullptr, nullptr, st);
            k_radix_scatter<<<n_tiles, SORT_THREADS, 0, st>>>(kin, vin, n, shift, n_tiles, hist, kout, vout);
            cur ^= 1;
        }
        order = (const uint32_t *)ctx->vals[cur].ptr;
        sorted_keys = (const uint64_t *)ctx->keys[cur].ptr;
    } else {
        order = (const uint32_t *)ctx->perm.ptr;  // perm holds the same values (int32 >= 0)
    }

    VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->info.ptr, 0, sizeof(float) * 16, st));
    k_gather_soa<<<cdiv(n, B), B, 0, st>>>(data10, attr, sh, n, ctx->sh_floats, sh_stride4, ctx->extent, order,
                                           (float4 *)ctx->geo0.ptr, (float4 *)ctx->geo1.ptr, (float4 *)ctx->geo2.ptr,
                                           (float4 *)ctx->sh4.ptr, (float4 *)ctx->xf.ptr, (float *)ctx->info.ptr,
                                           (int32_t *)ctx->perm.ptr,
                                           (int32_t *)ctx->inv_perm.ptr, (float4 *)ctx->leaf_lo.ptr,
                                           (float4 *)ctx->leaf_hi.ptr);
    if (n == 1) {
        ctx->root = ~0;
    } else {
        if (!refit_only) {
            k_hierarchy<<<cdiv(n - 1, B), B, 0, st>>>(sorted_keys, n, (float *)ctx->nodes.ptr,
                                                      (int32_t *)ctx->parent.ptr, (int32_t *)ctx->counters.ptr);
        } else {
            k_zero_i32<<<cdiv(n - 1, B), B, 0, st>>>((int32_t *)ctx->counters.ptr, n - 1);
        }
        k_refit<<<cdiv(n, B), B, 0, st>>>(n, (float *)ctx->nodes.ptr, (const int32_t *)ctx->parent.ptr,
                                          (const float4 *)ctx->leaf_lo.ptr, (const float4 *)ctx->leaf_hi.ptr,
                                          (int32_t *)ctx->counters.ptr);
        ctx->root = 0;
    }
    k_scene_info<<<1, 32, 0, st>>>(n, (const float *)ctx->nodes.ptr, (const float4 *)ctx->leaf_lo.ptr,
                                   (const float4 *)ctx->leaf_hi.ptr, (float *)ctx->info.ptr);
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    ctx->built = true;
    ctx->built_n = n64;
    return VP_OK;
}
