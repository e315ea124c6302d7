// vp_trace.cu -- per-ray volumetric-primitive integration on sm_100a.
//
// Replaces the Dr.Jit megakernels "Primitive splatting (Primal/Backward)" and "Primitive tracing
// (Primal/Backward)" together with the scene.ray_intersect() they call once per hit
// (reference: volprim/integrators/volprim_rf.py:103-192, volprim_tomography.py:47-127).
//
// Algorithm (one thread per ray; loops warp-uniform).  The ray is cut into consecutive intervals [t_start, t_start +
// delta]; delta adapts so that the lists below fill to ~70 % (tile walker) / hold about ten entries (per-ray walker).
//   phase 1   (collect)  LBVH walk restricted to the interval; leaves are only APPENDED to a candidate list in shared
//                        memory.  No test result feeds back into the walk (the per-lane k-buffer version measured 7
//                        active lanes of 32 here).  Tile walker (image-shaped launches, a warp = an 8x4 pixel tile):
//                        ONE walk per warp -- 32 queued nodes per step are tested by the 32 lanes against a capsule
//                        around the tile's ray segments.  Per-ray walker (explicit ray batches, and what a tile hands
//                        over to when its rays are incoherent): branch-free stack traversal per lane.
//   phase 1.5 (cull)     tile walker only: one candidate per lane against the tile's ray bundle in the primitive's
//                        unit-sphere space; more than half of the box-level candidates are touched by no ray.
//   phase 2   (test)     every candidate gets the ray/ellipsoid entry distance from the pre-transformed SoA record
//                        (48 B); entries inside the interval go to a per-lane list sorted by distance;
//   phase 3   (drain)    entries are taken in increasing distance; each is RE-EVALUATED against the current, re-based
//                        origin with the same fixed-order fp32 arithmetic as the parity oracle.  This applies the
//                        reference's "advance the origin by 1e-4 and query again with back-face culling" rule exactly
//                        (entries that fell behind the advanced origin are dropped) and keeps the
//                        origin bit-identical to the one the reference loop carries.
//   The ray stops on leaving the scene box (miss), max_depth, or (rf) beta <= t_cutoff.  A tile whose lane lists fill
//   up ends the interval at the nearest full lane's farthest entry (nothing is walked twice); candidate-list overflows
//   retry with a shorter interval; below a minimum width the per-ray walker runs a plain closest-hit search per step.
// The adjoint either replays recorded hit lists (no BVH access) or re-traces like the primal.
#include "vp_internal.cuh"
#include <cstdlib>
#include "vp_scan.cuh"

#include <cfloat>
#include <cmath>

// Value-path arithmetic.  Everything that ORDERS hits or applies the epsilon cull is evaluated with individually
// rounded operations (exact_isect); everything else -- the transmittance of a hit, the ordering keys of the candidate
// lists (re-decided by the exact test), the conservative culls -- may use approximate reciprocals, roots and ex2
// (2 ulp).  IEEE divisions / roots / expf in those places cost the issue-bound forward kernel 12 % (measured).
#ifndef VP_EVAL_FAST
#define VP_EVAL_FAST 1
#endif
#ifndef VP_APPROX_ORDER
#define VP_APPROX_ORDER 1
#endif
#ifndef VP_DIV_SHARED
#define VP_DIV_SHARED 1
#endif
#ifndef VP_RANKS_MATCH
#define VP_RANKS_MATCH 1
#endif

namespace {

#ifndef VP_CAND_CAP
#define VP_CAND_CAP 36   // 6 blocks x 36 KB of shared memory per SM
#endif
#ifndef VP_MIN_BLOCKS
#define VP_MIN_BLOCKS 6   // 80 registers / thread; measured on cfg2: 4 blocks +13 %, 5: +4 %, 7: +8 %, 8: +25 % time
#endif
constexpr int CAND_CAP = VP_CAND_CAP;     // candidate / hit list entries per ray (shared memory: 8 B each)
#ifndef VP_TARGET_HITS
#define VP_TARGET_HITS 10
#endif
constexpr int TARGET_HITS = VP_TARGET_HITS;  // the interval width adapts towards this many entries per interval
constexpr int STACK_MAX = 96;    // LBVH depth bound: 63 Morton bits + index tie-break bits
constexpr int TRACE_THREADS = 128;
constexpr int NODE_SENTINEL = 0x7fffffff;
constexpr size_t TRACE_SMEM = (size_t)CAND_CAP * TRACE_THREADS * 8;
// warp-cooperative (tile) walker: per-lane hit lists + per-warp node queue and tile candidate list
// (tunables; the values are the measured optimum over cfg2 / cfg3 / cfg5, see DESIGN.md sections 7-8)
#ifndef VP_FILL
#define VP_FILL 0.7f            // interval width aims at this fill of the fullest list
#endif
#ifndef VP_OVF_SHRINK
#define VP_OVF_SHRINK 0.35f     // width factor after a candidate-list / queue overflow
#endif
#ifndef VP_GROW_MAX
#define VP_GROW_MAX 2.f         // largest growth of the width from one interval to the next
#endif
#ifndef VP_CSHRINK_FLOOR
#define VP_CSHRINK_FLOOR 32.f   // the candidate count stops shrinking the width below delta0 / this
#endif
#ifndef VP_CAND_TARGET
#define VP_CAND_TARGET (VP_FILL * VP_TILE_CCAP)
#endif
#ifndef VP_TILE_HIT_CAP
#define VP_TILE_HIT_CAP 24
#endif
// 24 x 128 x 8 B of lists + 4 x (224 + 192) x 4 B of queues = 30.5 KB per block: six blocks fit the 196 KB shared-memory
// carve-out, which leaves the SM 60 KB of L1 instead of 28 KB (cfg2 / cfg3 -2 %; a 512-entry queue was never needed)
#ifndef VP_TILE_QCAP
#define VP_TILE_QCAP 224
#endif
#ifndef VP_TILE_CCAP
#define VP_TILE_CCAP 192
#endif
constexpr int TILE_HIT_CAP = VP_TILE_HIT_CAP;
constexpr int TILE_QCAP = VP_TILE_QCAP;
constexpr int TILE_CCAP = VP_TILE_CCAP;
constexpr size_t TILE_SMEM = (size_t)TILE_HIT_CAP * TRACE_THREADS * 8 + (size_t)(TRACE_THREADS / 32) * (TILE_QCAP + TILE_CCAP) * 4;
static_assert(TILE_QCAP >= TILE_CCAP, "the idle node queue holds two of the three candidate buckets");
constexpr int TILE_FALLBACK_CAP = TILE_HIT_CAP;  // the per-ray fallback may only touch the warp's OWN list columns
#define VP_INF __int_as_float(0x7f800000)
#ifdef VP_DEBUG_CHECKS   // bounds checks for debugging builds (compute-sanitizer is not available on the GPU pool)
#include <cstdio>
#define VP_CHECK(cond, code, a, b) do { if (!(cond)) printf("VP_CHECK %d failed: %d %d (block %d thread %d)\n", code, (int)(a), (int)(b), blockIdx.x, threadIdx.x); } while (0)
#else
#define VP_CHECK(cond, code, a, b) do { } while (0)
#endif

struct Isect {
    bool valid;
    float tn, tf;
    float3 ro, rd;  // R^T (o - c), R^T d  (fixed evaluation order)
};

// ray_ellipsoid_intersection (reference common.py:346-367, RT-Gems-2 branch) followed by
// mi.math.improved_solve_quadratic.  Individually rounded operations: must match oracle/volprim_oracle.c
// ray_ellipsoid() bit for bit, because these distances order the hits and drive the epsilon cull.
__device__ __forceinline__ Isect exact_isect(float3 o, float3 d, float4 g0, float4 g1, const Mat3 &R, float extent)
{
    Isect r;
    float3 sc = make_float3(__fmul_rn(g1.x, extent), __fmul_rn(g1.y, extent), __fmul_rn(g1.z, extent));
    float3 v = make_float3(__fsub_rn(o.x, g0.x), __fsub_rn(o.y, g0.y), __fsub_rn(o.z, g0.z));
    r.rd = vp_rot_t_mul_rn(R, d);
    r.ro = vp_rot_t_mul_rn(R, v);
#if VP_DIV_SHARED
    // six divisions over three denominators, then three more: shared reciprocals, same bits (vp_internal.cuh)
    const VpDivisor dx = vp_divisor(sc.x), dy = vp_divisor(sc.y), dz = vp_divisor(sc.z);
    float3 dd = make_float3(vp_div_rn(r.rd.x, dx), vp_div_rn(r.rd.y, dy), vp_div_rn(r.rd.z, dz));
    float3 oo = make_float3(vp_div_rn(r.ro.x, dx), vp_div_rn(r.ro.y, dy), vp_div_rn(r.ro.z, dz));
#else
    float3 dd = make_float3(__fdiv_rn(r.rd.x, sc.x), __fdiv_rn(r.rd.y, sc.y), __fdiv_rn(r.rd.z, sc.z));
    float3 oo = make_float3(__fdiv_rn(r.ro.x, sc.x), __fdiv_rn(r.ro.y, sc.y), __fdiv_rn(r.ro.z, sc.z));
#endif
    float a = vp_dot_rn(dd, dd);
    float b = -vp_dot_rn(oo, dd);
    float c = __fsub_rn(vp_dot_rn(oo, oo), 1.f);
#if VP_DIV_SHARED
    const VpDivisor da = vp_divisor(a);
    float ba = vp_div_rn(b, da);
#else
    float ba = __fdiv_rn(b, a);
#endif
    float3 l = make_float3(__fmaf_rn(ba, dd.x, oo.x), __fmaf_rn(ba, dd.y, oo.y), __fmaf_rn(ba, dd.z, oo.z));
    float discr = __fsub_rn(1.f, vp_dot_rn(l, l));
    r.valid = false;
    r.tn = r.tf = 0.f;
    if (!(discr >= 0.f) || a == 0.f) return r;
    float sq = __fsqrt_rn(__fmul_rn(a, discr));
    float q = __fadd_rn(b, copysignf(sq, b));
#if VP_DIV_SHARED
    float x0 = vp_div_rn(c, vp_divisor(q));
    float x1 = vp_div_rn(q, da);
#else
    float x0 = __fdiv_rn(c, q);
    float x1 = __fdiv_rn(q, a);
#endif
    r.tn = fminf(x0, x1);
    r.tf = fmaxf(x0, x1);
    r.valid = isfinite(x0) && isfinite(x1);
    return r;
}

// Work counters.  `hits` is the ray's depth and is taken from there at the end; `overflow` (traversal-stack overflow:
// must stay 0) and `retries` are touched on rare paths only.  The three per-step work counters (candidates, nodes,
// passes) are compiled in with -DVP_FULL_STATS (profiling builds): three registers kept alive through the whole walk
// are not free in a kernel that runs at its register cap.
struct Counters {
    uint32_t hits, overflow, retries;
#ifdef VP_FULL_STATS
    uint32_t candidates, nodes, passes;
#endif
};
#ifdef VP_FULL_STATS
#define VP_COUNT(x) x
#else
#define VP_COUNT(x) do { } while (0)
#endif

// Approximate entry distance from the pre-transformed record xf (rows of M = diag(1/(extent s)) R^T with the
// centre in .w).  Only used to ORDER candidates and to bin them into intervals; validity is conservative
// (slightly negative discriminants pass) because the drain phase repeats the test exactly.
// `o` is the walker origin advanced to the start of the current interval (o0 + t_base d) and the result is shifted back
// by t_base: seen from the interval start a primitive is tens, not thousands, of its own radii away, so the entry
// distance resolves the gaps the parity contract asks the hit order to resolve (from 4 units away a grazing entry on a
// 0.0015-unit primitive was off by > 1e-4 and swapped places with its neighbour).
__device__ __forceinline__ bool fast_isect(const DevScene &S, int pos, float3 o, float3 d, float t_base, float &tn)
{
    VP_CHECK(pos >= 0 && pos < S.n, 1, pos, S.n);
    const float4 *x = S.xf + 3ll * pos;
    float4 r0 = __ldg(x), r1 = __ldg(x + 1), r2 = __ldg(x + 2);
    float3 v = make_float3(o.x - r0.w, o.y - r1.w, o.z - r2.w);
    float3 oo = make_float3(r0.x * v.x + r0.y * v.y + r0.z * v.z, r1.x * v.x + r1.y * v.y + r1.z * v.z,
                            r2.x * v.x + r2.y * v.y + r2.z * v.z);
    float3 dd = make_float3(r0.x * d.x + r0.y * d.y + r0.z * d.z, r1.x * d.x + r1.y * d.y + r1.z * d.z,
                            r2.x * d.x + r2.y * d.y + r2.z * d.z);
    float a = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
    float b = -(oo.x * dd.x + oo.y * dd.y + oo.z * dd.z);
    float c = oo.x * oo.x + oo.y * oo.y + oo.z * oo.z - 1.f;
    const float ia = vp_rcp(a);
    float ba = b * ia;
    float lx = fmaf(ba, dd.x, oo.x), ly = fmaf(ba, dd.y, oo.y), lz = fmaf(ba, dd.z, oo.z);
    float discr = 1.f - (lx * lx + ly * ly + lz * lz);
    // Conservative validity: the interval origin sits on the world-coordinate grid, up to 2^-24 |o| beside the ray, and
    // the unit-sphere transform of a 0.0015-unit primitive magnifies that 700 times (3e-4 in the discriminant); |l|^2 is
    // also a small difference of terms of size |o'|^2.  A grazing hit must never be lost here -- false candidates only
    // cost one exact test in the drain (0.5 % of the candidates have |discr| < 1e-2).
    if (!(discr >= -(1e-2f + 2e-6f * (fabsf(oo.x) + fabsf(oo.y) + fabsf(oo.z)))) || !(a > 0.f)) return false;
#if VP_APPROX_ORDER
    const float ad = a * fmaxf(discr, 0.f);
    float sq = ad * vp_rsqrt(fmaxf(ad, 1e-37f));      // 2 ulp: the ordering key needs 1e-6, the exact test decides
#else
    float sq = sqrtf(a * fmaxf(discr, 0.f));
#endif
    float q = b + copysignf(sq, b);
    float x0 = c * vp_rcp(q), x1 = q * ia;
    tn = fminf(x0, x1) + t_base;
    return tn == tn;
}

// ray / box slab test against the interval [t_lo, t_hi]
__device__ __forceinline__ bool slab(float3 lo, float3 hi, float3 oi, float3 inv, float t_lo, float t_hi)
{
    // oi = o * inv, so each plane distance is a single FMA
    float ax = fmaf(lo.x, inv.x, -oi.x), bx = fmaf(hi.x, inv.x, -oi.x);
    float ay = fmaf(lo.y, inv.y, -oi.y), by = fmaf(hi.y, inv.y, -oi.y);
    float az = fmaf(lo.z, inv.z, -oi.z), bz = fmaf(hi.z, inv.z, -oi.z);
    float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), t_lo));
    float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fminf(fmaxf(az, bz), t_hi));
    // boxes are padded at build time; widen the exit by a few ulp for the reciprocal's rounding
    return t0 <= t1 * 1.000001f + 1e-30f;
}

// phase 3 of both walkers: hand the listed entries to on_hit in increasing distance.  Each entry is RE-EVALUATED
// against the current, re-based origin `o` with the oracle's fixed-order arithmetic (quirk Q1, see file header).
// L1 prefetch of everything the evaluation of primitive `pos` will read (3 geometry records + the SH block):
// issued one entry ahead so that the next hit's gathers overlap the current hit's arithmetic.
__device__ __forceinline__ void prefetch_prim(const DevScene &S, int pos)
{
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.geo0 + pos));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.geo1 + pos));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(S.geo2 + pos));
    if (S.sh_stride4 > 0) {
        const float4 *f = S.sh4 + (size_t)pos * S.sh_stride4;
        asm volatile("prefetch.global.L1 [%0];" ::"l"(f));
        if (S.sh_stride4 > 8) asm volatile("prefetch.global.L1 [%0];" ::"l"(f + 8));
    }
}

// phase 3 of both walkers hands the listed entries to on_hit in increasing distance; each entry is RE-EVALUATED
// against the current, re-based origin `o` with the oracle's fixed-order arithmetic (quirk Q1, see file header).
// sorted insert of (t, pos) into a per-lane list kept in increasing t (insertion sort: the lists are short)
template <int STRIDE>
__device__ __forceinline__ void list_insert(int *s_id, float *s_t, int n, float tn, int pos)
{
    VP_CHECK(n >= 0 && n < 64, 4, n, pos);
    int j = n;
    while (j > 0) {
        const float tp = s_t[(j - 1) * STRIDE];
        if (!(tp > tn)) break;
        s_t[j * STRIDE] = tp;
        s_id[j * STRIDE] = s_id[(j - 1) * STRIDE];
        --j;
    }
    s_t[j * STRIDE] = tn;
    s_id[j * STRIDE] = pos;
}

// same for the tile walker's lists: one 64-bit key (distance bits << 32 | position) per entry, so an insertion moves
// one shared-memory word pair instead of two separate arrays (positive floats order like their bit patterns)
template <int STRIDE>
__device__ __forceinline__ void list_insert_key(unsigned long long *s_key, int n, float tn, int pos)
{
    const unsigned long long key = ((unsigned long long)(tn > 0.f ? __float_as_uint(tn) : 0u) << 32) | (unsigned)pos;
    int j = n;
    while (j > 0) {
        const unsigned long long kp = s_key[(j - 1) * STRIDE];
        if (!(kp > key)) break;
        s_key[j * STRIDE] = kp;
        --j;
    }
    s_key[j * STRIDE] = key;
}

// An on_hit callable may carry a preload(pos) hook: it is called as soon as the entry's position is known, before the
// exact intersection, so that loads the hit will need (the head of its SH block) are in flight during those ~200
// instructions instead of being issued after them.
template <class F, class P>
struct HitWithPreload {
    F f;
    P p;
    template <class... Args>
    __device__ __forceinline__ bool operator()(Args &&...args) { return f(static_cast<Args &&>(args)...); }
    __device__ __forceinline__ void preload(int pos) { p(pos); }
};
template <class T, class = void>
struct has_preload { static constexpr bool value = false; };
template <class T>
struct has_preload<T, decltype(void(&T::preload))> { static constexpr bool value = true; };

template <class GetPos, class OnHit>
__device__ __forceinline__ void drain_list(const DevScene &S, int n_found, GetPos &&get_pos, const float3 &o,
                                           const float3 d, float maxt, bool &alive, bool &missed, OnHit &&on_hit)
{
    VP_CHECK(n_found >= 0 && n_found <= 64, 5, n_found, 0);
    // (no prefetch of entry k + 1 here: with ~200 KB of the SM's 256 KB given to shared memory the 768 resident
    // threads' prefetched lines evict each other from L1 before they are used -- measured +2 % time; the replay
    // adjoint, which uses no shared memory, keeps its software pipeline)
    for (int k = 0; alive && k < n_found; ++k) {
        const int pos = get_pos(k);
        VP_CHECK(pos >= 0 && pos < S.n, 6, pos, k);
        float4 g0 = __ldg(S.geo0 + pos), g1 = __ldg(S.geo1 + pos), g2 = __ldg(S.geo2 + pos);
        if constexpr (has_preload<typename std::remove_reference<OnHit>::type>::value) on_hit.preload(pos);
        Mat3 Rm = vp_quat_to_matrix_rn(g2);
        Isect is = exact_isect(o, d, g0, g1, Rm, S.extent);
        if (!is.valid || !(is.tn > 0.f)) continue;      // entry fell behind the advanced origin (Q1)
        if (!(is.tn <= maxt)) { missed = true; alive = false; break; }
        if (!on_hit(pos, g0, g1, g2, Rm, is)) { alive = false; break; }
    }
}

// Walks one ray front to back and calls on_hit(pos, g0, g1, g2, R, isect) for every accepted entry, in the
// reference's order.  on_hit evaluates the primitive, advances the origin `o` (captured by the caller) and
// returns false to terminate the ray.  WARP-CONVERGENT: all 32 lanes call it (lanes without a ray pass
// alive = false).  s_id / s_t are this thread's columns of the shared candidate list (stride TRACE_THREADS).
template <int CAP = CAND_CAP, int STRIDE = TRACE_THREADS, class OnHit>
__device__ __forceinline__ void walk_ray(const DevScene &S, int *s_id, float *s_t, const float3 &o, const float3 o0,
                                         const float3 d, const float maxt, bool alive, bool &missed, Counters &cn,
                                         OnHit &&on_hit, const float t_begin = 0.f)
{
    missed = false;
    if (S.n <= 0) { missed = alive; return; }
    float3 inv;
    inv.x = 1.f / (fabsf(d.x) > 1e-30f ? d.x : copysignf(1e-30f, d.x));
    inv.y = 1.f / (fabsf(d.y) > 1e-30f ? d.y : copysignf(1e-30f, d.y));
    inv.z = 1.f / (fabsf(d.z) > 1e-30f ? d.z : copysignf(1e-30f, d.z));
    const float3 oi = make_float3(o0.x * inv.x, o0.y * inv.y, o0.z * inv.z);
    // all interval distances are measured from the ORIGINAL origin o0; `o` is the re-based origin
    float t_start = 0.f, t_stop = VP_INF;
    const float delta0 = __ldg(S.info + 6);
    if (S.root >= 0) {
        const float3 lo = make_float3(__ldg(S.info), __ldg(S.info + 1), __ldg(S.info + 2));
        const float3 hi = make_float3(__ldg(S.info + 3), __ldg(S.info + 4), __ldg(S.info + 5));
        float ax = (lo.x - o0.x) * inv.x, bx = (hi.x - o0.x) * inv.x;
        float ay = (lo.y - o0.y) * inv.y, by = (hi.y - o0.y) * inv.y;
        float az = (lo.z - o0.z) * inv.z, bz = (hi.z - o0.z) * inv.z;
        t_start = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
        t_stop = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.000001f;
        if (alive && !(t_start <= t_stop)) { missed = true; alive = false; }
        t_start = fmaxf(fmaxf(t_start * 0.999999f - 1e-6f, 0.f), t_begin);
    }
    float delta = delta0;
    const float delta_min = delta0 * (1.f / 4096.f);
    bool closest_mode = false, strict_lo = false;
    int closest_left = 0;
    int stack[STACK_MAX];
    while (__any_sync(0xffffffffu, alive)) {
        // Entries are re-listed from a little before t_start (the ordering test's rounding error grows with t).  An
        // interval shorter than that look-back would list the same stale entries again and again and crawl (seen with a
        // dense cluster 2000 units from the camera), hence the floor on the interval width.
        const float slack = 1e-4f + 1e-5f * t_start;
        const float t_lo = strict_lo ? t_start : t_start - slack;
        const float delta_floor = fmaxf(delta_min, 2.f * slack);
        // origin of the ordering tests of this interval (see fast_isect)
#ifdef VP_ORDER_FAR
        const float t_base = 0.f;
        const float3 ob = o0;
#else
        const float t_base = fmaxf(t_lo, 0.f);
        const float3 ob = make_float3(fmaf(d.x, t_base, o0.x), fmaf(d.y, t_base, o0.y), fmaf(d.z, t_base, o0.z));
#endif
        // (the max keeps the interval from collapsing to nothing when delta is below one ulp of a large t_start)
        float t_end = (S.root >= 0 && !closest_mode) ? fmaxf(t_start + delta, t_start * 1.000001f) : VP_INF;
        int n_c = 0;
        bool overflow = false;
        float best_t = VP_INF;
        int best_pos = -1;
        VP_COUNT(if (alive) cn.passes++);
        // ---- phase 1: collect the leaves whose boxes meet [t_lo, t_end] ----
        // The loop body is branch-free per lane (predicated appends / push / pop): lanes differ only in their
        // trip count.  Leaves are appended from their parent and never become the current node.
        int node = (alive && !closest_mode && S.root >= 0) ? S.root : NODE_SENTINEL;
        int sp = 0;
        if (alive && S.root < 0) { s_id[0] = ~S.root; n_c = 1; }       // single-primitive scene
        while (node != NODE_SENTINEL) {
            VP_CHECK(node >= 0 && node < S.n - 1, 7, node, sp);
            const float4 *nd = S.nodes + 4ll * node;
            float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3);
            const int top = stack[sp > 0 ? sp - 1 : 0];                   // speculative: used only on a pop
            VP_COUNT(cn.nodes++);
            const bool hl = slab(make_float3(n0.x, n0.y, n0.z), make_float3(n0.w, n1.x, n1.y), oi, inv, t_lo, t_end);
            const bool hr = slab(make_float3(n1.z, n1.w, n2.x), make_float3(n2.y, n2.z, n2.w), oi, inv, t_lo, t_end);
            const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
#ifdef VP_PREFETCH
            if (left >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.nodes + 4ll * left));
            if (right >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(S.nodes + 4ll * right));
#endif
            if (hl && left < 0) {
                if (n_c < CAP) s_id[n_c * STRIDE] = ~left;
                ++n_c;
            }
            if (hr && right < 0) {
                if (n_c < CAP) s_id[n_c * STRIDE] = ~right;
                ++n_c;
            }
            const bool vl = hl && left >= 0, vr = hr && right >= 0;
            if (vl && vr) {
                if (sp < STACK_MAX) stack[sp++] = right;
                else cn.overflow++;
                node = left;
            } else {
                const bool pop = !vl && !vr;
                node = vl ? left : (vr ? right : (sp > 0 ? top : NODE_SENTINEL));
                sp -= (pop && sp > 0) ? 1 : 0;
            }
            if (n_c > CAP) node = NODE_SENTINEL;                     // list overflow: abandon the interval
        }
        if (n_c > CAP) { overflow = true; n_c = 0; }
        // closest-hit walk (fallback for lanes whose interval cannot be listed; rare)
        if (__any_sync(0xffffffffu, alive && closest_mode)) {
            int cnode = (alive && closest_mode) ? S.root : NODE_SENTINEL;
            sp = 0;
            while (cnode != NODE_SENTINEL) {
                if (cnode < 0) {
                    float tn;
                    VP_COUNT(cn.candidates++);
                    if (fast_isect(S, ~cnode, ob, d, t_base, tn) && tn > t_lo && tn < best_t) {
                        best_t = tn;
                        best_pos = ~cnode;
                        t_end = tn;  // nothing beyond the closest entry is needed
                    }
                    cnode = sp > 0 ? stack[--sp] : NODE_SENTINEL;
                    continue;
                }
                const float4 *nd = S.nodes + 4ll * cnode;
                float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3);
                VP_COUNT(cn.nodes++);
                const bool hl = slab(make_float3(n0.x, n0.y, n0.z), make_float3(n0.w, n1.x, n1.y), oi, inv, t_lo, t_end);
                const bool hr = slab(make_float3(n1.z, n1.w, n2.x), make_float3(n2.y, n2.z, n2.w), oi, inv, t_lo, t_end);
                const int left = __float_as_int(n3.x), right = __float_as_int(n3.y);
                if (hl && hr) {
                    cnode = left;
                    if (sp < STACK_MAX) stack[sp++] = right;
                } else if (hl) cnode = left;
                else if (hr) cnode = right;
                else cnode = sp > 0 ? stack[--sp] : NODE_SENTINEL;
            }
        }
        __syncwarp();
        // ---- phase 2: entry distances; keep the entries inside the interval ----
        int n_h = 0, n_new = 0;   // n_new: entries beyond t_start (the look-back ones do not count for the width)
        if (closest_mode) {
            if (best_pos >= 0) { s_id[0] = best_pos; s_t[0] = best_t; n_h = 1; }
        } else if (!overflow) {
            for (int k = 0; k < n_c; ++k) {
                const int pos = s_id[k * STRIDE];
                float tn;
                bool ok;
                if (S.root >= 0) ok = fast_isect(S, pos, ob, d, t_base, tn);
                else { ok = true; tn = 1.f; }
                if (ok && tn > t_lo && tn <= t_end) {
                    list_insert<STRIDE>(s_id, s_t, n_h, tn, pos);   // n_h <= k: entry k is already read
                    ++n_h;
                    n_new += tn > t_start ? 1 : 0;
                }
            }
            VP_COUNT(cn.candidates += n_c);
        }
        __syncwarp();
        // ---- phase 3: drain in increasing distance ----
        const int n_found = n_h;
        drain_list(S, n_found, [&](int k) { return s_id[k * STRIDE]; }, o, d, maxt, alive, missed, on_hit);
        // ---- next interval ----
        if (alive) {
            if (S.root < 0) { missed = true; alive = false; }
            else if (overflow) {
                delta *= 0.5f;
                if (delta < delta_floor) { closest_mode = true; closest_left = 8; }
            } else if (closest_mode) {
                if (best_pos < 0) { missed = true; alive = false; }   // nothing in front of t_start at all
                else {
                    t_start = best_t;      // strictly beyond the entry just handled
                    strict_lo = true;
                    if (--closest_left <= 0) { closest_mode = false; delta = delta_floor * 8.f; }
                }
            } else {
                strict_lo = false;
                t_start = t_end;
                delta = fmaxf(delta * ((n_new == 0) ? 4.f : fminf(fmaxf((float)TARGET_HITS / (float)n_new, 0.5f), 2.f)), delta_floor);
            }
            if (alive && t_start > t_stop) { missed = true; alive = false; }
        }
    }
}

// Capsule around the living lanes' ray segments [t0, t1]: mean segment A -> B and the largest deviation r of any
// lane's segment end points from it (distance between two linear motions is convex, so the ends bound the whole
// segment).  Boxes inflated by r and tested against the axis segment (parameter in [0, 1]) are a conservative cull.
struct Capsule {
    float3 AI, invD;
    float r;
};
__device__ __forceinline__ Capsule tile_capsule(bool alive, unsigned am, float3 o0, float3 d, float t0, float t1)
{
    constexpr unsigned FULL = 0xffffffffu;
    const float inv_n = vp_rcp((float)__popc(am));
    float3 P0 = make_float3(fmaf(d.x, t0, o0.x), fmaf(d.y, t0, o0.y), fmaf(d.z, t0, o0.z));
    float3 P1 = make_float3(fmaf(d.x, t1, o0.x), fmaf(d.y, t1, o0.y), fmaf(d.z, t1, o0.z));
    float3 A = alive ? P0 : make_float3(0.f, 0.f, 0.f), B = alive ? P1 : make_float3(0.f, 0.f, 0.f);
    for (int off = 16; off; off >>= 1) {
        A.x += __shfl_xor_sync(FULL, A.x, off); A.y += __shfl_xor_sync(FULL, A.y, off); A.z += __shfl_xor_sync(FULL, A.z, off);
        B.x += __shfl_xor_sync(FULL, B.x, off); B.y += __shfl_xor_sync(FULL, B.y, off); B.z += __shfl_xor_sync(FULL, B.z, off);
    }
    A.x *= inv_n; A.y *= inv_n; A.z *= inv_n; B.x *= inv_n; B.y *= inv_n; B.z *= inv_n;
    float r = 0.f;
    if (alive) {
        float e0 = (P0.x - A.x) * (P0.x - A.x) + (P0.y - A.y) * (P0.y - A.y) + (P0.z - A.z) * (P0.z - A.z);
        float e1 = (P1.x - B.x) * (P1.x - B.x) + (P1.y - B.y) * (P1.y - B.y) + (P1.z - B.z) * (P1.z - B.z);
        const float e = fmaxf(e0, e1);
        r = e * vp_rsqrt(fmaxf(e, 1e-37f));      // (the 1.0001 margin below covers the approximation)
    }
    for (int off = 16; off; off >>= 1) r = fmaxf(r, __shfl_xor_sync(FULL, r, off));
    Capsule c;
    c.r = r * 1.0001f + 1e-6f * (1.f + fabsf(A.x) + fabsf(A.y) + fabsf(A.z));
    float3 Dx = make_float3(B.x - A.x, B.y - A.y, B.z - A.z);
    c.invD.x = vp_rcp(fabsf(Dx.x) > 1e-30f ? Dx.x : copysignf(1e-30f, Dx.x));
    c.invD.y = vp_rcp(fabsf(Dx.y) > 1e-30f ? Dx.y : copysignf(1e-30f, Dx.y));
    c.invD.z = vp_rcp(fabsf(Dx.z) > 1e-30f ? Dx.z : copysignf(1e-30f, Dx.z));
    c.AI = make_float3(A.x * c.invD.x, A.y * c.invD.y, A.z * c.invD.z);
    return c;
}

// Tighter bound of the same ray segments for the per-candidate cull (phase 1.5 of the tile walker): the mean segment
// A + lambda D plus a BOX of deviations in the tile's own frame (u along the tile's pixel rows, w along the mean
// direction, v = w x u) -- an 8x4 pixel tile is twice as wide as it is high, the capsule's circle wastes half its area.
struct TilePrism {
    float3 A, D, u, v, w;
    float ra, rb, rw;
};
__device__ __forceinline__ TilePrism tile_prism(bool alive, unsigned am, float3 o0, float3 d, float t0, float t1)
{
    constexpr unsigned FULL = 0xffffffffu;
    const float inv_n = vp_rcp((float)__popc(am));
    const float3 P0 = make_float3(fmaf(d.x, t0, o0.x), fmaf(d.y, t0, o0.y), fmaf(d.z, t0, o0.z));
    const float3 P1 = make_float3(fmaf(d.x, t1, o0.x), fmaf(d.y, t1, o0.y), fmaf(d.z, t1, o0.z));
    float3 A = alive ? P0 : make_float3(0.f, 0.f, 0.f), B = alive ? P1 : make_float3(0.f, 0.f, 0.f);
    for (int off = 16; off; off >>= 1) {
        A.x += __shfl_xor_sync(FULL, A.x, off); A.y += __shfl_xor_sync(FULL, A.y, off); A.z += __shfl_xor_sync(FULL, A.z, off);
        B.x += __shfl_xor_sync(FULL, B.x, off); B.y += __shfl_xor_sync(FULL, B.y, off); B.z += __shfl_xor_sync(FULL, B.z, off);
    }
    A.x *= inv_n; A.y *= inv_n; A.z *= inv_n; B.x *= inv_n; B.y *= inv_n; B.z *= inv_n;
    TilePrism p;
    p.A = A;
    p.D = make_float3(B.x - A.x, B.y - A.y, B.z - A.z);
    const float dl = vp_rsqrt(fmaxf(p.D.x * p.D.x + p.D.y * p.D.y + p.D.z * p.D.z, 1e-30f));
    p.w = make_float3(p.D.x * dl, p.D.y * dl, p.D.z * dl);
    // u: the direction from the tile's first to its last pixel of a row (lanes 0 and 7), made orthogonal to w.  ANY
    // orthonormal frame keeps the bound valid; this one makes it tight for image tiles.
    float3 ur = make_float3(__shfl_sync(FULL, P1.x, 7) - __shfl_sync(FULL, P1.x, 0),
                            __shfl_sync(FULL, P1.y, 7) - __shfl_sync(FULL, P1.y, 0),
                            __shfl_sync(FULL, P1.z, 7) - __shfl_sync(FULL, P1.z, 0));
    float uw = ur.x * p.w.x + ur.y * p.w.y + ur.z * p.w.z;
    ur = make_float3(ur.x - uw * p.w.x, ur.y - uw * p.w.y, ur.z - uw * p.w.z);
    float ul = ur.x * ur.x + ur.y * ur.y + ur.z * ur.z;
    if (!(ul > 1e-20f)) {   // degenerate tile: any direction perpendicular to w
        ur = fabsf(p.w.x) < 0.6f ? make_float3(0.f, -p.w.z, p.w.y) : make_float3(-p.w.z, 0.f, p.w.x);
        ul = ur.x * ur.x + ur.y * ur.y + ur.z * ur.z;
    }
    const float uil = vp_rsqrt(ul);
    p.u = make_float3(ur.x * uil, ur.y * uil, ur.z * uil);
    p.v = make_float3(p.w.y * p.u.z - p.w.z * p.u.y, p.w.z * p.u.x - p.w.x * p.u.z, p.w.x * p.u.y - p.w.y * p.u.x);
    float ra = 0.f, rb = 0.f, rw = 0.f;
    if (alive) {
        const float3 e0 = make_float3(P0.x - A.x, P0.y - A.y, P0.z - A.z), e1 = make_float3(P1.x - B.x, P1.y - B.y, P1.z - B.z);
        ra = fmaxf(fabsf(e0.x * p.u.x + e0.y * p.u.y + e0.z * p.u.z), fabsf(e1.x * p.u.x + e1.y * p.u.y + e1.z * p.u.z));
        rb = fmaxf(fabsf(e0.x * p.v.x + e0.y * p.v.y + e0.z * p.v.z), fabsf(e1.x * p.v.x + e1.y * p.v.y + e1.z * p.v.z));
        rw = fmaxf(fabsf(e0.x * p.w.x + e0.y * p.w.y + e0.z * p.w.z), fabsf(e1.x * p.w.x + e1.y * p.w.y + e1.z * p.w.z));
    }
    for (int off = 16; off; off >>= 1) {
        ra = fmaxf(ra, __shfl_xor_sync(FULL, ra, off));
        rb = fmaxf(rb, __shfl_xor_sync(FULL, rb, off));
        rw = fmaxf(rw, __shfl_xor_sync(FULL, rw, off));
    }
    const float slack = 1e-6f * (1.f + fabsf(A.x) + fabsf(A.y) + fabsf(A.z));
    p.ra = ra * 1.001f + slack; p.rb = rb * 1.001f + slack; p.rw = rw * 1.001f + slack;
    return p;
}

// Can ANY ray of the tile touch the primitive's bounding ellipsoid?  In the primitive's unit-sphere space (x' = M (x -
// c), the record fast_isect uses) the mean line passes the origin at distance |P|, P = A' - (A'.D^)D^.  Along e = P/|P|
// a point of the prism lies at e.x' = |P| + g.q with g = M^T e and q the deviation, |g.q| <= ra|g.u| + rb|g.v| + rw|g.w|.
// If that cannot come down to 1 no ray of the tile reaches the ellipsoid.  Conservative: never rejects a real hit.
__device__ __forceinline__ bool prism_may_hit(const DevScene &S, int pos, const TilePrism &p, float &lam)
{
    lam = 0.f;   // where the mean line enters the ellipsoid (or passes closest), as a fraction of the segment
    const float4 *x = S.xf + 3ll * pos;
    const float4 r0 = __ldg(x), r1 = __ldg(x + 1), r2 = __ldg(x + 2);
    const float3 a = make_float3(p.A.x - r0.w, p.A.y - r1.w, p.A.z - r2.w);
    const float3 Ap = make_float3(r0.x * a.x + r0.y * a.y + r0.z * a.z, r1.x * a.x + r1.y * a.y + r1.z * a.z,
                                  r2.x * a.x + r2.y * a.y + r2.z * a.z);
    const float3 Dp = make_float3(r0.x * p.D.x + r0.y * p.D.y + r0.z * p.D.z, r1.x * p.D.x + r1.y * p.D.y + r1.z * p.D.z,
                                  r2.x * p.D.x + r2.y * p.D.y + r2.z * p.D.z);
    const float dd = Dp.x * Dp.x + Dp.y * Dp.y + Dp.z * Dp.z;
    if (!(dd > 1e-30f)) return true;
#if VP_APPROX_ORDER
    // approximate reciprocal / root: 2 ulp on k is covered by `err` below, lam only sorts candidates into buckets
    const float idd = vp_rcp(dd);
    const float k = (Ap.x * Dp.x + Ap.y * Dp.y + Ap.z * Dp.z) * idd;
#else
    const float k = (Ap.x * Dp.x + Ap.y * Dp.y + Ap.z * Dp.z) / dd;
#endif
    const float3 P = make_float3(Ap.x - k * Dp.x, Ap.y - k * Dp.y, Ap.z - k * Dp.z);
    const float pl2 = P.x * P.x + P.y * P.y + P.z * P.z;
    lam = -k;
    if (!(pl2 > 1.f)) {                     // the mean line itself passes through the bounding ellipsoid
#if VP_APPROX_ORDER
        const float x = (1.f - pl2) * idd;
        lam -= x * vp_rsqrt(fmaxf(x, 1e-37f));
#else
        lam -= sqrtf((1.f - pl2) / dd);
#endif
        return true;
    }
    const float ipl = vp_rsqrt(pl2), pl = pl2 * ipl;
    const float3 e = make_float3(P.x * ipl, P.y * ipl, P.z * ipl);
    const float3 g = make_float3(r0.x * e.x + r1.x * e.y + r2.x * e.z, r0.y * e.x + r1.y * e.y + r2.y * e.z,
                                 r0.z * e.x + r1.z * e.y + r2.z * e.z);
    const float h = p.ra * fabsf(g.x * p.u.x + g.y * p.u.y + g.z * p.u.z) + p.rb * fabsf(g.x * p.v.x + g.y * p.v.y + g.z * p.v.z)
                  + p.rw * fabsf(g.x * p.w.x + g.y * p.w.y + g.z * p.w.z);
    // rounding of P = A' - k D' grows with |A'| (long intervals over tiny primitives): widen the margin accordingly
    const float err = 2e-6f * (fabsf(Ap.x) + fabsf(Ap.y) + fabsf(Ap.z));
    return !(pl > 1.002f + 1.002f * h + err);
}

// child boxes of one node against a capsule
__device__ __forceinline__ void capsule_children(const DevScene &S, int node, const Capsule &c, bool &hl, bool &hr,
                                                 int &left, int &right)
{
    VP_CHECK(node >= 0 && node < S.n - 1, 2, node, S.n);
    const float4 *nd = S.nodes + 4ll * node;
    float4 n0 = __ldg(nd), n1 = __ldg(nd + 1), n2 = __ldg(nd + 2), n3 = __ldg(nd + 3);
    const float r = c.r;
    hl = slab(make_float3(n0.x - r, n0.y - r, n0.z - r), make_float3(n0.w + r, n1.x + r, n1.y + r), c.AI, c.invD, 0.f, 1.f);
    hr = slab(make_float3(n1.z - r, n1.w - r, n2.x - r), make_float3(n2.y + r, n2.z + r, n2.w + r), c.AI, c.invD, 0.f, 1.f);
    left = __float_as_int(n3.x);
    right = __float_as_int(n3.y);
}

// Warp-cooperative walker for COHERENT rays (the 32 lanes of a warp are an 8x4 pixel tile).
//
// The per-ray walker spends three quarters of its time in ~600 DEPENDENT node visits per ray.  Here the warp walks
// the tree ONCE per interval for the whole tile: the 32 ray segments of the interval are bounded by a capsule
// (mean segment + largest deviation r), every step lets the 32 lanes test 32 DIFFERENT queued nodes against that
// capsule (child boxes inflated by r), and ballots compact the surviving children into the shared queue / the
// tile's candidate list.  Each lane then tests the tile's candidates against its own ray (warp-uniform, broadcast
// loads) and drains its own hits exactly like the per-ray walker, so results are identical.
// Interval bookkeeping is warp-uniform.  Pathological overlap (lists overflow below the minimum interval width)
// hands the warp over to the per-ray walker.
template <class OnHit>
__device__ __forceinline__ void walk_tile(const DevScene &S, int *smem, const float3 &o, const float3 o0,
                                          const float3 d, const float maxt, bool alive, bool &missed, Counters &cn,
                                          OnHit &&on_hit)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // shared memory: [hit ids CAP x 128][hit t CAP x 128][per warp: node queue, tile candidates].  The per-ray
    // fallback runs per WARP while the other warps of the block keep walking, so it re-uses exactly this warp's
    // list columns (a wider layout would overwrite the neighbours' queues -- found the hard way)
    unsigned long long *s_key = reinterpret_cast<unsigned long long *>(smem) + threadIdx.x;
    // the per-ray fallback keeps (id, t) in the two halves of THIS thread's 64-bit slots (row stride 256 words), so
    // that it never touches a slot of another thread -- the other warps of the block keep using theirs
    int *s_id = smem + 2 * threadIdx.x;
    float *s_t = reinterpret_cast<float *>(smem) + 2 * threadIdx.x + 1;
    int *w_queue = smem + 2 * TILE_HIT_CAP * TRACE_THREADS + (threadIdx.x >> 5) * (TILE_QCAP + TILE_CCAP);
    int *w_cand = w_queue + TILE_QCAP;
    int *fb_id = s_id;
    float *fb_t = s_t;
    const unsigned lt = (1u << lane) - 1u;
    missed = false;
    if (S.n <= 0) { missed = alive; return; }
    // The per-ray walker takes over (from distance t_hand) when the tile walk returns true; it is instantiated at this
    // one place only -- the kernel is issue-bound and every inlined copy costs registers and instruction cache.
    float t_hand = 0.f;
    auto tile_part = [&]() -> bool {
    if (S.root < 0) return true;   // single primitive: nothing to share
    const float delta0 = __ldg(S.info + 6);
    const float delta_min = delta0 * (1.f / 4096.f);
    float t_in = VP_INF, t_out = -VP_INF;
    {
        float3 inv;
        inv.x = 1.f / (fabsf(d.x) > 1e-30f ? d.x : copysignf(1e-30f, d.x));
        inv.y = 1.f / (fabsf(d.y) > 1e-30f ? d.y : copysignf(1e-30f, d.y));
        inv.z = 1.f / (fabsf(d.z) > 1e-30f ? d.z : copysignf(1e-30f, d.z));
        const float3 lo = make_float3(__ldg(S.info), __ldg(S.info + 1), __ldg(S.info + 2));
        const float3 hi = make_float3(__ldg(S.info + 3), __ldg(S.info + 4), __ldg(S.info + 5));
        float ax = (lo.x - o0.x) * inv.x, bx = (hi.x - o0.x) * inv.x;
        float ay = (lo.y - o0.y) * inv.y, by = (hi.y - o0.y) * inv.y;
        float az = (lo.z - o0.z) * inv.z, bz = (hi.z - o0.z) * inv.z;
        t_in = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
        t_out = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz)) * 1.000001f;
        if (alive && !(t_in <= t_out)) { missed = true; alive = false; }
        t_in = fmaxf(t_in * 0.999999f - 1e-6f, 0.f);
    }
    // warp-uniform interval start: where the first living ray enters the scene box
    float t_start = alive ? t_in : VP_INF;
    for (int off = 16; off; off >>= 1) t_start = fminf(t_start, __shfl_xor_sync(FULL, t_start, off));
    float delta = delta0;
    // ---- tile root set: the subtrees the tile's rays can reach at all (capsule over the whole remaining range).
    // Every interval starts its walk from these <= 32 nodes (one per lane, in a register) instead of re-descending
    // from the root with a nearly empty queue. ----
    int rs_node = lane == 0 ? S.root : -1, rs_n = 1;
    {
        const unsigned am0 = __ballot_sync(FULL, alive);
        if (!am0) return false;
        float t_far = alive ? t_out : -VP_INF;
        for (int off = 16; off; off >>= 1) t_far = fmaxf(t_far, __shfl_xor_sync(FULL, t_far, off));
        // A tile much wider than the primitives (coarse images, incoherent "tiles") would list several times more
        // candidates per lane than a ray of its own needs: cost grows like (1 + r / h)^2 with r the capsule radius
        // and h the mean leaf half-extent.  Beyond r = 2 h (9 x the candidates) every lane walks for itself.
        {
            const Capsule c0 = tile_capsule(alive, am0, o0, d, t_start, t_start + delta0);
            if (c0.r > 2.f * __ldg(S.info + 8)) {
                return true;
            }
        }
        const Capsule cf = tile_capsule(alive, am0, o0, d, fmaxf(t_start - 1e-3f, 0.f), t_far);
        for (int iter = 0; iter < 12 && rs_n <= 16; ++iter) {
            bool hl = false, hr = false;
            int left = 0, right = 0;
            if (rs_node >= 0) capsule_children(S, rs_node, cf, hl, hr, left, right);
            const bool leaf_hit = (hl && left < 0) || (hr && right < 0);
            const bool expand = rs_node >= 0 && !leaf_hit;     // a node with a reachable leaf child stays as it is
            if (!__any_sync(FULL, expand)) break;
            const int c0 = expand ? (hl ? left : (hr ? right : -1)) : rs_node;
            const int c1 = (expand && hl && hr) ? right : -1;
            const unsigned m0 = __ballot_sync(FULL, c0 >= 0), m1 = __ballot_sync(FULL, c1 >= 0);
            const int n0 = __popc(m0), n1 = __popc(m1);
            if (n0 + n1 > 32) break;
            __syncwarp();
            if (c0 >= 0) w_queue[__popc(m0 & lt)] = c0;
            if (c1 >= 0) w_queue[n0 + __popc(m1 & lt)] = c1;
            __syncwarp();
            rs_n = n0 + n1;
            rs_node = lane < rs_n ? w_queue[lane] : -1;
            __syncwarp();
        }
    }
    while (true) {
        const unsigned am = __ballot_sync(FULL, alive);
        if (!am) break;
        const float t_lo = t_start - (1e-4f + 1e-5f * t_start);   // look-back and width floor: see walk_ray
        const float t_end = fmaxf(t_start + delta, t_start * 1.000001f);   // progress even below one ulp of t_start
        // origin of the ordering tests of this interval (see fast_isect)
#ifdef VP_ORDER_FAR
        const float t_base = 0.f;
        const float3 ob = o0;
#else
        const float t_base = fmaxf(t_lo, 0.f);
        const float3 ob = make_float3(fmaf(d.x, t_base, o0.x), fmaf(d.y, t_base, o0.y), fmaf(d.z, t_base, o0.z));
#endif
        VP_COUNT(if (alive) cn.passes++);
        const Capsule cap = tile_capsule(alive, am, o0, d, t_lo, t_end);
        // ---- phase 1 (cooperative): 32 queued nodes per step against the interval's capsule ----
        int qn = rs_n, tcn = 0;
        bool overflow = false;
        if (lane < rs_n) w_queue[lane] = rs_node;
        __syncwarp();
        while (qn > 0) {
            const int take = qn < 32 ? qn : 32;
            const int base = qn - take;
            const int node = lane < take ? w_queue[base + lane] : -1;
            qn = base;
            __syncwarp();
            bool iL = false, iR = false, lL = false, lR = false;
            int left = 0, right = 0;
            if (node >= 0) {
                bool hl, hr;
                capsule_children(S, node, cap, hl, hr, left, right);
                VP_COUNT(cn.nodes++);
                iL = hl && left >= 0; lL = hl && left < 0;
                iR = hr && right >= 0; lR = hr && right < 0;
            }
            const unsigned mIL = __ballot_sync(FULL, iL), mIR = __ballot_sync(FULL, iR);
            const unsigned mLL = __ballot_sync(FULL, lL), mLR = __ballot_sync(FULL, lR);
            const int nIL = __popc(mIL), nI = nIL + __popc(mIR), nLL = __popc(mLL), nL = nLL + __popc(mLR);
            if (qn + nI > TILE_QCAP || tcn + nL > TILE_CCAP) { overflow = true; break; }
            VP_CHECK(qn >= 0 && qn + nI <= TILE_QCAP && tcn + nL <= TILE_CCAP, 3, qn + nI, tcn + nL);
            if (iL) w_queue[qn + __popc(mIL & lt)] = left;
            if (iR) w_queue[qn + nIL + __popc(mIR & lt)] = right;
            if (lL) w_cand[tcn + __popc(mLL & lt)] = ~left;
            if (lR) w_cand[tcn + nLL + __popc(mLR & lt)] = ~right;
            qn += nI;
            tcn += nL;
            __syncwarp();
        }
        __syncwarp();
        // ---- phase 1.5: lane-parallel cull of the candidates no ray of the tile can touch (more than half of what
        // the box-vs-capsule walk lists: rotated anisotropic ellipsoids fill little of their boxes).  One candidate
        // per lane, compacted in place; costs ~3 issue slots per candidate, phase 2 costs ~50 per surviving one. ----
        int n_b0 = tcn, n_b1 = 0, n_b2 = 0;
        if (!overflow && tcn > 0) {
            const TilePrism pr = tile_prism(alive, am, o0, d, t_lo, t_end);
            // The survivors are split into three buckets by where the tile's mean line meets them (near third of the
            // interval: compacted in place; middle: front of the now idle node queue; far: back of it) and phase 2
            // visits the buckets in that order, so the per-lane sorted insertions mostly land near the list tail.
            int kept = 0;
            n_b1 = 0; n_b2 = 0;
            for (int k0 = 0; k0 < tcn; k0 += 32) {
                const int k = k0 + lane;
                const int pos = k < tcn ? w_cand[k] : -1;
                float lam;
                const bool keep = pos >= 0 && prism_may_hit(S, pos, pr, lam);
                const bool b0 = keep && lam < (1.f / 3.f), b2 = keep && lam >= (2.f / 3.f), b1 = keep && !b0 && !b2;
                const unsigned m0 = __ballot_sync(FULL, b0), m1 = __ballot_sync(FULL, b1), m2 = __ballot_sync(FULL, b2);
                if (b0) w_cand[kept + __popc(m0 & lt)] = pos;   // kept <= k0: never ahead of the reads
                if (b1) w_queue[n_b1 + __popc(m1 & lt)] = pos;
                if (b2) w_queue[TILE_QCAP - 1 - (n_b2 + __popc(m2 & lt))] = pos;
                kept += __popc(m0); n_b1 += __popc(m1); n_b2 += __popc(m2);
                __syncwarp();
            }
            n_b0 = kept;
            tcn = kept + n_b1 + n_b2;
        }
        // ---- phase 2: every lane tests the tile's candidates against its own ray ----
        // (warp-uniform loop, broadcast loads; four candidates per trip so that their loads overlap)
        int n_h = 0;
        bool dropped = false;   // this lane saw more entries than its list holds and kept the closest ones
        if (!overflow) {
            for (int k0 = 0; k0 < tcn; k0 += 4) {
                int pos[4];
                float tn[4];
                bool ok[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int idx = k0 + u < tcn ? k0 + u : k0;
                    const int *src = idx < n_b0 ? w_cand + idx
                                                : (idx < n_b0 + n_b1 ? w_queue + (idx - n_b0)
                                                                     : w_queue + (TILE_QCAP - 1 - (idx - n_b0 - n_b1)));
                    pos[u] = *src;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) ok[u] = fast_isect(S, pos[u], ob, d, t_base, tn[u]);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    bool take = alive && k0 + u < tcn && ok[u] && tn[u] > t_lo && tn[u] <= t_end;
                    if (take && n_h == TILE_HIT_CAP) {
                        // full: the list keeps its TILE_HIT_CAP closest entries (the farthest one falls out)
                        dropped = true;
                        const unsigned last_t = (unsigned)(s_key[(TILE_HIT_CAP - 1) * TRACE_THREADS] >> 32);
                        if (__float_as_uint(tn[u]) < last_t) n_h = TILE_HIT_CAP - 1;
                        else take = false;
                    }
                    if (take) {
                        list_insert_key<TRACE_THREADS>(s_key, n_h, tn[u], pos[u]);
                        ++n_h;
                    }
                }
            }
            VP_COUNT(if (alive) cn.candidates += tcn);
        }
        if (overflow) {   // the tile's candidate list or the node queue did not fit: shorter interval, walk again
            if (alive) cn.retries++;   // (statistics)
            delta *= VP_OVF_SHRINK;
            if (delta < fmaxf(delta_min, 2.f * (t_start - t_lo))) {   // cannot be listed: the per-ray walker has the closest-hit fallback
                t_hand = t_start;
                return true;
            }
            continue;
        }
        // A lane that dropped entries is complete only up to its farthest kept entry.  Nothing is walked again: the
        // WHOLE tile ends the interval there (the collected candidates cover it), every lane keeps its entries up to
        // that distance, and the next interval starts from it.
        float t_cut = dropped ? __uint_as_float((unsigned)(s_key[(TILE_HIT_CAP - 1) * TRACE_THREADS] >> 32)) : VP_INF;
        for (int off = 16; off; off >>= 1) t_cut = fminf(t_cut, __shfl_xor_sync(FULL, t_cut, off));
        float t_done = t_end;
        if (t_cut < t_end && !(t_cut - t_start > delta_min)) {
            // more entries than a list holds within the smallest interval (heavy overlap): no progress possible here,
            // the per-ray walker's closest-hit search takes over from the current position
            t_hand = t_start;
            return true;
        }
        if (t_cut < t_end) {
            t_done = t_cut;
            const unsigned cut_bits = __float_as_uint(t_cut);
            while (n_h > 0 && (unsigned)(s_key[(n_h - 1) * TRACE_THREADS] >> 32) > cut_bits) --n_h;
        }
        // ---- phase 3: drain (per lane) ----
        const int found_max = (int)__reduce_max_sync(FULL, (unsigned)n_h);
        drain_list(S, n_h, [&](int k) { return (int)(unsigned)(s_key[k * TRACE_THREADS] & 0xffffffffull); }, o, d, maxt, alive,
                   missed, on_hit);
        // ---- next interval (warp-uniform) ----
        // the interval grows or shrinks so that the fullest of the lists it has to fit -- the busiest lane's hit list
        // and the tile's candidate list -- lands at VP_FILL of its capacity; after a cut it restarts from the width
        // that actually fitted
        const float delta_floor = fmaxf(delta_min, 2e-4f + 2e-5f * t_start);   // twice the look-back
        if (t_done < t_end) delta = fmaxf((t_done - t_start) * 0.75f, delta_floor);
        else {
            const float fill_h = (float)found_max * (1.f / (VP_FILL * TILE_HIT_CAP));
            const float fill_c = (float)tcn * (1.f / (float)(VP_CAND_TARGET));
            const float f_h = (fill_h < 0.15f) ? VP_GROW_MAX : fminf(fmaxf(vp_rcp(fill_h), 0.5f), 2.f);
            // the candidate count only shrinks the interval while that can help: boxes that contain the whole
            // neighbourhood (nested primitives) stay candidates however short the interval gets
            const float f_c = fmaxf(vp_rcp(fmaxf(fill_c, 0.25f)), delta > delta0 * (1.f / VP_CSHRINK_FLOOR) ? 0.5f : 1.f);
            delta = fmaxf(delta * fminf(f_h, f_c), delta_floor);
        }
        t_start = t_done;
        if (alive && t_start > t_out) { missed = true; alive = false; }
    }
    return false;
    };
    if (tile_part()) {
        bool m2 = false;   // lanes that missed the scene box keep their `missed`
        __syncwarp();
        walk_ray<TILE_FALLBACK_CAP, 2 * TRACE_THREADS>(S, fb_id, fb_t, o, o0, d, maxt, alive, m2, cn, on_hit, t_hand);
        missed = missed || m2;
    }
}

// Origin the walkers measure their interval distances from.  Ordinarily the ray origin itself; for a camera far outside
// the scene (more than 8 box diagonals away) a point just before the scene box, so that the fp32 resolution of the
// ordering distances is that of the scene, not of the camera distance.  Warp-uniform shift for the tile walker (its
// intervals are shared by the 32 lanes).  Only ordering / culling use it: hits are accepted by the exact test against
// the ray's true, re-based origin.
template <bool TILE>
__device__ __forceinline__ float3 walker_origin(const DevScene &S, float3 o, float3 d, bool alive)
{
    if (S.n <= 0 || S.root < 0) return o;
    const float3 lo = make_float3(__ldg(S.info), __ldg(S.info + 1), __ldg(S.info + 2));
    const float3 hi = make_float3(__ldg(S.info + 3), __ldg(S.info + 4), __ldg(S.info + 5));
    const float ix = 1.f / (fabsf(d.x) > 1e-30f ? d.x : copysignf(1e-30f, d.x));
    const float iy = 1.f / (fabsf(d.y) > 1e-30f ? d.y : copysignf(1e-30f, d.y));
    const float iz = 1.f / (fabsf(d.z) > 1e-30f ? d.z : copysignf(1e-30f, d.z));
    const float ax = (lo.x - o.x) * ix, bx = (hi.x - o.x) * ix, ay = (lo.y - o.y) * iy, by = (hi.y - o.y) * iy;
    const float az = (lo.z - o.z) * iz, bz = (hi.z - o.z) * iz;
    float t_in = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
    const float t_ex = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    if (!alive || !(t_in <= t_ex)) t_in = VP_INF;                // misses the box: no say in the shift
    if (TILE) for (int off = 16; off; off >>= 1) t_in = fminf(t_in, __shfl_xor_sync(0xffffffffu, t_in, off));
    const float dlen = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
    const float diag = sqrtf((hi.x - lo.x) * (hi.x - lo.x) + (hi.y - lo.y) * (hi.y - lo.y) + (hi.z - lo.z) * (hi.z - lo.z));
    if (!(t_in < VP_INF) || !(t_in * dlen > 8.f * diag)) return o;
    const float shift = t_in - diag / fmaxf(dlen, 1e-30f);      // one diagonal before the box
    return make_float3(fmaf(d.x, shift, o.x), fmaf(d.y, shift, o.y), fmaf(d.z, shift, o.z));
}

// ---- closed-form primitive evaluation --------------------------------------------------------

template <int D>
__device__ __forceinline__ void sh_basis(float3 d, float (&Y)[(D + 1) * (D + 1)])
{
    // dr.sh_eval: Sloan's real-SH recurrences with Condon-Shortley sign, index l(l+1)+m (rf:90)
    float x = d.x, y = d.y, z = d.z;
    Y[0] = 0.28209479177387814f;
    if constexpr (D >= 1) {
        Y[2] = 0.48860251190291992f * z;
        Y[3] = -0.48860251190291992f * x;
        Y[1] = -0.48860251190291992f * y;
    }
    if constexpr (D >= 2) {
        float z2 = z * z;
        Y[6] = 0.94617469575756008f * z2 - 0.31539156525251999f;
        float tb = -1.0925484305920792f * z;
        Y[7] = tb * x;
        Y[5] = tb * y;
        float c1 = x * x - y * y, s1 = 2.f * x * y;
        Y[8] = 0.54627421529603959f * c1;
        Y[4] = 0.54627421529603959f * s1;
        if constexpr (D >= 3) {
            Y[12] = z * (1.8658816629505769f * z2 - 1.1195289977703462f);
            float tc = -2.2852289973223288f * z2 + 0.45704579946446572f;
            Y[13] = tc * x;
            Y[11] = tc * y;
            float td = 1.4453057213202769f * z;
            Y[14] = td * c1;
            Y[10] = td * s1;
            float c2 = x * c1 - y * s1, s2 = x * s1 + y * c1;
            Y[15] = -0.59004358992664352f * c2;
            Y[9] = -0.59004358992664352f * s2;
        }
    }
}

// 256-bit read-only global load (LDG.E.256 on sm_100a); p must be 32-byte aligned
__device__ __forceinline__ void ldg256(const float4 *p, float4 &a, float4 &b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
        : "l"(p));
}

// eval_sh_emission (rf:82-100): raw = sum_i Y_i f_i + 0.5 ; col = max(raw, 0).  128-bit loads of the
// primitive's contiguous coefficient block.
// how many leading float4 of the SH block the forward drain fetches before the exact intersection (VP_SH_HOIST)
#ifndef VP_SH_HOIST
#define VP_SH_HOIST 2
#endif
template <int D>
struct ShHead {
    static constexpr int N4 = (3 * (D + 1) * (D + 1) + 3) / 4;
    static constexpr int PRE = (N4 % 2 == 0 && VP_SH_HOIST <= N4) ? VP_SH_HOIST : 0;     // even block sizes only (256-bit loads)
    float4 v[PRE > 0 ? PRE : 1];
    __device__ __forceinline__ void load(const DevScene &S, int pos)
    {
        if constexpr (PRE > 0) {
            const float4 *f = S.sh4 + (size_t)pos * N4;
#pragma unroll
            for (int k = 0; k < PRE; k += 2)
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=f"(v[k].x), "=f"(v[k].y), "=f"(v[k].z), "=f"(v[k].w), "=f"(v[k + 1].x), "=f"(v[k + 1].y),
                               "=f"(v[k + 1].z), "=f"(v[k + 1].w)
                             : "l"(f + k));
        }
    }
};

template <int D>
__device__ __forceinline__ void sh_color(const DevScene &S, int pos, const float (&Y)[(D + 1) * (D + 1)], float (&raw)[3],
                                         const ShHead<D> *head = nullptr)
{
    constexpr int NB = (D + 1) * (D + 1);
    constexpr int C = 3 * NB;
    constexpr int N4 = (C + 3) / 4;
    const float4 *f = S.sh4 + (size_t)pos * N4;
    float acc[3] = { 0.f, 0.f, 0.f };
    float4 v[N4];
    if constexpr (ShHead<D>::PRE > 0) {
        if (head) {
            constexpr int PRE = ShHead<D>::PRE;
#pragma unroll
            for (int k = PRE; k < N4; k += 2) ldg256(f + k, v[k], v[k + 1]);
#pragma unroll
            for (int k = 0; k < PRE; ++k) v[k] = head->v[k];
#pragma unroll
            for (int k = 0; k < N4; ++k) {
                const float e[4] = { v[k].x, v[k].y, v[k].z, v[k].w };
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int idx = 4 * k + c;
                    if (idx < C) acc[idx % 3] = fmaf(Y[idx / 3], e[c], acc[idx % 3]);
                }
            }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) raw[ch] = acc[ch] + 0.5f;
            return;
        }
    }
#ifdef VP_SH_LDG256
    // sm_100 has 256-bit global loads: the coefficient block of a primitive (192 B at degree 3, 32-byte aligned when
    // N4 is even) takes half as many load instructions -- and half as many L1 wavefronts, the co-limiter of the drain
    if constexpr (N4 % 2 == 0) {
#pragma unroll
        for (int k = 0; k < N4; k += 2) ldg256(f + k, v[k], v[k + 1]);
    } else
#endif
    {
#pragma unroll
        for (int k = 0; k < N4; ++k) v[k] = __ldg(f + k);
    }
#pragma unroll
    for (int k = 0; k < N4; ++k) {
        const float e[4] = { v[k].x, v[k].y, v[k].z, v[k].w };
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int idx = 4 * k + c;
            if (idx < C) acc[idx % 3] = fmaf(Y[idx / 3], e[c], acc[idx % 3]);
        }
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) raw[ch] = acc[ch] + 0.5f;
}

struct RfEval {
    float T, G, op;
    float3 pp;  // p_peak
};

// eval_transmission (rf:63-80) with GaussianKernel.eval / EpanechnikovKernel.eval (common.py:153-159, 251-259)
template <int KERNEL>
__device__ __forceinline__ RfEval rf_eval(float3 o, float3 d, float4 g0, float4 g1, const Mat3 &R, const Isect &is)
{
    RfEval e;
    // (value path: reciprocal scales instead of nine IEEE divisions; agrees with the oracle to a few ulp)
    const float isx = 1.f / g1.x, isy = 1.f / g1.y, isz = 1.f / g1.z;
    float3 oo = make_float3(is.ro.x * isx, is.ro.y * isy, is.ro.z * isz);
    float3 dd = make_float3(is.rd.x * isx, is.rd.y * isy, is.rd.z * isz);
    float od = oo.x * dd.x + oo.y * dd.y + oo.z * dd.z, dd2 = dd.x * dd.x + dd.y * dd.y + dd.z * dd.z;
    float tp = -od / dd2;
    e.pp = make_float3(fmaf(d.x, tp, o.x), fmaf(d.y, tp, o.y), fmaf(d.z, tp, o.z));
    float3 v = make_float3(e.pp.x - g0.x, e.pp.y - g0.y, e.pp.z - g0.z);
    float3 w = vp_rot_t_mul_rn(R, v);
    if (KERNEL == VP_KERNEL_GAUSSIAN) {
        float ux = w.x * isx, uy = w.y * isy, uz = w.z * isz;
        float q = ux * ux + uy * uy + uz * uz;
        e.G = expf(-0.5f * q);
    } else {
        float ux = w.x * isx * (1.f / 3.f), uy = w.y * isy * (1.f / 3.f), uz = w.z * isz * (1.f / 3.f);
        float dist = sqrtf(ux * ux + uy * uy + uz * uz);
        e.G = fmaxf(0.75f * (1.f - dist * dist), 0.f);
    }
    e.op = g0.w;
    float a = e.op * e.G;
    if (!(a < 0.9999f)) a = 0.9999f;
    e.T = 1.f - a;
    return e;
}

// The primal pass needs only T of eval_transmission (rf:63-80).  With u = R^T (p_peak - c) / s = o' + t_peak d' (o', d'
// the ray in the primitive's scaled frame, which exact_isect has already rotated) the kernel value is a function of
// |u|^2 alone: approximate reciprocals and ex2 are enough here (errors ~1e-6 of T; nothing orders hits).
template <int KERNEL>
__device__ __forceinline__ float rf_transmittance(float4 g0, float4 g1, const Isect &is)
{
    const float isx = vp_rcp(g1.x), isy = vp_rcp(g1.y), isz = vp_rcp(g1.z);
    const float3 oo = make_float3(is.ro.x * isx, is.ro.y * isy, is.ro.z * isz);
    const float3 dd = make_float3(is.rd.x * isx, is.rd.y * isy, is.rd.z * isz);
    const float od = fmaf(oo.x, dd.x, fmaf(oo.y, dd.y, oo.z * dd.z)), dd2 = fmaf(dd.x, dd.x, fmaf(dd.y, dd.y, dd.z * dd.z));
    const float tp = -od * vp_rcp(dd2);
    const float ux = fmaf(tp, dd.x, oo.x), uy = fmaf(tp, dd.y, oo.y), uz = fmaf(tp, dd.z, oo.z);
    const float q = fmaf(ux, ux, fmaf(uy, uy, uz * uz));
    float G;
    if (KERNEL == VP_KERNEL_GAUSSIAN) G = vp_exp(-0.5f * q);
    else G = fmaxf(0.75f * (1.f - q * (1.f / 9.f)), 0.f);
    float a = g0.w * G;
    if (!(a < 0.9999f)) a = 0.9999f;
    return 1.f - a;
}

// quat_to_matrix / R^T v with fused multiply-adds (value paths of the adjoint; nothing here decides hit order)
__device__ __forceinline__ Mat3 quat_to_matrix_fast(float4 q)
{
    const float x = q.x, y = q.y, z = q.z, w = q.w;
    Mat3 R;
    R.m[0][0] = 1.f - 2.f * fmaf(y, y, z * z);
    R.m[0][1] = 2.f * fmaf(x, y, -z * w);
    R.m[0][2] = 2.f * fmaf(x, z, y * w);
    R.m[1][0] = 2.f * fmaf(x, y, z * w);
    R.m[1][1] = 1.f - 2.f * fmaf(x, x, z * z);
    R.m[1][2] = 2.f * fmaf(y, z, -x * w);
    R.m[2][0] = 2.f * fmaf(x, z, -y * w);
    R.m[2][1] = 2.f * fmaf(y, z, x * w);
    R.m[2][2] = 1.f - 2.f * fmaf(x, x, y * y);
    return R;
}
__device__ __forceinline__ float3 rot_t_mul_fast(const Mat3 &R, float3 v)
{
    return make_float3(fmaf(R.m[2][0], v.z, fmaf(R.m[1][0], v.y, R.m[0][0] * v.x)),
                       fmaf(R.m[2][1], v.z, fmaf(R.m[1][1], v.y, R.m[0][1] * v.x)),
                       fmaf(R.m[2][2], v.z, fmaf(R.m[1][2], v.y, R.m[0][2] * v.x)));
}

// eval_transmission from an origin that may be FAR from the primitive (the gather adjoint evaluates every hit from the
// ray's original origin: no per-hit origin has to be carried from the ray-major to the primitive-major pass).  From
// |o - c| / s ~ 1e3 the peak parameter t* = -(o'.d') / (d'.d') loses its low bits and p* slides ALONG the ray by up to
// ~1e-6 -- invisible in G (stationary in t) but first order in the gradient terms.  So the origin is first moved to the
// coarse peak (rounded to the world-coordinate grid, exactly like the reference's own re-based origins) and the peak is
// solved again from there, where |o - c| / s <= extent and fp32 is accurate.
template <int KERNEL>
__device__ __forceinline__ RfEval rf_eval_far(float3 o, float3 d, float4 g0, float4 g1, const Mat3 &R)
{
    RfEval e;
    // approximate reciprocals / ex2 (2 ulp): a value path, nothing here orders hits
    const float isx = vp_rcp(g1.x), isy = vp_rcp(g1.y), isz = vp_rcp(g1.z);
    const float3 rd = rot_t_mul_fast(R, d);
    const float3 dd = make_float3(rd.x * isx, rd.y * isy, rd.z * isz);
    const float inv_dd2 = vp_rcp(fmaf(dd.x, dd.x, fmaf(dd.y, dd.y, dd.z * dd.z)));
    float3 ro = rot_t_mul_fast(R, make_float3(o.x - g0.x, o.y - g0.y, o.z - g0.z));
    const float t0 = -fmaf(ro.x * isx, dd.x, fmaf(ro.y * isy, dd.y, ro.z * isz * dd.z)) * inv_dd2;
    const float3 o1 = make_float3(fmaf(d.x, t0, o.x), fmaf(d.y, t0, o.y), fmaf(d.z, t0, o.z));
    ro = rot_t_mul_fast(R, make_float3(o1.x - g0.x, o1.y - g0.y, o1.z - g0.z));
    const float t1 = -fmaf(ro.x * isx, dd.x, fmaf(ro.y * isy, dd.y, ro.z * isz * dd.z)) * inv_dd2;
    e.pp = make_float3(fmaf(d.x, t1, o1.x), fmaf(d.y, t1, o1.y), fmaf(d.z, t1, o1.z));
    const float3 w = rot_t_mul_fast(R, make_float3(e.pp.x - g0.x, e.pp.y - g0.y, e.pp.z - g0.z));
    if (KERNEL == VP_KERNEL_GAUSSIAN) {
        const float ux = w.x * isx, uy = w.y * isy, uz = w.z * isz;
        e.G = vp_exp(-0.5f * fmaf(ux, ux, fmaf(uy, uy, uz * uz)));
    } else {
        const float ux = w.x * isx * (1.f / 3.f), uy = w.y * isy * (1.f / 3.f), uz = w.z * isz * (1.f / 3.f);
        e.G = fmaxf(0.75f * (1.f - fmaf(ux, ux, fmaf(uy, uy, uz * uz))), 0.f);
    }
    e.op = g0.w;
    float a = e.op * e.G;
    if (!(a < 0.9999f)) a = 0.9999f;
    e.T = 1.f - a;
    return e;
}

// GaussianKernel.density_integral full range (common.py:199-206, 238-243); returns clamped rho
__device__ __forceinline__ float gauss_density_integral(float4 g1, const Isect &is)
{
    float3 w = is.rd, p = is.ro;
    // The literal expression cancels catastrophically in fp32 when |o - c| >> s; it is therefore evaluated
    // with individually rounded operations in exactly the oracle's association order (bit-identical inputs
    // to exp), instead of letting the compiler contract it differently from the CPU restatement.
    float sx2 = __fmul_rn(g1.x, g1.x), sy2 = __fmul_rn(g1.y, g1.y), sz2 = __fmul_rn(g1.z, g1.z);
    float wx2 = __fmul_rn(w.x, w.x), wy2 = __fmul_rn(w.y, w.y), wz2 = __fmul_rn(w.z, w.z);
    float px2 = __fmul_rn(p.x, p.x), py2 = __fmul_rn(p.y, p.y), pz2 = __fmul_rn(p.z, p.z);
    float C1 = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(sx2, sy2), wz2), __fmul_rn(__fmul_rn(sx2, sz2), wy2)),
                         __fmul_rn(__fmul_rn(sy2, sz2), wx2));
    float t1 = __fmul_rn(__fadd_rn(__fmul_rn(px2, sy2), __fmul_rn(py2, sx2)), wz2);
    float t2 = __fmul_rn(__fmul_rn(__fmul_rn(2.f, p.z), w.z),
                         __fadd_rn(__fmul_rn(__fmul_rn(p.y, sx2), w.y), __fmul_rn(__fmul_rn(p.x, sy2), w.x)));
    float t3 = __fmul_rn(wy2, __fadd_rn(__fmul_rn(px2, sz2), __fmul_rn(pz2, sx2)));
    float t4 = __fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.f, p.x), p.y), sz2), w.x), w.y);
    float t5 = __fmul_rn(wx2, __fadd_rn(__fmul_rn(py2, sz2), __fmul_rn(pz2, sy2)));
    float num = __fadd_rn(__fsub_rn(__fadd_rn(__fsub_rn(t1, t2), t3), t4), t5);
    float exponent = __fdiv_rn(num, __fmul_rn(2.f, C1));
    float denom = __fmul_rn(6.2831855f, __fsqrt_rn(C1));
    float density = __fdiv_rn(expf(-exponent), denom);
    if (!(density > 0.f) || !isfinite(density)) density = 0.f;
    return density;
}

// EpanechnikovKernel.density_integral full range (common.py:287-324); bandwidth s over the extent*s chord
__device__ __forceinline__ float epan_density_integral(float3 o, float3 d, float4 g0, float4 g1, const Mat3 &R,
                                                       const Isect &is)
{
    if (!is.valid || !(is.tn < is.tf) || !(is.tf > 0.f)) return 0.f;
    float3 a0 = make_float3(fmaf(d.x, is.tn, o.x) - g0.x, fmaf(d.y, is.tn, o.y) - g0.y, fmaf(d.z, is.tn, o.z) - g0.z);
    float3 a1 = make_float3(fmaf(d.x, is.tf, o.x) - g0.x, fmaf(d.y, is.tf, o.y) - g0.y, fmaf(d.z, is.tf, o.z) - g0.z);
    float3 p = vp_rot_t_mul_rn(R, a0), p1 = vp_rot_t_mul_rn(R, a1);
    float3 w = make_float3(p1.x - p.x, p1.y - p.y, p1.z - p.z);
    float t = sqrtf(w.x * w.x + w.y * w.y + w.z * w.z);
    w.x /= t; w.y /= t; w.z /= t;
    float sx2 = g1.x * g1.x, sy2 = g1.y * g1.y, sz2 = g1.z * g1.z;
    float t2 = t * t, t3 = t2 * t;
    float poly = sx2 * sy2 * t3 * (w.z * w.z) + 3.f * p.z * sx2 * sy2 * t2 * w.z + sx2 * sz2 * t3 * (w.y * w.y)
               + 3.f * p.y * sx2 * sz2 * t2 * w.y + sy2 * sz2 * t3 * (w.x * w.x) + 3.f * p.x * sy2 * sz2 * t2 * w.x
               + (((3.f * (p.x * p.x) - 3.f * sx2) * sy2 + 3.f * (p.y * p.y) * sx2) * sz2
                  + 3.f * (p.z * p.z) * sx2 * sy2) * t;
    float s3 = (g1.x * g1.x * g1.x) * (g1.y * g1.y * g1.y) * (g1.z * g1.z * g1.z);
    float density = -poly * 5.f / (8.f * 3.14159265358979f * s3);
    if (!(density > 0.f) || !isfinite(density)) density = 0.f;
    return density;
}

__device__ __forceinline__ float srgb_to_linear(float x)
{
    return x <= 0.04045f ? x / 12.92f : powf((x + 0.055f) / 1.055f, 2.4f);
}

// sampler.next_1d() of Mitsuba's `independent` sampler, random access (third-party, restated from the published
// algorithms; identical to oracle/volprim_oracle.c pcg32_float_at): PCG32 stream per ray, seeded with
// sample_tea_32(seed, ray index).  No per-ray generator state is carried through the walk: Russian roulette draws
// a handful of samples per ray, each recomputed from the seed (one 64-bit multiply-add per skipped sample).
__device__ __forceinline__ float pcg32_float_at(uint32_t seed, uint32_t idx, uint32_t n)
{
    uint32_t v0 = seed, v1 = idx, sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    const unsigned long long mult = 0x5851f42d4c957f2dull;
    const unsigned long long inc = ((unsigned long long)v1 << 1) | 1ull;
    unsigned long long state = inc;          // state = 0 * mult + inc
    state += (unsigned long long)v0;
    state = state * mult + inc;
    for (uint32_t i = 0; i < n; ++i) state = state * mult + inc;
    const uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    const uint32_t rot = (uint32_t)(state >> 59);
    const uint32_t out = (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
    return __uint_as_float((out >> 9) | 0x3f800000u) - 1.f;
}

// thread -> ray index; with an image hint the 32 lanes of a warp cover an 8x4 pixel tile
__device__ __forceinline__ int64_t ray_index(int64_t t, int W, int H)
{
    if (W <= 0) return t;
    int64_t per = (int64_t)W * H;
    int64_t img = t / per;
    int rem = (int)(t - img * per);
    int tile = rem >> 5, lane = rem & 31;
    int tiles_x = W >> 3;
    int ty = tile / tiles_x, tx = tile - ty * tiles_x;
    int x = tx * 8 + (lane & 7), y = ty * 4 + (lane >> 3);
    return img * per + (int64_t)y * W + x;
}

__device__ __forceinline__ void flush_counters(const Counters &cn, vp_stats *st)
{
    if (!st) return;
#ifdef VP_FULL_STATS
    uint32_t v[6] = { cn.hits, cn.candidates, cn.nodes, cn.passes, cn.overflow, cn.retries };
#else
    uint32_t v[6] = { cn.hits, 0u, 0u, 0u, cn.overflow, cn.retries };
#endif
#pragma unroll
    for (int k = 0; k < 6; ++k)
        for (int off = 16; off; off >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], off);
    if ((threadIdx.x & 31) == 0) {
        atomicAdd((unsigned long long *)&st->hits, (unsigned long long)v[0]);
#ifdef VP_FULL_STATS
        atomicAdd((unsigned long long *)&st->candidates, (unsigned long long)v[1]);
        atomicAdd((unsigned long long *)&st->node_visits, (unsigned long long)v[2]);
        atomicAdd((unsigned long long *)&st->passes, (unsigned long long)v[3]);
#endif
        if (v[4]) atomicAdd((unsigned long long *)&st->stack_overflows, (unsigned long long)v[4]);
        if (v[5]) atomicAdd((unsigned long long *)&st->interval_retries, (unsigned long long)v[5]);
    }
}

// Where a kernel takes its rays from: explicit arrays, or a perspective sensor evaluated in the kernel (fused
// Sensor.sample_ray; volprim/cameras.py:114-137 configures the Mitsuba `perspective` plugin this restates).
struct RaySrc {
    const float *o, *d, *maxt, *jitter;
    vp_camera cam;
    // computed once on the host in double:
    float tan_half;        // tan(fov_x / 2)
    float inv_w, inv_h;    // 1 / width, 1 / height
    float ly_scale;        // tan_half / aspect
    int32_t has_cam, spp;
    int64_t index_base;    // ray i of the call is sample (index_base + i) of the sensor (row bands, record bands)
};

// One ray of a perspective sensor: sample index gi = pixel * spp + sample, pixel row-major.  ONE routine for the
// stand-alone ray generation kernel and every fused path, so that all of them see bit-identical rays; a sample index
// below 2^31 (always, for a film) keeps the index arithmetic in 32 bits, and there is no IEEE division on the way --
// the primitive-major pass of the gather adjoint regenerates a ray per bucket entry.
__device__ __forceinline__ void camera_ray(const RaySrc &S, int64_t i, float3 &o, float3 &d, float &maxt)
{
    const vp_camera &cam = S.cam;
    const int64_t gi = S.index_base + i;
    int x, y;
    if (gi < 0x7fffffffll) {
        const uint32_t pix = (uint32_t)gi / (uint32_t)S.spp;
        y = (int)(pix / (uint32_t)cam.width);
        x = (int)(pix - (uint32_t)y * (uint32_t)cam.width);
    } else {
        const int64_t pix = gi / S.spp;
        y = (int)(pix / cam.width);
        x = (int)(pix - (int64_t)y * cam.width);
    }
    float jx = 0.5f, jy = 0.5f;
    if (S.jitter) { jx = S.jitter[2 * i]; jy = S.jitter[2 * i + 1]; }
    const float u = __fmul_rn(__fadd_rn((float)x, jx), S.inv_w);
    const float v = __fmul_rn(__fadd_rn((float)y, jy), S.inv_h);
    // Mitsuba perspective sensor: +z forward, +x to the LEFT of the image, +y up
    const float lx = __fmul_rn(__fadd_rn(__fsub_rn(1.f, __fmul_rn(2.f, u)), __fmul_rn(2.f, cam.cx)), S.tan_half);
    const float ly = __fmul_rn(__fadd_rn(__fsub_rn(1.f, __fmul_rn(2.f, v)), __fmul_rn(2.f, cam.cy)), S.ly_scale);
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(lx, lx), __fmul_rn(ly, ly)), 1.f);
    float inv_n = rsqrtf(n2);
    inv_n = __fmul_rn(inv_n, __fsub_rn(1.5f, __fmul_rn(__fmul_rn(0.5f, n2), __fmul_rn(inv_n, inv_n))));   // one Newton step
    const float3 dl = make_float3(__fmul_rn(lx, inv_n), __fmul_rn(ly, inv_n), inv_n);
    const float *m = cam.to_world;
    d = make_float3(__fadd_rn(__fadd_rn(__fmul_rn(m[0], dl.x), __fmul_rn(m[1], dl.y)), __fmul_rn(m[2], dl.z)),
                    __fadd_rn(__fadd_rn(__fmul_rn(m[4], dl.x), __fmul_rn(m[5], dl.y)), __fmul_rn(m[6], dl.z)),
                    __fadd_rn(__fadd_rn(__fmul_rn(m[8], dl.x), __fmul_rn(m[9], dl.y)), __fmul_rn(m[10], dl.z)));
    // 1 / dl.z = sqrt(n2) up to rounding: no division needed
    const float inv_z = __fmul_rn(n2, inv_n);
    const float near_t = __fmul_rn(cam.near_clip, inv_z), far_t = __fmul_rn(cam.far_clip, inv_z);
    o = make_float3(__fmaf_rn(d.x, near_t, m[3]), __fmaf_rn(d.y, near_t, m[7]), __fmaf_rn(d.z, near_t, m[11]));
    maxt = __fsub_rn(far_t, near_t);
}

__device__ __forceinline__ void load_ray(const RaySrc &S, int64_t r, float3 &o, float3 &d, float &maxt)
{
#ifndef VP_NO_CAMERA
    if (S.has_cam) { camera_ray(S, r, o, d, maxt); return; }
#endif
    o = make_float3(S.o[3 * r], S.o[3 * r + 1], S.o[3 * r + 2]);
    d = make_float3(S.d[3 * r], S.d[3 * r + 1], S.d[3 * r + 2]);
    maxt = S.maxt ? S.maxt[r] : FLT_MAX;
}

// What the ray-major pass of the gather adjoint leaves in the bucket of the hit primitive: one 32-byte entry per recorded
// hit -- (d colour rgb, d alpha, ray direction xyz, ray index) -- written with ONE 256-bit store (a full sector: no
// read-for-merge in L2) and read back with one 256-bit load.  Every recorded entry is written, zero gradient included,
// so the buckets need no initialisation and no validity flags.
struct GatherBuf {
    const uint32_t *offsets;   // [N + 1] bucket starts (exclusive prefix of the recorded hits per primitive)
    const uint32_t *rank;      // position of a record entry inside the bucket of its primitive (indexed like the ids)
    float4 *entries;           // two float4 per bucket slot
    const int64_t *total;      // record validity: total[0] <= capacity && total[1] == 0
    int64_t capacity;
};

__device__ __forceinline__ void stg256(float4 *p, float4 a, float4 b)
{
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                 "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

__device__ __forceinline__ void bucket_store(const GatherBuf &gb, uint32_t slot, float dc0, float dc1, float dc2, float dalpha,
                                             float3 d, int64_t r)
{
    stg256(gb.entries + 2ll * slot, make_float4(dc0, dc1, dc2, dalpha), make_float4(d.x, d.y, d.z, __uint_as_float((uint32_t)r)));
}

struct TraceArgs {
    int64_t R;
    RaySrc src;
    float *rgb, *T;
    uint32_t *nhits;
    int32_t *ids;            // forward: hit-id output (dense, strides rs / hs)
    float4 *hit_state;       // forward, optional (volprim_rf): (colour rgb, transmittance) of every recorded hit, same indexing
    int32_t cap;
    int64_t rs, hs;
    // adjoint
    const float *dL, *state_in;
    const int32_t *rec_ids;       // replay: dense (rs, hs, rec_counts, cap) or compressed rows (rec_offsets, stride 1)
    const uint32_t *rec_counts;
    const int64_t *rec_offsets;
    const float4 *rec_state; // replay of a record that carries the per-hit (colour, transmittance)
    float *g_data, *g_attr, *g_sh;
    GatherBuf gb;
    vp_stats *stats;
};

// ---- forward ----------------------------------------------------------------------------------
// RR: Russian roulette compiled in (a separate instantiation: the 64-bit generator arithmetic in the hit loop costs
// the roulette-free kernel 9 % through register pressure alone, and no shipped configuration enables it)
// REC: hit lists (and, for volprim_rf, the per-hit colour + transmittance) are written out; the plain instantiation
// carries no store in its hit loop at all.
template <int INTEG, int KERNEL, int D, bool TILE, bool RR, bool REC>
__global__ void __launch_bounds__(TRACE_THREADS, VP_MIN_BLOCKS) k_trace_forward(DevScene S, vp_params P, TraceArgs A)
{
    extern __shared__ float4 smem_raw[];
    int *s_id = reinterpret_cast<int *>(smem_raw) + threadIdx.x;
    float *s_t = reinterpret_cast<float *>(smem_raw) + CAND_CAP * TRACE_THREADS + threadIdx.x;
    const int64_t t = (int64_t)blockIdx.x * TRACE_THREADS + threadIdx.x;
    Counters cn = {};
    const bool in_range = t < A.R;
    const int64_t r = in_range ? ray_index(t, P.image_width, P.image_height) : 0;
    float3 o = make_float3(0.f, 0.f, 0.f), d = make_float3(0.f, 0.f, 1.f);
    float maxt = FLT_MAX;
    if (in_range) load_ray(A.src, r, o, d, maxt);
    const float3 o0 = walker_origin<TILE>(S, o, d, in_range);
    float beta = 1.f, L[3] = { 0.f, 0.f, 0.f };
    uint32_t depth = 0;
    bool missed = false;

    ShHead<(D >= 0 ? D : 0)> sh_head;
    auto on_hit_body = [&](int pos, float4 g0, float4 g1, float4 g2, const Mat3 &Rm, const Isect &is) -> bool {
        float T;
        if constexpr (INTEG == VP_INTEGRATOR_RF) {
#if VP_EVAL_FAST
            T = rf_transmittance<KERNEL>(g0, g1, is);
#else
            RfEval e = rf_eval<KERNEL>(o, d, g0, g1, Rm, is);
            T = e.T;
#endif
            float raw[3] = { 0.f, 0.f, 0.f };
            if constexpr (D >= 0) {
                // the SH basis is re-evaluated per hit (~35 instructions) instead of holding 16 registers across the
                // whole walk: lower register pressure buys a sixth resident block per SM
                float Y[(D >= 0) ? (D + 1) * (D + 1) : 1];
                sh_basis<(D >= 0 ? D : 0)>(d, Y);
                sh_color<(D >= 0 ? D : 0)>(S, pos, Y, raw, &sh_head);
            }
            const float omt = 1.f - T;
            float col[3];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                col[ch] = (D >= 0) ? fmaxf(raw[ch], 0.f) : 0.f;
                float le = beta * omt * col[ch];       // rf:140
                if (!isfinite(le)) le = 0.f;            // rf:141
                L[ch] += le;                            // rf:145
            }
            if constexpr (REC) {
                // what the adjoint's ray pass needs of this hit: with (colour, T) on record it is a pure recurrence
                if (A.hit_state && depth < (uint32_t)A.cap)
                    A.hit_state[r * A.rs + depth * A.hs] = make_float4(col[0], col[1], col[2], T);
            }
        } else {
            float rho = (KERNEL == VP_KERNEL_GAUSSIAN) ? gauss_density_integral(g1, is)
                                                       : epan_density_integral(o, d, g0, g1, Rm, is);
            T = expf(-rho * g0.w);                      // tomo:44
        }
        beta *= T;                                      // rf:146 / tomo:85
        // (no atomics in this kernel: a RED anywhere in the hit loop -- even one that never executes -- stops ptxas from
        // moving the loads of the drain across it and doubled the kernel's time; per-primitive hit counts are taken from
        // the compressed record in a flat pass of the adjoint instead)
        if constexpr (REC) {
            if (depth < (uint32_t)A.cap) A.ids[r * A.rs + depth * A.hs] = __float_as_int(g1.w);
        }
        // ray.o = si.p + ray.d * 1e-4                     rf:149 / tomo:114
        o.x = fmaf(d.x, P.eps_advance, fmaf(d.x, is.tn, o.x));
        o.y = fmaf(d.y, P.eps_advance, fmaf(d.y, is.tn, o.y));
        o.z = fmaf(d.z, P.eps_advance, fmaf(d.z, is.tn, o.z));
        depth += 1;
        if (INTEG == VP_INTEGRATOR_RF && !(beta > P.t_cutoff)) return false;  // rf:173-174
        if constexpr (INTEG == VP_INTEGRATOR_RF && RR) {                      // rf:177-183 (primal pass only)
            const float rr_prob = fmaxf(beta, 0.1f);
            if (depth >= P.rr_depth && beta < 0.1f) {
                beta *= __fdiv_rn(1.f, rr_prob);
                // the sampler advances once per loop iteration: this one drew sample rr_skip + depth - 1 of the ray's stream
                const float u = pcg32_float_at(P.rr_seed, (uint32_t)(A.src.index_base + r), P.rr_skip + depth - 1u);
                if (!(u < rr_prob)) return false;
            }
        }
        if (!(depth < P.max_depth)) return false;                             // rf:186
        return true;
    };
    auto on_pre = [&](int pos) {
        if constexpr (INTEG == VP_INTEGRATOR_RF && D >= 0) sh_head.load(S, pos);
    };
    HitWithPreload<decltype(on_hit_body) &, decltype(on_pre) &> on_hit{ on_hit_body, on_pre };
    if constexpr (TILE) walk_tile(S, reinterpret_cast<int *>(smem_raw), o, o0, d, maxt, in_range, missed, cn, on_hit);
    else walk_ray(S, s_id, s_t, o, o0, d, maxt, in_range, missed, cn, on_hit);

    if (in_range) {
        if (INTEG == VP_INTEGRATOR_TOMO && missed && !(depth == 0 && P.hide_emitters)) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) L[ch] += beta * P.env[ch];   // tomo:105-111
        }
        if (INTEG == VP_INTEGRATOR_RF && P.srgb_primitives) {
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) L[ch] = srgb_to_linear(L[ch]);  // rf:189-190
        }
        A.rgb[3 * r] = L[0];
        A.rgb[3 * r + 1] = L[1];
        A.rgb[3 * r + 2] = L[2];
        if (A.T) A.T[r] = beta;
        if (A.nhits) A.nhits[r] = depth;
        // (hit lists are not padded: a ray's entries beyond its count are never read -- 75 % of the dense block on
        // the headline configuration used to be -1 fill)
    }
    cn.hits = depth;
    flush_counters(cn, A.stats);
}

// ---- adjoint ----------------------------------------------------------------------------------

// chain d(loss)/dR (3x3) to the un-normalised quaternion (x, y, z, w)
__device__ __forceinline__ void chain_dR_to_quat(float4 q, const float (&dR)[3][3], float (&gq)[4])
{
    float x = q.x, y = q.y, z = q.z, w = q.w;
    gq[0] = 2.f * (y * dR[0][1] + z * dR[0][2] + y * dR[1][0] - 2.f * x * dR[1][1] - w * dR[1][2] + z * dR[2][0]
                   + w * dR[2][1] - 2.f * x * dR[2][2]);
    gq[1] = 2.f * (-2.f * y * dR[0][0] + x * dR[0][1] + w * dR[0][2] + x * dR[1][0] + z * dR[1][2] - w * dR[2][0]
                   + z * dR[2][1] - 2.f * y * dR[2][2]);
    gq[2] = 2.f * (-2.f * z * dR[0][0] - w * dR[0][1] + x * dR[0][2] + w * dR[1][0] - 2.f * z * dR[1][1] + y * dR[1][2]
                   + x * dR[2][0] + y * dR[2][1]);
    gq[3] = 2.f * (-z * dR[0][1] + y * dR[0][2] + z * dR[1][0] - x * dR[1][2] - y * dR[2][0] + x * dR[2][1]);
}

__device__ __forceinline__ void scatter_geo(float *g_data, int orig, const float (&g10)[10])
{
    // [N*10] records are 8-byte aligned: five 64-bit vector reductions
    float2 *dst = reinterpret_cast<float2 *>(g_data + 10ll * orig);
#pragma unroll
    for (int k = 0; k < 5; ++k) atomicAdd(dst + k, make_float2(g10[2 * k], g10[2 * k + 1]));
}

// Per-hit coefficients of the rf adjoint: everything that depends on the ray's running state
// (beta, remaining radiance).  Given these, the gradient of the hit is a function of (ray, primitive) alone.
struct RfCoeffs {
    float T, G, op;
    float3 pp;        // p_peak
    float dalpha;     // d loss / d alpha  (before the 0.9999 clamp of alpha)
    float dcol[3];    // d loss / d colour, zero where the colour is clamped at 0
};

template <int KERNEL, int D, bool FAR = false>
__device__ __forceinline__ RfCoeffs rf_adjoint_coeffs(const DevScene &S, int pos, float3 o, float3 d, float4 g0, float4 g1,
                                                      const Mat3 &Rm, const Isect &is, float beta, const float (&g)[3],
                                                      float (&L)[3], const float (&Y)[(D >= 0) ? (D + 1) * (D + 1) : 1])
{
    RfCoeffs c;
    RfEval e;
    if constexpr (FAR) e = rf_eval_far<KERNEL>(o, d, g0, g1, Rm);
    else {
        e = rf_eval<KERNEL>(o, d, g0, g1, Rm, is);
#if VP_EVAL_FAST
        // the transmittance the primal pass used, bit for bit: a re-walking adjoint must terminate where it did
        e.T = rf_transmittance<KERNEL>(g0, g1, is);
#endif
    }
    float raw[3] = { 0.f, 0.f, 0.f };
    if constexpr (D >= 0) sh_color<(D >= 0 ? D : 0)>(S, pos, Y, raw);
    const float omt = 1.f - e.T;
    float dalpha = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float col = (D >= 0) ? fmaxf(raw[ch], 0.f) : 0.f;
        float le = beta * omt * col;
        bool lef = isfinite(le);
        if (!lef) le = 0.f;
        L[ch] -= le;                                  // rf:145 (adjoint branch)
        float lo = le + L[ch] * e.T / e.T;            // rf:156-159
        c.dcol[ch] = 0.f;
        if (isfinite(lo)) {                           // rf:160
            if (lef) {
                dalpha += g[ch] * beta * col;
                if (raw[ch] > 0.f) c.dcol[ch] = g[ch] * beta * omt;
            }
            dalpha -= g[ch] * L[ch] / e.T;
        }
    }
    c.T = e.T; c.G = e.G; c.op = e.op; c.pp = e.pp; c.dalpha = dalpha;
    return c;
}

// geometry terms of one hit: with dq = d loss / d q (q = |diag(1/(k s)) R^T (p_peak - c)|^2; d q / d t_peak = 0 at the
// peak), dw_i = 2 w_i / (k s_i)^2 dq.  d center = -R dw, d scale_i = -w_i dw_i / s_i, d R_ab = v_a dw_b.
template <int KERNEL>
__device__ __forceinline__ bool rf_geo_terms(float4 g0, float4 g1, const Mat3 &Rm, float3 pp, float G, float op, float dalpha,
                                             float3 &v, float3 &w, float (&dw)[3])
{
    const float dG = dalpha * op;
    float dq, k2 = 1.f;
    if (KERNEL == VP_KERNEL_GAUSSIAN) dq = -0.5f * G * dG;
    else { dq = (G > 0.f) ? -0.75f * dG : 0.f; k2 = 9.f; }
    if (dq == 0.f) return false;
    v = make_float3(pp.x - g0.x, pp.y - g0.y, pp.z - g0.z);
    w = rot_t_mul_fast(Rm, v);
    // reciprocals of per-primitive constants: loop-invariant in the primitive-major pass
    const float ix2 = __fdividef(2.f, k2 * g1.x * g1.x), iy2 = __fdividef(2.f, k2 * g1.y * g1.y), iz2 = __fdividef(2.f, k2 * g1.z * g1.z);
    dw[0] = w.x * ix2 * dq; dw[1] = w.y * iy2 * dq; dw[2] = w.z * iz2 * dq;
    return true;
}

// scatter formulation: the gradient of one hit goes to the reference-layout buffers with vector reductions
template <int KERNEL, int D>
__device__ __forceinline__ void rf_scatter_hit(const TraceArgs &A, float4 g0, float4 g1, float4 g2, const Mat3 &Rm,
                                               const RfCoeffs &c, const float (&Y)[(D >= 0) ? (D + 1) * (D + 1) : 1])
{
    constexpr int NB = (D >= 0) ? (D + 1) * (D + 1) : 0;
    const int orig = __float_as_int(g1.w);
    if constexpr (D >= 0) {
        if (A.g_sh && (c.dcol[0] != 0.f || c.dcol[1] != 0.f || c.dcol[2] != 0.f)) {
            constexpr int C = 3 * NB;
            float *dst = A.g_sh + (size_t)orig * C;
            if constexpr (C % 4 == 0) {
#pragma unroll
                for (int k = 0; k < C / 4; ++k) {
                    float4 v;
                    v.x = Y[(4 * k + 0) / 3] * c.dcol[(4 * k + 0) % 3];
                    v.y = Y[(4 * k + 1) / 3] * c.dcol[(4 * k + 1) % 3];
                    v.z = Y[(4 * k + 2) / 3] * c.dcol[(4 * k + 2) % 3];
                    v.w = Y[(4 * k + 3) / 3] * c.dcol[(4 * k + 3) % 3];
                    atomicAdd(reinterpret_cast<float4 *>(dst) + k, v);
                }
            } else {
#pragma unroll
                for (int k = 0; k < C; ++k) atomicAdd(dst + k, Y[k / 3] * c.dcol[k % 3]);
            }
        }
    }
    if (c.op * c.G < 0.9999f) {                       // dr.minimum passes the gradient to opacity * density
        atomicAdd(A.g_attr + orig, c.dalpha * c.G);
        float3 v, w;
        float dw[3];
        if (rf_geo_terms<KERNEL>(g0, g1, Rm, c.pp, c.G, c.op, c.dalpha, v, w, dw)) {
            float g10[10];
            g10[3] = -w.x * dw[0] / g1.x;
            g10[4] = -w.y * dw[1] / g1.y;
            g10[5] = -w.z * dw[2] / g1.z;
            const float va[3] = { v.x, v.y, v.z };
            float dR[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                g10[a] = -(Rm.m[a][0] * dw[0] + Rm.m[a][1] * dw[1] + Rm.m[a][2] * dw[2]);
#pragma unroll
                for (int b = 0; b < 3; ++b) dR[a][b] = va[a] * dw[b];
            }
            float gq[4];
            chain_dR_to_quat(g2, dR, gq);
            g10[6] = gq[0]; g10[7] = gq[1]; g10[8] = gq[2]; g10[9] = gq[3];
            scatter_geo(A.g_data, orig, g10);
        }
    }
}

// gather formulation, ray-major pass: the per-hit coefficients go to the bucket of the hit primitive, at the slot the
// counting pass assigned to this record entry (no atomics here: they would pin the loads of the replay loop in place)
__device__ __forceinline__ void rf_bucket_hit(const TraceArgs &A, int orig, int64_t r, int64_t entry, float3 d, RfCoeffs c)
{
    if (!(c.op * c.G < 0.9999f)) c.dalpha = 0.f;      // clamped alpha: no gradient to opacity / geometry
    const uint32_t slot = __ldg(A.gb.offsets + orig) + __ldg(A.gb.rank + entry);
    bucket_store(A.gb, slot, c.dcol[0], c.dcol[1], c.dcol[2], c.dalpha, d, r);
}

// one tomography interaction of the adjoint; returns T.  L stays state_in until the ray escapes (tomo:92-101).
template <int KERNEL>
__device__ __forceinline__ float tomo_adjoint_hit(const DevScene &S, const TraceArgs &A, float3 o, float3 d, float4 g0,
                                                  float4 g1, float4 g2, const Mat3 &Rm, const Isect &is,
                                                  const float (&g)[3], const float (&L)[3])
{
    const int orig = __float_as_int(g1.w);
    float rho = (KERNEL == VP_KERNEL_GAUSSIAN) ? gauss_density_integral(g1, is)
                                               : epan_density_integral(o, d, g0, g1, Rm, is);
    float sigma = g0.w;
    float T = expf(-rho * sigma);
    float dT = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float lo = L[ch] * T / T;
        if (isfinite(lo)) dT += g[ch] * L[ch] / T;
    }
    atomicAdd(A.g_attr + orig, -rho * T * dT);
    if (rho > 0.f) {
        float drho = -sigma * T * dT;
        // closed forms (see oracle chain_density / DESIGN.md): A = sum w^2/s^2, B = sum p w/s^2, C = sum p^2/s^2,
        // h = C - B^2/A.  h and u = p - (B/A) w do not change when the origin slides along the ray, so they are
        // evaluated from the ENTRY point (|p| <= extent * s) where fp32 does not cancel; B/A is shifted back.
        float3 w = is.rd;
        float3 ae = make_float3(fmaf(d.x, is.tn, o.x) - g0.x, fmaf(d.y, is.tn, o.y) - g0.y, fmaf(d.z, is.tn, o.z) - g0.z);
        float3 p = vp_rot_t_mul_rn(Rm, ae);
        float s[3] = { g1.x, g1.y, g1.z }, wv[3] = { w.x, w.y, w.z }, pv[3] = { p.x, p.y, p.z };
        float Aq = 0.f, Be = 0.f, Cq = 0.f, wn2 = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float is2 = 1.f / (s[i] * s[i]);
            Aq += wv[i] * wv[i] * is2;
            Be += pv[i] * wv[i] * is2;
            Cq += pv[i] * pv[i] * is2;
            wn2 += wv[i] * wv[i];
        }
        const float BAe = Be / Aq;        // B/A seen from the entry point
        const float BA = BAe - is.tn;     // B/A seen from the current origin
        float h = Cq - Be * BAe;
        float ch_, cW;
        if (KERNEL == VP_KERNEL_GAUSSIAN) { ch_ = -0.5f; cW = 0.f; }
        else {
            float E2 = S.extent * S.extent;
            ch_ = -0.5f / (E2 - h) + (-2.f / 3.f) / (1.f - E2 / 3.f - 2.f * h / 3.f);
            cW = 1.f;
        }
        float scale = rho * drho;
        float dp[3], dw[3], g10[10];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float si2 = s[i] * s[i], si3 = si2 * s[i];
            float u = pv[i] - BAe * wv[i];
            dp[i] = ch_ * (2.f * u / si2);
            dw[i] = ch_ * (-2.f * BA * u / si2) - (wv[i] / si2) / Aq + cW * wv[i] / wn2;
            g10[3 + i] = (ch_ * (-2.f * u * u / si3) + (wv[i] * wv[i] / si3) / Aq - 1.f / s[i]) * scale;
        }
        const float va[3] = { o.x - g0.x, o.y - g0.y, o.z - g0.z }, da[3] = { d.x, d.y, d.z };
        float dR[3][3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            g10[a] = -(Rm.m[a][0] * dp[0] + Rm.m[a][1] * dp[1] + Rm.m[a][2] * dp[2]) * scale;
#pragma unroll
            for (int b = 0; b < 3; ++b) dR[a][b] = (va[a] * dp[b] + da[a] * dw[b]) * scale;
        }
        float gq[4];
        chain_dR_to_quat(g2, dR, gq);
        g10[6] = gq[0]; g10[7] = gq[1]; g10[8] = gq[2]; g10[9] = gq[3];
        scatter_geo(A.g_data, orig, g10);
    }
    return T;
}

#ifndef VP_REPLAY_BLOCKS
#define VP_REPLAY_BLOCKS 8   // the replay adjoint uses no shared memory; bound by reduction throughput, not occupancy
#endif
// Where the adjoint takes a ray's hit sequence from, and where the gradient of a hit goes
enum { ADJ_WALK = 0,        // re-trace like the primal, scatter with vector reductions
       ADJ_REPLAY_DENSE,    // replay a dense [cap, R] record, scatter
       ADJ_REPLAY_ROWS,     // replay a compressed-row record, scatter            (volprim_tomography)
       ADJ_REPLAY_BUCKET }; // replay a compressed-row record, per-hit coefficients to the primitive's bucket (volprim_rf)

__device__ __forceinline__ bool record_usable(const GatherBuf &gb)
{
    return !gb.total || (gb.total[0] <= gb.capacity && gb.total[1] == 0);
}

template <int INTEG, int KERNEL, int D, int MODE, bool TILE>
__global__ void __launch_bounds__(TRACE_THREADS, MODE != ADJ_WALK ? VP_REPLAY_BLOCKS : VP_MIN_BLOCKS) k_trace_adjoint(DevScene S, vp_params P, TraceArgs A)
{
    constexpr bool REPLAY = MODE != ADJ_WALK;
    extern __shared__ float4 smem_raw[];
    int *s_id = reinterpret_cast<int *>(smem_raw) + threadIdx.x;
    float *s_t = reinterpret_cast<float *>(smem_raw) + CAND_CAP * TRACE_THREADS + threadIdx.x;
    const int64_t t = (int64_t)blockIdx.x * TRACE_THREADS + threadIdx.x;
    Counters cn = {};
    if constexpr (MODE == ADJ_REPLAY_ROWS || MODE == ADJ_REPLAY_BUCKET)
        if (!record_usable(A.gb)) return;            // truncated record: the caller re-traces (see vp_hit_record)
    const bool in_range = t < A.R;
    const int64_t r = in_range ? ray_index(t, P.image_width, P.image_height) : 0;
    float g[3] = { 0.f, 0.f, 0.f };
    if (in_range) { g[0] = A.dL[3 * r]; g[1] = A.dL[3 * r + 1]; g[2] = A.dL[3 * r + 2]; }
    const bool alive = in_range && (g[0] != 0.f || g[1] != 0.f || g[2] != 0.f);   // rf:111-112
    float3 o = make_float3(0.f, 0.f, 0.f), d = make_float3(0.f, 0.f, 1.f);
    float maxt = FLT_MAX;
    float L[3] = { 0.f, 0.f, 0.f };
    if (alive) {
        load_ray(A.src, r, o, d, maxt);
        L[0] = A.state_in[3 * r]; L[1] = A.state_in[3 * r + 1]; L[2] = A.state_in[3 * r + 2];
    }
    float3 o0 = o;
    if constexpr (!REPLAY) o0 = walker_origin<TILE>(S, o, d, alive);
    float beta = 1.f;
    uint32_t depth = 0;
    constexpr int NY = (D >= 0) ? (D + 1) * (D + 1) : 1;
    float Y[NY];
    if constexpr (INTEG == VP_INTEGRATOR_RF && D >= 0) sh_basis<(D >= 0 ? D : 0)>(d, Y);
    else Y[0] = 0.f;

    int64_t entry = 0;        // index of the current hit in the compressed record (bucket mode)
    auto interact = [&](int pos, float4 g0, float4 g1, float4 g2, const Mat3 &Rm, const Isect &is) -> bool {
        float T;
        if constexpr (INTEG == VP_INTEGRATOR_RF) {
            // bucket mode evaluates every hit from the ray's ORIGINAL origin (rf_eval_far): no exact entry distance, no
            // origin advance -- the list already fixes which primitives are hit and in which order
            const RfCoeffs c = rf_adjoint_coeffs<KERNEL, D, MODE == ADJ_REPLAY_BUCKET>(S, pos, o, d, g0, g1, Rm, is, beta, g, L, Y);
            if constexpr (MODE == ADJ_REPLAY_BUCKET) rf_bucket_hit(A, __float_as_int(g1.w), r, entry, d, c);
            else rf_scatter_hit<KERNEL, D>(A, g0, g1, g2, Rm, c, Y);
            T = c.T;
        } else
            T = tomo_adjoint_hit<KERNEL>(S, A, o, d, g0, g1, g2, Rm, is, g, L);
        beta *= T;
        if constexpr (MODE != ADJ_REPLAY_BUCKET) {
            o.x = fmaf(d.x, P.eps_advance, fmaf(d.x, is.tn, o.x));
            o.y = fmaf(d.y, P.eps_advance, fmaf(d.y, is.tn, o.y));
            o.z = fmaf(d.z, P.eps_advance, fmaf(d.z, is.tn, o.z));
        }
        depth += 1;
        if (INTEG == VP_INTEGRATOR_RF && !(beta > P.t_cutoff)) return false;
        if (!(depth < P.max_depth)) return false;
        return true;
    };

    if constexpr (MODE == ADJ_REPLAY_BUCKET) {
        if (in_range && !alive) {      // rf:111-112 skips the ray; its bucket entries still have to exist (as zeros)
            for (int64_t e = A.rec_offsets[r]; e < A.rec_offsets[r + 1]; ++e)
                bucket_store(A.gb, __ldg(A.gb.offsets + A.rec_ids[e]) + __ldg(A.gb.rank + e), 0.f, 0.f, 0.f, 0.f, d, r);
        }
    }
    if constexpr (REPLAY) {
        if (alive) {
            // this ray's list: element k at list[k * stride]
            const int32_t *list;
            int64_t stride;
            uint32_t n;
            if constexpr (MODE == ADJ_REPLAY_DENSE) {
                list = A.rec_ids + r * A.rs;
                stride = A.hs;
                n = A.rec_counts[r];
                if (n > (uint32_t)A.cap) n = (uint32_t)A.cap;
            } else {
                const int64_t b = A.rec_offsets[r];
                list = A.rec_ids + b;
                stride = 1;
                n = (uint32_t)(A.rec_offsets[r + 1] - b);
                entry = b;
            }
            // software pipeline: the id -> position -> records chain of hit k+1 is started during hit k
            int orig_next = n > 0 ? list[0] : -1;
            int pos_next = orig_next >= 0 ? __ldg(S.inv_perm + orig_next) : 0;
            if (orig_next >= 0) prefetch_prim(S, pos_next);
            for (uint32_t k = 0; k < n; ++k) {
                const int orig = orig_next;
                if (orig < 0) break;
                const int pos = pos_next;
                orig_next = (k + 1 < n) ? list[(k + 1) * stride] : -1;
                pos_next = orig_next >= 0 ? __ldg(S.inv_perm + orig_next) : 0;
                if (orig_next >= 0) prefetch_prim(S, pos_next);
                float4 g0 = __ldg(S.geo0 + pos), g1 = __ldg(S.geo1 + pos), g2 = __ldg(S.geo2 + pos);
                if constexpr (MODE == ADJ_REPLAY_BUCKET) {
                    const Mat3 Rm = quat_to_matrix_fast(g2);
                    Isect is;
                    is.valid = true; is.tn = is.tf = 0.f;
                    is.ro = is.rd = make_float3(0.f, 0.f, 0.f);
                    interact(pos, g0, g1, g2, Rm, is);
                } else {
                    Mat3 Rm = vp_quat_to_matrix_rn(g2);
                    Isect is = exact_isect(o, d, g0, g1, Rm, S.extent);
                    interact(pos, g0, g1, g2, Rm, is);
                }
                ++entry;
            }
        }
    } else {
        bool missed = false;
        if constexpr (TILE) walk_tile(S, reinterpret_cast<int *>(smem_raw), o, o0, d, maxt, alive, missed, cn, interact);
        else walk_ray(S, s_id, s_t, o, o0, d, maxt, alive, missed, cn, interact);
    }
    cn.hits = depth;
    flush_counters(cn, A.stats);
}

// ---- gather adjoint, primitive-major pass ---------------------------------------------------------------------------
// One warp per primitive (original numbering, so that consecutive warps stream consecutive buckets and gradient rows).
// The lanes stride over the primitive's bucket; every lane re-evaluates its hits from (ray, primitive, per-hit
// coefficients) and keeps the sums in registers: C colour-coefficient gradients + 16 geometry sums (opacity, dw (3),
// w.dw (3), v (x) dw (9); the chain to centre / scale / quaternion is linear in them and applied once).  A butterfly
// reduce-scatter (NV - NV/32 shuffles for NV values) leaves NV/32 finished sums per lane, which are ADDED to the caller's
// buffers with plain coalesced read-modify-writes: one write per gradient float and launch, no reductions in L2.
struct GatherArgs {
    RaySrc src;
    const float *data10, *attr;     // the primitives in the reference layouts (context copies, original order)
    GatherBuf gb;
    int64_t p_begin, p_end;
    const uint32_t *extra_items;    // EXTRA pass: primitive of every extra work item
    const uint32_t *extra_offsets;  // [N + 1] exclusive prefix of the extra items per primitive
    const uint32_t *n_extra;        // device scalar: number of extra items
    float *g_data, *g_attr, *g_sh;
};

#ifndef VP_GATHER_CHUNK
#define VP_GATHER_CHUNK 256        // bucket entries one warp accumulates (8 rounds of 32 lanes)
#endif
constexpr uint32_t GATHER_CHUNK = VP_GATHER_CHUNK;

// Buckets are heavy-tailed (a primitive close to the camera collects tens of thousands of hits, the median one a few
// dozen): a warp takes at most GATHER_CHUNK entries.  The MAIN pass (EXTRA = false) gives every primitive of the
// range one warp for its first chunk and adds the sums with plain read-modify-writes; the EXTRA pass runs one warp per
// further chunk of the big buckets and adds with reductions (few: 59 per 256 hits instead of 27 per hit).
#ifndef VP_GATHER_BLOCKS
#define VP_GATHER_BLOCKS 4
#endif
template <int KERNEL, int D, bool EXTRA>
__global__ void __launch_bounds__(128, VP_GATHER_BLOCKS) k_adjoint_gather(GatherArgs A)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int NB = (D >= 0) ? (D + 1) * (D + 1) : 0;
    constexpr int C = 3 * NB;
    constexpr int NV = (C + 16 <= 32) ? 32 : 64;    // sums per primitive, padded to a multiple of the warp size
    constexpr int M = NV / 32;                      // finished sums per lane
    static_assert(C + 16 <= 64, "SH degree 3 at most");
    const int lane = threadIdx.x & 31;
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (!record_usable(A.gb)) return;
    int64_t p;
    uint32_t first;        // first bucket entry of this warp's chunk
    if constexpr (EXTRA) {
        if (w >= (int64_t)__ldg(A.n_extra)) return;
        p = __ldg(A.extra_items + w);
        if (p < A.p_begin || p >= A.p_end) return;
        first = ((uint32_t)w - __ldg(A.extra_offsets + p) + 1u) * GATHER_CHUNK;
    } else {
        p = A.p_begin + w;
        if (p >= A.p_end) return;
        first = 0;
    }
    const uint32_t base = __ldg(A.gb.offsets + p);
    const uint32_t size = __ldg(A.gb.offsets + p + 1) - base;
    if (first >= size) return;
    const uint32_t cnt = min(size - first, GATHER_CHUNK);
    const float4 *bent = A.gb.entries + 2ll * (base + first);
    const float *rec = A.data10 + 10 * p;
    const float4 g0 = make_float4(__ldg(rec), __ldg(rec + 1), __ldg(rec + 2), A.attr ? __ldg(A.attr + p) : 1.f);
    const float4 g1 = make_float4(__ldg(rec + 3), __ldg(rec + 4), __ldg(rec + 5), 0.f);
    const float4 g2 = make_float4(__ldg(rec + 6), __ldg(rec + 7), __ldg(rec + 8), __ldg(rec + 9));
    const Mat3 Rm = quat_to_matrix_fast(g2);
    float acc[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) acc[k] = 0.f;
    // main pass: the gradient rows this warp will add to are fetched NOW, so that the read of the final
    // read-modify-write has long arrived when the bucket is done (it used to cost 8 % of the kernel's stall samples)
    float old_sh[M], old_geo = 0.f;
#pragma unroll
    for (int k = 0; k < M; ++k) old_sh[k] = 0.f;
    if constexpr (!EXTRA) {
        if constexpr (C > 0) {
            if (A.g_sh) {
#pragma unroll
                for (int k = 0; k < M; ++k)
                    if (M * lane + k < C) old_sh[k] = A.g_sh[(size_t)p * C + M * lane + k];
            }
        }
        if (lane <= 10) old_geo = lane < 10 ? A.g_data[10 * p + lane] : A.g_attr[p];
    }

    // one entry ahead: the bucket entry of the next round is in flight while this one is evaluated
    uint32_t i = lane;
    float4 ea_next = make_float4(0.f, 0.f, 0.f, 0.f), eb_next = ea_next;
    if (i < cnt) ldg256(bent + 2ll * i, ea_next, eb_next);
    bool any = false;
    while (i < cnt) {
        const float4 st = ea_next, eb = eb_next;
        i += 32;
        if (i < cnt) ldg256(bent + 2ll * i, ea_next, eb_next);
        const bool has_col = st.x != 0.f || st.y != 0.f || st.z != 0.f;
        if (!has_col && st.w == 0.f) continue;      // a hit without gradient (or a ray without): nothing to add
        any = true;
        const float3 d = make_float3(eb.x, eb.y, eb.z);
        if constexpr (D >= 0) {
            if (has_col) {
                float Y[NB > 0 ? NB : 1];
                sh_basis<(D >= 0 ? D : 0)>(d, Y);
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    acc[3 * j + 0] = fmaf(Y[j], st.x, acc[3 * j + 0]);
                    acc[3 * j + 1] = fmaf(Y[j], st.y, acc[3 * j + 1]);
                    acc[3 * j + 2] = fmaf(Y[j], st.z, acc[3 * j + 2]);
                }
            }
        }
        if (st.w != 0.f) {
            // any point of the ray serves as the origin (rf_eval_far re-bases): the sensor's position, or the ray's own
            // origin for an explicit batch
            float3 o;
            if (A.src.has_cam) o = make_float3(A.src.cam.to_world[3], A.src.cam.to_world[7], A.src.cam.to_world[11]);
            else {
                const int64_t r = __float_as_uint(eb.w);
                o = make_float3(__ldg(A.src.o + 3 * r), __ldg(A.src.o + 3 * r + 1), __ldg(A.src.o + 3 * r + 2));
            }
            const RfEval e = rf_eval_far<KERNEL>(o, d, g0, g1, Rm);
            acc[C] = fmaf(st.w, e.G, acc[C]);                       // d opacity
            float3 v, wv;
            float dw[3];
            if (rf_geo_terms<KERNEL>(g0, g1, Rm, e.pp, e.G, e.op, st.w, v, wv, dw)) {
                acc[C + 1] += dw[0]; acc[C + 2] += dw[1]; acc[C + 3] += dw[2];
                acc[C + 4] = fmaf(wv.x, dw[0], acc[C + 4]);
                acc[C + 5] = fmaf(wv.y, dw[1], acc[C + 5]);
                acc[C + 6] = fmaf(wv.z, dw[2], acc[C + 6]);
                const float va[3] = { v.x, v.y, v.z };
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) acc[C + 7 + 3 * a + b] = fmaf(va[a], dw[b], acc[C + 7 + 3 * a + b]);
            }
        }
    }
    if (!__any_sync(FULL, any)) return;
    // butterfly reduce-scatter: after the step with lane distance s, a lane holds the partial sums of the half of the
    // index range selected by its bit s; in the end lane l owns indices M * l .. M * l + M - 1
#pragma unroll
    for (int half = NV / 2, sft = 16; sft >= 1; half >>= 1, sft >>= 1) {
        const bool up = (lane & sft) != 0;
#pragma unroll
        for (int k = 0; k < half; ++k) {
            const float keep = up ? acc[k + half] : acc[k];
            const float send = up ? acc[k] : acc[k + half];
            acc[k] = keep + __shfl_xor_sync(FULL, send, sft);
        }
    }
    // colour coefficients: coalesced read-modify-write (main pass) / reductions (extra chunks of a big bucket)
    if constexpr (C > 0) {
        if (A.g_sh) {
            float *dst = A.g_sh + (size_t)p * C;
#pragma unroll
            for (int k = 0; k < M; ++k) {
                const int idx = M * lane + k;
                if (idx < C) {
                    if constexpr (EXTRA) atomicAdd(dst + idx, acc[k]);
                    else dst[idx] = old_sh[k] + acc[k];
                }
            }
        }
    }
    // geometry sums -> the chain to centre / scale / quaternion is linear in them: every lane gets all sixteen, lanes
    // 0..10 each finish ONE output float and add it to its slot (eleven independent read-modify-writes in flight)
    float gs[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) gs[j] = __shfl_sync(FULL, acc[(C + j) % M], (C + j) / M);
    float g10[10];
    float dR[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        g10[a] = -(Rm.m[a][0] * gs[1] + Rm.m[a][1] * gs[2] + Rm.m[a][2] * gs[3]);
#pragma unroll
        for (int b = 0; b < 3; ++b) dR[a][b] = gs[7 + 3 * a + b];
    }
    g10[3] = -gs[4] / g1.x;
    g10[4] = -gs[5] / g1.y;
    g10[5] = -gs[6] / g1.z;
    float gq[4];
    chain_dR_to_quat(g2, dR, gq);
    g10[6] = gq[0]; g10[7] = gq[1]; g10[8] = gq[2]; g10[9] = gq[3];
    float mine = gs[0];
#pragma unroll
    for (int k = 0; k < 10; ++k) mine = lane == k ? g10[k] : mine;      // lane 10 keeps gs[0] (d attribute)
    if (lane <= 10) {
        float *dst = lane < 10 ? A.g_data + 10 * p + lane : A.g_attr + p;
        if constexpr (EXTRA) atomicAdd(dst, mine);
        else *dst = old_geo + mine;
    }
}

// counting pass of the gather adjoint: one thread per record entry; rank[e] = position of the entry inside the bucket
// of its primitive (arrival order), counts[p] = bucket size.  Flat and coalesced; the only atomics of the adjoint.
__global__ void __launch_bounds__(256) k_bucket_ranks(const int32_t *__restrict__ ids, const int64_t *__restrict__ total,
                                                      int64_t capacity, uint32_t *__restrict__ counts, uint32_t *__restrict__ rank)
{
    const int64_t n = total[0];
    if (n > capacity || total[1] != 0) return;
    // four entries per thread and round: the atomics return a value, so their latency is hidden by having several in flight
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < n; e0 += 4 * stride) {
        int32_t id[4];
        uint32_t rk[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) id[k] = e0 + k * stride < n ? __ldcs(ids + e0 + k * stride) : -1;
#pragma unroll
        for (int k = 0; k < 4; ++k) rk[k] = id[k] >= 0 ? atomicAdd(counts + id[k], 1u) : 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (id[k] >= 0) rank[e0 + k * stride] = rk[k];
    }
}

// the same over a DENSE hit-major record ([id_cap][n_rays], entry (k, r) valid for k < counts[r]): one thread per ray,
// reads and rank writes coalesced across the warp.  The 32 rays of a warp are neighbours and share primitives at equal
// depth, and the pass is bound by the throughput of value-returning atomics: lanes with the same primitive are grouped
// (__match_any_sync), the group's first lane adds the group size once and hands out consecutive ranks.  Four depths are
// in flight per round so that the atomics' latency overlaps.
__global__ void __launch_bounds__(256) k_bucket_ranks_dense(const int32_t *__restrict__ ids, const uint32_t *__restrict__ ray_counts,
                                                            int64_t n_rays, const int64_t *__restrict__ total, int64_t capacity,
                                                            uint32_t *__restrict__ counts, uint32_t *__restrict__ rank)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int U = 4;
    if (total[0] > capacity || total[1] != 0) return;
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const uint32_t c = r < n_rays ? ray_counts[r] : 0u;       // (no early exit: the warp stays convergent for the matches)
    const uint32_t cmax = __reduce_max_sync(FULL, c);
    for (uint32_t k0 = 0; k0 < cmax; k0 += U) {
        int32_t id[U];
        unsigned grp[U];
        uint32_t base[U];
#pragma unroll
        for (int u = 0; u < U; ++u) id[u] = (k0 + u < c) ? __ldcs(ids + (int64_t)(k0 + u) * n_rays + r) : -1;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned act = __ballot_sync(FULL, id[u] >= 0);
            grp[u] = 0u;
            base[u] = 0u;
            if (id[u] >= 0) {
#if VP_RANKS_MATCH
                grp[u] = __match_any_sync(act, id[u]);
#else
                grp[u] = 1u << lane;
#endif
                if ((grp[u] & lt) == 0u) base[u] = atomicAdd(counts + id[u], (uint32_t)__popc(grp[u]));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (id[u] >= 0) {
                const uint32_t b = __shfl_sync(grp[u], base[u], __ffs(grp[u]) - 1);
                rank[(int64_t)(k0 + u) * n_rays + r] = b + (uint32_t)__popc(grp[u] & lt);
            }
        }
    }
}

// Ray-major pass over a dense record carrying the per-hit (colour, transmittance): one thread per ray, every load
// coalesced (lane = ray, same hit index); the PRB recurrences of volprim_rf.py:137-165 alone.
// one hit of the recurrences of volprim_rf.py:137-165 from its recorded (colour, transmittance)
__device__ __forceinline__ void prb_step(float4 st, const float (&g)[3], float (&L)[3], float &beta, float &dalpha, float (&dcol)[3])
{
    const float T = st.w, omt = 1.f - T;
    const float col[3] = { st.x, st.y, st.z };
    // one approximate reciprocal per hit instead of six IEEE divisions (T = 0 -> inf -> NaN below, as the literal form)
    const float inv_t = vp_rcp(T);
    dalpha = 0.f;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float le = beta * omt * col[ch];
        const bool lef = isfinite(le);
        if (!lef) le = 0.f;
        L[ch] -= le;                                  // rf:145 (adjoint branch)
        const float lo = le + L[ch] * T * inv_t;      // rf:156-159
        dcol[ch] = 0.f;
        if (isfinite(lo)) {                           // rf:160
            if (lef) {
                dalpha += g[ch] * beta * col[ch];
                if (col[ch] > 0.f) dcol[ch] = g[ch] * beta * omt;
            }
            dalpha -= g[ch] * L[ch] * inv_t;
        }
    }
    if (!(T > 1.f - 0.9999f)) dalpha = 0.f;           // alpha was clamped (rf:76): no gradient to opacity / geometry
    beta *= T;
}

__global__ void __launch_bounds__(128) k_adjoint_rows_dense(RaySrc src, int64_t R, int32_t image_w, int32_t image_h, const float *__restrict__ dL,
                                                            const float *__restrict__ state_in,
                                                            const uint32_t *__restrict__ ray_counts,
                                                            const int32_t *__restrict__ ids, const float4 *__restrict__ hit_state,
                                                            GatherBuf gb)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= R || !record_usable(gb)) return;
    const int64_t r = ray_index(t, image_w, image_h);      // 8x4 pixel tiles: neighbours hit the same primitives
    const float g[3] = { dL[3 * r], dL[3 * r + 1], dL[3 * r + 2] };
    const bool live = g[0] != 0.f || g[1] != 0.f || g[2] != 0.f;      // rf:111-112: a ray without gradient leaves zero entries
    float L[3] = { state_in[3 * r], state_in[3 * r + 1], state_in[3 * r + 2] };
    float beta = 1.f;
    float3 o, d;
    float maxt;
    load_ray(src, r, o, d, maxt);
    const uint32_t n = ray_counts[r];
    float4 st_next = n ? __ldcs(hit_state + r) : make_float4(0.f, 0.f, 0.f, 1.f);
    for (uint32_t k = 0; k < n; ++k) {
        const int64_t e = (int64_t)k * R + r;
        const float4 st = st_next;
        if (k + 1 < n) st_next = __ldcs(hit_state + e + R);
        const uint32_t slot = __ldg(gb.offsets + __ldcs(ids + e)) + __ldcs(gb.rank + e);
        float dalpha, dcol[3];
        prb_step(st, g, L, beta, dalpha, dcol);
        if (!live) { dalpha = 0.f; dcol[0] = dcol[1] = dcol[2] = 0.f; }
        bucket_store(gb, slot, dcol[0], dcol[1], dcol[2], dalpha, d, r);
    }
}

// extra work items of the primitive-major pass: ceil(size / CHUNK) - 1 per bucket
__global__ void k_extra_counts(const uint32_t *__restrict__ offsets, int64_t n, uint32_t *__restrict__ extra)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t c = offsets[p + 1] - offsets[p];
    extra[p] = c > GATHER_CHUNK ? (c - 1) / GATHER_CHUNK : 0u;
}

__global__ void k_extra_items(const uint32_t *__restrict__ extra_offsets, int64_t n, uint32_t *__restrict__ items, uint32_t max_items)
{
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const uint32_t a = extra_offsets[p], b = extra_offsets[p + 1];
    for (uint32_t k = a; k < b && k < max_items; ++k) items[k] = (uint32_t)p;
}

// ---- hit records: dense hit-major scratch -> compressed rows ----------------------------------------------------------
// per-ray entry counts of a band (clamped at the record's per-ray cap) + number of rays whose list was cut
__global__ void k_record_counts(const uint32_t *__restrict__ nhits, int64_t n, int32_t cap, uint32_t *__restrict__ counts,
                                int64_t *__restrict__ total, bool sum_entries)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool cut = false;
    uint32_t c = 0;
    if (r < n) {
        const uint32_t h = nhits[r];
        cut = h > (uint32_t)cap;
        c = cut ? (uint32_t)cap : h;
        counts[r] = c;
    }
    const unsigned m = __ballot_sync(0xffffffffu, cut);
    if (m && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)(total + 1), (unsigned long long)__popc(m));
    if (sum_entries) {      // dense records: total[0] = number of entries (compressed rows get it from their scan)
        const uint32_t s = __reduce_add_sync(0xffffffffu, c);
        if (s && (threadIdx.x & 31) == 0) atomicAdd((unsigned long long *)total, (unsigned long long)s);
    }
}

// Dense hit-major block -> compressed rows.  One warp per 32 consecutive rays; the block is read coalesced (lane = ray,
// same hit index), transposed through shared memory, and every ray's row is written coalesced (lane = hit index).
__global__ void __launch_bounds__(128) k_compact_hits(const int32_t *__restrict__ dense, int64_t n, const uint32_t *__restrict__ counts,
                                                      const int64_t *__restrict__ offsets, int32_t *__restrict__ ids, int64_t capacity)
{
    __shared__ int32_t tile[4][32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * 4 + wid) * 32;
    if (r0 >= n) return;
    const int64_t r = r0 + lane;
    const uint32_t c = r < n ? counts[r] : 0u;
    const int64_t off = r < n ? offsets[r] : 0;
    const uint32_t cmax = __reduce_max_sync(0xffffffffu, c);
    for (uint32_t k0 = 0; k0 < cmax; k0 += 32) {
        // rows k0 .. k0 + 31 of the dense block, columns r0 .. r0 + 31
        const uint32_t kn = min(32u, cmax - k0);
        for (uint32_t k = 0; k < kn; ++k)
            tile[wid][k][lane] = (k0 + k < c) ? dense[(int64_t)(k0 + k) * n + r] : -1;
        __syncwarp();
        for (int j = 0; j < 32; ++j) {
            const uint32_t cj = __shfl_sync(0xffffffffu, c, j);
            const int64_t oj = __shfl_sync(0xffffffffu, off, j);
            if (k0 + lane < cj && oj + cj <= capacity) ids[oj + k0 + lane] = tile[wid][lane][j];
        }
        __syncwarp();
    }
}

// the same for the per-hit (colour, transmittance) records: 16-byte elements, two warps per block
__global__ void __launch_bounds__(64) k_compact_state(const float4 *__restrict__ dense, int64_t n, const uint32_t *__restrict__ counts,
                                                      const int64_t *__restrict__ offsets, float4 *__restrict__ out, int64_t capacity)
{
    __shared__ float4 tile[2][32][33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * 2 + wid) * 32;
    if (r0 >= n) return;
    const int64_t r = r0 + lane;
    const uint32_t c = r < n ? counts[r] : 0u;
    const int64_t off = r < n ? offsets[r] : 0;
    const uint32_t cmax = __reduce_max_sync(0xffffffffu, c);
    for (uint32_t k0 = 0; k0 < cmax; k0 += 32) {
        const uint32_t kn = min(32u, cmax - k0);
        for (uint32_t k = 0; k < kn; ++k)
            if (k0 + k < c) tile[wid][k][lane] = dense[(int64_t)(k0 + k) * n + r];
        __syncwarp();
        for (int j = 0; j < 32; ++j) {
            const uint32_t cj = __shfl_sync(0xffffffffu, c, j);
            const int64_t oj = __shfl_sync(0xffffffffu, off, j);
            if (k0 + lane < cj && oj + cj <= capacity) out[oj + k0 + lane] = tile[wid][lane][j];
        }
        __syncwarp();
    }
}

// Ray-major pass of the gather adjoint when the record carries the per-hit (colour, transmittance): the PRB recurrences
// of volprim_rf.py:137-165 alone -- no primitive is loaded, nothing is intersected or shaded again.  A warp owns 32
// consecutive rays.  Their rows of the record are staged through shared memory 16 entries at a time (two rows per
// step, every row segment read coalesced); lane l then walks the staged entries of ray l and leaves (d colour, d alpha)
// in the bucket slot of every hit.
constexpr int ROWS_K = 16;          // entries of a row staged per round
constexpr int ROWS_WARPS = 2;
__global__ void __launch_bounds__(32 * ROWS_WARPS) k_adjoint_rows(RaySrc src, int64_t R, const float *__restrict__ dL,
                                                                  const float *__restrict__ state_in,
                                                                  const int64_t *__restrict__ offsets,
                                                                  const int32_t *__restrict__ ids,
                                                                  const float4 *__restrict__ hit_state, GatherBuf gb)
{
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ float4 s_state[ROWS_WARPS][32][ROWS_K + 1];
    __shared__ int32_t s_id[ROWS_WARPS][32][ROWS_K + 1];
    __shared__ uint32_t s_rank[ROWS_WARPS][32][ROWS_K + 1];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t r0 = ((int64_t)blockIdx.x * ROWS_WARPS + wid) * 32;
    if (r0 >= R || !record_usable(gb)) return;
    const int64_t r = r0 + lane;
    float g[3] = { 0.f, 0.f, 0.f }, L[3] = { 0.f, 0.f, 0.f };
    float3 o = make_float3(0.f, 0.f, 0.f), d = make_float3(0.f, 0.f, 1.f);
    int64_t e0 = 0;
    uint32_t n = 0;
    if (r < R) {
        float maxt;
        load_ray(src, r, o, d, maxt);
        g[0] = dL[3 * r]; g[1] = dL[3 * r + 1]; g[2] = dL[3 * r + 2];
        e0 = offsets[r];
        n = (uint32_t)(offsets[r + 1] - e0);
        L[0] = state_in[3 * r]; L[1] = state_in[3 * r + 1]; L[2] = state_in[3 * r + 2];
    }
    const bool live = g[0] != 0.f || g[1] != 0.f || g[2] != 0.f;      // rf:111-112: a ray without gradient leaves zero entries
    float beta = 1.f;
    const uint32_t nmax = __reduce_max_sync(FULL, n);
    const int half = lane >> 4, sub = lane & 15;
    for (uint32_t k0 = 0; k0 < nmax; k0 += ROWS_K) {
        // stage entries k0 .. k0 + 15 of all 32 rows: lanes 0-15 take row 2 i, lanes 16-31 row 2 i + 1
#pragma unroll 4
        for (int i = 0; i < 16; ++i) {
            const int j = 2 * i + half;
            const uint32_t nj = __shfl_sync(FULL, n, j);
            const int64_t ej = __shfl_sync(FULL, e0, j);
            if (k0 + sub < nj) {
                const int64_t e = ej + k0 + sub;
                s_state[wid][j][sub] = __ldcs(hit_state + e);
                s_id[wid][j][sub] = __ldcs(ids + e);
                s_rank[wid][j][sub] = __ldcs(gb.rank + e);
            }
        }
        __syncwarp();
        const uint32_t kn = n > k0 ? min((uint32_t)ROWS_K, n - k0) : 0u;
        for (uint32_t k = 0; k < kn; ++k) {
            const uint32_t slot = __ldg(gb.offsets + s_id[wid][lane][k]) + s_rank[wid][lane][k];
            float dalpha, dcol[3];
            prb_step(s_state[wid][lane][k], g, L, beta, dalpha, dcol);
            if (!live) { dalpha = 0.f; dcol[0] = dcol[1] = dcol[2] = 0.f; }
            bucket_store(gb, slot, dcol[0], dcol[1], dcol[2], dalpha, d, r);
        }
        __syncwarp();
    }
}

__global__ void k_raygen(RaySrc S, int64_t total, float *__restrict__ ro, float *__restrict__ rd, float *__restrict__ rmaxt)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    float3 o, d;
    float maxt;
    camera_ray(S, i, o, d, maxt);
    ro[3 * i] = o.x; ro[3 * i + 1] = o.y; ro[3 * i + 2] = o.z;
    rd[3 * i] = d.x; rd[3 * i + 1] = d.y; rd[3 * i + 2] = d.z;
    if (rmaxt) rmaxt[i] = maxt;
}

template <int INTEG, int KERNEL, int D>
void launch_forward(const DevScene &S, const vp_params &P, const TraceArgs &A, cudaStream_t st)
{
    const unsigned blocks = (unsigned)((A.R + TRACE_THREADS - 1) / TRACE_THREADS);
    constexpr bool RF = INTEG == VP_INTEGRATOR_RF;
    const bool rr = RF && P.use_rr, rec = A.ids != nullptr;
    // image-shaped launches walk the tree once per 8x4 tile (warp-cooperative); explicit ray batches per ray.
    // Instantiations: plain, recording, Russian roulette (+ recording only for the dense debug lists of explicit batches).
#define VP_FWD(TILE, RR, REC, SMEM) k_trace_forward<INTEG, KERNEL, D, TILE, RR, REC><<<blocks, TRACE_THREADS, SMEM, st>>>(S, P, A)
    if (P.image_width > 0) {
        if (rr && rec) VP_FWD(true, RF, true, TILE_SMEM);
        else if (rr) VP_FWD(true, RF, false, TILE_SMEM);
        else if (rec) VP_FWD(true, false, true, TILE_SMEM);
        else VP_FWD(true, false, false, TILE_SMEM);
    } else {
        if (rr && rec) VP_FWD(false, RF, true, TRACE_SMEM);
        else if (rr) VP_FWD(false, RF, false, TRACE_SMEM);
        else if (rec) VP_FWD(false, false, true, TRACE_SMEM);
        else VP_FWD(false, false, false, TRACE_SMEM);
    }
#undef VP_FWD
}

template <int INTEG, int KERNEL, int D>
void launch_adjoint(const DevScene &S, const vp_params &P, const TraceArgs &A, cudaStream_t st)
{
    int64_t blocks = (A.R + TRACE_THREADS - 1) / TRACE_THREADS;
    if (A.rec_offsets) {
        if constexpr (INTEG == VP_INTEGRATOR_RF)
            k_trace_adjoint<INTEG, KERNEL, D, ADJ_REPLAY_BUCKET, false><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(S, P, A);
        else
            k_trace_adjoint<INTEG, KERNEL, D, ADJ_REPLAY_ROWS, false><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(S, P, A);
    } else if (A.rec_ids) k_trace_adjoint<INTEG, KERNEL, D, ADJ_REPLAY_DENSE, false><<<(unsigned)blocks, TRACE_THREADS, 0, st>>>(S, P, A);
    else if (P.image_width > 0) k_trace_adjoint<INTEG, KERNEL, D, ADJ_WALK, true><<<(unsigned)blocks, TRACE_THREADS, TILE_SMEM, st>>>(S, P, A);
    else k_trace_adjoint<INTEG, KERNEL, D, ADJ_WALK, false><<<(unsigned)blocks, TRACE_THREADS, TRACE_SMEM, st>>>(S, P, A);
}

template <int INTEG, int KERNEL, int D>
void launch_gather(const GatherArgs &G, int64_t max_extra, cudaStream_t st)
{
    if constexpr (INTEG == VP_INTEGRATOR_RF) {
        const int64_t warps = G.p_end - G.p_begin;
        if (warps <= 0) return;
        k_adjoint_gather<KERNEL, D, false><<<(unsigned)((warps * 32 + 127) / 128), 128, 0, st>>>(G);
        if (max_extra > 0) k_adjoint_gather<KERNEL, D, true><<<(unsigned)((max_extra * 32 + 127) / 128), 128, 0, st>>>(G);
    }
}

// WHAT: 0 = forward, 1 = adjoint (ray-major), 2 = gather (primitive-major pass of the rf adjoint)
template <int WHAT>
int dispatch(vp_ctx *ctx, const DevScene &S, const vp_params &P, const TraceArgs &A, const GatherArgs *G, cudaStream_t st,
             int64_t max_extra = 0)
{
#define VP_LAUNCH(I, K, D)                                        \
    do {                                                          \
        if (WHAT == 0) launch_forward<I, K, D>(S, P, A, st);      \
        else if (WHAT == 1) launch_adjoint<I, K, D>(S, P, A, st); \
        else launch_gather<I, K, D>(*G, max_extra, st);           \
        return VP_OK;                                             \
    } while (0)
    const bool gauss = P.kernel == VP_KERNEL_GAUSSIAN;
    if (P.integrator == VP_INTEGRATOR_TOMO) {
        if (gauss) VP_LAUNCH(VP_INTEGRATOR_TOMO, VP_KERNEL_GAUSSIAN, -1);
        VP_LAUNCH(VP_INTEGRATOR_TOMO, VP_KERNEL_EPANECHNIKOV, -1);
    }
    switch (S.sh_degree) {
    case -1: if (gauss) VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_GAUSSIAN, -1); VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_EPANECHNIKOV, -1);
    case 0: if (gauss) VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_GAUSSIAN, 0); VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_EPANECHNIKOV, 0);
    case 1: if (gauss) VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_GAUSSIAN, 1); VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_EPANECHNIKOV, 1);
    case 2: if (gauss) VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_GAUSSIAN, 2); VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_EPANECHNIKOV, 2);
    case 3: if (gauss) VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_GAUSSIAN, 3); VP_LAUNCH(VP_INTEGRATOR_RF, VP_KERNEL_EPANECHNIKOV, 3);
    default: break;
    }
#undef VP_LAUNCH
    return vp_fail(ctx, VP_E_INVALID, "sh_coeffs: only SH degrees 0..3 (3, 12, 27 or 48 floats per primitive) are supported");
}

int check_common(vp_ctx *ctx, const vp_params *p, int64_t R, const char *who)
{
    if (!p) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": params is NULL");
    if (!ctx->built) return vp_fail(ctx, VP_E_STATE, std::string(who) + ": acceleration structure not built (call vp_build)");
    if (p->integrator != VP_INTEGRATOR_RF && p->integrator != VP_INTEGRATOR_TOMO)
        return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": unknown integrator");
    if (p->kernel != VP_KERNEL_GAUSSIAN && p->kernel != VP_KERNEL_EPANECHNIKOV)
        return vp_fail(ctx, VP_E_INVALID, "Unknown kernel type! Should be one of \"gaussian\" or \"epanechnikov\".");
    if (R < 0 || R > 0x7fffffffll * TRACE_THREADS) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": bad ray count");
    if (p->image_width > 0) {
        if (p->image_width % 8 || p->image_height % 4 || p->image_height <= 0 ||
            R % ((int64_t)p->image_width * p->image_height))
            return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": image hint needs width % 8 == 0, height % 4 == 0 and "
                                                                 "a ray count that is a multiple of width*height");
    }
    return VP_OK;
}

// gradient buffers take 64-bit (record) and 128-bit (colour coefficient) vector reductions in the scatter formulation
int check_grad_alignment(vp_ctx *ctx, const float *g_data, const float *g_sh, int sh_floats, const char *who)
{
    if ((uintptr_t)g_data % 8) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": g_data10 must be 8-byte aligned");
    if (g_sh && sh_floats % 4 == 0 && (uintptr_t)g_sh % 16)
        return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": g_sh must be 16-byte aligned");
    return VP_OK;
}

// Resolve a vp_ray_source into the kernels' RaySrc; *image_w / *image_h receive the tile-walker hint of a sensor.
int make_ray_src(vp_ctx *ctx, const vp_ray_source *rs, int64_t R, RaySrc &out, int &image_w, int &image_h, const char *who)
{
    if (!rs) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": rays is NULL");
    out = RaySrc{};
    image_w = image_h = 0;
    if (rs->camera) {
        const vp_camera &c = *rs->camera;
        if (c.width <= 0 || c.height <= 0 || rs->spp <= 0) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": bad film size or spp");
        const int rows = rs->row_count > 0 ? rs->row_count : c.height - rs->row_begin;
        if (rs->row_begin < 0 || rows <= 0 || rs->row_begin + rows > c.height)
            return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": row band outside the film");
        if (R != (int64_t)c.width * rows * rs->spp)
            return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": n_rays must equal width * rows * spp of the sensor");
        out.cam = c;
        const double th = std::tan((double)c.fov_x_deg * 0.5 * 0.017453292519943295);
        out.tan_half = (float)th;
        out.inv_w = (float)(1.0 / c.width);
        out.inv_h = (float)(1.0 / c.height);
        out.ly_scale = (float)(th * (double)c.height / (double)c.width);
        out.has_cam = 1;
        out.spp = rs->spp;
        out.jitter = rs->jitter;
        out.index_base = (int64_t)rs->row_begin * c.width * rs->spp;
        if (((int64_t)c.width * rs->spp) % 8 == 0 && rows % 4 == 0) { image_w = c.width * rs->spp; image_h = rows; }
    } else {
        if (R > 0 && (!rs->ray_o || !rs->ray_d)) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": ray_o and ray_d (or a camera) are required");
        out.o = rs->ray_o; out.d = rs->ray_d; out.maxt = rs->ray_maxt;
        out.spp = 1;
    }
    return VP_OK;
}

// the part of a ray source that belongs to rays [first, first + n) of the call
RaySrc slice_ray_src(const RaySrc &s, int64_t first)
{
    RaySrc o = s;
    if (s.has_cam) {
        o.index_base = s.index_base + first;
        if (s.jitter) o.jitter = s.jitter + 2 * first;
    } else {
        o.o = s.o + 3 * first; o.d = s.d + 3 * first;
        if (s.maxt) o.maxt = s.maxt + first;
    }
    return o;
}

int check_record(vp_ctx *ctx, const vp_hit_record *rec, const char *who)
{
    if (!rec->ids || !rec->total || (rec->dense ? (!rec->counts || !rec->state) : !rec->ray_offsets))
        return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": record arrays are missing (rows: ray_offsets, ids, total; dense: ids, "
                                                             "state, counts, total)");
    if (rec->capacity <= 0 || rec->capacity >= (1ll << 32) || rec->id_cap <= 0)
        return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": record needs 0 < capacity < 2^32 and id_cap > 0");
    if (rec->state && (uintptr_t)rec->state % 16) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": record state must be 16-byte aligned");
    return VP_OK;
}

}  // namespace

int vp_trace_forward_impl(vp_ctx *ctx, const vp_params *p, int64_t R, const float *o, const float *d, const float *maxt,
                          float *rgb, float *T, uint32_t *nhits, int32_t *ids, int32_t cap, int64_t rs, int64_t hs,
                          cudaStream_t st)
{
    int rc = check_common(ctx, p, R, "vp_trace_forward");
    if (rc) return rc;
    if (R == 0) return VP_OK;
    if (!o || !d || !rgb) return vp_fail(ctx, VP_E_INVALID, "vp_trace_forward: ray_o, ray_d and out_rgb are required");
    if (ids && cap <= 0) return vp_fail(ctx, VP_E_INVALID, "vp_trace_forward: out_hit_ids needs id_cap > 0");
    if ((rc = vp_ensure(ctx, ctx->stats, sizeof(vp_stats)))) return rc;
    VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->stats.ptr, 0, sizeof(vp_stats), st));
    DevScene S = vp_dev_scene(ctx);
    TraceArgs A = {};
    A.R = R; A.src.o = o; A.src.d = d; A.src.maxt = maxt; A.src.spp = 1;
    A.rgb = rgb; A.T = T; A.nhits = nhits;
    A.ids = ids; A.cap = ids ? cap : 0; A.rs = rs; A.hs = hs;
    A.stats = (vp_stats *)ctx->stats.ptr;
    rc = dispatch<0>(ctx, S, *p, A, nullptr, st);
    if (rc) return rc;
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}

int vp_trace_adjoint_impl(vp_ctx *ctx, const vp_params *p, int64_t R, const float *o, const float *d, const float *maxt,
                          const float *dL, const float *state_in, const int32_t *ids, const uint32_t *counts,
                          int32_t cap, int64_t rs, int64_t hs, float *g_data, float *g_attr, float *g_sh,
                          cudaStream_t st)
{
    int rc = check_common(ctx, p, R, "vp_trace_adjoint");
    if (rc) return rc;
    if (R == 0) return VP_OK;
    if (!o || !d || !dL || !state_in || !g_data || !g_attr)
        return vp_fail(ctx, VP_E_INVALID, "vp_trace_adjoint: rays, d_L, state_in, g_data10 and g_attr are required");
    if (ids && (!counts || cap <= 0)) return vp_fail(ctx, VP_E_INVALID, "vp_trace_adjoint: hit_ids needs hit_counts and id_cap > 0");
    if (p->integrator == VP_INTEGRATOR_RF && ctx->sh_floats > 0 && !g_sh)
        return vp_fail(ctx, VP_E_INVALID, "vp_trace_adjoint: g_sh is required when the primitives carry sh_coeffs");
    if ((rc = check_grad_alignment(ctx, g_data, g_sh, ctx->sh_floats, "vp_trace_adjoint"))) return rc;
    if ((rc = vp_ensure(ctx, ctx->stats, sizeof(vp_stats)))) return rc;
    VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->stats.ptr, 0, sizeof(vp_stats), st));
    DevScene S = vp_dev_scene(ctx);
    TraceArgs A = {};
    A.R = R; A.src.o = o; A.src.d = d; A.src.maxt = maxt; A.src.spp = 1;
    A.dL = dL; A.state_in = state_in;
    A.rec_ids = ids; A.rec_counts = counts; A.cap = cap; A.rs = rs; A.hs = hs;
    A.g_data = g_data; A.g_attr = g_attr; A.g_sh = g_sh;
    A.stats = (vp_stats *)ctx->stats.ptr;
    rc = dispatch<1>(ctx, S, *p, A, nullptr, st);
    if (rc) return rc;
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}

int vp_raygen_impl(vp_ctx *ctx, const vp_camera *cam, int32_t spp, const float *jitter, float *o, float *d, float *maxt,
                   cudaStream_t st)
{
    if (!cam || !o || !d) return vp_fail(ctx, VP_E_INVALID, "vp_raygen_perspective: camera, ray_o and ray_d are required");
    if (cam->width <= 0 || cam->height <= 0 || spp <= 0) return vp_fail(ctx, VP_E_INVALID, "vp_raygen_perspective: bad film size or spp");
    int64_t total = (int64_t)cam->width * cam->height * spp;
    vp_ray_source rs = {};
    rs.camera = cam; rs.spp = spp; rs.jitter = jitter;
    RaySrc src;
    int iw, ih;
    int rc = make_ray_src(ctx, &rs, total, src, iw, ih, "vp_raygen_perspective");
    if (rc) return rc;
    k_raygen<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, total, o, d, maxt);
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}

// ---- sensor-fused / record-replay entry points -------------------------------------------------------------------------
int vp_render_forward_impl(vp_ctx *ctx, const vp_params *p_in, const vp_ray_source *rays, int64_t R, float *rgb, float *T,
                           uint32_t *nhits, const vp_hit_record *rec, cudaStream_t st)
{
    if (!p_in) return vp_fail(ctx, VP_E_INVALID, "vp_render_forward: params is NULL");
    vp_params P = *p_in;
    RaySrc src;
    int iw, ih;
    int rc = make_ray_src(ctx, rays, R, src, iw, ih, "vp_render_forward");
    if (rc) return rc;
    if (src.has_cam) { P.image_width = iw; P.image_height = ih; }
    if ((rc = check_common(ctx, &P, R, "vp_render_forward"))) return rc;
    if (rec && (rc = check_record(ctx, rec, "vp_render_forward"))) return rc;
    if (R > 0 && !rgb) return vp_fail(ctx, VP_E_INVALID, "vp_render_forward: out_rgb is required");
    if ((rc = vp_ensure(ctx, ctx->stats, sizeof(vp_stats)))) return rc;
    VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->stats.ptr, 0, sizeof(vp_stats), st));
    DevScene S = vp_dev_scene(ctx);
    TraceArgs A = {};
    A.rgb = rgb; A.T = T; A.nhits = nhits;
    A.stats = (vp_stats *)ctx->stats.ptr;
    if (!rec) {
        if (R == 0) return VP_OK;
        A.R = R; A.src = src;
        if ((rc = dispatch<0>(ctx, S, P, A, nullptr, st))) return rc;
        VP_CUDA_CHECK(ctx, cudaGetLastError());
        return VP_OK;
    }
    VP_CUDA_CHECK(ctx, cudaMemsetAsync(rec->total, 0, 2 * sizeof(int64_t), st));
    if (rec->dense) {
        // dense hit-major record in the caller's buffers ([id_cap][n_rays] ids and (colour, T)): ONE launch, nothing is
        // moved afterwards, and the adjoint reads it coalesced.  5x the bytes of compressed rows, for the fastest step.
        if (P.integrator != VP_INTEGRATOR_RF) return vp_fail(ctx, VP_E_INVALID, "vp_render_forward: dense state records are a volprim_rf feature");
        if (R == 0) return VP_OK;
        uint32_t *nh = nhits;
        if (!nh) {
            if ((rc = vp_ensure(ctx, ctx->rec_nhits, sizeof(uint32_t) * (size_t)R))) return rc;
            nh = (uint32_t *)ctx->rec_nhits.ptr;
        }
        A.R = R; A.src = src; A.nhits = nh;
        A.ids = rec->ids; A.hit_state = (float4 *)rec->state; A.cap = rec->id_cap; A.rs = 1; A.hs = R;
        if ((rc = dispatch<0>(ctx, S, P, A, nullptr, st))) return rc;
        k_record_counts<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(nh, R, rec->id_cap, rec->counts, rec->total, true);
        VP_CUDA_CHECK(ctx, cudaGetLastError());
        return VP_OK;
    }
    // ---- recording: row bands bound the dense hit-major scratch ----
    if (R == 0) {
        VP_CUDA_CHECK(ctx, cudaMemsetAsync(rec->ray_offsets, 0, sizeof(int64_t), st));
        return VP_OK;
    }
    const int cap = rec->id_cap;
    const bool with_state = rec->state != nullptr && P.integrator == VP_INTEGRATOR_RF;
    int64_t band;                                     // rays per band
    const int64_t fit = ctx->record_scratch_bytes / ((int64_t)cap * (with_state ? 20 : 4));
    const int64_t per_image = P.image_width > 0 ? (int64_t)P.image_width * P.image_height : 0;
    int64_t rows_per_band = 0;
    if (per_image > 0) {
        rows_per_band = (fit / P.image_width) & ~3ll;
        if (rows_per_band < 4) rows_per_band = 4;
        if (rows_per_band > P.image_height) rows_per_band = P.image_height;
        band = rows_per_band * P.image_width;
    } else {
        band = fit & ~(int64_t)(TRACE_THREADS - 1);
        if (band < TRACE_THREADS) band = TRACE_THREADS;
        if (band > R) band = R;
    }
    if ((rc = vp_ensure(ctx, ctx->rec_dense, sizeof(int32_t) * (size_t)cap * (size_t)band))) return rc;
    if (with_state && (rc = vp_ensure(ctx, ctx->rec_dense_state, sizeof(float4) * (size_t)cap * (size_t)band))) return rc;
    if ((rc = vp_ensure(ctx, ctx->rec_counts, sizeof(uint32_t) * (size_t)band))) return rc;
    if ((rc = vp_ensure(ctx, ctx->scan_tmp, sizeof(uint64_t) * (size_t)(vpscan::n_tiles(band) + 1)))) return rc;
    uint32_t *nh_all = nhits;
    if (!nh_all) {
        if ((rc = vp_ensure(ctx, ctx->rec_nhits, sizeof(uint32_t) * (size_t)R))) return rc;
        nh_all = (uint32_t *)ctx->rec_nhits.ptr;
    }
    int64_t first = 0;
    while (first < R) {
        int64_t cnt = band;
        vp_params Pb = P;
        if (per_image > 0) {
            const int64_t in_img = first % per_image;               // bands never straddle two images
            const int64_t rows_left = (per_image - in_img) / P.image_width;
            const int64_t rows = rows_left < rows_per_band ? rows_left : rows_per_band;
            cnt = rows * P.image_width;
            Pb.image_height = (int32_t)rows;
        } else if (first + cnt > R) cnt = R - first;
        TraceArgs B = A;
        B.R = cnt;
        B.src = slice_ray_src(src, first);
        B.rgb = rgb + 3 * first;
        B.T = T ? T + first : nullptr;
        B.nhits = nh_all + first;
        B.ids = (int32_t *)ctx->rec_dense.ptr; B.cap = cap; B.rs = 1; B.hs = cnt;
        B.hit_state = with_state ? (float4 *)ctx->rec_dense_state.ptr : nullptr;
        if ((rc = dispatch<0>(ctx, S, Pb, B, nullptr, st))) return rc;
        const unsigned blocks = (unsigned)((cnt + 255) / 256), cblocks = (unsigned)((cnt + 127) / 128);
        k_record_counts<<<blocks, 256, 0, st>>>(B.nhits, cnt, cap, (uint32_t *)ctx->rec_counts.ptr, rec->total, false);
        vpscan::exclusive_scan<unsigned long long>((const uint32_t *)ctx->rec_counts.ptr, cnt,
                                                   (unsigned long long *)rec->ray_offsets + first,
                                                   (unsigned long long *)ctx->scan_tmp.ptr, (unsigned long long *)rec->total,
                                                   (unsigned long long *)rec->ray_offsets + first + cnt, st);
        k_compact_hits<<<cblocks, 128, 0, st>>>((const int32_t *)ctx->rec_dense.ptr, cnt, (const uint32_t *)ctx->rec_counts.ptr,
                                               rec->ray_offsets + first, rec->ids, rec->capacity);
        if (with_state)
            k_compact_state<<<(unsigned)((cnt + 63) / 64), 64, 0, st>>>((const float4 *)ctx->rec_dense_state.ptr, cnt,
                                                                         (const uint32_t *)ctx->rec_counts.ptr, rec->ray_offsets + first,
                                                                         (float4 *)rec->state, rec->capacity);
        first += cnt;
    }
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}

namespace {
int adjoint_common(vp_ctx *ctx, const vp_params *p_in, const vp_ray_source *rays, int64_t R, const vp_hit_record *rec,
                   float *g_data, float *g_attr, float *g_sh, const char *who, vp_params &P, RaySrc &src)
{
    if (!p_in) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": params is NULL");
    P = *p_in;
    int iw, ih;
    int rc = make_ray_src(ctx, rays, R, src, iw, ih, who);
    if (rc) return rc;
    // the ray-major replay walks the rays in 8x4 pixel tiles like the primal did (neighbouring rays replay the same
    // primitives: their records and colour blocks are shared through L1); explicit batches are indexed linearly
    P.image_width = src.has_cam ? iw : 0;
    P.image_height = src.has_cam ? ih : 0;
    if ((rc = check_common(ctx, &P, R, who))) return rc;
    if (!rec) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": a hit record is required (vp_trace_adjoint re-traces)");
    if ((rc = check_record(ctx, rec, who))) return rc;
    if (!g_data || !g_attr) return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": g_data10 and g_attr are required");
    if (P.integrator == VP_INTEGRATOR_RF && ctx->sh_floats > 0 && !g_sh)
        return vp_fail(ctx, VP_E_INVALID, std::string(who) + ": g_sh is required when the primitives carry sh_coeffs");
    return check_grad_alignment(ctx, g_data, g_sh, ctx->sh_floats, who);
}

GatherBuf gather_buf(vp_ctx *ctx, const vp_hit_record *rec)
{
    GatherBuf gb;
    gb.offsets = (const uint32_t *)ctx->adj_offsets.ptr;
    gb.rank = (const uint32_t *)ctx->adj_rank.ptr;
    gb.entries = (float4 *)ctx->adj_state.ptr;
    gb.total = rec->total;
    gb.capacity = rec->capacity;
    return gb;
}
}  // namespace

int vp_adjoint_begin_impl(vp_ctx *ctx, const vp_params *p_in, const vp_ray_source *rays, int64_t R, const float *dL,
                          const float *state_in, const vp_hit_record *rec, float *g_data, float *g_attr, float *g_sh,
                          cudaStream_t st)
{
    vp_params P;
    RaySrc src;
    int rc = adjoint_common(ctx, p_in, rays, R, rec, g_data, g_attr, g_sh, "vp_adjoint_begin", P, src);
    if (rc) return rc;
    if (R > 0 && (!dL || !state_in)) return vp_fail(ctx, VP_E_INVALID, "vp_adjoint_begin: d_L and state_in are required");
    if (R >= (1ll << 32)) return vp_fail(ctx, VP_E_INVALID, "vp_adjoint_begin: more than 2^32 rays per call are not supported");
    if ((rc = vp_ensure(ctx, ctx->stats, sizeof(vp_stats)))) return rc;
    VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->stats.ptr, 0, sizeof(vp_stats), st));
    DevScene S = vp_dev_scene(ctx);
    TraceArgs A = {};
    A.R = R; A.src = src; A.dL = dL; A.state_in = state_in;
    A.rec_ids = rec->ids; A.rec_offsets = rec->ray_offsets;
    A.g_data = g_data; A.g_attr = g_attr; A.g_sh = g_sh;
    A.stats = (vp_stats *)ctx->stats.ptr;
    if (P.integrator == VP_INTEGRATOR_RF) {
        // counting pass: bucket sizes + the slot of every record entry inside its bucket, bucket offsets, and the
        // extra work items of the buckets that exceed one warp's chunk
        const int64_t n = ctx->n;
        const int64_t max_extra = rec->capacity / GATHER_CHUNK + 1;
        if ((rc = vp_ensure(ctx, ctx->adj_offsets, sizeof(uint32_t) * (size_t)(n + 1)))) return rc;
        if ((rc = vp_ensure(ctx, ctx->adj_extra, sizeof(uint32_t) * (size_t)(n + 1)))) return rc;
        if ((rc = vp_ensure(ctx, ctx->adj_items, sizeof(uint32_t) * (size_t)max_extra))) return rc;
        // (a dense record addresses its rank array like its ids: [id_cap][n_rays])
        const int64_t rank_slots = rec->dense ? (int64_t)rec->id_cap * R : rec->capacity;
        if ((rc = vp_ensure(ctx, ctx->adj_rank, sizeof(uint32_t) * (size_t)(rank_slots > 0 ? rank_slots : 1)))) return rc;
        if ((rc = vp_ensure(ctx, ctx->adj_state, 2 * sizeof(float4) * (size_t)rec->capacity))) return rc;
        if ((rc = vp_ensure(ctx, ctx->scan_tmp, sizeof(uint64_t) * (size_t)(vpscan::n_tiles(n) + 1)))) return rc;
        uint32_t *offsets = (uint32_t *)ctx->adj_offsets.ptr, *extra = (uint32_t *)ctx->adj_extra.ptr;
        VP_CUDA_CHECK(ctx, cudaMemsetAsync(offsets, 0, sizeof(uint32_t) * (size_t)(n + 1), st));
        // debugging aid (VOLPRIM_POISON=1): every bucket slot starts as NaN, so a slot the ray pass failed to write shows
        // up as a NaN gradient instead of silently reusing what an earlier pass left there
        static const bool poison = std::getenv("VOLPRIM_POISON") != nullptr;
        if (poison) VP_CUDA_CHECK(ctx, cudaMemsetAsync(ctx->adj_state.ptr, 0xff, 2 * sizeof(float4) * (size_t)rec->capacity, st));
        if (n > 0) {
            int64_t blocks = (rec->capacity + 255) / 256;
            if (blocks > 148 * 32) blocks = 148 * 32;
            if (rec->dense)
                k_bucket_ranks_dense<<<(unsigned)((R + 255) / 256), 256, 0, st>>>(rec->ids, rec->counts, R, rec->total, rec->capacity,
                                                                                 offsets, (uint32_t *)ctx->adj_rank.ptr);
            else
                k_bucket_ranks<<<(unsigned)blocks, 256, 0, st>>>(rec->ids, rec->total, rec->capacity, offsets, (uint32_t *)ctx->adj_rank.ptr);
            vpscan::exclusive_scan<uint32_t>(offsets, n, offsets, (uint32_t *)ctx->scan_tmp.ptr, nullptr, offsets + n, st);
            k_extra_counts<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(offsets, n, extra);
            vpscan::exclusive_scan<uint32_t>(extra, n, extra, (uint32_t *)ctx->scan_tmp.ptr, nullptr, extra + n, st);
            k_extra_items<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(extra, n, (uint32_t *)ctx->adj_items.ptr, (uint32_t)max_extra);
        }
    }
    A.gb = gather_buf(ctx, rec);
    if (R == 0) return VP_OK;
    if (P.integrator == VP_INTEGRATOR_RF && rec->dense) {
        k_adjoint_rows_dense<<<(unsigned)((R + 127) / 128), 128, 0, st>>>(src, R, P.image_width, P.image_height, dL, state_in, rec->counts,
                                                                         rec->ids, (const float4 *)rec->state, A.gb);
        VP_CUDA_CHECK(ctx, cudaGetLastError());
        return VP_OK;
    }
    if (P.integrator == VP_INTEGRATOR_RF && rec->state) {
        k_adjoint_rows<<<(unsigned)((R + 32 * ROWS_WARPS - 1) / (32 * ROWS_WARPS)), 32 * ROWS_WARPS, 0, st>>>(
            src, R, dL, state_in, rec->ray_offsets, rec->ids, (const float4 *)rec->state, A.gb);
        VP_CUDA_CHECK(ctx, cudaGetLastError());
        return VP_OK;
    }
    if ((rc = dispatch<1>(ctx, S, P, A, nullptr, st))) return rc;
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}

int vp_adjoint_finish_impl(vp_ctx *ctx, const vp_params *p_in, const vp_ray_source *rays, int64_t R, const vp_hit_record *rec,
                           int64_t p_begin, int64_t p_end, float *g_data, float *g_attr, float *g_sh, cudaStream_t st)
{
    vp_params P;
    RaySrc src;
    int rc = adjoint_common(ctx, p_in, rays, R, rec, g_data, g_attr, g_sh, "vp_adjoint_finish", P, src);
    if (rc) return rc;
    if (P.integrator != VP_INTEGRATOR_RF) return VP_OK;   // volprim_tomography scattered everything in vp_adjoint_begin
    if (p_begin < 0 || p_end > ctx->n || p_begin > p_end) return vp_fail(ctx, VP_E_INVALID, "vp_adjoint_finish: bad primitive range");
    if (!ctx->adj_offsets.ptr || !ctx->adj_state.ptr) return vp_fail(ctx, VP_E_STATE, "vp_adjoint_finish: call vp_adjoint_begin first");
    if (p_begin == p_end) return VP_OK;
    DevScene S = vp_dev_scene(ctx);
    GatherArgs G = {};
    G.src = src;
    G.data10 = (const float *)ctx->raw_data.ptr;
    G.attr = ctx->have_attr ? (const float *)ctx->raw_attr.ptr : nullptr;
    G.gb = gather_buf(ctx, rec);
    G.p_begin = p_begin; G.p_end = p_end;
    G.extra_items = (const uint32_t *)ctx->adj_items.ptr;
    G.extra_offsets = (const uint32_t *)ctx->adj_extra.ptr;
    G.n_extra = G.extra_offsets + ctx->n;
    G.g_data = g_data; G.g_attr = g_attr; G.g_sh = g_sh;
    TraceArgs A = {};
    if ((rc = dispatch<2>(ctx, S, P, A, &G, st, rec->capacity / GATHER_CHUNK + 1))) return rc;
    VP_CUDA_CHECK(ctx, cudaGetLastError());
    return VP_OK;
}
