// vp_scan.cuh -- multi-block exclusive prefix sum over uint32 counts (internal).
//
// Used by the LBVH radix sort (digit histograms), by the hit-record compaction (per-ray hit counts -> CSR offsets)
// and by the gather adjoint (per-primitive hit counts -> bucket offsets).  Three launches: per-tile sums, one block
// scanning the tile sums (with an optional device-side carry so that consecutive calls continue one running total),
// per-tile scan + offset.  Every element is read twice and written once: a few microseconds at these sizes.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vpscan {

constexpr int THREADS = 256;
constexpr int ITEMS = 16;
constexpr int TILE = THREADS * ITEMS;

template <class T>
__device__ __forceinline__ T warp_inclusive(T v, int lane)
{
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, v, off);
        if (lane >= off) v += t;
    }
    return v;
}

// block-wide exclusive scan of one value per thread; returns the exclusive prefix, `total` = block sum
template <class T, int NT>
__device__ __forceinline__ T block_exclusive(T v, T &total)
{
    __shared__ T warp_sums[NT / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const T incl = warp_inclusive(v, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    T before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) {
        const T s = warp_sums[w];
        if (w < wid) before += s;
        all += s;
    }
    __syncthreads();   // warp_sums may be reused by the caller's next call
    total = all;
    return before + incl - v;
}

template <class OUT>
__global__ void __launch_bounds__(THREADS) k_tile_sums(const uint32_t *__restrict__ in, int64_t n, OUT *__restrict__ tile_sums)
{
    const int64_t base = (int64_t)blockIdx.x * TILE;
    OUT s = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int64_t i = base + (int64_t)k * THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    OUT total;
    block_exclusive<OUT, THREADS>(s, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile sums in place; *carry (nullable) is added to every offset and then advanced
// by the grand total; *total_out (nullable) receives carry + grand total
template <class OUT>
__global__ void __launch_bounds__(1024) k_scan_tile_sums(OUT *__restrict__ tile_sums, int n_tiles, OUT *__restrict__ carry,
                                                         OUT *__restrict__ total_out)
{
    __shared__ OUT run_s;
    if (threadIdx.x == 0) run_s = carry ? *carry : (OUT)0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const OUT v = i < n_tiles ? tile_sums[i] : (OUT)0;
        OUT total;
        const OUT excl = block_exclusive<OUT, 1024>(v, total);
        const OUT run = run_s;
        if (i < n_tiles) tile_sums[i] = run + excl;
        __syncthreads();
        if (threadIdx.x == 0) run_s = run + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (carry) *carry = run_s;
        if (total_out) *total_out = run_s;
    }
}

// per-tile exclusive scan + tile offset; `out` may alias `in` when OUT is uint32_t (every thread reads its own 16
// items before it writes them)
template <class OUT>
__global__ void __launch_bounds__(THREADS) k_tile_scan(const uint32_t *in, int64_t n, const OUT *__restrict__ tile_offsets,
                                                       OUT *out)
{
    const int64_t base = (int64_t)blockIdx.x * TILE + (int64_t)threadIdx.x * ITEMS;
    uint32_t v[ITEMS];
    OUT s = 0;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0u;
        s += v[k];
    }
    OUT total;
    OUT excl = block_exclusive<OUT, THREADS>(s, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
}

inline int64_t n_tiles(int64_t n) { return (n + TILE - 1) / TILE; }

// out[i] = carry + sum_{j<i} in[j]   (i < n);  tile_tmp needs n_tiles(n) elements of OUT.
// carry / total_out are optional DEVICE scalars (see k_scan_tile_sums).
template <class OUT>
inline void exclusive_scan(const uint32_t *in, int64_t n, OUT *out, OUT *tile_tmp, OUT *carry, OUT *total_out, cudaStream_t st)
{
    if (n <= 0) {
        if (total_out) k_scan_tile_sums<OUT><<<1, 1024, 0, st>>>(tile_tmp, 0, carry, total_out);
        return;
    }
    const int nt = (int)n_tiles(n);
    k_tile_sums<OUT><<<nt, THREADS, 0, st>>>(in, n, tile_tmp);
    k_scan_tile_sums<OUT><<<1, 1024, 0, st>>>(tile_tmp, nt, carry, total_out);
    k_tile_scan<OUT><<<nt, THREADS, 0, st>>>(in, n, tile_tmp, out);
}

}  // namespace vpscan
