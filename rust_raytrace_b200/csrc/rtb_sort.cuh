// rtb_sort.cuh — the two data-parallel primitives the BVH builder needs, hand-written for sm_100a:
//   * stable LSD radix sort of (64-bit key, 32-bit value) pairs, 8 bits per pass
//   * exclusive prefix sum of 32- or 64-bit integers
// They replace cub::DeviceRadixSort / cub::DeviceScan (round-1 interim).  Both are HBM-streaming kernels: a sort pass
// reads and writes 12 bytes per pair twice (histogram + scatter), a scan reads and writes each element once plus the
// per-block totals; at the sizes of interest (<= a few million primitives, <= 24 MB per array) they are launch- and
// latency-bound, not bandwidth-bound.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "rtb_internal.cuh"

namespace rtbsort {

constexpr int RS_THREADS = 256;              // 8 warps per block
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_PER_WARP = 512;             // consecutive keys ranked by one warp, 32 at a time, in order (stability)
constexpr int RS_TILE = RS_WARPS * RS_PER_WARP;   // 4096 keys per block
constexpr int RS_DIGITS = 256;

__device__ __forceinline__ uint32_t digit_of(unsigned long long key, int shift) { return (uint32_t)(key >> shift) & 255u; }

// Per-warp digit counts of the warp's sub-tile into wh[RS_DIGITS] (shared, warp-private).  Lanes with equal digits
// are grouped with __match_any_sync, the lowest lane of each group adds the group size.
__device__ __forceinline__ void warp_histogram(const unsigned long long* __restrict__ keys, uint32_t begin, uint32_t end,
                                               int shift, uint32_t* wh, unsigned lane) {
    for (uint32_t i = begin + lane; i - lane < end; i += 32u) {          // whole warp iterates together
        const bool valid = i < end;
        const uint32_t d = valid ? digit_of(keys[i], shift) : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid && (unsigned)(__ffs(peers) - 1) == lane) wh[d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
}

// pass kernel 1: per-block digit histogram -> hist[d * n_blocks + block]
static __global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const unsigned long long* __restrict__ keys, uint32_t n, int shift,
                                                       uint32_t n_blocks, uint32_t* __restrict__ hist) {
    __shared__ uint32_t wh[RS_WARPS][RS_DIGITS];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * RS_DIGITS; d += RS_THREADS) (&wh[0][0])[d] = 0u;
    __syncthreads();
    const uint32_t begin = min(n, blockIdx.x * (uint32_t)RS_TILE + warp * (uint32_t)RS_PER_WARP);
    const uint32_t end = min(n, begin + (uint32_t)RS_PER_WARP);
    warp_histogram(keys, begin, end, shift, wh[warp], lane);
    __syncthreads();
    for (int d = threadIdx.x; d < RS_DIGITS; d += RS_THREADS) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) s += wh[w][d];
        hist[(size_t)d * n_blocks + blockIdx.x] = s;
    }
}

// pass kernel 3: stable scatter.  offs = exclusive scan of hist (global start of (digit, block)).
static __global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const unsigned long long* __restrict__ keys_in,
                                                          const uint32_t* __restrict__ vals_in, uint32_t n, int shift,
                                                          uint32_t n_blocks, const uint32_t* __restrict__ offs,
                                                          unsigned long long* __restrict__ keys_out,
                                                          uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t wh[RS_WARPS][RS_DIGITS];    // first: per-warp counts, then: running output cursor per (warp, digit)
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < RS_WARPS * RS_DIGITS; d += RS_THREADS) (&wh[0][0])[d] = 0u;
    __syncthreads();
    const uint32_t begin = min(n, blockIdx.x * (uint32_t)RS_TILE + warp * (uint32_t)RS_PER_WARP);
    const uint32_t end = min(n, begin + (uint32_t)RS_PER_WARP);
    warp_histogram(keys_in, begin, end, shift, wh[warp], lane);
    __syncthreads();
    // cursor(warp, d) = global start of (d, block) + counts of the lower warps of this block
    for (int d = threadIdx.x; d < RS_DIGITS; d += RS_THREADS) {
        uint32_t run = offs[(size_t)d * n_blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) { const uint32_t c = wh[w][d]; wh[w][d] = run; run += c; }
    }
    __syncthreads();
    uint32_t* cur = wh[warp];
    for (uint32_t i = begin + lane; i - lane < end; i += 32u) {
        const bool valid = i < end;
        unsigned long long k = 0ull;
        uint32_t d = 0xffffffffu;
        if (valid) { k = keys_in[i]; d = digit_of(k, shift); }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        if (valid) {
            const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
            const uint32_t dst = cur[d] + rank;
            keys_out[dst] = k;
            vals_out[dst] = vals_in[i];
        }
        __syncwarp();
        if (valid && (unsigned)(__ffs(peers) - 1) == lane) cur[d] += (uint32_t)__popc(peers);
        __syncwarp();
    }
}

// ---- exclusive sum -----------------------------------------------------------------------------------------
constexpr int SC_THREADS = 256;
constexpr int SC_ITEMS = 8;                        // per thread
constexpr int SC_TILE = SC_THREADS * SC_ITEMS;     // 2048 elements per block

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* smem /* [SC_THREADS / 32] */) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    T incl = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const T up = __shfl_up_sync(0xffffffffu, incl, off);
        if ((int)lane >= off) incl += up;
    }
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (unsigned)(blockDim.x >> 5) ? smem[lane] : T(0);
        const T w0 = w;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const T up = __shfl_up_sync(0xffffffffu, w, off);
            if ((int)lane >= off) w += up;
        }
        if (lane < (unsigned)(blockDim.x >> 5)) smem[lane] = w - w0;
        if (lane == 31) *total = w;
    }
    __syncthreads();
    const T r = smem[warp] + incl - v;
    __syncthreads();
    return r;
}

// level 1: exclusive scan inside each 2048-element tile, tile total -> totals[block]
template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_scan_tiles(const T* __restrict__ in, T* __restrict__ out, uint32_t n,
                                                          T* __restrict__ totals) {
    __shared__ T smem[SC_THREADS / 32];
    __shared__ T tot;
    const uint32_t base = blockIdx.x * (uint32_t)SC_TILE + threadIdx.x * (uint32_t)SC_ITEMS;
    T v[SC_ITEMS];
    T sum = T(0);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { v[k] = (base + k < n) ? in[base + k] : T(0); sum += v[k]; }
    T run = block_exclusive_scan<T>(sum, &tot, smem);
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) { if (base + k < n) out[base + k] = run; run += v[k]; }
    if (threadIdx.x == 0) totals[blockIdx.x] = tot;
}

// level 2: one block scans the tile totals in place (exclusive), looping when there are more than 2048 of them
template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_scan_totals(T* __restrict__ totals, uint32_t n_tiles) {
    __shared__ T smem[SC_THREADS / 32];
    __shared__ T tot;
    T carry = T(0);
    for (uint32_t t0 = 0; t0 < n_tiles; t0 += SC_TILE) {
        const uint32_t base = t0 + threadIdx.x * (uint32_t)SC_ITEMS;
        T v[SC_ITEMS];
        T sum = T(0);
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) { v[k] = (base + k < n_tiles) ? totals[base + k] : T(0); sum += v[k]; }
        T run = carry + block_exclusive_scan<T>(sum, &tot, smem);
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) { if (base + k < n_tiles) totals[base + k] = run; run += v[k]; }
        carry += tot;
        __syncthreads();
    }
}

// level 3: add the scanned tile totals
template <typename T>
__global__ void __launch_bounds__(SC_THREADS) k_scan_add(T* __restrict__ out, uint32_t n, const T* __restrict__ totals) {
    const T add = totals[blockIdx.x];
    const uint32_t base = blockIdx.x * (uint32_t)SC_TILE + threadIdx.x * (uint32_t)SC_ITEMS;
#pragma unroll
    for (int k = 0; k < SC_ITEMS; ++k) if (base + k < n) out[base + k] += add;
}

inline uint32_t scan_tiles(uint32_t n) { return (n + SC_TILE - 1) / SC_TILE; }
// bytes of scratch exclusive_sum needs for n elements of T
template <typename T>
inline size_t scan_tmp_bytes(uint32_t n) { return sizeof(T) * (size_t)(scan_tiles(n) + 1); }

// out[i] = sum of in[0..i) ; in and out may alias.  3 launches (1 when n fits one tile and 2 otherwise would save
// little at these sizes).  Returns the number of kernels launched.
template <typename T>
inline int exclusive_sum(const T* in, T* out, uint32_t n, void* tmp, cudaStream_t stream) {
    if (n == 0) return 0;
    T* totals = (T*)tmp;
    const uint32_t tiles = scan_tiles(n);
    k_scan_tiles<T><<<tiles, SC_THREADS, 0, stream>>>(in, out, n, totals);
    if (tiles == 1) return 1;
    k_scan_totals<T><<<1, SC_THREADS, 0, stream>>>(totals, tiles);
    k_scan_add<T><<<tiles, SC_THREADS, 0, stream>>>(out, n, totals);
    return 3;
}

inline uint32_t sort_blocks(uint32_t n) { return (n + RS_TILE - 1) / RS_TILE; }
// scratch: histogram (256 x blocks) + the scan's own scratch
inline size_t sort_tmp_bytes(uint32_t n) {
    const size_t h = sizeof(uint32_t) * (size_t)RS_DIGITS * sort_blocks(n);
    return ((h + 255) & ~(size_t)255) + scan_tmp_bytes<uint32_t>((uint32_t)(RS_DIGITS * sort_blocks(n)));
}

// Stable sort of (key, value) pairs by bits [0, key_bits) of the key.  The result ends up in (keys_b, vals_b) when the
// number of passes is odd and in (keys_a, vals_a) otherwise; *result_in_b says which.  Returns launches.
inline int radix_sort_pairs(unsigned long long* keys_a, unsigned long long* keys_b, uint32_t* vals_a, uint32_t* vals_b,
                            uint32_t n, int key_bits, void* tmp, cudaStream_t stream, bool* result_in_b) {
    *result_in_b = false;
    if (n == 0) return 0;
    const uint32_t blocks = sort_blocks(n);
    uint32_t* hist = (uint32_t*)tmp;
    void* scan_tmp = (char*)tmp + ((sizeof(uint32_t) * (size_t)RS_DIGITS * blocks + 255) & ~(size_t)255);
    int launches = 0;
    bool in_a = true;
    for (int shift = 0; shift < key_bits; shift += 8) {
        unsigned long long* ki = in_a ? keys_a : keys_b; unsigned long long* ko = in_a ? keys_b : keys_a;
        uint32_t* vi = in_a ? vals_a : vals_b; uint32_t* vo = in_a ? vals_b : vals_a;
        k_rs_hist<<<blocks, RS_THREADS, 0, stream>>>(ki, n, shift, blocks, hist);
        launches += 1 + exclusive_sum<uint32_t>(hist, hist, (uint32_t)RS_DIGITS * blocks, scan_tmp, stream);
        k_rs_scatter<<<blocks, RS_THREADS, 0, stream>>>(ki, vi, n, shift, blocks, hist, ko, vo);
        ++launches;
        in_a = !in_a;
    }
    *result_in_b = !in_a;
    return launches;
}

}  // namespace rtbsort
