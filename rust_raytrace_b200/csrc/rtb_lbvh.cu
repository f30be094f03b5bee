// rtb_lbvh.cu — GPU Morton-code LBVH builder (replaces the reference's octree build,
// build_bounding_box raytrace.rs:790-845, with an accelerator that returns the same
// closest hit; see DESIGN.md for the equivalence argument).
//
// Pipeline (all on one stream; host round trips only to read counts that size the next allocations / grids):
//   k_prim_bounds   per-primitive AABB from `corners`, scene AABB by atomic min/max
//   k_split_refs    reference splitting: long primitives enter as several clipped references (count, scan, emit)
//   k_morton        63-bit Morton code of the AABB centre
//   radix sort      (key = morton, value = reference): 8 stable LSD passes of 8 bits, rtb_sort.cuh
//   topology        k_sah_level_small/big (binned-SAH top-down, default) | k_ploc_* (PLOC) | k_hierarchy (Karras 2012
//                   radix tree, also the fallback for over-deep trees): children, parent, covered range per node
//   k_refit         bottom-up AABB union with one atomic arrival counter per internal node
//   k_node_kind     subtrees of <= RTB_LEAF_MAX references become leaves unless the SAH prefers the split
//   k_emit_nodes    32-byte BVH2 nodes with adjacent sibling pairs (one-kernel renderers, rtb_scene_download_bvh)
//   k_emit_tris     gather the 19 intersect floats + shading record of each reference into leaf order
//   k_cost4/k_mark4/k_emit_nodes4   4-wide collapse chosen by dynamic programming, 128-byte BVH4 nodes (wavefront renderer)
//   k_c8_level/k_emit_tris8         8-wide compressed collapse, 80-byte nodes, one level per launch (A/B, only with RTB_BVH8=1)

#include <algorithm>
#include <cfloat>
#include <limits>
#include <string>
#include <vector>
#include <cstdlib>

#include "rtb_internal.cuh"
#include "rtb_sort.cuh"

namespace {

// Order-preserving float <-> uint mapping for atomicMin/Max on floats.
__device__ __forceinline__ uint32_t f2o(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float o2f(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    return __uint_as_float(u);
}

struct BuildScratch {
    uint32_t scene_lo[3];   // ordered-uint encoded
    uint32_t scene_hi[3];
    uint32_t max_abs;       // ordered-uint of max |coordinate|
    uint32_t height;        // tree height (edges from the root to the deepest Karras leaf)
    uint32_t max_leaf;
    uint32_t n_leaves;
    uint32_t depth4;        // depth of the deepest 4-wide node (root = 0)
    uint32_t stack_need;    // exact worst-case BVH4 traversal stack: max over root-to-node paths of sum (entries - 1); 0 = not computed
};

__global__ void k_init_scratch(BuildScratch* s) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { s->scene_lo[k] = 0xffffffffu; s->scene_hi[k] = 0u; }
        s->max_abs = 0u; s->height = 0u; s->max_leaf = 0u; s->n_leaves = 0u; s->depth4 = 0u; s->stack_need = 0u;
    }
}

__global__ void k_prim_bounds(const RtbTriangle* __restrict__ tris, const uint32_t* __restrict__ keep, uint32_t n,
                              float4* __restrict__ plo, float4* __restrict__ phi, BuildScratch* s) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const RtbTriangle& t = tris[keep[i]];
        for (int c = 0; c < 3; ++c)
            for (int k = 0; k < 3; ++k) {
                float v = t.corners[3 * c + k];
                lo[k] = fminf(lo[k], v);
                hi[k] = fmaxf(hi[k], v);
            }
        plo[i] = make_float4(lo[0], lo[1], lo[2], 0.f);
        phi[i] = make_float4(hi[0], hi[1], hi[2], 0.f);
    }
    // warp-level reduce, then one atomic per warp
    for (int k = 0; k < 3; ++k) {
        float l = lo[k], h = hi[k];
        for (int off = 16; off > 0; off >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, off));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, off));
        }
        if ((threadIdx.x & 31) == 0 && l <= h) {
            atomicMin(&s->scene_lo[k], f2o(l));
            atomicMax(&s->scene_hi[k], f2o(h));
            atomicMax(&s->max_abs, f2o(fmaxf(fabsf(l), fabsf(h))));
        }
    }
}


// ---- reference splitting (early split clipping) -------------------------------------------------------
// A primitive whose AABB is long compared to the scene (the fan triangles of make_disk, raytrace.rs:531-592: radius-long
// slivers on a tilted plane, every one's AABB covering a good part of the disk) is entered into the tree as several
// REFERENCES, each with the tight bounds of the triangle clipped to one cell of a recursive midpoint split of its AABB
// along the longest axis.  The exact test still runs on the whole triangle, so a reference that is found twice gives
// the same t and the closest-hit result (min t, lowest index on ties) is unchanged; only the boxes get tighter.
// Measured on the 4K teapot frame (CPU experiment tools/experiments/bvh_quality.cpp, then B200): 6,720 primitives ->
// 7,1xx references, exact triangle tests per bounce ray -43 %, per primary ray -42 %, node visits +3 %.
constexpr int SPLIT_MAX_DEPTH = 6;      // <= 64 references per primitive

// Bounds of (triangle ∩ box) by Sutherland-Hodgman clipping against the six planes; false when they do not overlap
// in an area (touching in a point or an edge: the neighbouring cell, whose closed half-space contains it, keeps it).
__device__ bool clip_tri_bounds(const float* __restrict__ c, const float* blo, const float* bhi, float* olo, float* ohi) {
    float P[10][3], Q[10][3];
    int np = 3;
    for (int v = 0; v < 3; ++v) for (int k = 0; k < 3; ++k) P[v][k] = c[3 * v + k];
    for (int ax = 0; ax < 3; ++ax)
        for (int side = 0; side < 2; ++side) {
            const float pos = side ? bhi[ax] : blo[ax];
            int nq = 0;
            for (int i = 0; i < np; ++i) {
                const int j = (i + 1 == np) ? 0 : i + 1;
                const float da = P[i][ax] - pos, db = P[j][ax] - pos;
                const bool ia = side ? (da <= 0.f) : (da >= 0.f), ib = side ? (db <= 0.f) : (db >= 0.f);
                if (ia && nq < 10) { for (int k = 0; k < 3; ++k) Q[nq][k] = P[i][k]; ++nq; }
                if (ia != ib && nq < 10) {
                    const float t = da / (da - db);
                    for (int k = 0; k < 3; ++k) Q[nq][k] = P[i][k] + (P[j][k] - P[i][k]) * t;
                    Q[nq][ax] = pos;
                    ++nq;
                }
            }
            np = nq;
            if (np < 3) return false;
            for (int i = 0; i < np; ++i) for (int k = 0; k < 3; ++k) P[i][k] = Q[i][k];
        }
    for (int k = 0; k < 3; ++k) { olo[k] = FLT_MAX; ohi[k] = -FLT_MAX; }
    for (int i = 0; i < np; ++i)
        for (int k = 0; k < 3; ++k) { olo[k] = fminf(olo[k], P[i][k]); ohi[k] = fmaxf(ohi[k], P[i][k]); }
    for (int k = 0; k < 3; ++k) { olo[k] = fmaxf(olo[k], blo[k]); ohi[k] = fminf(ohi[k], bhi[k]); }
    return true;
}

// EMIT = false: count the references of every primitive; EMIT = true: write them (same walk, same arithmetic).
template <bool EMIT>
__global__ void k_split_refs(const RtbTriangle* __restrict__ tris, const uint32_t* __restrict__ keep, uint32_t n,
                             const float4* __restrict__ plo, const float4* __restrict__ phi,
                             const BuildScratch* __restrict__ s, float inv_div, uint32_t* __restrict__ count,
                             const uint32_t* __restrict__ offset, float4* __restrict__ rlo, float4* __restrict__ rhi,
                             uint32_t* __restrict__ rkeep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float ext = 0.f;
    for (int k = 0; k < 3; ++k) ext = fmaxf(ext, o2f(s->scene_hi[k]) - o2f(s->scene_lo[k]));
    const float max_len = ext * inv_div;
    const float4 l = plo[i], h = phi[i];
    uint32_t out = EMIT ? offset[i] : 0u;
    uint32_t made = 0;
    if (fmaxf(fmaxf(h.x - l.x, h.y - l.y), h.z - l.z) <= max_len || (tris[keep[i]].kind & RTB_PRIM_SPHERE)) {   // a sphere's corners are its AABB, not a triangle
        if (EMIT) { rlo[out] = l; rhi[out] = h; rkeep[out] = keep[i]; }
        made = 1;
    } else {
        const float* c = tris[keep[i]].corners;
        float slo[SPLIT_MAX_DEPTH + 2][3], shi[SPLIT_MAX_DEPTH + 2][3];
        int sdepth[SPLIT_MAX_DEPTH + 2];
        int sp = 0;
        slo[0][0] = l.x; slo[0][1] = l.y; slo[0][2] = l.z; shi[0][0] = h.x; shi[0][1] = h.y; shi[0][2] = h.z;
        sdepth[0] = 0; sp = 1;
        while (sp > 0) {
            --sp;
            float blo[3], bhi[3], tlo[3], thi[3];
            for (int k = 0; k < 3; ++k) { blo[k] = slo[sp][k]; bhi[k] = shi[sp][k]; }
            const int depth = sdepth[sp];
            if (!clip_tri_bounds(c, blo, bhi, tlo, thi)) continue;
            int ax = 0;
            float len = thi[0] - tlo[0];
            if (thi[1] - tlo[1] > len) { ax = 1; len = thi[1] - tlo[1]; }
            if (thi[2] - tlo[2] > len) { ax = 2; len = thi[2] - tlo[2]; }
            if (depth >= SPLIT_MAX_DEPTH || len <= max_len) {
                if (EMIT) {
                    rlo[out + made] = make_float4(tlo[0], tlo[1], tlo[2], 0.f);
                    rhi[out + made] = make_float4(thi[0], thi[1], thi[2], 0.f);
                    rkeep[out + made] = keep[i];
                }
                ++made;
                continue;
            }
            const float mid = tlo[ax] + 0.5f * len;
            for (int k = 0; k < 3; ++k) { slo[sp][k] = tlo[k]; shi[sp][k] = thi[k]; slo[sp + 1][k] = tlo[k]; shi[sp + 1][k] = thi[k]; }
            shi[sp][ax] = mid; slo[sp + 1][ax] = mid;
            sdepth[sp] = depth + 1; sdepth[sp + 1] = depth + 1;
            sp += 2;
        }
        if (made == 0) {   // numerically degenerate primitive: keep its plain box
            if (EMIT) { rlo[out] = l; rhi[out] = h; rkeep[out] = keep[i]; }
            made = 1;
        }
    }
    if (!EMIT) count[i] = made;
}

__device__ __forceinline__ uint64_t spread21(uint32_t x) {
    uint64_t v = x & 0x1fffffu;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void k_morton(const float4* __restrict__ plo, const float4* __restrict__ phi, uint32_t n,
                         const BuildScratch* __restrict__ s, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float code[3];
    float4 l = plo[i], h = phi[i];
    float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
    for (int k = 0; k < 3; ++k) {
        float slo = o2f(s->scene_lo[k]), shi = o2f(s->scene_hi[k]);
        float ext = shi - slo;
        float u = ext > 0.f ? (c[k] - slo) / ext : 0.f;
        code[k] = fminf(fmaxf(u * 2097152.0f, 0.f), 2097151.0f);
    }
    keys[i] = (spread21((uint32_t)code[0]) << 2) | (spread21((uint32_t)code[1]) << 1) | spread21((uint32_t)code[2]);
    vals[i] = i;
}

// Karras 2012.  Internal nodes 0..n-2, leaves are referred to as n-1+k.
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                            int2* __restrict__ range, int* __restrict__ parent) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? (n - 1 + gamma) : gamma;
    int right = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi);
    parent[left] = i;
    parent[right] = i;
    if (i == 0) parent[0] = -1;
}


// ---------------------------------------------------------------------------------------------------------
// PLOC — parallel locally-ordered clustering (Meister & Bittner 2018) on the Morton-sorted primitives
// (RTB_BUILDER=ploc; the default of most of round 1, faster to build than the binned-SAH tree, 6-18 % slower to trace).  Clusters sit in Morton order; every cluster looks PLOC_R places left and right for the
// neighbour whose union box has the smallest surface area; mutual nearest neighbours merge into a new node; the
// array is compacted; repeat until one cluster is left.  Compared with the plain radix tree (k_hierarchy) the result
// needs 14 % fewer BVH4 node visits per bounce ray on the teapot scene and 22-25 % fewer on the 1 M-triangle field
// (CPU experiment tools/experiments/bvh_quality.cpp; a full binned-SAH build gets 15 % / 37 %).
//
// PLOC node ids: leaves 0..n-1 (Morton order), internal n..2n-2 in creation order, root = 2n-2.  k_ploc_finish
// relabels to the convention the rest of the builder uses (internal 0..n-2 with root 0, leaf n-1+k where k is now
// the DFS position), so that every node again covers a contiguous primitive range.
// ---------------------------------------------------------------------------------------------------------
struct PlocState { uint32_t m; uint32_t n_internal; };

__device__ __forceinline__ float union_area(float4 al, float4 ah, float4 bl, float4 bh) {
    const float dx = fmaxf(ah.x, bh.x) - fminf(al.x, bl.x);
    const float dy = fmaxf(ah.y, bh.y) - fminf(al.y, bl.y);
    const float dz = fmaxf(ah.z, bh.z) - fminf(al.z, bl.z);
    return dx * dy + dy * dz + dz * dx;
}

__global__ void k_ploc_init(uint32_t n, const uint32_t* __restrict__ sorted_vals, const float4* __restrict__ plo,
                            const float4* __restrict__ phi, float4* __restrict__ lo, float4* __restrict__ hi,
                            uint32_t* __restrict__ cluster, uint32_t* __restrict__ size, int* __restrict__ parent,
                            PlocState* st) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) { st->m = n; st->n_internal = 0u; }
    if (k >= n) return;
    const uint32_t prim = sorted_vals[k];
    lo[k] = plo[prim]; hi[k] = phi[prim];
    cluster[k] = k; size[k] = 1u; parent[k] = -1;
}

constexpr int PLOC_BLOCK = 256;
constexpr int PLOC_R_MAX = 32;

// nearest neighbour of every cluster within +-R places; the candidate boxes of a block are staged in shared memory
__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_nn(const uint32_t* __restrict__ cluster, const PlocState* st,
                                                       const float4* __restrict__ lo, const float4* __restrict__ hi,
                                                       int R, uint32_t* __restrict__ nn) {
    __shared__ float4 slo[PLOC_BLOCK + 2 * PLOC_R_MAX], shi[PLOC_BLOCK + 2 * PLOC_R_MAX];
    const int m = (int)st->m;
    const int b0 = blockIdx.x * PLOC_BLOCK;
    if (b0 >= m) return;
    for (int t = threadIdx.x; t < PLOC_BLOCK + 2 * R; t += PLOC_BLOCK) {
        const int j = b0 - R + t;
        if (j >= 0 && j < m) { const uint32_t c = cluster[j]; slo[t] = lo[c]; shi[t] = hi[c]; }
    }
    __syncthreads();
    const int i = b0 + threadIdx.x;
    if (i >= m) return;
    const float4 al = slo[threadIdx.x + R], ah = shi[threadIdx.x + R];
    float best = FLT_MAX;
    int bj = -1;
    for (int j = max(0, i - R); j <= min(m - 1, i + R); ++j) {
        if (j == i) continue;
        const int t = j - b0 + R;
        const float a = union_area(al, ah, slo[t], shi[t]);
        if (a < best) { best = a; bj = j; }       // ties: the lowest index
    }
    nn[i] = (uint32_t)bj;
}

// flags for the compaction scan: low word = "stays in the array", high word = "creates a node"
__global__ void k_ploc_flags(const uint32_t* __restrict__ nn, const PlocState* st, unsigned long long* __restrict__ flags,
                             uint32_t m_upper) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = st->m;
    if (i >= m) { if (i < m_upper) flags[i] = 0ull; return; }   // the host's bound on m may be a few rounds old
    const uint32_t j = nn[i];
    const bool mutual = (j < m) && nn[j] == i;
    flags[i] = mutual ? (i < j ? ((1ull << 32) | 1ull) : 0ull) : 1ull;
}

__global__ void k_ploc_apply(uint32_t n, const uint32_t* __restrict__ cin, const uint32_t* __restrict__ nn,
                             const unsigned long long* __restrict__ pos, const PlocState* st_in, PlocState* st_out,
                             uint32_t* __restrict__ cout, int2* __restrict__ pchildren, int* __restrict__ parent,
                             float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* __restrict__ size) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t m = st_in->m, base = st_in->n_internal;
    if (i >= m) return;
    const uint32_t j = nn[i];
    const bool mutual = (j < m) && nn[j] == i;
    const unsigned long long p = pos[i];
    const uint32_t out = (uint32_t)p, created = (uint32_t)(p >> 32);
    uint32_t keeps = 1u, creates = 0u;
    if (mutual) {
        if (i < j) {
            const uint32_t a = cin[i], b = cin[j];
            const uint32_t id = n + base + created;
            pchildren[id - n] = make_int2((int)a, (int)b);
            parent[a] = (int)id; parent[b] = (int)id; parent[id] = -1;
            const float4 al = lo[a], ah = hi[a], bl = lo[b], bh = hi[b];
            lo[id] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
            hi[id] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
            size[id] = size[a] + size[b];
            cout[out] = id;
            creates = 1u;
        } else {
            keeps = 0u;
        }
    } else {
        cout[out] = cin[i];
    }
    if (i == m - 1) { st_out->m = out + keeps; st_out->n_internal = base + created + creates; }
}


// The last rounds (<= PLOC_TAIL clusters) in ONE block: nearest neighbours, mutual-pair merge and compaction loop in
// shared memory with __syncthreads between the phases, instead of four launches and a host round trip per round —
// for the 6,720-triangle teapot scene that is 15 of the ~19 rounds.
constexpr int PLOC_TAIL = 1024;
__global__ void __launch_bounds__(PLOC_TAIL) k_ploc_tail(uint32_t n, const uint32_t* __restrict__ cin, const PlocState* st_in,
                                                        PlocState* st_out, int R, int2* __restrict__ pchildren,
                                                        int* __restrict__ parent, float4* __restrict__ lo,
                                                        float4* __restrict__ hi, uint32_t* __restrict__ size) {
    __shared__ uint32_t C[2][PLOC_TAIL];
    __shared__ uint32_t nn[PLOC_TAIL];
    __shared__ uint32_t wsum_keep[32], wsum_make[32];
    __shared__ uint32_t tot_keep, tot_make;
    const int i = threadIdx.x, lane = i & 31, warp = i >> 5;
    int m = (int)st_in->m;
    uint32_t base = st_in->n_internal;
    if (i < m) C[0][i] = cin[i];
    int cur = 0;
    __syncthreads();
    while (m > 1) {
        // nearest neighbour
        uint32_t bj = 0xffffffffu;
        if (i < m) {
            const uint32_t ci = C[cur][i];
            const float4 al = lo[ci], ah = hi[ci];
            float best = FLT_MAX;
            for (int j = max(0, i - R); j <= min(m - 1, i + R); ++j) {
                if (j == i) continue;
                const uint32_t cj = C[cur][j];
                const float a = union_area(al, ah, lo[cj], hi[cj]);
                if (a < best) { best = a; bj = (uint32_t)j; }
            }
            nn[i] = bj;
        }
        __syncthreads();
        const bool mutual = (i < m) && bj < (uint32_t)m && nn[bj] == (uint32_t)i;
        const uint32_t make = (mutual && (uint32_t)i < bj) ? 1u : 0u;
        const uint32_t keep = (i < m) ? ((mutual && !make) ? 0u : 1u) : 0u;
        // block-wide exclusive scan of (keep, make)
        uint32_t sk = keep, sm = make;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t uk = __shfl_up_sync(0xffffffffu, sk, off), um = __shfl_up_sync(0xffffffffu, sm, off);
            if (lane >= off) { sk += uk; sm += um; }
        }
        if (lane == 31) { wsum_keep[warp] = sk; wsum_make[warp] = sm; }
        __syncthreads();
        if (warp == 0) {
            uint32_t a = wsum_keep[lane], b = wsum_make[lane];
            const uint32_t a0 = a, b0 = b;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t ua = __shfl_up_sync(0xffffffffu, a, off), ub = __shfl_up_sync(0xffffffffu, b, off);
                if (lane >= off) { a += ua; b += ub; }
            }
            wsum_keep[lane] = a - a0; wsum_make[lane] = b - b0;       // exclusive warp offsets
            if (lane == 31) { tot_keep = a; tot_make = b; }
        }
        __syncthreads();
        const uint32_t out = wsum_keep[warp] + sk - keep, created = wsum_make[warp] + sm - make;
        if (i < m) {
            if (make) {
                const uint32_t a = C[cur][i], b = C[cur][bj];
                const uint32_t id = n + base + created;
                pchildren[id - n] = make_int2((int)a, (int)b);
                parent[a] = (int)id; parent[b] = (int)id; parent[id] = -1;
                const float4 al = lo[a], ah = hi[a], bl = lo[b], bh = hi[b];
                lo[id] = make_float4(fminf(al.x, bl.x), fminf(al.y, bl.y), fminf(al.z, bl.z), 0.f);
                hi[id] = make_float4(fmaxf(ah.x, bh.x), fmaxf(ah.y, bh.y), fmaxf(ah.z, bh.z), 0.f);
                size[id] = size[a] + size[b];
                C[cur ^ 1][out] = id;
            } else if (keep) {
                C[cur ^ 1][out] = C[cur][i];
            }
        }
        __syncthreads();
        m = (int)tot_keep;
        base += tot_make;
        cur ^= 1;
        __syncthreads();
    }
    if (i == 0) { st_out->m = 1u; st_out->n_internal = base; }
}

// PLOC ids -> the builder's node convention + DFS primitive order.  One thread per internal PLOC node.
__global__ void k_ploc_finish(uint32_t n, const int2* __restrict__ pchildren, const int* __restrict__ pparent,
                              const uint32_t* __restrict__ size, const uint32_t* __restrict__ sorted_vals,
                              int2* __restrict__ children, int2* __restrict__ range, int* __restrict__ parent,
                              uint32_t* __restrict__ vals_out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n - 1) return;
    const uint32_t x = n + t;
    // DFS offset of x: sizes of the left siblings on the way to the root
    uint32_t off = 0;
    for (uint32_t c = x;;) {
        const int p = pparent[c];
        if (p < 0) break;
        const int2 ch = pchildren[(uint32_t)p - n];
        if ((uint32_t)ch.y == c) off += size[ch.x];
        c = (uint32_t)p;
    }
    const int2 ch = pchildren[t];
    const uint32_t l = (uint32_t)ch.x, r = (uint32_t)ch.y;
    const uint32_t off_l = off, off_r = off + size[l];
    const int kid = (int)(2u * n - 2u - x);
    const int kl = l < n ? (int)(n - 1u + off_l) : (int)(2u * n - 2u - l);
    const int kr = r < n ? (int)(n - 1u + off_r) : (int)(2u * n - 2u - r);
    children[kid] = make_int2(kl, kr);
    range[kid] = make_int2((int)off, (int)(off + size[x] - 1u));
    parent[kl] = kid; parent[kr] = kid;
    if (kid == 0) parent[0] = -1;
    if (l < n) vals_out[off_l] = sorted_vals[l];
    if (r < n) vals_out[off_r] = sorted_vals[r];
}


// ---------------------------------------------------------------------------------------------------------
// Binned-SAH top-down build (the default topology builder): the quality yardstick of tools/experiments/bvh_quality.cpp on
// the GPU.  4K teapot frame 2.28 -> 2.13 ms, 1 M-triangle field 1.35 -> 1.11 ms against PLOC (RTB_BUILDER=ploc).
// Breadth-first, one kernel per tree level, one WARP per node: centroid bounds -> 16 bins per axis (shared-memory
// atomics) -> the 45 candidate planes evaluated by 45 lanes (cost = A(L) n_L + A(R) n_R) -> stable partition of the
// node's reference range from one index buffer into the other (left block, then right block; n_L is known from the
// bin counts) -> two children.  Splitting goes down to single references; the SAH leaf rule of k_node_kind then
// collapses subtrees of <= RTB_LEAF_MAX as for the other builders.  Output is in the builder's node convention
// (internal ids 0..n-2 handed out by an atomic counter with root 0, leaf id n-1+k for the reference at position k),
// every node covers a contiguous range.  The top levels run on few warps (a 1 M-reference root is one warp looping
// 31 K times), so this build takes longer than PLOC; it is not part of the frame time (main.rs:160 vs :191).
// ---------------------------------------------------------------------------------------------------------
constexpr int SAH_BINS = 16;
constexpr int SAH_CAND = 3 * (SAH_BINS - 1);
constexpr int SAH_BIG_TEAM = 512;        // threads per node for nodes of more than SAH_BIG references (one CTA each)
constexpr uint32_t SAH_BIG = 2048;
constexpr uint32_t SAH_SAMPLES = 16384;  // big nodes of >= 2 x this many references bin a sample of about this size
constexpr int SAH_WARPS = 4;             // small nodes: one warp each, this many per CTA
struct SahItem { int node; uint32_t lo, hi; };     // references [lo, hi) of the level's input index buffer
struct SahState { uint32_t n_next_small, n_next_big, next_internal, pad; };
struct SahBins {
    uint32_t cnt[3][SAH_BINS];
    uint32_t lo[3][SAH_BINS][3], hi[3][SAH_BINS][3];     // ordered-uint encoded box per bin
    float cost[SAH_CAND]; uint32_t nl[SAH_CAND];
    float red[6][SAH_BIG_TEAM / 32];                      // cross-warp reduction of the centroid bounds
    uint32_t wl[SAH_BIG_TEAM / 32], wr[SAH_BIG_TEAM / 32];
    int best; uint32_t n_left;
};

__device__ __forceinline__ int sah_bin(float c, float cmin, float k) {     // k = SAH_BINS / extent
    const int b = (int)((c - cmin) * k);
    return b < 0 ? 0 : (b > SAH_BINS - 1 ? SAH_BINS - 1 : b);
}

// One node, split by a team of TEAM threads (a warp, or a whole CTA of SAH_BIG_TEAM threads for the big nodes near the root).
template <int TEAM>
__device__ void sah_split_node(const SahItem item, unsigned tid, SahBins& sb, const uint32_t* __restrict__ src,
                               uint32_t* __restrict__ dst, const float4* __restrict__ plo, const float4* __restrict__ phi,
                               uint32_t n, SahItem* __restrict__ next_small, SahItem* __restrict__ next_big, SahState* st,
                               int2* __restrict__ children, int2* __restrict__ range, int* __restrict__ parent,
                               uint32_t* __restrict__ leaf_vals) {
    const unsigned FULLM = 0xffffffffu;
    const unsigned lane = tid & 31u, wid = tid >> 5;
    constexpr int NW = TEAM / 32;
    auto team_sync = [&]() { if (TEAM == 32) __syncwarp(); else __syncthreads(); };
    const uint32_t lo = item.lo, hi = item.hi, count = hi - lo;
    // Very big nodes choose their plane from a sample (every `stride`-th reference, >= SAH_SAMPLES of them): the two
    // binning passes then cost next to nothing and the node is read once, by the partition.  The sampled centroid
    // bounds need not contain every centroid: sah_bin clamps, and the partition uses the same function.
    const uint32_t stride = (TEAM > 32 && count >= 2u * SAH_SAMPLES) ? count / SAH_SAMPLES : 1u;

    // centroid bounds
    float cmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, cmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = lo + tid * stride; i < hi; i += TEAM * stride) {
        const uint32_t r = src[i];
        const float4 l = plo[r], h = phi[r];
        const float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
        for (int a = 0; a < 3; ++a) { cmin[a] = fminf(cmin[a], c[a]); cmax[a] = fmaxf(cmax[a], c[a]); }
    }
    for (int a = 0; a < 3; ++a)
        for (int off = 16; off > 0; off >>= 1) {
            cmin[a] = fminf(cmin[a], __shfl_xor_sync(FULLM, cmin[a], off));
            cmax[a] = fmaxf(cmax[a], __shfl_xor_sync(FULLM, cmax[a], off));
        }
    if (TEAM > 32) {
        if (lane == 0) for (int a = 0; a < 3; ++a) { sb.red[a][wid] = cmin[a]; sb.red[3 + a][wid] = cmax[a]; }
        __syncthreads();
        for (int a = 0; a < 3; ++a)
            for (int k = 0; k < NW; ++k) { cmin[a] = fminf(cmin[a], sb.red[a][k]); cmax[a] = fmaxf(cmax[a], sb.red[3 + a][k]); }
    }
    float kbin[3];
    for (int a = 0; a < 3; ++a) kbin[a] = cmax[a] > cmin[a] ? (float)SAH_BINS / (cmax[a] - cmin[a]) : 0.f;

    // binning
    for (int i = tid; i < 3 * SAH_BINS; i += TEAM) {
        (&sb.cnt[0][0])[i] = 0u;
        for (int k = 0; k < 3; ++k) { (&sb.lo[0][0][0])[3 * i + k] = 0xffffffffu; (&sb.hi[0][0][0])[3 * i + k] = 0u; }
    }
    team_sync();
    for (uint32_t i = lo + tid * stride; i < hi; i += TEAM * stride) {
        const uint32_t r = src[i];
        const float4 l = plo[r], h = phi[r];
        const float c[3] = {0.5f * (l.x + h.x), 0.5f * (l.y + h.y), 0.5f * (l.z + h.z)};
        const uint32_t el[3] = {f2o(l.x), f2o(l.y), f2o(l.z)}, eh[3] = {f2o(h.x), f2o(h.y), f2o(h.z)};
        for (int a = 0; a < 3; ++a) {
            if (kbin[a] == 0.f) continue;
            const int b = sah_bin(c[a], cmin[a], kbin[a]);
            atomicAdd(&sb.cnt[a][b], 1u);
            for (int k = 0; k < 3; ++k) { atomicMin(&sb.lo[a][b][k], el[k]); atomicMax(&sb.hi[a][b][k], eh[k]); }
        }
    }
    team_sync();

    // the 3 x (SAH_BINS - 1) candidate planes: cost = A(L) n_L + A(R) n_R
    for (int cand = tid; cand < SAH_CAND; cand += TEAM) {
        const int a = cand / (SAH_BINS - 1), k = cand % (SAH_BINS - 1);     // left = bins 0..k
        float area[2] = {0.f, 0.f};
        uint32_t cnt[2] = {0u, 0u};
        for (int side = 0; side < 2; ++side) {
            uint32_t blo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, bhi[3] = {0u, 0u, 0u};
            for (int b = side ? k + 1 : 0; b <= (side ? SAH_BINS - 1 : k); ++b) {
                if (!sb.cnt[a][b]) continue;
                cnt[side] += sb.cnt[a][b];
                for (int q = 0; q < 3; ++q) { blo[q] = min(blo[q], sb.lo[a][b][q]); bhi[q] = max(bhi[q], sb.hi[a][b][q]); }
            }
            if (cnt[side]) {
                const float dx = o2f(bhi[0]) - o2f(blo[0]), dy = o2f(bhi[1]) - o2f(blo[1]), dz = o2f(bhi[2]) - o2f(blo[2]);
                area[side] = dx * dy + dy * dz + dz * dx;
            }
        }
        const bool valid = kbin[a] != 0.f && cnt[0] && cnt[1];
        sb.cost[cand] = valid ? area[0] * (float)cnt[0] + area[1] * (float)cnt[1] : FLT_MAX;
        sb.nl[cand] = cnt[0];
    }
    team_sync();
    if (tid == 0) {
        int best = -1;
        float bc = FLT_MAX;
        for (int c = 0; c < SAH_CAND; ++c) if (sb.cost[c] < bc) { bc = sb.cost[c]; best = c; }
        sb.best = best;
        sb.n_left = best >= 0 ? sb.nl[best] : count / 2u;
    }
    team_sync();
    const int best = sb.best;
    uint32_t n_left = sb.n_left;

    // stable partition of [lo, hi) from src into dst: left block, then right block
    if (best < 0) {                       // all centroids coincide: split in the middle, order unchanged
        for (uint32_t i = lo + tid; i < hi; i += TEAM) dst[i] = src[i];
    } else {
        const int a = best / (SAH_BINS - 1), k = best % (SAH_BINS - 1);
        // a warp knows n_left from the bin counts and writes two forward blocks (stable); a CTA (whose counts may come
        // from a sample) fills the right block from the back and counts n_left as it goes
        uint32_t base_l = lo, base_r = TEAM > 32 ? 0u : lo + n_left;
        for (uint32_t i0 = lo; i0 < hi; i0 += TEAM) {
            const uint32_t i = i0 + tid;
            uint32_t r = 0;
            bool left = false;
            if (i < hi) {
                r = src[i];
                const float4 l = plo[r], h = phi[r];
                const float c = a == 0 ? 0.5f * (l.x + h.x) : (a == 1 ? 0.5f * (l.y + h.y) : 0.5f * (l.z + h.z));
                left = sah_bin(c, cmin[a], kbin[a]) <= k;
            }
            const unsigned ml = __ballot_sync(FULLM, i < hi && left), mr = __ballot_sync(FULLM, i < hi && !left);
            const unsigned lt = (1u << lane) - 1u;
            uint32_t off_l = 0, off_r = 0, tot_l = __popc(ml), tot_r = __popc(mr);
            if (TEAM > 32) {
                if (lane == 0) { sb.wl[wid] = tot_l; sb.wr[wid] = tot_r; }
                __syncthreads();
                tot_l = tot_r = 0;
                for (int q = 0; q < NW; ++q) {
                    if (q < (int)wid) { off_l += sb.wl[q]; off_r += sb.wr[q]; }
                    tot_l += sb.wl[q]; tot_r += sb.wr[q];
                }
            }
            if (i < hi) {
                const uint32_t rr = base_r + off_r + __popc(mr & lt);
                dst[left ? base_l + off_l + __popc(ml & lt) : (TEAM > 32 ? hi - 1u - rr : rr)] = r;
            }
            base_l += tot_l; base_r += tot_r;
            if (TEAM > 32) __syncthreads();
        }
        if (TEAM > 32) n_left = base_l - lo;
    }
    __threadfence_block();
    team_sync();

    if (tid == 0) {
        const uint32_t mid = lo + n_left;
        int ids[2];
        const uint32_t clo[2] = {lo, mid}, chi[2] = {mid, hi};
        for (int c = 0; c < 2; ++c) {
            const uint32_t cc = chi[c] - clo[c];
            if (cc == 1u) {
                ids[c] = (int)(n - 1u + clo[c]);
                leaf_vals[clo[c]] = dst[clo[c]];
            } else {
                ids[c] = (int)atomicAdd(&st->next_internal, 1u);
                SahItem ni; ni.node = ids[c]; ni.lo = clo[c]; ni.hi = chi[c];
                if (cc > SAH_BIG) next_big[atomicAdd(&st->n_next_big, 1u)] = ni;
                else next_small[atomicAdd(&st->n_next_small, 1u)] = ni;
            }
            parent[ids[c]] = item.node;
        }
        children[item.node] = make_int2(ids[0], ids[1]);
        range[item.node] = make_int2((int)lo, (int)hi - 1);
        if (item.node == 0) parent[0] = -1;
    }
}

__global__ void __launch_bounds__(32 * SAH_WARPS)
k_sah_level_small(const SahItem* __restrict__ items, uint32_t n_items, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                  const float4* __restrict__ plo, const float4* __restrict__ phi, uint32_t n, SahItem* __restrict__ next_small,
                  SahItem* __restrict__ next_big, SahState* st, int2* __restrict__ children, int2* __restrict__ range,
                  int* __restrict__ parent, uint32_t* __restrict__ leaf_vals) {
    __shared__ SahBins sb[SAH_WARPS];
    const uint32_t it = blockIdx.x * SAH_WARPS + (threadIdx.x >> 5);
    if (it >= n_items) return;
    sah_split_node<32>(items[it], threadIdx.x & 31u, sb[threadIdx.x >> 5], src, dst, plo, phi, n, next_small, next_big, st, children,
                       range, parent, leaf_vals);
}
__global__ void __launch_bounds__(SAH_BIG_TEAM)
k_sah_level_big(const SahItem* __restrict__ items, uint32_t n_items, const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                const float4* __restrict__ plo, const float4* __restrict__ phi, uint32_t n, SahItem* __restrict__ next_small,
                SahItem* __restrict__ next_big, SahState* st, int2* __restrict__ children, int2* __restrict__ range,
                int* __restrict__ parent, uint32_t* __restrict__ leaf_vals) {
    __shared__ SahBins sb;
    if (blockIdx.x >= n_items) return;
    sah_split_node<SAH_BIG_TEAM>(items[blockIdx.x], threadIdx.x, sb, src, dst, plo, phi, n, next_small, next_big, st, children, range,
                                 parent, leaf_vals);
}

// Bottom-up refit.  blo/bhi are indexed by Karras node id (internal 0..n-2, leaf n-1+k).
__global__ void k_refit(const float4* __restrict__ plo, const float4* __restrict__ phi,
                        const uint32_t* __restrict__ sorted_vals, int n, const int2* __restrict__ children,
                        const int* __restrict__ parent, float4* blo, float4* bhi, uint32_t* arrive,
                        BuildScratch* s) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t prim = sorted_vals[k];
    int node = n - 1 + k;
    blo[node] = plo[prim];
    bhi[node] = phi[prim];
    __threadfence();
    uint32_t depth = 0;
    int p = parent[node];
    while (p >= 0) {
        ++depth;
        if (atomicAdd(&arrive[p], 1u) == 0u) break;  // the sibling subtree is not finished yet
        __threadfence();
        int2 ch = children[p];
        float4 l0 = __ldcg(&blo[ch.x]), l1 = __ldcg(&blo[ch.y]);
        float4 h0 = __ldcg(&bhi[ch.x]), h1 = __ldcg(&bhi[ch.y]);
        blo[p] = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), 0.f);
        bhi[p] = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        __threadfence();
        p = parent[p];
    }
    (void)depth;
    (void)s;
}

__global__ void k_leaf_depth(int n, const int* __restrict__ parent, BuildScratch* s) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t depth = 0;
    for (int p = parent[n - 1 + k]; p >= 0; p = parent[p]) ++depth;
    for (int off = 16; off > 0; off >>= 1) depth = max(depth, __shfl_xor_sync(0xffffffffu, depth, off));
    if ((threadIdx.x & 31) == 0) atomicMax(&s->height, depth);
}

// What becomes of every node of the binary tree when subtrees of <= RTB_LEAF_MAX primitives are collapsed into
// leaves.  A small subtree is collapsed unless the surface-area heuristic prefers to keep its top split
// (cost of one more box test, 1 x A(node), against the triangle tests it saves: A(l)|l| + A(r)|r| vs A(node)|node|);
// on the teapot scene that rule trades +3 % node visits for -17 % exact triangle tests, on the 1 M field -50 %.
enum : uint8_t { KIND_INTERNAL = 0, KIND_LEAF = 1, KIND_SWALLOWED = 2 };

__device__ __forceinline__ float box_area(float4 l, float4 h) {
    const float dx = h.x - l.x, dy = h.y - l.y, dz = h.z - l.z;
    return dx * dy + dy * dz + dz * dx;
}

__device__ __forceinline__ bool collapses(int c, int n, const int2* __restrict__ children, const int2* __restrict__ range,
                                          const float4* __restrict__ blo, const float4* __restrict__ bhi, int sah) {
    if (c >= n - 1) return true;                                   // a single primitive
    const int size = range[c].y - range[c].x + 1;
    if (size > RTB_LEAF_MAX) return false;
    if (!sah) return true;
    const int2 ch = children[c];
    const int sl = ch.x >= n - 1 ? 1 : range[ch.x].y - range[ch.x].x + 1;
    const int sr = ch.y >= n - 1 ? 1 : range[ch.y].y - range[ch.y].x + 1;
    const float a = box_area(blo[c], bhi[c]);
    // sah = cost of one more box test in percent of an exact triangle test (RTB_SAH_LEAVES; 1 means 100)
    const float c_node = sah == 1 ? 1.0f : 0.01f * (float)sah;
    const float split = a * c_node + box_area(blo[ch.x], bhi[ch.x]) * (float)sl + box_area(blo[ch.y], bhi[ch.y]) * (float)sr;
    return !(split < a * (float)size);
}

__global__ void k_node_kind(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                            const int* __restrict__ parent, const float4* __restrict__ blo,
                            const float4* __restrict__ bhi, int sah, uint8_t* __restrict__ kind) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * n - 1) return;
    // swallowed iff some ancestor collapses; only ancestors of <= RTB_LEAF_MAX primitives can
    bool swallowed = false;
    for (int p = parent[c]; p >= 0 && (range[p].y - range[p].x + 1) <= RTB_LEAF_MAX; p = parent[p])
        if (collapses(p, n, children, range, blo, bhi, sah)) { swallowed = true; break; }
    kind[c] = swallowed ? KIND_SWALLOWED : (collapses(c, n, children, range, blo, bhi, sah) ? KIND_LEAF : KIND_INTERNAL);
}

__global__ void k_split_flags(const uint8_t* __restrict__ kind, int n_internal, uint32_t* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    flags[i] = kind[i] == KIND_INTERNAL ? 1u : 0u;
}

__device__ __forceinline__ void write_node(float4* nodes, uint32_t pos, float4 lo, float4 hi, float pad, uint32_t a,
                                           uint32_t b) {
    nodes[2 * pos + 0] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, __uint_as_float(a));
    nodes[2 * pos + 1] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, __uint_as_float(b));
}

// One thread per Karras node (internal and leaf).  A node is emitted iff its parent covers more
// than RTB_LEAF_MAX primitives; sibling pairs land at 2 + 2*slot(parent) and +1.
__global__ void k_emit_nodes(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                             const int* __restrict__ parent, const float4* __restrict__ blo,
                             const float4* __restrict__ bhi, const uint32_t* __restrict__ slot,
                             const uint8_t* __restrict__ kind, float4* __restrict__ nodes, BuildScratch* s) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int total = 2 * n - 1;
    if (c >= total) return;
    if (kind[c] == KIND_SWALLOWED) return;                      // inside a collapsed leaf
    const float pad = o2f(s->max_abs) * (1.0f / 131072.0f);   // 2^-17 of the largest |coordinate|
    int first, count;
    if (c >= n - 1) { first = c - (n - 1); count = 1; }
    else { first = range[c].x; count = range[c].y - range[c].x + 1; }
    uint32_t pos;
    if (c == 0) pos = 0;
    else {
        int p = parent[c];
        pos = 2u + 2u * slot[p] + (children[p].y == c ? 1u : 0u);
    }
    if (kind[c] == KIND_LEAF) {
        write_node(nodes, pos, blo[c], bhi[c], pad, (uint32_t)first, (uint32_t)count);
        atomicMax(&s->max_leaf, (uint32_t)count);
        atomicAdd(&s->n_leaves, 1u);
    } else {
        write_node(nodes, pos, blo[c], bhi[c], pad, 2u + 2u * slot[c], 0u);
    }
    if (c == 0) {  // padding node 1: an empty box that can never be reached
        nodes[2] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, __uint_as_float(0u));
        nodes[3] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, __uint_as_float(0u));
    }
}

// ---- 4-wide collapse -------------------------------------------------------------------------------
// A BVH4 node is rooted at the root and at every internal node that covers more than RTB_LEAF_MAX primitives
// and sits at EVEN depth of the binary tree; its (up to 4) entries are its grandchildren, or a child
// itself where that child is already a leaf.  Node = 8 x float4 (128 B, one cache line), SoA over the
// entries: c.x[4] e.x[4] c.y[4] e.y[4] c.z[4] e.z[4] code[4] pad — child boxes as centre and half extent (RTB_NODE_CE, rtb_internal.cuh;
// lo / hi planes with RTB_NODE_CE=0), an empty slot's box all NaN.  code: 0 = empty slot,
// 0x80000000 | first<<3 | count = leaf, otherwise the index of the child BVH4 node.
__global__ void k_flag4(int n_internal, const uint8_t* __restrict__ kind, const int* __restrict__ parent,
                        uint32_t* __restrict__ flags4, BuildScratch* s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    uint32_t depth = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) ++depth;
    const bool kept = (i == 0) || (kind[i] == KIND_INTERNAL && (depth & 1u) == 0u);
    flags4[i] = kept ? 1u : 0u;
    if (kept) atomicMax(&s->depth4, depth >> 1);
}


// ---- 4-wide collapse by dynamic programming (default; RTB_COLLAPSE=even selects the rule above) -----------------------
// Which descendants of a BVH2 node v become the (up to four) entries of its BVH4 node is chosen to minimise
//   cost4(v) = A(v) + min over cuts C of v's subtree, |C| <= 4, of  sum_{c in C} cost4(c),   cost4(leaf) = 1.2 A(leaf) |leaf|
// (one node visit per unit of area hit, 1.2 per exact triangle test).  The eight cuts with at most four members are
// enumerated bottom-up (k_cost4, the arrival-counter walk of k_refit); the BVH4 roots are then marked top-down, one pass
// per BVH4 level (k_mark4).  CPU experiment (tools/experiments/bvh_quality.cpp): 4.7 % fewer node visits per bounce ray
// and 21 % fewer BVH4 nodes than "grandchildren at even depth" on the binned-SAH tree of the teapot scene.
__device__ __forceinline__ float cost_of(const float* cost4, int c) { return __ldcg(cost4 + c); }

__global__ void k_cost4(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                        const int* __restrict__ parent, const float4* __restrict__ blo, const float4* __restrict__ bhi,
                        const uint8_t* __restrict__ kind, float* cost4, int4* __restrict__ cut, uint32_t* arrive, float c_tri) {
    const int c0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (c0 >= 2 * n - 1 || kind[c0] != KIND_LEAF) return;
    const int cnt = c0 >= n - 1 ? 1 : range[c0].y - range[c0].x + 1;
    cost4[c0] = c_tri * box_area(blo[c0], bhi[c0]) * (float)cnt;
    __threadfence();
    auto internal = [&](int x) { return kind[x] == KIND_INTERNAL; };
    for (int v = parent[c0]; v >= 0; v = parent[v]) {
        if (atomicAdd(&arrive[v], 1u) == 0u) break;          // the sibling subtree is not finished yet
        __threadfence();
        const int2 ch = children[v];
        const int l = ch.x, r = ch.y;
        float best = cost_of(cost4, l) + cost_of(cost4, r);
        int4 bc = make_int4(l, r, -1, -1);
        auto consider = [&](int a, int b, int c, int d) {
            const float s = cost_of(cost4, a) + cost_of(cost4, b) + cost_of(cost4, c) + (d >= 0 ? cost_of(cost4, d) : 0.f);
            if (s < best) { best = s; bc = make_int4(a, b, c, d); }
        };
        int2 cl = make_int2(-1, -1), cr = make_int2(-1, -1);
        if (internal(l)) { cl = children[l]; consider(cl.x, cl.y, r, -1); }
        if (internal(r)) { cr = children[r]; consider(l, cr.x, cr.y, -1); }
        if (internal(l) && internal(r)) consider(cl.x, cl.y, cr.x, cr.y);
        if (internal(l)) {
            if (internal(cl.x)) { const int2 g = children[cl.x]; consider(g.x, g.y, cl.y, r); }
            if (internal(cl.y)) { const int2 g = children[cl.y]; consider(cl.x, g.x, g.y, r); }
        }
        if (internal(r)) {
            if (internal(cr.x)) { const int2 g = children[cr.x]; consider(l, g.x, g.y, cr.y); }
            if (internal(cr.y)) { const int2 g = children[cr.y]; consider(l, cr.x, g.x, g.y); }
        }
        cut[v] = bc;
        cost4[v] = box_area(blo[v], bhi[v]) + best;
        __threadfence();
    }
}

// lvl4[v] = BVH4 depth of v where v is a BVH4 root, -1 elsewhere.  Pass `level` marks the entries of the roots of that level.
// It also accumulates the exact stack bound: when the traversal stands at a BVH4 node, the stack holds at most the
// other entries of every node on the path to it, so need(v) = need(parent4(v)) + entries(v) - 1 (3 per level would be the
// crude bound; the shared-memory stack of the persistent kernels is sized from this).
__global__ void k_mark4(int n_internal, int level, const uint8_t* __restrict__ kind, const int4* __restrict__ cut,
                        int* lvl4, int* sacc, uint32_t* __restrict__ flags4, BuildScratch* s) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_internal || lvl4[v] != level || kind[v] != KIND_INTERNAL) return;
    const int4 c = cut[v];
    const int e[4] = {c.x, c.y, c.z, c.w};
    int n_ent = 0;
    for (int k = 0; k < 4; ++k) n_ent += e[k] >= 0 ? 1 : 0;
    const int need = sacc[v] + n_ent - 1;
    atomicMax(&s->stack_need, (uint32_t)need);
    bool any = false;
    for (int k = 0; k < 4; ++k)
        if (e[k] >= 0 && e[k] < n_internal && kind[e[k]] == KIND_INTERNAL) {
            lvl4[e[k]] = level + 1; sacc[e[k]] = need; flags4[e[k]] = 1u; any = true;
        }
    if (any) atomicMax(&s->depth4, (uint32_t)level + 1u);
}
__global__ void k_mark4_init(int n_internal, int* lvl4, int* sacc, uint32_t* flags4) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_internal) return;
    lvl4[v] = v == 0 ? 0 : -1;
    sacc[v] = 0;
    flags4[v] = v == 0 ? 1u : 0u;
}

__global__ void k_emit_nodes4(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                              const float4* __restrict__ blo, const float4* __restrict__ bhi,
                              const uint32_t* __restrict__ flags4, const uint32_t* __restrict__ idx4,
                              const uint8_t* __restrict__ kind, float4* __restrict__ nodes4,
                              const BuildScratch* __restrict__ s, const int4* __restrict__ cut) {
    const int n_internal = n - 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_internal > 0 ? n_internal : 1)) return;
    if (n_internal > 0 && !flags4[i]) return;
    // (centre, half extent) boxes: the slab test rounds three times per plane instead of twice, so the pad is doubled
    const float pad = o2f(s->max_abs) * (RTB_NODE_CE ? 1.0f / 65536.0f : 1.0f / 131072.0f);
    auto leaf_final = [&](int c) { return kind[c] == KIND_LEAF; };
    int ent[4];
    int n_ent = 0;
    if (n_internal == 0) ent[n_ent++] = 0;                       // the single Karras leaf
    else if (leaf_final(i)) ent[n_ent++] = i;                    // tiny scene: the root itself is a leaf
    else if (cut) {                                              // entries chosen by k_cost4
        const int4 c = cut[i];
        const int e4[4] = {c.x, c.y, c.z, c.w};
        for (int k = 0; k < 4; ++k) if (e4[k] >= 0) ent[n_ent++] = e4[k];
    } else {
        const int2 ch = children[i];
        const int cs[2] = {ch.x, ch.y};
        for (int k = 0; k < 2; ++k) {
            if (leaf_final(cs[k])) ent[n_ent++] = cs[k];
            else { const int2 g = children[cs[k]]; ent[n_ent++] = g.x; ent[n_ent++] = g.y; }
        }
    }
    float lo[3][4], hi[3][4];
    uint32_t code[4];
    for (int e = 0; e < 4; ++e) {
        if (e < n_ent) {
            const int c = ent[e];
            const float4 l = blo[c], h = bhi[c];
            lo[0][e] = l.x - pad; lo[1][e] = l.y - pad; lo[2][e] = l.z - pad;
            hi[0][e] = h.x + pad; hi[1][e] = h.y + pad; hi[2][e] = h.z + pad;
#if RTB_NODE_CE
            // lo[] holds the centre, hi[] the half extent, rounded up so that [c - e, c + e] contains the padded box
            for (int a = 0; a < 3; ++a) {
                const float c0 = 0.5f * (lo[a][e] + hi[a][e]);
                const float e0 = fmaxf(__fsub_ru(hi[a][e], c0), __fsub_ru(c0, lo[a][e]));
                lo[a][e] = c0; hi[a][e] = e0;
            }
#endif
            if (leaf_final(c)) {
                const uint32_t first = c >= n - 1 ? (uint32_t)(c - (n - 1)) : (uint32_t)range[c].x;
                const uint32_t cnt = c >= n - 1 ? 1u : (uint32_t)(range[c].y - range[c].x + 1);
                code[e] = 0x80000000u | (first << 3) | cnt;
            } else {
                code[e] = idx4[c];
            }
        } else {
            // empty slot: an all-NaN box.  fminf/fmaxf drop NaN operands, so the slab test ends with t_far = NaN and
            // `t_near <= t_far` is false for every ray; an inverted box would NOT do (min/max re-order its planes)
            lo[0][e] = lo[1][e] = lo[2][e] = hi[0][e] = hi[1][e] = hi[2][e] = __int_as_float(0x7fc00000);
            code[e] = 0u;
        }
    }
    float4* out = nodes4 + 8u * (n_internal > 0 ? idx4[i] : 0u);
    for (int a = 0; a < 3; ++a) {
        out[2 * a + 0] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        out[2 * a + 1] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
    out[6] = make_float4(__uint_as_float(code[0]), __uint_as_float(code[1]), __uint_as_float(code[2]), __uint_as_float(code[3]));
    out[7] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ---- 8-wide compressed collapse (A/B accelerator of the wavefront renderer, RTB_BVH8=1 + RTB_FLAG_BVH8) ---------------------
// An 80-byte node with up to eight children, after Ylitie, Karras & Laine, "Efficient Incoherent Ray Traversal on GPUs
// Through Compressed Wide BVHs" (HPG 2017), built from the same binary tree (blo/bhi already refitted):
//   n0  p.x p.y p.z            origin of the node's quantisation frame (= lower corner of the padded node box)
//       ex | ey<<8 | ez<<16 | imask<<24     per-axis scale 2^(e-127) (biased exponent bytes); imask: slots that hold child NODES
//   n1  child_base  tri_base  meta[0..3]  meta[4..7]
//         meta: 0 = empty slot; child node: 0b001_11sss (low 5 bits = 24 + slot); leaf: unary count (1, 3, 7) << 5 | offset
//         of its first reference relative to tri_base (a node's leaf slots hold <= 24 references, contiguous in `tri8`)
//   n2..n4  qlo_x[8] qlo_y[8] qlo_z[8] qhi_x[8] qhi_y[8] qhi_z[8]   child boxes, 8 bits per plane: p + q * 2^(e-127),
//         lower planes rounded down, upper planes rounded up (of the box padded like the BVH2 / BVH4 boxes)
// Child nodes of one node are contiguous (child_base + rank among the node's internal slots), so a traversal stack entry
// is (child_base, hit bits | imask) — 8 bytes for up to eight children — and slots are ASSIGNED by octant (the child in
// the (-,-,-) corner of the node goes to slot 0, ...) so that `slot ^ ray octant` orders the hits front to back with no sort.
// Which descendants become the eight children: start from the two BVH2 children, keep opening the entry with the largest
// surface area among those covering more than 3 references; then, while slots are free, open the largest 2- or 3-reference
// entries too — an empty slot costs the same eight box tests, a single-reference slot is that reference's own (quantised)
// bounding box, so most exact triangle tests are preceded by a box test of the triangle itself.
struct C8Item { int node; uint32_t out; };
struct C8State { uint32_t n_next, n_nodes, n_tris, depth; };

__device__ __forceinline__ int c8_size(int c, int n, const int2* __restrict__ range) {
    return c >= n - 1 ? 1 : range[c].y - range[c].x + 1;
}

__global__ void k_c8_level(const C8Item* __restrict__ items, uint32_t n_items, C8Item* __restrict__ next, C8State* st,
                           int n, const int2* __restrict__ children, const int2* __restrict__ range,
                           const float4* __restrict__ blo, const float4* __restrict__ bhi, const BuildScratch* __restrict__ s,
                           uint4* __restrict__ nodes8, uint32_t* __restrict__ perm8, uint32_t level) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    const C8Item it = items[i];
    const float pad = o2f(s->max_abs) * (1.0f / 131072.0f);   // 2^-17 of the largest |coordinate|, as for the other node arrays
    int ent[8];
    int cnt = 0;
    if (it.node < 0) {                 // tiny scene without an internal BVH2 node: the single Karras leaf
        ent[cnt++] = n - 1;
    } else {
        const int2 ch = children[it.node];
        ent[cnt++] = ch.x; ent[cnt++] = ch.y;
    }
    // open entries: first the big ones (they cannot be leaf slots), then 2-3 reference groups while slots are free
    for (int pass = 0; pass < 2; ++pass) {
        while (cnt < 8) {
            int best = -1; float best_a = -1.f;
            for (int k = 0; k < cnt; ++k) {
                const int sz = c8_size(ent[k], n, range);
                if (pass == 0 ? sz <= 3 : sz < 2) continue;
                const float a = box_area(blo[ent[k]], bhi[ent[k]]);
                if (a > best_a) { best_a = a; best = k; }
            }
            if (best < 0) break;
            const int2 ch = children[ent[best]];
            ent[best] = ch.x; ent[cnt++] = ch.y;
        }
    }
    // node frame
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float4 el[8], eh[8];
    for (int k = 0; k < cnt; ++k) {
        el[k] = blo[ent[k]]; eh[k] = bhi[ent[k]];
        el[k].x -= pad; el[k].y -= pad; el[k].z -= pad; eh[k].x += pad; eh[k].y += pad; eh[k].z += pad;
        lo[0] = fminf(lo[0], el[k].x); lo[1] = fminf(lo[1], el[k].y); lo[2] = fminf(lo[2], el[k].z);
        hi[0] = fmaxf(hi[0], eh[k].x); hi[1] = fmaxf(hi[1], eh[k].y); hi[2] = fmaxf(hi[2], eh[k].z);
    }
    uint32_t eb[3];
    float inv_scale[3];
    for (int a = 0; a < 3; ++a) {
        // smallest power of two with 254 * scale >= extent (254, not 255: headroom for the rounded-up subtraction below)
        const float t = __fdiv_ru(__fsub_ru(hi[a], lo[a]), 254.0f);
        uint32_t bits = __float_as_uint(t);
        uint32_t e = bits >> 23;                                   // t >= 0
        if (bits & 0x007fffffu) ++e;
        e = min(max(e, 1u), 254u);
        eb[a] = e;
        inv_scale[a] = __uint_as_float((254u - e) << 23);         // 2^(127-e), exact
    }
    // octant slot assignment: greedily give (child, slot) pairs with the best alignment of the child's offset from the
    // node centre with the slot's corner direction
    const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
    int slot_of[8], child_in[8];
    for (int k = 0; k < 8; ++k) { slot_of[k] = -1; child_in[k] = -1; }
    for (int round = 0; round < cnt; ++round) {
        float bestv = -FLT_MAX; int bk = -1, bs = -1;
        for (int k = 0; k < cnt; ++k) {
            if (slot_of[k] >= 0) continue;
            const float dx = 0.5f * (el[k].x + eh[k].x) - cx, dy = 0.5f * (el[k].y + eh[k].y) - cy, dz = 0.5f * (el[k].z + eh[k].z) - cz;
            for (int sl = 0; sl < 8; ++sl) {
                if (child_in[sl] >= 0) continue;
                const float v = ((sl & 1) ? dx : -dx) + ((sl & 2) ? dy : -dy) + ((sl & 4) ? dz : -dz);
                if (v > bestv) { bestv = v; bk = k; bs = sl; }
            }
        }
        slot_of[bk] = bs; child_in[bs] = bk;
    }
    // children: nodes (more than 3 references) get consecutive indices in slot order, leaf slots consecutive references
    uint32_t n_int = 0, n_tri = 0;
    for (int sl = 0; sl < 8; ++sl) {
        if (child_in[sl] < 0) continue;
        const int sz = c8_size(ent[child_in[sl]], n, range);
        if (sz > 3) ++n_int; else n_tri += (uint32_t)sz;
    }
    const uint32_t child_base = n_int ? atomicAdd(&st->n_nodes, n_int) : 0u;
    const uint32_t item_base = n_int ? atomicAdd(&st->n_next, n_int) : 0u;
    const uint32_t tri_base = n_tri ? atomicAdd(&st->n_tris, n_tri) : 0u;
    uint32_t meta[8], qlo[3][8], qhi[3][8];
    uint32_t imask = 0, ri = 0, rt = 0;
    for (int sl = 0; sl < 8; ++sl) {
        meta[sl] = 0u;
        for (int a = 0; a < 3; ++a) { qlo[a][sl] = 0u; qhi[a][sl] = 0u; }
        const int k = child_in[sl];
        if (k < 0) continue;
        const int c = ent[k];
        const int sz = c8_size(c, n, range);
        if (sz > 3) {
            imask |= 1u << sl;
            meta[sl] = (1u << 5) | (24u + (uint32_t)sl);
            next[item_base + ri].node = c; next[item_base + ri].out = child_base + ri;
            ++ri;
        } else {
            meta[sl] = (((1u << sz) - 1u) << 5) | rt;
            const int first = c >= n - 1 ? c - (n - 1) : range[c].x;
            for (int j = 0; j < sz; ++j) perm8[tri_base + rt + j] = (uint32_t)(first + j);
            rt += (uint32_t)sz;
        }
        const float l3[3] = {el[k].x, el[k].y, el[k].z}, h3[3] = {eh[k].x, eh[k].y, eh[k].z};
        for (int a = 0; a < 3; ++a) {
            // p + q * scale <= lower plane, p + q * scale >= upper plane (the division by a power of two is exact)
            const float ql = floorf(__fsub_rd(l3[a], lo[a]) * inv_scale[a]);
            const float qh = ceilf(__fsub_ru(h3[a], lo[a]) * inv_scale[a]);
            qlo[a][sl] = (uint32_t)fminf(fmaxf(ql, 0.0f), 255.0f);
            qhi[a][sl] = (uint32_t)fminf(fmaxf(qh, 0.0f), 255.0f);
        }
    }
    auto pack4 = [](const uint32_t* b) { return b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24); };
    uint4* out = nodes8 + 5u * (size_t)it.out;
    out[0] = make_uint4(__float_as_uint(lo[0]), __float_as_uint(lo[1]), __float_as_uint(lo[2]),
                        eb[0] | (eb[1] << 8) | (eb[2] << 16) | (imask << 24));
    out[1] = make_uint4(child_base, tri_base, pack4(meta), pack4(meta + 4));
    out[2] = make_uint4(pack4(qlo[0]), pack4(qlo[0] + 4), pack4(qlo[1]), pack4(qlo[1] + 4));
    out[3] = make_uint4(pack4(qlo[2]), pack4(qlo[2] + 4), pack4(qhi[0]), pack4(qhi[0] + 4));
    out[4] = make_uint4(pack4(qhi[1]), pack4(qhi[1] + 4), pack4(qhi[2]), pack4(qhi[2] + 4));
    if (i == 0) atomicMax(&st->depth, level);
}

// references gathered into BVH8 order: position i holds leaf-order reference perm8[i]
__global__ void k_emit_tris8(const float4* __restrict__ tri, const float4* __restrict__ shade,
                             const uint32_t* __restrict__ perm8, uint32_t n, float4* __restrict__ tri8,
                             float4* __restrict__ shade8) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t k = perm8[i];
    for (int j = 0; j < RTB_TRI_F4; ++j) tri8[(size_t)RTB_TRI_F4 * i + j] = tri[(size_t)RTB_TRI_F4 * k + j];
    for (int j = 0; j < RTB_SHADE_F4; ++j) shade8[(size_t)RTB_SHADE_F4 * i + j] = shade[(size_t)RTB_SHADE_F4 * k + j];
}

__global__ void k_emit_tris(const RtbTriangle* __restrict__ tris, const uint32_t* __restrict__ keep,
                            const uint32_t* __restrict__ sorted_vals, uint32_t n, float4* __restrict__ tri,
                            float4* __restrict__ shade, uint32_t* __restrict__ prim_order) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t orig = keep[sorted_vals[k]];
    const RtbTriangle& t = tris[orig];
    float4* q = tri + (size_t)RTB_TRI_F4 * k;
    q[0] = make_float4(t.norm[0], t.norm[1], t.norm[2], t.bounding_r2);
    q[1] = make_float4(t.incenter[0], t.incenter[1], t.incenter[2], __uint_as_float(orig));
    q[2] = make_float4(t.sides[0], t.sides[1], t.sides[2], t.side_lens[0]);
    q[3] = make_float4(t.sides[3], t.sides[4], t.sides[5], t.side_lens[1]);
    q[4] = make_float4(t.sides[6], t.sides[7], t.sides[8], t.side_lens[2]);
    float4* sh = shade + (size_t)RTB_SHADE_F4 * k;
    sh[0] = make_float4(t.color[0], t.color[1], t.color[2], t.alpha);
    sh[1] = make_float4(__uint_as_float(t.kind), t.scattering, t.edge_thickness, 0.f);
    prim_order[k] = orig;
}

// Build scratch comes from the device's stream-ordered pool (release threshold raised in rtb_init): with plain
// cudaMalloc / cudaFree a 1 M-triangle build took 30 ms most of the time and 285 ms whenever the driver decided to
// give memory back to the system and fetch it again.
thread_local cudaStream_t t_alloc_stream = nullptr;
template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFreeAsync(p, t_alloc_stream); }
    cudaError_t alloc(size_t n) { return cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), t_alloc_stream); }
};

inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

static void free_result(BuildResult* r) {
    cudaFree(r->d_nodes); cudaFree(r->d_nodes4); cudaFree(r->d_tri); cudaFree(r->d_shade); cudaFree(r->d_prim_order);
    cudaFree(r->d_nodes8); cudaFree(r->d_tri8); cudaFree(r->d_shade8);
    *r = BuildResult();
}

static int build_impl(const RtbTriangle* d_tris, const uint32_t* d_keep, uint32_t n, cudaStream_t stream,
                      bool force_karras, BuildResult* out);

int rtb_build_lbvh(const RtbTriangle* d_tris, const uint32_t* d_keep, uint32_t n, cudaStream_t stream,
                   BuildResult* out) {
    int rc = build_impl(d_tris, d_keep, n, stream, false, out);
    if (rc == RTB_ERR_INVALID && out->tree_height + 2 > RTB_STACK) {
        // a PLOC tree can, on adversarial input, come out deeper than the traversal stack allows: the radix tree
        // over 63-bit codes cannot (height <= 63 + log2 of the largest run of equal codes)
        free_result(out);
        rc = build_impl(d_tris, d_keep, n, stream, true, out);
    }
    if (rc != RTB_OK) free_result(out);
    return rc;
}

static int build_impl(const RtbTriangle* d_tris, const uint32_t* d_keep, uint32_t n, cudaStream_t stream,
                      bool force_karras, BuildResult* out) {
    *out = BuildResult();
    t_alloc_stream = stream;
    // device time of the build = the bounds/split phase (s0..s1) + everything after the allocations (e0..e1)
    cudaEvent_t e0, e1, s0, s1, s2, s3;
    RTB_CUDA(cudaEventCreate(&e0));
    RTB_CUDA(cudaEventCreate(&e1));
    RTB_CUDA(cudaEventCreate(&s0));
    RTB_CUDA(cudaEventCreate(&s1));
    RTB_CUDA(cudaEventCreate(&s2));
    RTB_CUDA(cudaEventCreate(&s3));
    bool split_emitted = false;
    const uint32_t B = 256;
    uint32_t launches = 0;

    // experiment knobs, read once (thread-safe function-local static)
    struct BuildEnv { int builder, ploc_r, sah_leaves, split_div, collapse_dp, collapse_ct; };
    static const BuildEnv benv = [] {
        BuildEnv e;
        const char* b = getenv("RTB_BUILDER");        // "karras" = plain radix tree, "ploc" = PLOC, default binned SAH top-down
        e.builder = (b && b[0] == 'k') ? 0 : ((b && b[0] == 'p') ? 1 : 2);
        const char* r = getenv("RTB_PLOC_R");
        e.ploc_r = r ? std::min(PLOC_R_MAX, std::max(1, atoi(r))) : 16;
        const char* l = getenv("RTB_SAH_LEAVES");
        e.sah_leaves = l ? atoi(l) : 1;
        const char* c4 = getenv("RTB_COLLAPSE");       // "even" = grandchildren at even depth, default: dynamic programming
        e.collapse_dp = (c4 && c4[0] == 'e') ? 0 : 1;
        const char* ct = getenv("RTB_COLLAPSE_CT");    // cost of an exact triangle test in percent of a node visit
        e.collapse_ct = ct ? std::max(1, atoi(ct)) : 120;
        const char* d = getenv("RTB_SPLIT_DIV");      // split references longer than scene extent / this; 0 = off
        e.split_div = d ? std::max(0, atoi(d)) : 16;
        return e;
    }();

    // primitive boxes and scene bounds; long primitives are entered as several references (k_split_refs)
    DevBuf<BuildScratch> scratch;
    DevBuf<float4> plo, phi;
    DevBuf<uint32_t> ref_keep;
    RTB_CUDA(scratch.alloc(1));
    RTB_CUDA(plo.alloc(n)); RTB_CUDA(phi.alloc(n));
    DevBuf<uint32_t> cnt, off;
    DevBuf<uint8_t> scan_tmp;
    if (n > 0 && benv.split_div > 0) {
        RTB_CUDA(cnt.alloc(n + 1)); RTB_CUDA(off.alloc(n + 1));
        RTB_CUDA(scan_tmp.alloc(rtbsort::scan_tmp_bytes<uint32_t>(n + 1)));
    }
    RTB_CUDA(cudaEventRecord(s0, stream));
    k_init_scratch<<<1, 32, 0, stream>>>(scratch.p); ++launches;
    out->n_refs = n;
    if (n > 0) {
        k_prim_bounds<<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, n, plo.p, phi.p, scratch.p); ++launches;
        if (benv.split_div > 0) {
            const float inv_div = 1.0f / (float)benv.split_div;
            k_split_refs<false><<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, n, plo.p, phi.p, scratch.p, inv_div, cnt.p,
                                                              nullptr, nullptr, nullptr, nullptr); ++launches;
            RTB_CUDA(cudaMemsetAsync(cnt.p + n, 0, sizeof(uint32_t), stream));
            launches += (uint32_t)rtbsort::exclusive_sum<uint32_t>(cnt.p, off.p, n + 1, scan_tmp.p, stream);
            uint32_t n_refs = n;
            RTB_CUDA(cudaEventRecord(s1, stream));
            RTB_CUDA(cudaMemcpyAsync(&n_refs, off.p + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            RTB_CUDA(cudaStreamSynchronize(stream));
            if (n_refs > n) {
                DevBuf<float4> rlo, rhi;
                RTB_CUDA(rlo.alloc(n_refs)); RTB_CUDA(rhi.alloc(n_refs)); RTB_CUDA(ref_keep.alloc(n_refs));
                RTB_CUDA(cudaEventRecord(s2, stream));
                split_emitted = true;
                k_split_refs<true><<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, n, plo.p, phi.p, scratch.p, inv_div, nullptr,
                                                                 off.p, rlo.p, rhi.p, ref_keep.p); ++launches;
                RTB_CUDA(cudaEventRecord(s3, stream));
                RTB_CUDA(cudaStreamSynchronize(stream));      // the old boxes are freed at the end of this scope
                std::swap(plo.p, rlo.p); std::swap(phi.p, rhi.p);
                d_keep = ref_keep.p;
                n = n_refs;
                out->n_refs = n_refs;
            }
        }
    }

    if (!(n > 0 && benv.split_div > 0)) RTB_CUDA(cudaEventRecord(s1, stream));

    // final arrays (owned by the scene afterwards)
    const uint32_t n_alloc = n ? n : 1;
    RTB_CUDA(cudaMalloc(&out->d_tri, sizeof(float4) * RTB_TRI_F4 * n_alloc));
    RTB_CUDA(cudaMalloc(&out->d_shade, sizeof(float4) * RTB_SHADE_F4 * n_alloc));
    RTB_CUDA(cudaMalloc(&out->d_prim_order, sizeof(uint32_t) * n_alloc));

    if (n == 0) {  // empty scene: a single leaf with no primitives
        RTB_CUDA(cudaMalloc(&out->d_nodes, sizeof(float4) * 4));
        float4 h[4] = {make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f),
                       make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f), make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f)};
        RTB_CUDA(cudaMemcpyAsync(out->d_nodes, h, sizeof h, cudaMemcpyHostToDevice, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
        out->n_nodes = 2;
        RTB_CUDA(cudaMalloc(&out->d_nodes4, sizeof(float4) * 8));
        // four empty slots: all-NaN boxes (the traversal recognises an empty slot by its box alone), codes 0
        float4 h4[8];
        const float qnan = std::numeric_limits<float>::quiet_NaN();
        for (int a = 0; a < 6; ++a) h4[a] = make_float4(qnan, qnan, qnan, qnan);
        h4[6] = h4[7] = make_float4(0.f, 0.f, 0.f, 0.f);
        RTB_CUDA(cudaMemcpyAsync(out->d_nodes4, h4, sizeof h4, cudaMemcpyHostToDevice, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
        out->n_nodes4 = 1;
        RTB_CUDA(cudaMalloc(&out->d_nodes8, sizeof(uint4) * 5));
        RTB_CUDA(cudaMemsetAsync(out->d_nodes8, 0, sizeof(uint4) * 5, stream));       // every meta byte 0: eight empty slots
        RTB_CUDA(cudaMalloc(&out->d_tri8, sizeof(float4) * RTB_TRI_F4));
        RTB_CUDA(cudaMalloc(&out->d_shade8, sizeof(float4) * RTB_SHADE_F4));
        RTB_CUDA(cudaStreamSynchronize(stream));
        out->n_nodes8 = 1; out->depth8 = 0;
        cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(s0); cudaEventDestroy(s1); cudaEventDestroy(s2); cudaEventDestroy(s3);
        return RTB_OK;
    }

    DevBuf<float4> blo, bhi;
    DevBuf<uint64_t> keys, keys_sorted;
    DevBuf<uint32_t> vals, vals_sorted, arrive, flags, slot, flags4, idx4;
    DevBuf<int2> children, range;
    DevBuf<int> parent;
    DevBuf<uint8_t> sort_tmp, kind;
    // PLOC scratch
    DevBuf<PlocState> pstate;
    DevBuf<float4> qlo, qhi;
    DevBuf<uint32_t> cl_a, cl_b, nn, psize, vals_dfs, vals_dfs_sah;
    DevBuf<unsigned long long> pflags, ppos;
    DevBuf<int2> pchildren;
    DevBuf<int> pparent;
    const uint32_t n_int = n - 1, n_all = 2 * n - 1;
    RTB_CUDA(blo.alloc(n_all)); RTB_CUDA(bhi.alloc(n_all));
    RTB_CUDA(keys.alloc(n)); RTB_CUDA(keys_sorted.alloc(n));
    RTB_CUDA(vals.alloc(n)); RTB_CUDA(vals_sorted.alloc(n));
    RTB_CUDA(arrive.alloc(n_int)); RTB_CUDA(flags.alloc(n_int + 1)); RTB_CUDA(slot.alloc(n_int + 1));
    RTB_CUDA(children.alloc(n_int)); RTB_CUDA(range.alloc(n_int)); RTB_CUDA(parent.alloc(n_all));
    RTB_CUDA(flags4.alloc(n_int + 1)); RTB_CUDA(idx4.alloc(n_int + 1));
    RTB_CUDA(kind.alloc(n_all));
    const int builder = benv.builder, ploc_r = benv.ploc_r, sah_leaves = benv.sah_leaves;
    const bool use_ploc = builder == 1 && n_int > 0 && !force_karras;
    const bool use_sah = builder == 2 && n_int > 0 && !force_karras;
    const bool collapse_dp = benv.collapse_dp != 0;
    DevBuf<float> cost4;
    DevBuf<int4> cut4;
    DevBuf<int> lvl4, sacc4;
    if (collapse_dp) { RTB_CUDA(cost4.alloc(n_all)); RTB_CUDA(cut4.alloc(n_int)); RTB_CUDA(lvl4.alloc(n_int)); RTB_CUDA(sacc4.alloc(n_int)); }
    // binned-SAH scratch: per-level work lists (small / big nodes), the second index buffer, the leaf order
    DevBuf<SahItem> ws[2], wbig[2];
    DevBuf<SahState> sst;
    DevBuf<uint32_t> ib;
    if (use_sah) {
        for (int k = 0; k < 2; ++k) { RTB_CUDA(ws[k].alloc(n)); RTB_CUDA(wbig[k].alloc(n / SAH_BIG + 2)); }
        RTB_CUDA(sst.alloc(1)); RTB_CUDA(ib.alloc(n)); RTB_CUDA(vals_dfs_sah.alloc(n));
    }
    if (use_ploc) {
        RTB_CUDA(pstate.alloc(2)); RTB_CUDA(qlo.alloc(n_all)); RTB_CUDA(qhi.alloc(n_all));
        RTB_CUDA(cl_a.alloc(n)); RTB_CUDA(cl_b.alloc(n)); RTB_CUDA(nn.alloc(n)); RTB_CUDA(psize.alloc(n_all));
        RTB_CUDA(vals_dfs.alloc(n)); RTB_CUDA(pflags.alloc(n)); RTB_CUDA(ppos.alloc(n));
        RTB_CUDA(pchildren.alloc(n_int)); RTB_CUDA(pparent.alloc(n_all));
    }

    const size_t tmp_bytes = std::max(rtbsort::sort_tmp_bytes(n), std::max(rtbsort::scan_tmp_bytes<uint32_t>(n_int + 1),
                                                                          rtbsort::scan_tmp_bytes<unsigned long long>(n)));
    RTB_CUDA(sort_tmp.alloc(tmp_bytes));

    RTB_CUDA(cudaEventRecord(e0, stream));
    k_morton<<<cdiv(n, B), B, 0, stream>>>(plo.p, phi.p, n, scratch.p, keys.p, vals.p); ++launches;
    {
        bool in_b = false;
        launches += (uint32_t)rtbsort::radix_sort_pairs((unsigned long long*)keys.p, (unsigned long long*)keys_sorted.p, vals.p,
                                                        vals_sorted.p, n, 63, sort_tmp.p, stream, &in_b);
        if (!in_b) { std::swap(keys.p, keys_sorted.p); std::swap(vals.p, vals_sorted.p); }   // result -> *_sorted
    }
    uint32_t total_split = 0, tree_height_host = 0;
    const uint32_t* leaf_vals = vals_sorted.p;     // primitive of every leaf, in leaf order
    if (use_ploc) {
        RTB_CUDA(cudaMemsetAsync(arrive.p, 0, sizeof(uint32_t) * n_int, stream));
        k_ploc_init<<<cdiv(n, B), B, 0, stream>>>(n, vals_sorted.p, plo.p, phi.p, qlo.p, qhi.p, cl_a.p, psize.p, pparent.p,
                                                  pstate.p); ++launches;
        // Rounds run in groups of PLOC_GROUP between host reads of the cluster count: the grids of a group are sized for
        // the count at its start (a stale upper bound, the kernels read the true count on the device), the two
        // PlocState slots alternate as input and output.
        constexpr int PLOC_GROUP = 3;
        uint32_t m_upper = n;
        uint32_t* cin = cl_a.p; uint32_t* cout = cl_b.p;
        int it = 0;
        while (m_upper > (uint32_t)PLOC_TAIL) {
            for (int k = 0; k < PLOC_GROUP; ++k, ++it) {
                const PlocState* si = pstate.p + (it & 1);
                PlocState* so = pstate.p + ((it + 1) & 1);
                k_ploc_nn<<<cdiv(m_upper, PLOC_BLOCK), PLOC_BLOCK, 0, stream>>>(cin, si, qlo.p, qhi.p, ploc_r, nn.p);
                k_ploc_flags<<<cdiv(m_upper, B), B, 0, stream>>>(nn.p, si, pflags.p, m_upper);
                rtbsort::exclusive_sum<unsigned long long>(pflags.p, ppos.p, m_upper, sort_tmp.p, stream);
                k_ploc_apply<<<cdiv(m_upper, B), B, 0, stream>>>(n, cin, nn.p, ppos.p, si, so, cout, pchildren.p, pparent.p,
                                                             qlo.p, qhi.p, psize.p);
                launches += 6;
                std::swap(cin, cout);
            }
            PlocState hs;
            RTB_CUDA(cudaMemcpyAsync(&hs, pstate.p + (it & 1), sizeof hs, cudaMemcpyDeviceToHost, stream));
            RTB_CUDA(cudaStreamSynchronize(stream));
            if (hs.m >= m_upper) { rtb_set_error("PLOC made no progress"); return RTB_ERR_CUDA; }
            m_upper = hs.m;
        }
        if (m_upper > 1) {
            k_ploc_tail<<<1, PLOC_TAIL, 0, stream>>>(n, cin, pstate.p + (it & 1), pstate.p + ((it + 1) & 1), ploc_r, pchildren.p,
                                                     pparent.p, qlo.p, qhi.p, psize.p); ++launches;
        }
        k_ploc_finish<<<cdiv(n_int, B), B, 0, stream>>>(n, pchildren.p, pparent.p, psize.p, vals_sorted.p, children.p,
                                                        range.p, parent.p, vals_dfs.p); ++launches;
        leaf_vals = vals_dfs.p;
    } else if (use_sah) {
        RTB_CUDA(cudaMemsetAsync(arrive.p, 0, sizeof(uint32_t) * n_int, stream));
        // level L reads the index buffer idx[L & 1] and writes idx[(L + 1) & 1]; the Morton order is the starting order
        uint32_t* idx[2] = {vals_sorted.p, ib.p};
        SahItem root; root.node = 0; root.lo = 0; root.hi = n;
        SahState hs; hs.n_next_small = hs.n_next_big = 0; hs.next_internal = 1; hs.pad = 0;
        uint32_t n_small = n > SAH_BIG ? 0u : 1u, n_big = n > SAH_BIG ? 1u : 0u;
        RTB_CUDA(cudaMemcpyAsync(n_big ? wbig[0].p : ws[0].p, &root, sizeof root, cudaMemcpyHostToDevice, stream));
        RTB_CUDA(cudaMemcpyAsync(sst.p, &hs, sizeof hs, cudaMemcpyHostToDevice, stream));
        for (int level = 0; n_small + n_big > 0; ++level) {
            if (level + 2 > RTB_STACK) {      // a chain deeper than the traversal stack: let the caller fall back to the radix tree
                RTB_CUDA(cudaStreamSynchronize(stream));
                out->tree_height = RTB_STACK;
                rtb_set_error("SAH tree deeper than the traversal stack");
                return RTB_ERR_INVALID;
            }
            const int c = level & 1, x = c ^ 1;
            if (n_big) {
                k_sah_level_big<<<n_big, SAH_BIG_TEAM, 0, stream>>>(wbig[c].p, n_big, idx[c], idx[x], plo.p, phi.p, n, ws[x].p, wbig[x].p,
                                                                    sst.p, children.p, range.p, parent.p, vals_dfs_sah.p); ++launches;
            }
            if (n_small) {
                k_sah_level_small<<<cdiv(n_small, SAH_WARPS), 32 * SAH_WARPS, 0, stream>>>(ws[c].p, n_small, idx[c], idx[x], plo.p, phi.p, n,
                                                                                          ws[x].p, wbig[x].p, sst.p, children.p, range.p,
                                                                                          parent.p, vals_dfs_sah.p); ++launches;
            }
            RTB_CUDA(cudaMemcpyAsync(&hs, sst.p, sizeof hs, cudaMemcpyDeviceToHost, stream));
            RTB_CUDA(cudaStreamSynchronize(stream));
            n_small = hs.n_next_small; n_big = hs.n_next_big;
            RTB_CUDA(cudaMemsetAsync(sst.p, 0, 2 * sizeof(uint32_t), stream));      // the two list counters, not next_internal
        }
        leaf_vals = vals_dfs_sah.p;
    } else if (n_int > 0) {
        RTB_CUDA(cudaMemsetAsync(arrive.p, 0, sizeof(uint32_t) * n_int, stream));
        k_hierarchy<<<cdiv(n_int, B), B, 0, stream>>>(keys_sorted.p, (int)n, children.p, range.p, parent.p); ++launches;
    } else {
        int minus1 = -1;
        RTB_CUDA(cudaMemcpyAsync(parent.p, &minus1, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    k_refit<<<cdiv(n, B), B, 0, stream>>>(plo.p, phi.p, leaf_vals, (int)n, children.p, parent.p, blo.p, bhi.p,
                                          arrive.p, scratch.p); ++launches;
    k_leaf_depth<<<cdiv(n, B), B, 0, stream>>>((int)n, parent.p, scratch.p); ++launches;
    k_node_kind<<<cdiv(n_all, B), B, 0, stream>>>((int)n, children.p, range.p, parent.p, blo.p, bhi.p, sah_leaves, kind.p);
    ++launches;
    if (n_int > 0) {
        k_split_flags<<<cdiv(n_int, B), B, 0, stream>>>(kind.p, (int)n_int, flags.p); ++launches;
        RTB_CUDA(cudaMemsetAsync(flags.p + n_int, 0, sizeof(uint32_t), stream));
        launches += (uint32_t)rtbsort::exclusive_sum<uint32_t>(flags.p, slot.p, n_int + 1, sort_tmp.p, stream);
        RTB_CUDA(cudaMemcpyAsync(&total_split, slot.p + n_int, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        RTB_CUDA(cudaMemcpyAsync(&tree_height_host, &scratch.p->height, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
    }
    out->n_nodes = 2 + 2 * total_split;
    RTB_CUDA(cudaMalloc(&out->d_nodes, sizeof(float4) * 2 * out->n_nodes));
    k_emit_nodes<<<cdiv(n_all, B), B, 0, stream>>>((int)n, children.p, range.p, parent.p, blo.p, bhi.p, slot.p, kind.p,
                                                   out->d_nodes, scratch.p); ++launches;
    k_emit_tris<<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, leaf_vals, n, out->d_tri, out->d_shade,
                                              out->d_prim_order); ++launches;
    // 4-wide collapse of the same tree (used by the wavefront renderer)
    uint32_t total4 = 1;
    if (n_int > 0) {
        if (collapse_dp) {
            RTB_CUDA(cudaMemsetAsync(arrive.p, 0, sizeof(uint32_t) * n_int, stream));
            k_cost4<<<cdiv(n_all, B), B, 0, stream>>>((int)n, children.p, range.p, parent.p, blo.p, bhi.p, kind.p, cost4.p, cut4.p,
                                                      arrive.p, 0.01f * (float)benv.collapse_ct); ++launches;
            k_mark4_init<<<cdiv(n_int, B), B, 0, stream>>>((int)n_int, lvl4.p, sacc4.p, flags4.p); ++launches;
            // one marking pass per possible BVH4 level: a BVH4 level spans at least one BVH2 level (height is on the host
            // from the read-back after k_node_kind)
            for (uint32_t level = 0; level < tree_height_host; ++level) {
                k_mark4<<<cdiv(n_int, B), B, 0, stream>>>((int)n_int, (int)level, kind.p, cut4.p, lvl4.p, sacc4.p, flags4.p, scratch.p); ++launches;
            }
        } else {
            k_flag4<<<cdiv(n_int, B), B, 0, stream>>>((int)n_int, kind.p, parent.p, flags4.p, scratch.p); ++launches;
        }
        RTB_CUDA(cudaMemsetAsync(flags4.p + n_int, 0, sizeof(uint32_t), stream));
        launches += (uint32_t)rtbsort::exclusive_sum<uint32_t>(flags4.p, idx4.p, n_int + 1, sort_tmp.p, stream);
        RTB_CUDA(cudaMemcpyAsync(&total4, idx4.p + n_int, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
    }
    out->n_nodes4 = total4;
    RTB_CUDA(cudaMalloc(&out->d_nodes4, sizeof(float4) * 8 * (size_t)total4));
    k_emit_nodes4<<<cdiv(n_int > 0 ? n_int : 1, B), B, 0, stream>>>((int)n, children.p, range.p, blo.p, bhi.p, flags4.p,
                                                                  idx4.p, kind.p, out->d_nodes4, scratch.p, collapse_dp ? cut4.p : nullptr); ++launches;
    // 8-wide compressed collapse of the same tree, one launch per BVH8 level.  Measured slower than the BVH4 on this path
    // (cache-resident scenes: the decode costs more issue slots than the compression saves), so it is
    // built only on request: RTB_BVH8=1 in the environment when the scene is created (read per call), used with RTB_FLAG_BVH8.
    const char* want8 = getenv("RTB_BVH8");
    if (want8 && atoi(want8) != 0) {
        DevBuf<C8Item> it_a, it_b;
        DevBuf<C8State> c8st;
        DevBuf<uint4> nodes8_tmp;
        DevBuf<uint32_t> perm8;
        const uint32_t max_nodes8 = n_int + 1, max_items = n / 4 + 8;      // a child node covers > 3 references of its own
        RTB_CUDA(it_a.alloc(max_items)); RTB_CUDA(it_b.alloc(max_items)); RTB_CUDA(c8st.alloc(1));
        RTB_CUDA(nodes8_tmp.alloc(5 * (size_t)max_nodes8)); RTB_CUDA(perm8.alloc(n));
        C8Item root; root.node = n_int > 0 ? 0 : -1; root.out = 0;
        C8State hs8; hs8.n_next = 0; hs8.n_nodes = 1; hs8.n_tris = 0; hs8.depth = 0;
        RTB_CUDA(cudaMemcpyAsync(it_a.p, &root, sizeof root, cudaMemcpyHostToDevice, stream));
        RTB_CUDA(cudaMemcpyAsync(c8st.p, &hs8, sizeof hs8, cudaMemcpyHostToDevice, stream));
        C8Item* cur = it_a.p; C8Item* nxt = it_b.p;
        uint32_t n_items = 1, level = 0;
        while (n_items > 0) {
            k_c8_level<<<cdiv(n_items, 64), 64, 0, stream>>>(cur, n_items, nxt, c8st.p, (int)n, children.p, range.p, blo.p, bhi.p,
                                                           scratch.p, nodes8_tmp.p, perm8.p, level); ++launches;
            RTB_CUDA(cudaMemcpyAsync(&hs8, c8st.p, sizeof hs8, cudaMemcpyDeviceToHost, stream));
            RTB_CUDA(cudaStreamSynchronize(stream));
            n_items = hs8.n_next;
            if (n_items > max_items) { rtb_set_error("BVH8 collapse: work list overflow"); return RTB_ERR_CUDA; }
            RTB_CUDA(cudaMemsetAsync(&c8st.p->n_next, 0, sizeof(uint32_t), stream));
            std::swap(cur, nxt);
            ++level;
        }
        if (hs8.n_tris != n || hs8.n_nodes > max_nodes8) { rtb_set_error("BVH8 collapse: inconsistent counts"); return RTB_ERR_CUDA; }
        out->n_nodes8 = hs8.n_nodes; out->depth8 = level;
        RTB_CUDA(cudaMalloc(&out->d_nodes8, sizeof(uint4) * 5 * (size_t)hs8.n_nodes));
        RTB_CUDA(cudaMemcpyAsync(out->d_nodes8, nodes8_tmp.p, sizeof(uint4) * 5 * (size_t)hs8.n_nodes, cudaMemcpyDeviceToDevice, stream));
        RTB_CUDA(cudaMalloc(&out->d_tri8, sizeof(float4) * RTB_TRI_F4 * (size_t)n));
        RTB_CUDA(cudaMalloc(&out->d_shade8, sizeof(float4) * RTB_SHADE_F4 * (size_t)n));
        k_emit_tris8<<<cdiv(n, B), B, 0, stream>>>(out->d_tri, out->d_shade, perm8.p, n, out->d_tri8, out->d_shade8); ++launches;
        RTB_CUDA(cudaStreamSynchronize(stream));      // the scratch buffers of this scope are freed stream-ordered
    }
    RTB_CUDA(cudaEventRecord(e1, stream));
    BuildScratch h;
    RTB_CUDA(cudaMemcpyAsync(&h, scratch.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    RTB_CUDA(cudaStreamSynchronize(stream));
    RTB_CUDA(cudaGetLastError());
    float ms_split = 0.f, ms_emit = 0.f;
    RTB_CUDA(cudaEventElapsedTime(&out->ms_build, e0, e1));
    RTB_CUDA(cudaEventElapsedTime(&ms_split, s0, s1));
    if (split_emitted) RTB_CUDA(cudaEventElapsedTime(&ms_emit, s2, s3));
    out->ms_build += ms_split + ms_emit;
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(s0); cudaEventDestroy(s1); cudaEventDestroy(s2); cudaEventDestroy(s3);

    auto dec = [](uint32_t o) { uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; float f; memcpy(&f, &u, 4); return f; };
    for (int k = 0; k < 3; ++k) { out->lo[k] = dec(h.scene_lo[k]); out->hi[k] = dec(h.scene_hi[k]); }
    out->tree_height = h.height;
    out->depth4 = h.depth4;
    out->stack4_need = h.stack_need;
    out->max_leaf = h.max_leaf;
    out->n_leaves = h.n_leaves;
    out->launches = launches;
    if (h.height + 2 > RTB_STACK) {
        rtb_set_error("LBVH height " + std::to_string(h.height) + " exceeds the traversal stack (" +
                      std::to_string(RTB_STACK) + ")");
        return RTB_ERR_INVALID;
    }
    return RTB_OK;
}

// Self-test of the builder's sort and scan against the host (std::stable_sort / a running sum): n random pairs with
// `key_bits` significant key bits (few bits = many equal keys = the stability check).  RTB_OK or RTB_ERR_INVALID.
int rtb_sort_selftest(uint32_t n, int key_bits, uint64_t seed) {
    t_alloc_stream = nullptr;            // scratch on the default stream (a previous build's stream may be gone)
    std::vector<unsigned long long> keys(n);
    std::vector<uint32_t> vals(n);
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 1;
    const unsigned long long mask = key_bits >= 64 ? ~0ull : ((1ull << key_bits) - 1ull);
    for (uint32_t i = 0; i < n; ++i) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        keys[i] = x & mask; vals[i] = i;
    }
    DevBuf<unsigned long long> ka, kb, s64;
    DevBuf<uint32_t> va, vb, s32;
    DevBuf<uint8_t> tmp;
    RTB_CUDA(ka.alloc(n)); RTB_CUDA(kb.alloc(n)); RTB_CUDA(va.alloc(n)); RTB_CUDA(vb.alloc(n));
    RTB_CUDA(s64.alloc(n)); RTB_CUDA(s32.alloc(n));
    RTB_CUDA(tmp.alloc(std::max(rtbsort::sort_tmp_bytes(n), rtbsort::scan_tmp_bytes<unsigned long long>(n))));
    RTB_CUDA(cudaMemcpy(ka.p, keys.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(va.p, vals.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    bool in_b = false;
    rtbsort::radix_sort_pairs(ka.p, kb.p, va.p, vb.p, n, key_bits, tmp.p, 0, &in_b);
    std::vector<unsigned long long> gk(n);
    std::vector<uint32_t> gv(n);
    RTB_CUDA(cudaMemcpy(gk.data(), in_b ? kb.p : ka.p, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
    RTB_CUDA(cudaMemcpy(gv.data(), in_b ? vb.p : va.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    std::vector<uint32_t> order(n);
    for (uint32_t i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    for (uint32_t i = 0; i < n; ++i)
        if (gv[i] != order[i] || gk[i] != keys[order[i]]) {
            rtb_set_error("radix sort mismatch at " + std::to_string(i));
            return RTB_ERR_INVALID;
        }
    // scans: 64-bit over the keys' low 20 bits, 32-bit over the values' low 4 bits
    std::vector<unsigned long long> h64(n);
    std::vector<uint32_t> h32(n);
    for (uint32_t i = 0; i < n; ++i) { h64[i] = keys[i] & 0xfffffull; h32[i] = vals[i] & 15u; }
    RTB_CUDA(cudaMemcpy(s64.p, h64.data(), sizeof(unsigned long long) * n, cudaMemcpyHostToDevice));
    RTB_CUDA(cudaMemcpy(s32.p, h32.data(), sizeof(uint32_t) * n, cudaMemcpyHostToDevice));
    rtbsort::exclusive_sum<unsigned long long>(s64.p, s64.p, n, tmp.p, 0);
    std::vector<unsigned long long> g64(n);
    RTB_CUDA(cudaMemcpy(g64.data(), s64.p, sizeof(unsigned long long) * n, cudaMemcpyDeviceToHost));
    rtbsort::exclusive_sum<uint32_t>(s32.p, s32.p, n, tmp.p, 0);
    std::vector<uint32_t> g32(n);
    RTB_CUDA(cudaMemcpy(g32.data(), s32.p, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
    unsigned long long r64 = 0; uint32_t r32 = 0;
    for (uint32_t i = 0; i < n; ++i) {
        if (g64[i] != r64 || g32[i] != r32) { rtb_set_error("exclusive sum mismatch at " + std::to_string(i)); return RTB_ERR_INVALID; }
        r64 += h64[i]; r32 += h32[i];
    }
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
