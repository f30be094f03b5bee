// rtb_lbvh.cu — GPU Morton-code LBVH builder (replaces the reference's octree build,
// build_bounding_box raytrace.rs:790-845, with an accelerator that returns the same
// closest hit; see DESIGN.md for the equivalence argument).
//
// Pipeline (all on one stream, no host round trip until the final info read-back):
//   k_prim_bounds   per-primitive AABB from `corners`, scene AABB by atomic min/max
//   k_morton        63-bit Morton code of the AABB centre
//   radix sort      (key = morton, value = primitive)            [cub::DeviceRadixSort — interim]
//   k_hierarchy     Karras 2012 radix-tree: children, parent, covered range per internal node
//   k_refit         bottom-up AABB union with one atomic arrival counter per internal node
//   exclusive scan  over "this internal node covers > RTB_LEAF_MAX primitives" [cub::DeviceScan — interim]
//   k_emit_nodes    collapse small subtrees into leaves, write 32-byte nodes with adjacent sibling pairs
//   k_emit_tris     gather the 19 intersect floats + shading record of each primitive into leaf order
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <cfloat>

#include "rtb_internal.cuh"

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
};

__global__ void k_init_scratch(BuildScratch* s) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        for (int k = 0; k < 3; ++k) { s->scene_lo[k] = 0xffffffffu; s->scene_hi[k] = 0u; }
        s->max_abs = 0u; s->height = 0u; s->max_leaf = 0u; s->n_leaves = 0u; s->depth4 = 0u;
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

__global__ void k_split_flags(const int2* __restrict__ range, int n_internal, uint32_t* __restrict__ flags) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    flags[i] = (range[i].y - range[i].x + 1) > RTB_LEAF_MAX ? 1u : 0u;
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
                             float4* __restrict__ nodes, BuildScratch* s) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    int total = 2 * n - 1;
    if (c >= total) return;
    const float pad = o2f(s->max_abs) * (1.0f / 131072.0f);   // 2^-17 of the largest |coordinate|
    int first, count;
    if (c >= n - 1) { first = c - (n - 1); count = 1; }
    else { first = range[c].x; count = range[c].y - range[c].x + 1; }
    uint32_t pos;
    if (c == 0) pos = 0;
    else {
        int p = parent[c];
        int psize = range[p].y - range[p].x + 1;
        if (psize <= RTB_LEAF_MAX) return;   // swallowed by a collapsed leaf
        pos = 2u + 2u * slot[p] + (children[p].y == c ? 1u : 0u);
    }
    if (count <= RTB_LEAF_MAX) {
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
// entries: lo.x[4] hi.x[4] lo.y[4] hi.y[4] lo.z[4] hi.z[4] code[4] pad.  code: 0 = empty slot,
// 0x80000000 | first<<3 | count = leaf, otherwise the index of the child BVH4 node.
__global__ void k_flag4(int n_internal, const int2* __restrict__ range, const int* __restrict__ parent,
                        uint32_t* __restrict__ flags4, BuildScratch* s) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    uint32_t depth = 0;
    for (int p = parent[i]; p >= 0; p = parent[p]) ++depth;
    const int size = range[i].y - range[i].x + 1;
    const bool kept = (i == 0) || (size > RTB_LEAF_MAX && (depth & 1u) == 0u);
    flags4[i] = kept ? 1u : 0u;
    if (kept) atomicMax(&s->depth4, depth >> 1);
}

__global__ void k_emit_nodes4(int n, const int2* __restrict__ children, const int2* __restrict__ range,
                              const float4* __restrict__ blo, const float4* __restrict__ bhi,
                              const uint32_t* __restrict__ flags4, const uint32_t* __restrict__ idx4,
                              float4* __restrict__ nodes4, const BuildScratch* __restrict__ s) {
    const int n_internal = n - 1;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (n_internal > 0 ? n_internal : 1)) return;
    if (n_internal > 0 && !flags4[i]) return;
    const float pad = o2f(s->max_abs) * (1.0f / 131072.0f);
    auto leaf_final = [&](int c) { return c >= n - 1 || (range[c].y - range[c].x + 1) <= RTB_LEAF_MAX; };
    int ent[4];
    int n_ent = 0;
    if (n_internal == 0) ent[n_ent++] = 0;                       // the single Karras leaf
    else if (leaf_final(i)) ent[n_ent++] = i;                    // tiny scene: the root itself is a leaf
    else {
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
            if (leaf_final(c)) {
                const uint32_t first = c >= n - 1 ? (uint32_t)(c - (n - 1)) : (uint32_t)range[c].x;
                const uint32_t cnt = c >= n - 1 ? 1u : (uint32_t)(range[c].y - range[c].x + 1);
                code[e] = 0x80000000u | (first << 3) | cnt;
            } else {
                code[e] = idx4[c];
            }
        } else {
            lo[0][e] = lo[1][e] = lo[2][e] = 1e30f;
            hi[0][e] = hi[1][e] = hi[2][e] = -1e30f;
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

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
};

inline uint32_t cdiv(uint32_t a, uint32_t b) { return (a + b - 1) / b; }

}  // namespace

int rtb_build_lbvh(const RtbTriangle* d_tris, const uint32_t* d_keep, uint32_t n, cudaStream_t stream,
                   BuildResult* out) {
    *out = BuildResult();
    cudaEvent_t e0, e1;
    RTB_CUDA(cudaEventCreate(&e0));
    RTB_CUDA(cudaEventCreate(&e1));

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
        RTB_CUDA(cudaMemsetAsync(out->d_nodes4, 0, sizeof(float4) * 8, stream));   // all codes 0 = empty
        RTB_CUDA(cudaStreamSynchronize(stream));
        out->n_nodes4 = 1;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        return RTB_OK;
    }

    DevBuf<BuildScratch> scratch;
    DevBuf<float4> plo, phi, blo, bhi;
    DevBuf<uint64_t> keys, keys_sorted;
    DevBuf<uint32_t> vals, vals_sorted, arrive, flags, slot, flags4, idx4;
    DevBuf<int2> children, range;
    DevBuf<int> parent;
    DevBuf<uint8_t> cub_tmp;
    const uint32_t n_int = n - 1, n_all = 2 * n - 1;
    RTB_CUDA(scratch.alloc(1));
    RTB_CUDA(plo.alloc(n)); RTB_CUDA(phi.alloc(n));
    RTB_CUDA(blo.alloc(n_all)); RTB_CUDA(bhi.alloc(n_all));
    RTB_CUDA(keys.alloc(n)); RTB_CUDA(keys_sorted.alloc(n));
    RTB_CUDA(vals.alloc(n)); RTB_CUDA(vals_sorted.alloc(n));
    RTB_CUDA(arrive.alloc(n_int)); RTB_CUDA(flags.alloc(n_int + 1)); RTB_CUDA(slot.alloc(n_int + 1));
    RTB_CUDA(children.alloc(n_int)); RTB_CUDA(range.alloc(n_int)); RTB_CUDA(parent.alloc(n_all));
    RTB_CUDA(flags4.alloc(n_int + 1)); RTB_CUDA(idx4.alloc(n_int + 1));

    size_t sort_bytes = 0, scan_bytes = 0;
    RTB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n,
                                             0, 63, stream));
    RTB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, flags.p, slot.p, (int)(n_int + 1), stream));
    size_t tmp_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    RTB_CUDA(cub_tmp.alloc(tmp_bytes));

    const uint32_t B = 256;
    uint32_t launches = 0;
    RTB_CUDA(cudaEventRecord(e0, stream));
    k_init_scratch<<<1, 32, 0, stream>>>(scratch.p); ++launches;
    k_prim_bounds<<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, n, plo.p, phi.p, scratch.p); ++launches;
    k_morton<<<cdiv(n, B), B, 0, stream>>>(plo.p, phi.p, n, scratch.p, keys.p, vals.p); ++launches;
    RTB_CUDA(cub::DeviceRadixSort::SortPairs(cub_tmp.p, tmp_bytes, keys.p, keys_sorted.p, vals.p, vals_sorted.p, (int)n,
                                             0, 63, stream));
    launches += 8;  // cub's onesweep: histogram + ~7 digit passes for 63 bits (approximate)
    uint32_t total_split = 0;
    if (n_int > 0) {
        RTB_CUDA(cudaMemsetAsync(arrive.p, 0, sizeof(uint32_t) * n_int, stream));
        k_hierarchy<<<cdiv(n_int, B), B, 0, stream>>>(keys_sorted.p, (int)n, children.p, range.p, parent.p); ++launches;
    } else {
        int minus1 = -1;
        RTB_CUDA(cudaMemcpyAsync(parent.p, &minus1, sizeof(int), cudaMemcpyHostToDevice, stream));
    }
    k_refit<<<cdiv(n, B), B, 0, stream>>>(plo.p, phi.p, vals_sorted.p, (int)n, children.p, parent.p, blo.p, bhi.p,
                                          arrive.p, scratch.p); ++launches;
    k_leaf_depth<<<cdiv(n, B), B, 0, stream>>>((int)n, parent.p, scratch.p); ++launches;
    if (n_int > 0) {
        k_split_flags<<<cdiv(n_int, B), B, 0, stream>>>(range.p, (int)n_int, flags.p); ++launches;
        RTB_CUDA(cudaMemsetAsync(flags.p + n_int, 0, sizeof(uint32_t), stream));
        RTB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, flags.p, slot.p, (int)(n_int + 1), stream));
        launches += 1;
        RTB_CUDA(cudaMemcpyAsync(&total_split, slot.p + n_int, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
    }
    out->n_nodes = 2 + 2 * total_split;
    RTB_CUDA(cudaMalloc(&out->d_nodes, sizeof(float4) * 2 * out->n_nodes));
    k_emit_nodes<<<cdiv(n_all, B), B, 0, stream>>>((int)n, children.p, range.p, parent.p, blo.p, bhi.p, slot.p,
                                                   out->d_nodes, scratch.p); ++launches;
    k_emit_tris<<<cdiv(n, B), B, 0, stream>>>(d_tris, d_keep, vals_sorted.p, n, out->d_tri, out->d_shade,
                                              out->d_prim_order); ++launches;
    // 4-wide collapse of the same tree (used by the wavefront renderer)
    uint32_t total4 = 1;
    if (n_int > 0) {
        k_flag4<<<cdiv(n_int, B), B, 0, stream>>>((int)n_int, range.p, parent.p, flags4.p, scratch.p); ++launches;
        RTB_CUDA(cudaMemsetAsync(flags4.p + n_int, 0, sizeof(uint32_t), stream));
        RTB_CUDA(cub::DeviceScan::ExclusiveSum(cub_tmp.p, tmp_bytes, flags4.p, idx4.p, (int)(n_int + 1), stream));
        launches += 1;
        RTB_CUDA(cudaMemcpyAsync(&total4, idx4.p + n_int, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
        RTB_CUDA(cudaStreamSynchronize(stream));
    }
    out->n_nodes4 = total4;
    RTB_CUDA(cudaMalloc(&out->d_nodes4, sizeof(float4) * 8 * (size_t)total4));
    k_emit_nodes4<<<cdiv(n_int > 0 ? n_int : 1, B), B, 0, stream>>>((int)n, children.p, range.p, blo.p, bhi.p, flags4.p,
                                                                  idx4.p, out->d_nodes4, scratch.p); ++launches;
    RTB_CUDA(cudaEventRecord(e1, stream));
    BuildScratch h;
    RTB_CUDA(cudaMemcpyAsync(&h, scratch.p, sizeof h, cudaMemcpyDeviceToHost, stream));
    RTB_CUDA(cudaStreamSynchronize(stream));
    RTB_CUDA(cudaGetLastError());
    RTB_CUDA(cudaEventElapsedTime(&out->ms_build, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);

    auto dec = [](uint32_t o) { uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o; float f; memcpy(&f, &u, 4); return f; };
    for (int k = 0; k < 3; ++k) { out->lo[k] = dec(h.scene_lo[k]); out->hi[k] = dec(h.scene_hi[k]); }
    out->tree_height = h.height;
    out->depth4 = h.depth4;
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
