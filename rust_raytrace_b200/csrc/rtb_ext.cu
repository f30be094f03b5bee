// rtb_ext.cu — EXTENSION renderer: analytic spheres and shadow rays.
//
// Neither exists in the mounted reference any more: the sphere primitive survives only as the
// CollisionFace::{Side, Face} vestiges (raytrace.rs:311-318) and circles_2k.png of an older revision; the shadow test is
// the commented-out block of color_ray (raytrace.rs:1203-1224) with LightSource::get_shadow_ray (:594-610).  BASELINE
// config 1 ("circles scene, primary + shadow rays") asks for both, so they are defined here and in the oracle
// (oracle/rt_oracle.cpp, same arithmetic, compared bit for bit):
//
//   sphere      |o + t d - c|^2 = r^2 with |d| = 1: b = dot(o-c, d), disc = b*b - (|o-c|^2 - r*r), t = -b - sqrt(disc),
//               or -b + sqrt(disc) when that is negative (origin inside); t < 0 rejects (Triangle::intersects' rule);
//               outward normal unit(p - c), flipped for the back face like Triangle::normal (:441-449)
//   shadow      `shadowed` is evaluated first for EVERY hit (as in the commented code): a point of the light cube
//               orig + rand()*len2 per axis, light ray from p + n * 0.005 * (rand() + 1) towards it (make_ray normalises
//               again); shadowed iff ANY other object intersects that ray anywhere (`intersects(..).is_some()`, no
//               distance limit); then `if !shadowed {color} else {black}` in all three SurfaceKind arms (:1228-1252)
//
// A sphere travels through the builder as a pseudo-`Triangle` (norm = 0, incenter = centre, bounding_r2 = r*r, corners =
// its AABB, kind | RTB_PRIM_SPHERE) so that the BVH pipeline is unchanged; the leaf test tells the two apart by the
// all-zero normal.  This file is the one-kernel renderer for such scenes (one thread per pixel, the structure of
// rtb_trace.cu's k_trace, BVH2 nodes): round 1's only renderer for them, now the A/B baseline (RTB_FLAG_MEGAKERNEL) of the
// wavefront renderer's EXT variants (rtb_wavefront.cu), which run the same arithmetic from rtb_device.cuh.
#include "rtb_device.cuh"

using namespace rtbdev;

namespace {

struct SlabRay { float ix, iy, iz, ox, oy, oz; };
__device__ __forceinline__ SlabRay slab_ray(V3 o, V3 d) {
    SlabRay s;
    s.ix = fminf(fmaxf(1.0f / d.x, -1e30f), 1e30f);
    s.iy = fminf(fmaxf(1.0f / d.y, -1e30f), 1e30f);
    s.iz = fminf(fmaxf(1.0f / d.z, -1e30f), 1e30f);
    s.ox = -o.x * s.ix; s.oy = -o.y * s.iy; s.oz = -o.z * s.iz;
    return s;
}
__device__ __forceinline__ bool slab(const SlabRay& r, float4 lo, float4 hi, float tbest, float* tn_out) {
    const float x0 = fmaf(lo.x, r.ix, r.ox), x1 = fmaf(hi.x, r.ix, r.ox);
    const float y0 = fmaf(lo.y, r.iy, r.oy), y1 = fmaf(hi.y, r.iy, r.oy);
    const float z0 = fmaf(lo.z, r.iz, r.oz), z1 = fmaf(hi.z, r.iz, r.oz);
    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000005f;
    *tn_out = tn;
    return (tn <= tf) && (tn <= tbest);
}

// BVH2 traversal (32-byte nodes, sibling pairs adjacent).  ANY = shadow query: stop at the first accepted primitive
// whose original index differs from `exclude`; otherwise closest hit (min t, lowest index on exact ties).
template <bool ANY, bool STATS>
__device__ Hit traverse(const SceneDev& sc, V3 o, V3 d, uint32_t exclude, unsigned long long& n_node,
                        unsigned long long& n_tri) {
    Hit h; h.t = FLT_MAX; h.slot = -1; h.orig = 0xffffffffu;
    if (sc.n_prims == 0u) return h;
    const SlabRay sr = slab_ray(o, d);
    float tbest = FLT_MAX;
    auto leaf = [&](uint32_t first, uint32_t cnt) -> bool {
        for (uint32_t k = first; k < first + cnt; ++k) {
            float t;
            if (STATS) ++n_tri;
            const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
            const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
            if (ANY) {
                if (orig != exclude && prim_test(q, o, d, false, 0.0f, &t)) { h.t = t; h.slot = (int)k; h.orig = orig; return true; }
            } else if (prim_test(q, o, d, h.slot >= 0, h.t, &t)) {
                if (h.slot < 0 || t < h.t || (t == h.t && orig < h.orig)) {
                    h.t = t; h.slot = (int)k; h.orig = orig;
                    if (t < tbest) tbest = t;
                }
            }
        }
        return false;
    };
    const float4 r0 = __ldg(sc.nodes + 0), r1 = __ldg(sc.nodes + 1);
    if (__float_as_uint(r1.w) != 0u) { leaf(__float_as_uint(r0.w), __float_as_uint(r1.w)); return h; }   // single-leaf scene
    uint32_t stack[RTB_STACK];
    int sp = 0;
    uint32_t node = __float_as_uint(r0.w);          // left node of the pair under the root
    for (;;) {
        const float4 a0 = __ldg(sc.nodes + 2 * node + 0), a1 = __ldg(sc.nodes + 2 * node + 1);
        const float4 b0 = __ldg(sc.nodes + 2 * node + 2), b1 = __ldg(sc.nodes + 2 * node + 3);
        if (STATS) n_node += 2;
        float ta, tb;
        const bool hit_a = slab(sr, a0, a1, tbest, &ta), hit_b = slab(sr, b0, b1, tbest, &tb);
        uint32_t next0 = 0xffffffffu, next1 = 0xffffffffu;
        float tn0 = 0.f, tn1 = 0.f;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            if (!(side ? hit_b : hit_a)) continue;
            const uint32_t a = __float_as_uint(side ? b0.w : a0.w), cnt = __float_as_uint(side ? b1.w : a1.w);
            const float tn = side ? tb : ta;
            if (cnt == 0u) {
                if (next0 == 0xffffffffu) { next0 = a; tn0 = tn; } else { next1 = a; tn1 = tn; }
            } else {
                if (tn > tbest) continue;
                if (leaf(a, cnt)) return h;
            }
        }
        if (next1 != 0xffffffffu) {
            if (tn1 < tn0) { const uint32_t s = next0; next0 = next1; next1 = s; }
            RTB_DASSERT(sp < RTB_STACK && next1 + 1u < sc.n_nodes);
            stack[sp++] = next1;
            node = next0;
            continue;
        }
        if (next0 != 0xffffffffu) { node = next0; continue; }
        if (sp == 0) break;
        node = stack[--sp];
    }
    return h;
}

// color_ray (raytrace.rs:1199-1254) with the shadow block live and sphere normals (pieces in rtb_device.cuh, shared with
// the wavefront renderer).  Returns 0 = terminal colour, 1 = bounce: (*color, *alpha) for the mix stack and the next ray.
template <bool STATS>
__device__ int shade_ext(const SceneDev& sc, const ExtParams& ex, const Hit& h, V3 o, V3 d, Rng& g, V3* color, float* alpha,
                         V3* no, V3* nd, unsigned long long& n_node, unsigned long long& n_tri) {
    const ExtHit e = ext_hit_geometry(sc, h.slot, h.t, o, d);
    bool shadowed = false;
    if (ex.has_light) {
        V3 so, sd;
        ext_shadow_ray(ex, e, g, &so, &sd);
        const Hit sh = traverse<true, STATS>(sc, so, sd, h.orig, n_node, n_tri);
        shadowed = sh.slot >= 0;
    }
    return ext_shade(e, d, shadowed, g, color, alpha, no, nd);
}

template <bool STATS>
__global__ void __launch_bounds__(128) k_trace_ext(const SceneDev sc, const ViewDev vw, const ExtParams ex,
                                                   float4* __restrict__ rgba, uint32_t* __restrict__ prim_out,
                                                   float* __restrict__ t_out, TraceCounters* __restrict__ counters) {
    const uint32_t tile = blockIdx.x;
    const uint32_t ty_local = tile / vw.tiles_x, tx = tile - ty_local * vw.tiles_x;
    const uint32_t my_ty = ty_local + vw.band_begin;
    const uint32_t ty = my_ty * vw.tile_world + vw.tile_rank;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const uint32_t col = tx * RTB_TILE_W + lx, row = ty * RTB_TILE_H + ly;
    unsigned long long n_rays = 0, n_node = 0, n_tri = 0;
    if (col < vw.width && row < vw.height) {
        const uint64_t pix = (uint64_t)row * vw.width + col;
        V3 acc = mk(0.0f, 0.0f, 0.0f);
        uint32_t first_prim = 0;
        float first_t = 0.0f;
        for (uint32_t smp = vw.s_begin; smp < vw.s_end; ++smp) {
            Rng g;
            g.seed(vw.seed_mixed, pix, smp);
            float u_off = 0.5f, v_off = 0.5f;
            if (vw.spp != 1) { u_off = g.next_f32(); v_off = g.next_f32(); }
            V3 o, d;
            gen_primary(vw, row, col, u_off, v_off, &o, &d);
            V3 cstack[RTB_MAX_DEPTH];
            float astack[RTB_MAX_DEPTH];
            int level = 0;
            V3 term = mk(0.0f, 0.0f, 0.0f);
            uint32_t depth = vw.maxdepth;
            for (;;) {                                           // project_ray :1256-1295, iteratively
                ++n_rays;
                const Hit h = traverse<false, STATS>(sc, o, d, 0xffffffffu, n_node, n_tri);
                if (level == 0 && smp == 0) { first_prim = h.slot >= 0 ? h.orig : 0u; first_t = h.slot >= 0 ? h.t : 0.0f; }
                if (h.slot < 0) { term = sky_color(); break; }
                V3 color, no, nd;
                float alpha = 0.0f;
                if (shade_ext<STATS>(sc, ex, h, o, d, g, &color, &alpha, &no, &nd, n_node, n_tri) == 0) { term = color; break; }
                cstack[level] = color;
                astack[level] = alpha;
                ++level;
                if (--depth == 0) { term = mk(0.0f, 0.0f, 0.0f); break; }   // project_ray(depth 0): black, not counted
                o = no; d = nd;
            }
            V3 cres = term;
            for (int k = level - 1; k >= 0; --k) cres = mix_color(cstack[k], cres, astack[k]);
            acc = vadd(acc, cres);
        }
        if (!(vw.flags & RTB_FLAG_SUM_ONLY)) acc = vmul(acc, __fdiv_rn(1.0f, (float)vw.spp));
        const uint32_t out_row = vw.compact ? (my_ty * RTB_TILE_H + ly) : row;
        const size_t oi = (size_t)out_row * vw.width + col;
        rgba[oi] = make_float4(acc.x, acc.y, acc.z, 0.0f);
        if (prim_out) prim_out[oi] = first_prim;
        if (t_out) t_out[oi] = first_t;
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_xor_sync(0xffffffffu, n_rays, off);
        if (STATS) {
            n_node += __shfl_xor_sync(0xffffffffu, n_node, off);
            n_tri += __shfl_xor_sync(0xffffffffu, n_tri, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&counters->rays, n_rays);
        if (STATS) { atomicAdd(&counters->node_tests, n_node); atomicAdd(&counters->tri_tests, n_tri); }
    }
}

}  // namespace

int rtb_launch_trace_ext(const SceneDev& sc, const ViewDev& vw, const ExtParams& ex, float4* d_rgba, uint32_t* d_prim,
                         float* d_t, TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches) {
    const uint32_t n_tiles = vw.tiles_x * vw.my_tile_rows;
    if (n_tiles == 0) return RTB_OK;
    if (vw.flags & RTB_FLAG_STATS)
        k_trace_ext<true><<<n_tiles, 128, 0, stream>>>(sc, vw, ex, d_rgba, d_prim, d_t, d_counters);
    else
        k_trace_ext<false><<<n_tiles, 128, 0, stream>>>(sc, vw, ex, d_rgba, d_prim, d_t, d_counters);
    if (launches) ++*launches;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
