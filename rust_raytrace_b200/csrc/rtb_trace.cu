// rtb_trace.cu — the hot path: per-pixel primary-ray generation, BVH traversal, exact
// ray/triangle test, bounce shading and per-pixel accumulation in ONE kernel.
//
// Reference functions restated here (raytrace_lib/src/raytrace.rs):
//   Viewport::pixel_ray            :1374-1394     -> gen_primary()
//   make_ray                       :201-210       -> (direction normalisation only; inv_dir is octree-only)
//   Triangle::intersects           :400-439       -> tri_test() + classify_hit()
//   get_box_min_time_intersection  :1013-1050     -> the leaf loop of closest_hit() (strict <, lowest index on ties)
//   get_object_intersection_for_ray:910-1010      -> replaced by BVH traversal with the same closest-hit result
//   project_ray / color_ray        :1199-1295     -> the bounce loop of k_trace() (iterative, folded innermost-first)
//   reflect_ray / lambertian_ray / random_vec / mix_color :188-192, :278-301
//   walk_ray_set                   :1413-1427     -> sample loop + acc * (1/spp)
//
// Arithmetic contract: everything that feeds a value the reference also computes uses the
// explicit round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn), which
// nvcc never contracts into FMA, in the reference's operation order (dot = ((p0+p1)+p2)+0).
// Only the AABB slab test, which has no counterpart in the reference's result, is free-form.
#include "rtb_device.cuh"

using namespace rtbdev;

namespace {

template <bool STATS>
__device__ __forceinline__ Hit closest_hit(const SceneDev& sc, V3 o, V3 d, bool brute, unsigned long long& n_node,
                                           unsigned long long& n_tri) {
    Hit h; h.t = FLT_MAX; h.slot = -1; h.orig = 0xffffffffu;
    if (sc.n_prims == 0u) return h;   // empty scene (only the dummy triangle): everything is sky
    // slab-test constants (not part of the exactness contract; boxes are padded, see rtb_lbvh.cu)
    // |1/d| is clamped to 1e30 so that a zero direction component gives +-huge instead of inf: with inf the
    // fused form lo*inv - o*inv turns into inf - inf = NaN on one plane only and the slab collapses.
    const float ix = fminf(fmaxf(1.0f / d.x, -1e30f), 1e30f);
    const float iy = fminf(fmaxf(1.0f / d.y, -1e30f), 1e30f);
    const float iz = fminf(fmaxf(1.0f / d.z, -1e30f), 1e30f);
    const float ox = -o.x * ix, oy = -o.y * iy, oz = -o.z * iz;
    float tbest = FLT_MAX;   // +inf-like culling bound until the first hit

    uint32_t stack[RTB_STACK];
    int sp = 0;
    uint32_t node = 0;
    // root: leaf or internal?  (brute = validation mode: one linear scan over every primitive)
    {
        const float4 r1 = __ldg(sc.nodes + 1);
        if (brute || __float_as_uint(r1.w) != 0u) {
            // single-leaf scene
            const uint32_t first = brute ? 0u : __float_as_uint(__ldg(sc.nodes + 0).w);
            const uint32_t cnt = brute ? sc.n_prims : __float_as_uint(r1.w);
            for (uint32_t k = first; k < first + cnt; ++k) {
                float t;
                if (STATS) ++n_tri;
                const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
                if (tri_test(q, o, d, h.slot >= 0, h.t, &t)) {
                    const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
                    if (h.slot < 0 || t < h.t || (t == h.t && orig < h.orig)) { h.t = t; h.slot = (int)k; h.orig = orig; }
                }
            }
            return h;
        }
        node = __float_as_uint(__ldg(sc.nodes + 0).w);   // left child of the root
    }
    // `node` always names the LEFT node of a sibling pair to be tested.
    for (;;) {
        const float4 a0 = __ldg(sc.nodes + 2 * node + 0), a1 = __ldg(sc.nodes + 2 * node + 1);
        const float4 b0 = __ldg(sc.nodes + 2 * node + 2), b1 = __ldg(sc.nodes + 2 * node + 3);
        if (STATS) n_node += 2;
        // slab tests, t in [0, tbest]
        float ta, tb;
        bool hit_a, hit_b;
        {
            float x0 = fmaf(a0.x, ix, ox), x1 = fmaf(a1.x, ix, ox);
            float y0 = fmaf(a0.y, iy, oy), y1 = fmaf(a1.y, iy, oy);
            float z0 = fmaf(a0.z, iz, oz), z1 = fmaf(a1.z, iz, oz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
            tf = tf * 1.0000005f;
            hit_a = (tn <= tf) && (tn <= tbest);
            ta = tn;
        }
        {
            float x0 = fmaf(b0.x, ix, ox), x1 = fmaf(b1.x, ix, ox);
            float y0 = fmaf(b0.y, iy, oy), y1 = fmaf(b1.y, iy, oy);
            float z0 = fmaf(b0.z, iz, oz), z1 = fmaf(b1.z, iz, oz);
            float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
            float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
            tf = tf * 1.0000005f;
            hit_b = (tn <= tf) && (tn <= tbest);
            tb = tn;
        }
        // Resolve leaves immediately, collect at most two internal children to descend into.
        uint32_t next0 = 0xffffffffu, next1 = 0xffffffffu;   // left-child indices of internal children
        float tn0 = 0.f, tn1 = 0.f;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            const bool hit = side ? hit_b : hit_a;
            if (!hit) continue;
            const uint32_t a = __float_as_uint(side ? b0.w : a0.w), cnt = __float_as_uint(side ? b1.w : a1.w);
            const float tn = side ? tb : ta;
            if (cnt == 0u) {
                if (next0 == 0xffffffffu) { next0 = a; tn0 = tn; } else { next1 = a; tn1 = tn; }
            } else {
                if (tn > tbest) continue;   // the other child may have tightened the bound
                for (uint32_t k = a; k < a + cnt; ++k) {
                    float t;
                    if (STATS) ++n_tri;
                    const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
                    if (tri_test(q, o, d, h.slot >= 0, h.t, &t)) {
                        const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
                        if (h.slot < 0 || t < h.t || (t == h.t && orig < h.orig)) {
                            h.t = t; h.slot = (int)k; h.orig = orig;
                            if (t < tbest) tbest = t;   // NaN t never tightens the bound
                        }
                    }
                }
            }
        }
        if (next1 != 0xffffffffu) {
            // both internal: near one first, far one on the stack
            if (tn1 < tn0) { uint32_t s = next0; next0 = next1; next1 = s; }
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

// ---------------------------------------------------------------------------
// the megakernel: one thread = one pixel, all samples and all bounces
// ---------------------------------------------------------------------------
template <bool STATS>
__global__ void __launch_bounds__(128) k_trace(const SceneDev sc, const ViewDev vw, float4* __restrict__ rgba,
                                               uint32_t* __restrict__ prim_out, float* __restrict__ t_out,
                                               TraceCounters* __restrict__ counters) {
    // tile -> pixel mapping: block = 16x8 pixels, warp = 8x4 pixels
    const uint32_t tile = blockIdx.x;
    const uint32_t ty_local = tile / vw.tiles_x, tx = tile - ty_local * vw.tiles_x;
    const uint32_t my_ty = ty_local + vw.band_begin;
    const uint32_t ty = my_ty * vw.tile_world + vw.tile_rank;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const uint32_t col = tx * RTB_TILE_W + lx, row = ty * RTB_TILE_H + ly;
    const bool active = (col < vw.width) && (row < vw.height);

    unsigned long long n_rays = 0, n_node = 0, n_tri = 0;

    if (active) {
        const V3 blue = sky_color();
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
            for (;;) {
                // project_ray :1256-1295 with depth > 0
                ++n_rays;
                const Hit h = closest_hit<STATS>(sc, o, d, (vw.flags & RTB_FLAG_BRUTE) != 0u, n_node, n_tri);
                if (level == 0 && smp == 0) { first_prim = h.slot >= 0 ? h.orig : 0u; first_t = h.slot >= 0 ? h.t : 0.0f; }
                if (h.slot < 0) { term = blue; break; }
                V3 color, no, nd;
                float alpha = 0.0f;
                if (shade_hit(sc, h.slot, h.t, o, d, g, &color, &alpha, &no, &nd) == 0) { term = color; break; }
                cstack[level] = color;
                astack[level] = alpha;
                ++level;
                --depth;
                if (depth == 0) { term = mk(0.0f, 0.0f, 0.0f); break; }  // project_ray(depth 0): black, not counted
                o = no; d = nd;
            }
            // unwind the recursion: mix_color(c, sub, a) = c*(1-a) + sub*a  (:299-301), innermost first
            V3 cres = term;
            for (int k = level - 1; k >= 0; --k) cres = mix_color(cstack[k], cres, astack[k]);
            acc = vadd(acc, cres);
        }
        if (!(vw.flags & RTB_FLAG_SUM_ONLY)) acc = vmul(acc, __fdiv_rn(1.0f, (float)vw.spp));   // :1426

        const uint32_t out_row = vw.compact ? (my_ty * RTB_TILE_H + ly) : row;
        const size_t oi = (size_t)out_row * vw.width + col;
        rgba[oi] = make_float4(acc.x, acc.y, acc.z, 0.0f);
        if (prim_out) prim_out[oi] = first_prim;
        if (t_out) t_out[oi] = first_t;
    }

    // one atomic per warp for the ray counter (the reference's "Rays" stat)
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

// write_png's `(c*255.) as u8` (raytrace.rs:1468-1473): truncating, saturating, NaN -> 0.  One thread = four pixels =
// three 32-bit stores (`rgb` is 4-byte aligned: every caller passes the start of a band or of an allocation).
__device__ __forceinline__ uint32_t quant_u8(float c) {
    const float v = __fmul_rn(c, 255.0f);
    return v >= 255.0f ? 255u : (v > 0.0f ? (uint32_t)v : 0u);
}
__global__ void k_quantize(const float4* __restrict__ rgba, uint64_t npix, uint8_t* __restrict__ rgb) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4u;
    if (i0 >= npix) return;
    if (i0 + 4u <= npix) {
        uint32_t b[12];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 c = __ldcs(rgba + i0 + k);
            b[3 * k] = quant_u8(c.x); b[3 * k + 1] = quant_u8(c.y); b[3 * k + 2] = quant_u8(c.z);
        }
        uint32_t* out = reinterpret_cast<uint32_t*>(rgb + 3u * i0);
#pragma unroll
        for (int w = 0; w < 3; ++w) out[w] = b[4 * w] | (b[4 * w + 1] << 8) | (b[4 * w + 2] << 16) | (b[4 * w + 3] << 24);
    } else {
        for (uint64_t i = i0; i < npix; ++i) {
            const float4 c = rgba[i];
            rgb[3 * i] = (uint8_t)quant_u8(c.x); rgb[3 * i + 1] = (uint8_t)quant_u8(c.y); rgb[3 * i + 2] = (uint8_t)quant_u8(c.z);
        }
    }
}

// Cross-GPU reduce over peer memory: every pointer in `bufs` may live on another GPU (NVLink P2P).
// Sums in GPU-index order (deterministic), scales by 1/spp (walk_ray_set :1426).
__global__ void k_peer_reduce(const RtbPeerBufs bufs, int n_bufs, float inv_spp, uint64_t first,
                              uint64_t count, float4* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (uint64_t)gridDim.x * blockDim.x) {
        float4 s = bufs.p[0][first + i];
        for (int g = 1; g < n_bufs; ++g) {
            const float4 v = bufs.p[g][first + i];
            s.x = __fadd_rn(s.x, v.x); s.y = __fadd_rn(s.y, v.y); s.z = __fadd_rn(s.z, v.z);
        }
        out[first + i] = make_float4(__fmul_rn(s.x, inv_spp), __fmul_rn(s.y, inv_spp), __fmul_rn(s.z, inv_spp), 0.0f);
    }
}

}  // namespace

int rtb_launch_trace(const SceneDev& sc, const ViewDev& vw, float4* d_rgba, uint32_t* d_prim, float* d_t,
                     TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches) {
    const uint32_t n_tiles = vw.tiles_x * vw.my_tile_rows;
    if (n_tiles == 0) return RTB_OK;
    if (vw.flags & RTB_FLAG_STATS)
        k_trace<true><<<n_tiles, 128, 0, stream>>>(sc, vw, d_rgba, d_prim, d_t, d_counters);
    else
        k_trace<false><<<n_tiles, 128, 0, stream>>>(sc, vw, d_rgba, d_prim, d_t, d_counters);
    if (launches) ++*launches;
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

// `data[col] = acc * (1/spp)` of walk_ray_set (raytrace.rs:1426) for a sample sum that was reduced across ranks
__global__ void k_scale(float4* __restrict__ rgba, uint64_t npix, float inv_spp) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npix; i += (uint64_t)gridDim.x * blockDim.x) {
        const float4 c = rgba[i];
        rgba[i] = make_float4(__fmul_rn(c.x, inv_spp), __fmul_rn(c.y, inv_spp), __fmul_rn(c.z, inv_spp), 0.f);
    }
}

int rtb_launch_scale(float4* d_rgba, uint64_t npix, float inv_spp, cudaStream_t stream) {
    if (npix == 0) return RTB_OK;
    k_scale<<<148 * 8, 256, 0, stream>>>(d_rgba, npix, inv_spp);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_launch_quantize(const float4* d_rgba, uint64_t npix, uint8_t* d_rgb, cudaStream_t stream) {
    if (npix == 0) return RTB_OK;
    k_quantize<<<(unsigned)((npix + 1023) / 1024), 256, 0, stream>>>(d_rgba, npix, d_rgb);       // 4 pixels per thread
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_launch_peer_reduce(const RtbPeerBufs& bufs, int n_bufs, float inv_spp, uint64_t first,
                           uint64_t count, float4* d_out, cudaStream_t stream) {
    if (count == 0) return RTB_OK;
    k_peer_reduce<<<148 * 8, 256, 0, stream>>>(bufs, n_bufs, inv_spp, first, count, d_out);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
