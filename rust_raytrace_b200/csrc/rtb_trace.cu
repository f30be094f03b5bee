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
#include <cfloat>

#include "rtb_internal.cuh"

namespace {

// ---------------------------------------------------------------------------
// exact f32 vector math (raytrace.rs:35-96)
// ---------------------------------------------------------------------------
struct V3 { float x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vmul(V3 a, float s) { return mk(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    float s = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    s = __fadd_rn(s, __fmul_rn(a.z, b.z));
    return __fadd_rn(s, 0.0f);   // lane 3 of the reference's f32x4 (0*0)
}
__device__ __forceinline__ V3 vunit(V3 a) {
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(vdot(a, a)));
    return vmul(a, inv);
}

// ---------------------------------------------------------------------------
// RNG: pcg32 keyed by (seed, pixel, sample); floats as rand 0.8's Standard f32.
// Same integer spec as oracle/rt_oracle.cpp so stochastic paths compare bit for bit.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state;
    __device__ __forceinline__ void seed(uint64_t seed, uint64_t pixel, uint32_t sample) {
        uint64_t s = splitmix64(seed);
        s = splitmix64(s ^ pixel);
        s = splitmix64(s ^ (uint64_t)sample);
        state = s;
    }
    __device__ __forceinline__ float next_f32() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        uint32_t r = __funnelshift_r(xorshifted, xorshifted, rot);
        return __fmul_rn((float)(r >> 8), 1.0f / 16777216.0f);
    }
};
__device__ __forceinline__ V3 random_vec(Rng& g) {   // raytrace.rs:188-192
    float a = __fsub_rn(g.next_f32(), 0.5f);
    float b = __fsub_rn(g.next_f32(), 0.5f);
    float c = __fsub_rn(g.next_f32(), 0.5f);
    return vunit(mk(a, b, c));
}

// ---------------------------------------------------------------------------
// closest hit
// ---------------------------------------------------------------------------
struct Hit {
    float t;
    int slot;        // leaf-order primitive slot, -1 = miss
    uint32_t orig;   // original triangle index of `slot`
};

// The acceptance part of Triangle::intersects (raytrace.rs:402-422) for one primitive.
// Returns true and t when the reference would return Some(..).  `has`/`best` allow skipping work
// that cannot change the running minimum (t > best can never win under strict-< / lowest-index).
__device__ __forceinline__ bool tri_test(const float4* __restrict__ q, V3 o, V3 d, bool has, float best, float* t_out) {
    const float4 q0 = __ldg(q + 0), q1 = __ldg(q + 1);
    const V3 n = mk(q0.x, q0.y, q0.z), c = mk(q1.x, q1.y, q1.z);
    const float t = __fdiv_rn(vdot(n, vsub(c, o)), vdot(n, d));
    if (t < 0.0f) return false;
    if (has && t > best) return false;
    const V3 ip = vsub(vadd(vmul(d, t), o), c);
    if (vdot(ip, ip) > q0.w) return false;
    const float4 q2 = __ldg(q + 2);
    if (vdot(ip, mk(q2.x, q2.y, q2.z)) > q2.w) return false;
    const float4 q3 = __ldg(q + 3);
    if (vdot(ip, mk(q3.x, q3.y, q3.z)) > q3.w) return false;
    const float4 q4 = __ldg(q + 4);
    if (vdot(ip, mk(q4.x, q4.y, q4.z)) > q4.w) return false;
    *t_out = t;
    return true;
}

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
    const uint32_t my_ty = tile / vw.tiles_x, tx = tile - my_ty * vw.tiles_x;
    const uint32_t ty = my_ty * vw.tile_world + vw.tile_rank;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t lx = (warp & 1) * 8 + (lane & 7), ly = (warp >> 1) * 4 + (lane >> 3);
    const uint32_t col = tx * RTB_TILE_W + lx, row = ty * RTB_TILE_H + ly;
    const bool active = (col < vw.width) && (row < vw.height);

    unsigned long long n_rays = 0, n_node = 0, n_tri = 0;

    if (active) {
        const V3 v_orig = mk(vw.orig[0], vw.orig[1], vw.orig[2]);
        const V3 v_cam = mk(vw.cam[0], vw.cam[1], vw.cam[2]);
        // pixel_ray :1379-1380
        const V3 vu_delta = vmul(mk(vw.vu[0], vw.vu[1], vw.vu[2]), __fdiv_rn(1.0f, (float)vw.width));
        const V3 vv_delta = vmul(mk(vw.vv[0], vw.vv[1], vw.vv[2]), __fdiv_rn(1.0f, (float)vw.height));
        const V3 blue = mk(__fdiv_rn(128.0f, 255.0f), __fdiv_rn(180.0f, 255.0f), __fdiv_rn(255.0f, 255.0f));
        const uint64_t pix = (uint64_t)row * vw.width + col;

        V3 acc = mk(0.0f, 0.0f, 0.0f);
        uint32_t first_prim = 0;
        float first_t = 0.0f;

        for (uint32_t smp = vw.s_begin; smp < vw.s_end; ++smp) {
            Rng g;
            g.seed(vw.seed, pix, smp);
            float u_off = 0.5f, v_off = 0.5f;
            if (vw.spp != 1) { u_off = g.next_f32(); v_off = g.next_f32(); }
            // pixel_ray :1388-1393 (px = (row, col): px_x = row, px_y = col)
            const V3 vu_frac = vmul(vu_delta, __fadd_rn((float)col, u_off));
            const V3 vv_frac = vmul(vv_delta, __fadd_rn((float)row, v_off));
            V3 o = vadd(vadd(v_orig, vu_frac), vv_frac);
            V3 d = vunit(vunit(vsub(o, v_cam)));   // unit() at :1393, again inside make_ray :202

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
                // classify the hit: Triangle::intersects :406-436 recomputed for the winner
                const float4* q = sc.tri + (size_t)RTB_TRI_F4 * (uint32_t)h.slot;
                const float4 q0 = __ldg(q + 0), q1 = __ldg(q + 1);
                const float4 s0 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)h.slot);
                const float4 s1 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)h.slot + 1);
                const V3 n = mk(q0.x, q0.y, q0.z);
                const V3 p = vadd(vmul(d, h.t), o);
                const V3 ip = vsub(p, mk(q1.x, q1.y, q1.z));
                const float edge_k = __fsub_rn(1.0f, s1.z);
                bool hit_edge = false;
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const float4 qs = __ldg(q + 2 + i);
                    const float dist = vdot(ip, mk(qs.x, qs.y, qs.z));
                    if (dist > __fmul_rn(qs.w, edge_k)) hit_edge = true;
                }
                if (hit_edge) { term = mk(0.0f, 0.0f, 0.0f); break; }   // getsurface :450-459
                const uint32_t kind = __float_as_uint(s1.x);
                const V3 color = mk(s0.x, s0.y, s0.z);
                if (kind == RTB_SOLID) { term = color; break; }
                const bool back = vdot(d, n) > 0.0f;                     // :425-435
                const V3 nn = back ? vmul(n, -1.0f) : n;                 // normal() :441-449
                V3 no, nd;
                if (kind == RTB_MATTE) {                                 // lambertian_ray :292-297
                    const V3 rv = random_vec(g);
                    no = vadd(p, vmul(rv, 0.001f));
                    nd = vunit(vadd(nn, rv));
                } else {                                                 // reflect_ray :278-290
                    const float ddot = fabsf(vdot(d, nn));
                    const V3 dir_p = vmul(nn, ddot);
                    const V3 dir_o = vadd(d, dir_p);
                    const V3 reflect = vadd(dir_p, dir_o);
                    const V3 rv = vmul(random_vec(g), s1.y);
                    const V3 rd = vunit(vadd(reflect, rv));
                    no = vadd(p, vmul(rd, 0.001f));
                    nd = vunit(rd);                                      // make_ray normalises again
                }
                cstack[level] = color;
                astack[level] = s0.w;
                ++level;
                --depth;
                if (depth == 0) { term = mk(0.0f, 0.0f, 0.0f); break; }  // project_ray(depth 0): black, not counted
                o = no; d = nd;
            }
            // unwind the recursion: mix_color(c, sub, a) = c*(1-a) + sub*a  (:299-301), innermost first
            V3 cres = term;
            for (int k = level - 1; k >= 0; --k)
                cres = vadd(vmul(cstack[k], __fsub_rn(1.0f, astack[k])), vmul(cres, astack[k]));
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

// write_png's `(c*255.) as u8` (raytrace.rs:1468-1473): truncating, saturating, NaN -> 0
__global__ void k_quantize(const float4* __restrict__ rgba, uint64_t npix, uint8_t* __restrict__ rgb) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npix) return;
    const float4 c = rgba[i];
    const float v[3] = {__fmul_rn(c.x, 255.0f), __fmul_rn(c.y, 255.0f), __fmul_rn(c.z, 255.0f)};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t q = 0;
        if (v[k] >= 255.0f) q = 255u;
        else if (v[k] > 0.0f) q = (uint32_t)v[k];
        rgb[3 * i + k] = (uint8_t)q;
    }
}

// Cross-GPU reduce over peer memory: every pointer in `bufs` may live on another GPU (NVLink P2P).
// Sums in GPU-index order (deterministic), scales by 1/spp (walk_ray_set :1426).
__global__ void k_peer_reduce(const float4* const* __restrict__ bufs, int n_bufs, float inv_spp, uint64_t first,
                              uint64_t count, float4* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count;
         i += (uint64_t)gridDim.x * blockDim.x) {
        float4 s = bufs[0][first + i];
        for (int g = 1; g < n_bufs; ++g) {
            const float4 v = bufs[g][first + i];
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

int rtb_launch_quantize(const float4* d_rgba, uint64_t npix, uint8_t* d_rgb, cudaStream_t stream) {
    if (npix == 0) return RTB_OK;
    k_quantize<<<(unsigned)((npix + 255) / 256), 256, 0, stream>>>(d_rgba, npix, d_rgb);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}

int rtb_launch_peer_reduce(const float4* const* d_bufs_on_device, int n_bufs, float inv_spp, uint64_t first,
                           uint64_t count, float4* d_out, cudaStream_t stream) {
    if (count == 0) return RTB_OK;
    k_peer_reduce<<<148 * 8, 256, 0, stream>>>(d_bufs_on_device, n_bufs, inv_spp, first, count, d_out);
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
