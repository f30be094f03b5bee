// rtb_device.cuh — device-side building blocks shared by the megakernel (rtb_trace.cu) and the
// wavefront pipeline (rtb_wavefront.cu): the exact f32 vector math, the RNG and the exact
// ray/triangle test.
//
// Arithmetic contract: everything that feeds a value the reference also computes uses the explicit
// round-to-nearest intrinsics (__fmul_rn/__fadd_rn/__fsub_rn/__fdiv_rn/__fsqrt_rn), which nvcc never
// contracts into FMA, in the reference's operation order.  Reference: raytrace_lib/src/raytrace.rs.
#pragma once

#include <cfloat>

#include "rtb_internal.cuh"

namespace rtbdev {

// ---------------------------------------------------------------------------
// exact f32 vector math (raytrace.rs:35-96)
// ---------------------------------------------------------------------------
struct V3 { float x, y, z; };

__device__ __forceinline__ V3 mk(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ V3 vadd(V3 a, V3 b) { return mk(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vsub(V3 a, V3 b) { return mk(__fsub_rn(a.x, b.x), __fsub_rn(a.y, b.y), __fsub_rn(a.z, b.z)); }
__device__ __forceinline__ V3 vmul(V3 a, float s) { return mk(__fmul_rn(a.x, s), __fmul_rn(a.y, s), __fmul_rn(a.z, s)); }
// dot = lane-wise product then ORDERED reduce over the 4 lanes of the reference's f32x4 (:66, :76):
// ((p0 + p1) + p2) + p3 with p3 = 0*0.
__device__ __forceinline__ float vdot(V3 a, V3 b) {
    float s = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    s = __fadd_rn(s, __fmul_rn(a.z, b.z));
    return __fadd_rn(s, 0.0f);
}
// the same dot for a value that is only COMPARED (x > y): the final "+ 0*0" can only turn -0 into +0, which no comparison sees
__device__ __forceinline__ float vdot_cmp(V3 a, V3 b) {
    float s = __fadd_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y));
    return __fadd_rn(s, __fmul_rn(a.z, b.z));
}
__device__ __forceinline__ V3 vunit(V3 a) {   // :93-96: v * (1/len)
    float inv = __fdiv_rn(1.0f, __fsqrt_rn(vdot(a, a)));
    return vmul(a, inv);
}

// ---------------------------------------------------------------------------
// 256-bit read-only load (sm_100: LDG.E.256.CONSTANT): two adjacent float4 of a 32-byte aligned address in ONE
// instruction.  A BVH4 node (128 B) is then 4 loads instead of 7, and the L1 data pipe — which spends one wavefront per
// distinct 128-byte line per load instruction when every lane of a divergent warp reads its own node — has 43 % less to do.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const float4* __restrict__ p, float4& a, float4& b) {
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

// ---------------------------------------------------------------------------
// RNG: pcg32 keyed by (seed, pixel, sample); floats as rand 0.8's Standard f32 (24-bit mantissa).
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
    // seed_mixed = splitmix64(seed), done once per frame on the host (ViewDev::seed_mixed)
    __device__ __forceinline__ void seed(uint64_t seed_mixed, uint64_t pixel, uint32_t sample) {
        uint64_t s = seed_mixed;
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
// Primary ray (Viewport::pixel_ray :1374-1394 + make_ray :201-210)
// ---------------------------------------------------------------------------
__device__ __forceinline__ void gen_primary(const ViewDev& vw, uint32_t row, uint32_t col, float u_off, float v_off,
                                            V3* o, V3* d) {
    const V3 vu_delta = mk(vw.vu_delta[0], vw.vu_delta[1], vw.vu_delta[2]);   // vu * (1/width), vv * (1/height): make_view
    const V3 vv_delta = mk(vw.vv_delta[0], vw.vv_delta[1], vw.vv_delta[2]);
    const V3 vu_frac = vmul(vu_delta, __fadd_rn((float)col, u_off));   // px = (row, col): px_y = col
    const V3 vv_frac = vmul(vv_delta, __fadd_rn((float)row, v_off));
    *o = vadd(vadd(mk(vw.orig[0], vw.orig[1], vw.orig[2]), vu_frac), vv_frac);
    *d = vunit(vunit(vsub(*o, mk(vw.cam[0], vw.cam[1], vw.cam[2]))));  // unit() at :1393, again in make_ray
}

// ---------------------------------------------------------------------------
// The acceptance part of Triangle::intersects (raytrace.rs:402-422) for one primitive.
// Returns true and t when the reference would return Some(..).  `has`/`best` allow skipping work
// that cannot change the running minimum (t > best can never win under strict-< / lowest-index).
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool tri_test_pre(const float4* __restrict__ q, float4 q0, float4 q1, V3 o, V3 d, bool has,
                                             float best, float* t_out) {
    const V3 n = mk(q0.x, q0.y, q0.z), c = mk(q1.x, q1.y, q1.z);
    const float t = __fdiv_rn(vdot(n, vsub(c, o)), vdot(n, d));
    if (t < 0.0f) return false;
    if (has && t > best) return false;
    const V3 ip = vsub(vadd(vmul(d, t), o), c);
    if (vdot_cmp(ip, ip) > q0.w) return false;
    // the three edge records together: one latency instead of up to three (ncu r1_v5: these dependent loads held
    // 11 % of the bounce kernel's stall samples at ~4 active lanes)
    const float4 q2 = __ldg(q + 2), q3 = __ldg(q + 3), q4 = __ldg(q + 4);
    if (vdot_cmp(ip, mk(q2.x, q2.y, q2.z)) > q2.w) return false;
    if (vdot_cmp(ip, mk(q3.x, q3.y, q3.z)) > q3.w) return false;
    if (vdot_cmp(ip, mk(q4.x, q4.y, q4.z)) > q4.w) return false;
    *t_out = t;
    return true;
}
__device__ __forceinline__ bool tri_test(const float4* __restrict__ q, V3 o, V3 d, bool has, float best, float* t_out) {
    return tri_test_pre(q, __ldg(q + 0), __ldg(q + 1), o, d, has, best, t_out);
}

struct Hit {
    float t;
    int slot;        // leaf-order primitive slot, -1 = miss
    uint32_t orig;   // original triangle index of `slot`
};

// Surface response at a hit (Triangle::intersects :406-436 recomputed for the winner, getsurface :450-459,
// normal :441-449, color_ray :1228-1252).  Returns:
//   0 = terminal, *color is the path's terminal colour (edge -> black, Solid -> colour)
//   1 = bounce: *color/*alpha go on the mix stack, (*no, *nd) is the next ray
__device__ __forceinline__ int shade_hit(const SceneDev& sc, int slot, float t, V3 o, V3 d, Rng& g, V3* color,
                                         float* alpha, V3* no, V3* nd) {
    const float4* q = sc.tri + (size_t)RTB_TRI_F4 * (uint32_t)slot;
    const float4 q0 = __ldg(q + 0), q1 = __ldg(q + 1);
    const float4 s0 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)slot);
    const float4 s1 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)slot + 1);
    const V3 n = mk(q0.x, q0.y, q0.z);
    const V3 p = vadd(vmul(d, t), o);
    const V3 ip = vsub(p, mk(q1.x, q1.y, q1.z));
    const float edge_k = __fsub_rn(1.0f, s1.z);
    bool hit_edge = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float4 qs = __ldg(q + 2 + i);
        const float dist = vdot_cmp(ip, mk(qs.x, qs.y, qs.z));
        if (dist > __fmul_rn(qs.w, edge_k)) hit_edge = true;
    }
    if (hit_edge) { *color = mk(0.0f, 0.0f, 0.0f); return 0; }
    const uint32_t kind = __float_as_uint(s1.x);
    *color = mk(s0.x, s0.y, s0.z);
    if (kind == RTB_SOLID) return 0;
    *alpha = s0.w;
    const bool back = vdot_cmp(d, n) > 0.0f;                 // :425-435
    const V3 nn = back ? vmul(n, -1.0f) : n;
    if (kind == RTB_MATTE) {                                 // lambertian_ray :292-297
        const V3 rv = random_vec(g);
        *no = vadd(p, vmul(rv, 0.001f));
        *nd = vunit(vadd(nn, rv));
    } else {                                                 // reflect_ray :278-290
        const float ddot = fabsf(vdot(d, nn));
        const V3 dir_p = vmul(nn, ddot);
        const V3 dir_o = vadd(d, dir_p);
        const V3 reflect = vadd(dir_p, dir_o);
        const V3 rv = vmul(random_vec(g), s1.y);
        const V3 rd = vunit(vadd(reflect, rv));
        *no = vadd(p, vmul(rd, 0.001f));
        *nd = vunit(rd);                                     // make_ray normalises again
    }
    return 1;
}

// ---------------------------------------------------------------------------
// EXTENSION: analytic spheres and shadow rays, shared by the one-kernel extension renderer
// (rtb_ext.cu) and the wavefront renderer (rtb_wavefront.cu, EXT variants).  Semantics: header of rtb_ext.cu.
// ---------------------------------------------------------------------------
// exact test of one leaf record (sphere or triangle) whose first two float4 are already loaded; `has`/`best` skip
// candidates that cannot win.  A sphere travels as (0, 0, 0, r*r), (centre, id): told apart by the all-zero normal.
__device__ __forceinline__ bool prim_test_pre(const float4* __restrict__ q, float4 q0, float4 q1, V3 o, V3 d, bool has,
                                              float best, float* t_out) {
    if (q0.x == 0.0f && q0.y == 0.0f && q0.z == 0.0f) {
        const V3 oc = vsub(o, mk(q1.x, q1.y, q1.z));
        const float b = vdot(oc, d);
        const float c = __fsub_rn(vdot(oc, oc), q0.w);
        const float disc = __fsub_rn(__fmul_rn(b, b), c);
        if (disc < 0.0f) return false;
        const float sq = __fsqrt_rn(disc);
        float t = __fsub_rn(-b, sq);
        if (t < 0.0f) { t = __fadd_rn(-b, sq); if (t < 0.0f) return false; }
        if (has && t > best) return false;
        *t_out = t;
        return true;
    }
    return tri_test_pre(q, q0, q1, o, d, has, best, t_out);
}
__device__ __forceinline__ bool prim_test(const float4* __restrict__ q, V3 o, V3 d, bool has, float best, float* t_out) {
    return prim_test_pre(q, __ldg(q), __ldg(q + 1), o, d, has, best, t_out);
}

// Geometry of a hit on a triangle or sphere record: the hit point, the normal turned against the ray (Triangle::normal
// :441-449, outward unit(p - c) for a sphere), and whether the hit lies on a wire-frame edge (:419-422).
struct ExtHit { V3 p, nn; bool hit_edge; float4 s0, s1; };
__device__ __forceinline__ ExtHit ext_hit_geometry(const SceneDev& sc, int slot, float t, V3 o, V3 d) {
    ExtHit e;
    const float4* q = sc.tri + (size_t)RTB_TRI_F4 * (uint32_t)slot;
    const float4 q0 = __ldg(q + 0), q1 = __ldg(q + 1);
    e.s0 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)slot);
    e.s1 = __ldg(sc.shade + (size_t)RTB_SHADE_F4 * (uint32_t)slot + 1);
    const bool is_sphere = (__float_as_uint(e.s1.x) & RTB_PRIM_SPHERE) != 0u;
    e.p = vadd(vmul(d, t), o);
    V3 n;
    e.hit_edge = false;
    if (is_sphere) {
        n = vunit(vsub(e.p, mk(q1.x, q1.y, q1.z)));
    } else {
        n = mk(q0.x, q0.y, q0.z);
        const V3 ip = vsub(e.p, mk(q1.x, q1.y, q1.z));
        const float edge_k = __fsub_rn(1.0f, e.s1.z);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float4 qs = __ldg(q + 2 + i);
            if (vdot(ip, mk(qs.x, qs.y, qs.z)) > __fmul_rn(qs.w, edge_k)) e.hit_edge = true;
        }
    }
    const bool back = vdot(d, n) > 0.0f;
    e.nn = back ? vmul(n, -1.0f) : n;
    return e;
}
// LightSource::get_shadow_ray (raytrace.rs:600-610): four RNG draws, a ray from p + n * 0.005 * (rand + 1) towards a random
// point of the light cube (make_ray normalises the direction again)
__device__ __forceinline__ void ext_shadow_ray(const ExtParams& ex, const ExtHit& e, Rng& g, V3* so, V3* sd) {
    const float rx = g.next_f32(), ry = g.next_f32(), rz = g.next_f32();
    const V3 adj = mk(__fadd_rn(ex.light[0], __fmul_rn(rx, ex.light[3])), __fadd_rn(ex.light[1], __fmul_rn(ry, ex.light[3])),
                      __fadd_rn(ex.light[2], __fmul_rn(rz, ex.light[3])));
    const V3 dir = vunit(vsub(adj, e.p));
    const V3 smudge = vmul(e.nn, __fmul_rn(0.005f, __fadd_rn(g.next_f32(), 1.0f)));
    *so = vadd(e.p, smudge);
    *sd = vunit(dir);
}
// color_ray (raytrace.rs:1228-1252) with the shadow block live: black instead of the surface colour where `shadowed`.
// Returns 0 = terminal colour, 1 = bounce: (*color, *alpha) for the mix stack and the next ray (*no, *nd).
__device__ __forceinline__ int ext_shade(const ExtHit& e, V3 d, bool shadowed, Rng& g, V3* color, float* alpha, V3* no, V3* nd) {
    if (e.hit_edge) { *color = mk(0.0f, 0.0f, 0.0f); return 0; }   // getsurface: edges are Solid black (:450-459)
    const uint32_t kind = __float_as_uint(e.s1.x) & 0xffu;
    *color = shadowed ? mk(0.0f, 0.0f, 0.0f) : mk(e.s0.x, e.s0.y, e.s0.z);
    if (kind == RTB_SOLID) return 0;
    *alpha = e.s0.w;
    if (kind == RTB_MATTE) {                                     // lambertian_ray :292-297
        const V3 rv = random_vec(g);
        *no = vadd(e.p, vmul(rv, 0.001f));
        *nd = vunit(vadd(e.nn, rv));
    } else {                                                     // reflect_ray :278-290
        const float ddot = fabsf(vdot(d, e.nn));
        const V3 dir_p = vmul(e.nn, ddot);
        const V3 dir_o = vadd(d, dir_p);
        const V3 reflect = vadd(dir_p, dir_o);
        const V3 rv = vmul(random_vec(g), e.s1.y);
        const V3 rd = vunit(vadd(reflect, rv));
        *no = vadd(e.p, vmul(rd, 0.001f));
        *nd = vunit(rd);
    }
    return 1;
}

// mix_color(c, sub, a) = c*(1-a) + sub*a  (:299-301)
__device__ __forceinline__ V3 mix_color(V3 c, V3 sub, float a) {
    return vadd(vmul(c, __fsub_rn(1.0f, a)), vmul(sub, a));
}

__device__ __forceinline__ V3 sky_color() {   // make_color((128,180,255)) :1264
    return mk(__fdiv_rn(128.0f, 255.0f), __fdiv_rn(180.0f, 255.0f), __fdiv_rn(255.0f, 255.0f));
}

}  // namespace rtbdev
