/*
 * rt_oracle.cpp — CPU oracle for the rust_raytrace hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see rt_oracle.h).  A restatement, function by
 * function, of /root/reference/raytrace_lib/src/raytrace.rs (cited below as
 * `rs:N`) and obj_parser.rs, written so that every f32 operation happens in the
 * reference's order:
 *
 *   - Vec3 is a 4-lane f32 SIMD value whose lane 3 is 0; `dot`/`len2` are a
 *     lane-wise multiply followed by an ORDERED reduce_sum, i.e.
 *     ((p0 + p1) + p2) + p3 with p3 = 0*0 (rs:65-77).  No FMA contraction
 *     (rustc never contracts), IEEE sqrt and divide, unit() = v * (1/len).
 *   - build with -ffp-contract=off and without -ffast-math (oracle/Makefile).
 *
 * PARITY UNPINNED by reference tests except `face_collision` (rs:735-750).
 *
 * Randomness: the reference draws from an OS-seeded ThreadRng (rs:188-192,
 * :1385), which cannot be reproduced.  The oracle substitutes a counter based
 * stream (pcg32 keyed by seed/pixel/sample, 24-bit mantissa floats exactly as
 * rand 0.8's Standard f32) drawn in the reference's call order, so that the
 * CUDA path can be compared with it bit for bit on stochastic materials too.
 */
#include "rt_oracle.h"

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <algorithm>

namespace {

/* ------------------------------------------------------------------ */
/* L0 math — rs:22-122                                                  */
/* ------------------------------------------------------------------ */
struct V3 { float x, y, z; };

static inline V3 mk(float x, float y, float z) { V3 r = {x, y, z}; return r; }
static inline V3 vadd(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }            /* rs:37 */
static inline V3 vsub(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }            /* rs:44 */
static inline V3 vmul(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }               /* rs:51 */
static inline V3 vmulper(V3 a, V3 b) { return mk(a.x * b.x, a.y * b.y, a.z * b.z); }         /* rs:58 */
/* ordered reduce over 4 lanes, lane 3 product is +0 (rs:66, rs:76) */
static inline float vdot(V3 a, V3 b) {
    float p0 = a.x * b.x, p1 = a.y * b.y, p2 = a.z * b.z;
    float s = p0 + p1;
    s = s + p2;
    s = s + 0.0f;
    return s;
}
static inline float vlen2(V3 a) { return vdot(a, a); }                                        /* rs:65 */
static inline float vlen(V3 a) { return sqrtf(vlen2(a)); }                                    /* rs:70 */
static inline V3 vcross(V3 a, V3 b) {                                                         /* rs:80-90 */
    return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline V3 vunit(V3 a) { float inv = 1.0f / vlen(a); return vmul(a, inv); }             /* rs:93-96 */

static V3 vorthogonal(V3 a, int guard = 0) {                                                  /* rs:98-108 */
    if (fabsf(a.x) > 0.1f) return vunit(mk(-1.0f * (a.y + a.z) / a.x, 1.0f, 1.0f));
    if (fabsf(a.y) > 0.1f) return vunit(mk(1.0f, -1.0f * (a.x + a.z) / a.y, 1.0f));
    if (fabsf(a.z) > 0.1f) return vunit(mk(1.0f, 1.0f, -1.0f * (a.x + a.y) / a.z));
    if (guard > 4) return mk(1.0f, 0.0f, 0.0f); /* the reference would recurse forever on 0 */
    return vorthogonal(vunit(a), guard + 1);
}
struct Basis { V3 b0, b1, b2; };
static inline V3 change_basis(V3 v, const Basis& b) {                                         /* rs:117-121 */
    return mk(vdot(b.b0, v), vdot(b.b1, v), vdot(b.b2, v));
}
static inline V3 make_color(uint8_t r, uint8_t g, uint8_t b) {                                /* rs:176-180 */
    return mk((float)r / 255.0f, (float)g / 255.0f, (float)b / 255.0f);
}

/* ------------------------------------------------------------------ */
/* RNG substitute (see header comment); rand 0.8 Standard f32 mapping   */
/* ------------------------------------------------------------------ */
static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
struct Rng {
    uint64_t state;
    void seed(uint64_t seed, uint64_t pixel, uint32_t sample) {
        uint64_t s = splitmix64(seed);
        s = splitmix64(s ^ pixel);
        s = splitmix64(s ^ (uint64_t)sample);
        state = s;
    }
    uint32_t next_u32() {
        uint64_t old = state;
        state = old * 6364136223846793005ull + 1442695040888963407ull;
        uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
        uint32_t rot = (uint32_t)(old >> 59);
        return (xorshifted >> rot) | (xorshifted << ((32u - rot) & 31u));
    }
    float next_f32() { return (float)(next_u32() >> 8) * (1.0f / 16777216.0f); }
};

/* ------------------------------------------------------------------ */
/* L1 rays — rs:194-301                                                 */
/* ------------------------------------------------------------------ */
struct Ray { V3 orig, dir, inv_dir; };

static inline Ray make_ray(V3 orig, V3 dir) {                                                 /* rs:201-210 */
    Ray r;
    V3 du = vunit(dir);
    r.orig = orig;
    r.dir = du;
    r.inv_dir = mk(1.0f / du.x, 1.0f / du.y, 1.0f / du.z);
    return r;
}
static inline V3 ray_at(const Ray& r, float t) { return vadd(vmul(r.dir, t), r.orig); }      /* rs:227-229 */

static bool ray_intersect_helper(V3 a, V3 v, V3 b, V3 u, float* t1, float* t2) {              /* rs:212-224 */
    float det = u.x * v.y - u.y * v.x;
    if (fabsf(det) < 0.0001f) return false;
    float dx = b.x - a.x;
    float dy = b.y - a.y;
    *t1 = (dy * u.x - dx * u.y) / det;
    *t2 = (dy * v.x - dx * v.y) / det;
    return true;
}
static bool ray_intersect(const Ray& s, const Ray& r, V3* out) {                              /* rs:231-267 */
    float t1, t2;
    if (!ray_intersect_helper(s.orig, s.dir, r.orig, r.dir, &t1, &t2)) {
        if (!ray_intersect_helper(mk(s.orig.x, s.orig.z, s.orig.y), mk(s.dir.x, s.dir.z, s.dir.y),
                                  mk(r.orig.x, r.orig.z, r.orig.y), mk(r.dir.x, r.dir.z, r.dir.y), &t1, &t2)) {
            if (!ray_intersect_helper(mk(s.orig.y, s.orig.z, s.orig.x), mk(s.dir.y, s.dir.z, s.dir.x),
                                      mk(r.orig.y, r.orig.z, r.orig.x), mk(r.dir.y, r.dir.z, r.dir.x), &t1, &t2))
                return false;
        }
    }
    V3 p1 = ray_at(s, t1);
    V3 p2 = ray_at(r, t2);
    V3 d = vsub(p2, p1);
    if (vlen2(d) < 0.01f) { *out = p1; return true; }
    return false;
}

static inline V3 random_vec(Rng& g) {                                                         /* rs:188-192 */
    float a = g.next_f32() - 0.5f;
    float b = g.next_f32() - 0.5f;
    float c = g.next_f32() - 0.5f;
    return vunit(mk(a, b, c));
}
static Ray reflect_ray(V3 orig, V3 norm, V3 dir, float fuzz, Rng& g) {                        /* rs:278-290 */
    float ddot = fabsf(vdot(dir, norm));
    V3 dir_p = vmul(norm, ddot);
    V3 dir_o = vadd(dir, dir_p);
    V3 reflect = vadd(dir_p, dir_o);
    V3 rand_vec = vmul(random_vec(g), fuzz);
    V3 reflect_dir = vunit(vadd(reflect, rand_vec));
    return make_ray(vadd(orig, vmul(reflect_dir, 0.001f)), vunit(vadd(reflect, rand_vec)));
}
static Ray lambertian_ray(V3 orig, V3 norm, Rng& g) {                                         /* rs:292-297 */
    V3 rand_vec = random_vec(g);
    return make_ray(vadd(orig, vmul(rand_vec, 0.001f)), vadd(norm, rand_vec));
}
static inline V3 mix_color(V3 c1, V3 c2, float a) {                                           /* rs:299-301 */
    return vadd(vmul(c1, 1.0f - a), vmul(c2, a));
}

/* ------------------------------------------------------------------ */
/* Triangle — rs:326-461                                                */
/* ------------------------------------------------------------------ */
struct Tri {
    V3 incenter, norm;
    float bounding_r2;
    V3 sides[3];
    float side_lens[3];
    V3 corners[3];
    uint32_t kind;
    V3 color;
    float alpha, scattering;
    float edge_thickness;
};

enum Face { F_NONE = 0, F_FRONT = 1, F_BACK = 2, F_EDGEFRONT = 3, F_EDGEBACK = 4 };

static bool make_triangle(const V3 pts[3], uint32_t kind, V3 color, float alpha, float scattering,
                          float edge_thickness, Tri* out) {                                    /* rs:340-383 */
    V3 a = pts[0], b = pts[1], c = pts[2];
    V3 ab = vsub(b, a), ac = vsub(c, a), bc = vsub(c, b);
    V3 bac_bisect = vadd(ac, ab);
    V3 abc_bisect = vadd(bc, vmul(ab, -1.0f));
    Ray bac_bi_ray = make_ray(a, bac_bisect);
    Ray abc_bi_ray = make_ray(b, abc_bisect);
    V3 incenter;
    if (!ray_intersect(bac_bi_ray, abc_bi_ray, &incenter)) return false; /* reference: unwrap() panic */

    Tri t;
    for (int idx = 0; idx < 3; idx++) {
        V3 vedge = vsub(pts[(idx + 1) % 3], pts[idx]);
        V3 po = vsub(incenter, pts[idx]);
        V3 pc = vmul(vedge, vdot(vedge, po) / vlen2(vedge));
        V3 oc = vsub(pc, po);
        t.sides[idx] = vunit(oc);
        t.side_lens[idx] = vlen(oc);
    }
    t.norm = vunit(vcross(t.sides[0], t.sides[1]));
    t.incenter = incenter;
    float r2 = 0.0f;
    for (int i = 0; i < 3; i++) r2 = fmaxf(r2, vlen2(vsub(pts[i], incenter)));               /* rs:375 */
    t.bounding_r2 = r2;
    for (int i = 0; i < 3; i++) t.corners[i] = pts[i];
    t.kind = kind; t.color = color; t.alpha = alpha; t.scattering = scattering;
    t.edge_thickness = edge_thickness;
    *out = t;
    return true;
}

/* Triangle::intersects rs:400-439.  NaN semantics follow from the literal comparisons. */
static inline int tri_intersects(const Tri& tr, const Ray& r, float* t_out, V3* p_out) {
    float t = vdot(tr.norm, vsub(tr.incenter, r.orig)) / vdot(tr.norm, r.dir);
    if (t < 0.0f) return F_NONE;
    V3 p = ray_at(r, t);
    V3 ip = vsub(p, tr.incenter);
    if (vlen2(ip) > tr.bounding_r2) return F_NONE;
    bool hit_edge = false;
    for (int i = 0; i < 3; i++) {
        float dist = vdot(ip, tr.sides[i]);
        float side_len = tr.side_lens[i];
        if (dist > side_len) return F_NONE;
        else if (dist > (side_len * (1.0f - tr.edge_thickness))) hit_edge = true;
    }
    int face;
    if (hit_edge) face = (vdot(r.dir, tr.norm) > 0.0f) ? F_EDGEBACK : F_EDGEFRONT;
    else          face = (vdot(r.dir, tr.norm) > 0.0f) ? F_BACK : F_FRONT;
    *t_out = t; *p_out = p;
    return face;
}
static inline V3 tri_normal(const Tri& tr, int face) {                                        /* rs:441-449 */
    if (face == F_FRONT || face == F_EDGEFRONT) return tr.norm;
    return vmul(tr.norm, -1.0f);
}

/* ------------------------------------------------------------------ */
/* EXTENSION: analytic sphere.  The mounted      */
/* reference has no sphere primitive any more (only the      */
/* vestiges CollisionFace::{Side,Face}, rs:311-318, and circles_2k.png  */
/* of an older revision), so this definition is the build's own; it     */
/* follows Triangle::intersects' conventions (t < 0 rejects, literal    */
/* comparisons, normal flipped for the back face).                      */
/* ------------------------------------------------------------------ */
struct Sph { V3 c; float r; uint32_t kind; V3 color; float alpha, scattering; };

static inline V3 sph_normal_out(const Sph& s, V3 p) { return vunit(vsub(p, s.c)); }
static inline int sph_intersects(const Sph& s, const Ray& r, float* t_out, V3* p_out) {
    V3 oc = vsub(r.orig, s.c);
    float b = vdot(oc, r.dir);                       /* |dir| = 1: t^2 + 2bt + c = 0 */
    float c = vlen2(oc) - s.r * s.r;
    float disc = b * b - c;
    if (disc < 0.0f) return F_NONE;
    float sq = sqrtf(disc);
    float t = (-b) - sq;                             /* near root, else the far one (origin inside) */
    if (t < 0.0f) { t = (-b) + sq; if (t < 0.0f) return F_NONE; }
    V3 p = ray_at(r, t);
    *t_out = t; *p_out = p;
    return (vdot(r.dir, sph_normal_out(s, p)) > 0.0f) ? F_BACK : F_FRONT;
}

/* ------------------------------------------------------------------ */
/* scene generators — rs:464-592                                        */
/* ------------------------------------------------------------------ */
static const float PI_F = 3.14159265358979323846f;
static const float FRAC_PI_2_F = 1.57079632679489661923f;

struct Surf { uint32_t kind; V3 color; float alpha, scattering; };

static int make_sphere(V3 orig, float r, uint32_t num_lat, uint32_t num_lon, Surf s, float edge,
                       std::vector<Tri>& tris) {                                               /* rs:464-529 */
    if (num_lat % 2 != 0) return -1;
    int n = 0;
    for (uint32_t lat_idx = 0; lat_idx < num_lat; lat_idx++) {
        for (uint32_t lon_idx = 0; lon_idx < num_lon; lon_idx++) {
            float phi1 = (((lat_idx % 2 == 0) ? (float)lat_idx / (float)num_lat * PI_F
                                              : (float)(lat_idx + 1) / (float)num_lat * PI_F) - FRAC_PI_2_F) * -1.0f;
            float phi23 = (((lat_idx % 2 == 0) ? (float)(lat_idx + 1) / (float)num_lat * PI_F
                                               : (float)lat_idx / (float)num_lat * PI_F) - FRAC_PI_2_F) * -1.0f;
            float smudge = (lat_idx % 2 == 0) ? 0.0f : 0.5f;
            float theta1 = ((float)lon_idx + smudge) / (float)num_lon * 2.0f * PI_F;
            float theta2 = ((float)lon_idx + 0.5f + smudge) / (float)num_lon * 2.0f * PI_F;
            float theta3 = ((float)lon_idx - 0.5f + smudge) / (float)num_lon * 2.0f * PI_F;
            float theta4 = ((float)lon_idx + 1.0f + smudge) / (float)num_lon * 2.0f * PI_F;

            float phi14sin = sinf(phi1), phi14cos = cosf(phi1);
            V3 p1 = vadd(orig, mk(r * phi14sin, r * phi14cos * cosf(theta1), r * phi14cos * sinf(theta1)));
            V3 p4 = vadd(orig, mk(r * phi14sin, r * phi14cos * cosf(theta4), r * phi14cos * sinf(theta4)));
            float phi23sin = sinf(phi23), phi23cos = cosf(phi23);
            V3 p2 = vadd(orig, mk(r * phi23sin, r * phi23cos * cosf(theta2), r * phi23cos * sinf(theta2)));
            V3 p3 = vadd(orig, mk(r * phi23sin, r * phi23cos * cosf(theta3), r * phi23cos * sinf(theta3)));

            Tri t;
            V3 a[3] = {p1, p2, p3};
            if (!make_triangle(a, s.kind, s.color, s.alpha, s.scattering, edge, &t)) return -1;
            tris.push_back(t); n++;
            if (lat_idx != 0 && lat_idx != (num_lat - 1)) {
                V3 b[3] = {p1, p2, p4};
                if (!make_triangle(b, s.kind, s.color, s.alpha, s.scattering, edge, &t)) return -1;
                tris.push_back(t); n++;
            }
        }
    }
    return n;
}

static int make_disk(V3 orig, V3 norm, float r, float d, uint32_t num_tris, Surf s, Surf side, float edge,
                     std::vector<Tri>& tris) {                                                 /* rs:531-592 */
    V3 norm_orth0 = vmul(vunit(vorthogonal(norm)), r);
    V3 norm_orth1 = vmul(vunit(vcross(norm, norm_orth0)), r);
    const float smudge = 0.0f;
    int n = 0;
    for (uint32_t idx = 0; idx < num_tris; idx++) {
        V3 norm_pd = vmul(norm, d);
        V3 norm_md = vmul(norm, -1.0f * d);
        float theta1 = (float)idx / (float)num_tris * 2.0f * PI_F - smudge;
        float theta2 = ((float)idx + 1.0f) / (float)num_tris * 2.0f * PI_F + smudge;
        float theta3 = ((float)idx + 0.5f) / (float)num_tris * 2.0f * PI_F - smudge;
        float theta4 = ((float)idx + 1.5f) / (float)num_tris * 2.0f * PI_F + smudge;

        V3 p1p = vadd(orig, norm_pd);
        V3 p2p = vadd(vadd(vadd(orig, norm_pd), vmul(norm_orth0, sinf(theta1))), vmul(norm_orth1, cosf(theta1)));
        V3 p3p = vadd(vadd(vadd(orig, norm_pd), vmul(norm_orth0, sinf(theta2))), vmul(norm_orth1, cosf(theta2)));
        V3 p1m = vadd(orig, norm_md);
        V3 p2m = vadd(vadd(vadd(orig, norm_md), vmul(norm_orth0, sinf(theta3))), vmul(norm_orth1, cosf(theta3)));
        V3 p3m = vadd(vadd(vadd(orig, norm_md), vmul(norm_orth0, sinf(theta4))), vmul(norm_orth1, cosf(theta4)));

        Tri t;
        V3 top[3] = {p1p, p2p, p3p};
        if (!make_triangle(top, s.kind, s.color, s.alpha, s.scattering, edge, &t)) return -1;
        tris.push_back(t);
        V3 bot[3] = {p1m, p2m, p3m};
        if (!make_triangle(bot, s.kind, s.color, s.alpha, s.scattering, edge, &t)) return -1;
        tris.push_back(t);
        V3 s1[3] = {p2p, p3p, p2m};
        if (!make_triangle(s1, side.kind, side.color, side.alpha, side.scattering, edge, &t)) return -1;
        tris.push_back(t);
        V3 s2[3] = {p2m, p3m, p3p};
        if (!make_triangle(s2, side.kind, side.color, side.alpha, side.scattering, edge, &t)) return -1;
        tris.push_back(t);
        n += 4;
    }
    return n;
}

/* ------------------------------------------------------------------ */
/* camera — rs:1305-1394                                                */
/* ------------------------------------------------------------------ */
static inline float to_radians(float deg) { const float k = PI_F / 180.0f; return deg * k; }

static Basis create_transform(V3 dir_in, float d_roll) {                                      /* rs:1320-1341 */
    V3 dir = vunit(dir_in);
    float roll = -1.0f * atan2f(-1.0f * dir.y, dir.z);
    float pitch = -1.0f * asinf(dir.x);
    float yaw = -1.0f * d_roll;
    Basis b;
    b.b0 = mk(cosf(yaw) * cosf(pitch), sinf(yaw) * cosf(pitch), -1.0f * sinf(pitch));
    b.b1 = mk(cosf(yaw) * sinf(pitch) * sinf(roll) - sinf(yaw) * cosf(roll),
              sinf(yaw) * sinf(pitch) * sinf(roll) + cosf(yaw) * cosf(roll),
              cosf(pitch) * sinf(roll));
    b.b2 = mk(cosf(yaw) * sinf(pitch) * cosf(roll) + sinf(yaw) * sinf(roll),
              sinf(yaw) * sinf(pitch) * cosf(roll) - cosf(yaw) * sinf(roll),
              cosf(pitch) * cosf(roll));
    return b;
}

struct View { uint32_t width, height; V3 orig, cam, vu, vv; uint32_t maxdepth, spp; };

static View create_viewport(uint32_t pw, uint32_t ph, float size0, float size1, V3 pos, V3 dir, float fov,
                            float c_roll, uint32_t maxdepth, uint32_t samples) {               /* rs:1343-1370 */
    float dist = size0 / (2.0f * tanf(to_radians(fov) / 2.0f));
    Basis rot = create_transform(dir, c_roll);
    V3 orig = vadd(pos, mk(1.0f * size1 / 2.0f, -1.0f * size0 / 2.0f, 0.0f));
    V3 cam_r = change_basis(mk(0.0f, 0.0f, dist), rot);
    V3 cam = vsub(pos, cam_r);
    V3 vu_r = change_basis(mk(0.0f, size0, 0.0f), rot);
    V3 vv_r = change_basis(mk(-1.0f * size1, 0.0f, 0.0f), rot);
    View v;
    v.width = pw; v.height = ph; v.orig = orig; v.cam = cam; v.vu = vu_r; v.vv = vv_r;
    v.maxdepth = maxdepth; v.spp = samples;
    return v;
}

/* pixel_ray rs:1374-1394; px = (row, col). */
static inline Ray pixel_ray(const View& v, uint32_t row, uint32_t col, Rng* g) {
    float px_x = (float)row;
    float px_y = (float)col;
    V3 vu_delta = vmul(v.vu, 1.0f / (float)v.width);
    V3 vv_delta = vmul(v.vv, 1.0f / (float)v.height);
    float u_off = 0.5f, v_off = 0.5f;
    if (v.spp != 1) { u_off = g->next_f32(); v_off = g->next_f32(); }
    V3 vu_frac = vmul(vu_delta, px_y + u_off);
    V3 vv_frac = vmul(vv_delta, px_x + v_off);
    V3 px_u = vadd(vadd(v.orig, vu_frac), vv_frac);
    return make_ray(px_u, vunit(vsub(px_u, v.cam)));
}

/* ------------------------------------------------------------------ */
/* octree build — rs:636-856                                            */
/* ------------------------------------------------------------------ */
static inline bool box_contains_point(V3 orig, float len2, V3 p) {                            /* rs:636-643 */
    V3 op = vsub(p, orig);
    return fabsf(op.x) < len2 && fabsf(op.y) < len2 && fabsf(op.z) < len2;
}

static inline float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

static bool face_contains_triangle(V3 p, V3 norm, float len2, const Tri& t) {                 /* rs:645-729 */
    float h1 = vdot(norm, vadd(p, vmul(norm, len2)));
    float h2 = vdot(t.norm, t.incenter);
    V3 n1 = norm, n2 = t.norm;
    float c1 = (h1 - h2 * (vdot(n1, n2))) / (1.0f - (vdot(n1, n2)) * (vdot(n1, n2)));
    float c2 = (h2 - h1 * (vdot(n1, n2))) / (1.0f - (vdot(n1, n2)) * (vdot(n1, n2)));
    Ray line_tmp = make_ray(vadd(vmul(n1, c1), vmul(n2, c2)), vcross(n1, n2));

    const float FMAX = std::numeric_limits<float>::max();
    float tmin = FMAX;
    for (int ax = 0; ax < 3; ax++) {
        if (comp(norm, ax) == 0.0f) {
            float t1 = (comp(p, ax) - len2 - comp(line_tmp.orig, ax)) * comp(line_tmp.inv_dir, ax);
            float t2 = (comp(p, ax) + len2 - comp(line_tmp.orig, ax)) * comp(line_tmp.inv_dir, ax);
            tmin = fminf(tmin, fminf(t1, t2));
        }
    }
    Ray line = (tmin > 0.0f) ? line_tmp : make_ray(ray_at(line_tmp, tmin * 2.0f), line_tmp.dir);

    tmin = -FMAX;
    float tmax = FMAX;
    for (int ax = 0; ax < 3; ax++) {
        if (comp(norm, ax) == 0.0f) {
            float t1 = (comp(p, ax) - len2 - comp(line.orig, ax)) * comp(line.inv_dir, ax);
            float t2 = (comp(p, ax) + len2 - comp(line.orig, ax)) * comp(line.inv_dir, ax);
            tmin = fmaxf(tmin, fminf(t1, t2));
            tmax = fminf(tmax, fmaxf(t1, t2));
        }
    }
    if (tmax < tmin) return false;

    float t1 = vdot(vsub(t.corners[0], line.orig), line.dir) / vlen2(line.dir);
    float t2 = vdot(vsub(t.corners[1], line.orig), line.dir) / vlen2(line.dir);
    float t3 = vdot(vsub(t.corners[2], line.orig), line.dir) / vlen2(line.dir);
    V3 p1 = ray_at(line, t1), p2 = ray_at(line, t2), p3 = ray_at(line, t3);
    return vdot(vsub(p1, t.corners[0]), vsub(p2, t.corners[1])) < 0.0f ||
           vdot(vsub(p1, t.corners[0]), vsub(p3, t.corners[2])) < 0.0f ||
           vdot(vsub(p2, t.corners[1]), vsub(p3, t.corners[2])) < 0.0f;
}

static bool box_contains_polygon(V3 orig, float len2, const Tri& t) {                          /* rs:753-779 */
    if (box_contains_point(orig, len2, t.incenter)) return true;
    for (int i = 0; i < 3; i++)
        if (box_contains_point(orig, len2, t.corners[i])) return true;
    const V3 face_norms[6] = {mk(1, 0, 0), mk(-1, 0, 0), mk(0, 1, 0), mk(0, -1, 0), mk(0, 0, 1), mk(0, 0, -1)};
    for (int i = 0; i < 6; i++)
        if (face_contains_triangle(orig, face_norms[i], len2, t)) return true;
    return false;
}

struct BBox {                                                                                 /* rs:618-623 */
    V3 orig;
    float len2;
    uint32_t depth;
    bool is_leaf;
    std::vector<BBox*> boxes;     /* BBSubobj::Boxes, octant order with empty octants removed */
    std::vector<uint32_t> tris;   /* BBSubobj::Tris */
    ~BBox() { for (BBox* b : boxes) delete b; }
};

static BBox* build_bb_helper(const std::vector<Tri>& tris, const std::vector<uint32_t>& objs, V3 orig,
                             float len2, uint32_t depth, uint32_t maxdepth, uint32_t minobjs,
                             int par_levels) {                                                 /* rs:795-845 */
    std::vector<uint32_t> subobjs;
    for (uint32_t idx : objs)
        if (box_contains_polygon(orig, len2, tris[idx])) subobjs.push_back(idx);

    if (subobjs.empty()) return nullptr;
    if (subobjs.size() < minobjs || depth >= maxdepth) {
        BBox* b = new BBox();
        b->orig = orig; b->len2 = len2; b->depth = depth; b->is_leaf = true;
        b->tris.swap(subobjs);
        return b;
    }
    float newlen2 = len2 / 2.0f;
    BBox* kids[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    auto build_child = [&](int i) {
        float xoff = ((i & 1) == 0) ? -1.0f * newlen2 : newlen2;
        float yoff = ((i & 2) == 0) ? -1.0f * newlen2 : newlen2;
        float zoff = ((i & 4) == 0) ? -1.0f * newlen2 : newlen2;
        kids[i] = build_bb_helper(tris, subobjs, vadd(orig, mk(xoff, yoff, zoff)), newlen2, depth + 1,
                                  maxdepth, minobjs, par_levels - 1);
    };
    if (par_levels > 0) {   /* host-side parallelism only; the tree is identical */
        std::thread th[8];
        for (int i = 0; i < 8; i++) th[i] = std::thread(build_child, i);
        for (int i = 0; i < 8; i++) th[i].join();
    } else {
        for (int i = 0; i < 8; i++) build_child(i);
    }
    BBox* b = new BBox();
    b->orig = orig; b->len2 = len2; b->depth = depth; b->is_leaf = false;
    for (int i = 0; i < 8; i++) if (kids[i]) b->boxes.push_back(kids[i]);
    if (b->boxes.empty()) { delete b; return nullptr; }
    return b;
}

/* ------------------------------------------------------------------ */
/* octree traversal — rs:858-1050                                       */
/* ------------------------------------------------------------------ */
struct Counters { uint64_t rays = 0, box_tests = 0, tri_tests = 0, node_visits = 0, nan_t = 0; };

static inline bool bb_collides(const BBox& b, const Ray& r, float* otmin, float* otmax) {     /* rs:861-907 */
    float tmin = -std::numeric_limits<float>::max();
    float tmax = std::numeric_limits<float>::max();
    V3 tmp1 = vmulper(vsub(b.orig, r.orig), r.inv_dir);
    V3 tmp2 = vmul(r.inv_dir, b.len2);
    V3 t1s = vsub(tmp1, tmp2);
    V3 t2s = vadd(tmp1, tmp2);
    if (r.dir.x != 0.0f) {
        if (r.inv_dir.x > 0.0f) { tmin = t1s.x; tmax = t2s.x; }
        else                    { tmin = t2s.x; tmax = t1s.x; }
    }
    if (r.dir.y != 0.0f) {
        if (r.inv_dir.y > 0.0f) { tmin = fmaxf(tmin, t1s.y); tmax = fminf(tmax, t2s.y); }
        else                    { tmin = fmaxf(tmin, t2s.y); tmax = fminf(tmax, t1s.y); }
    }
    if (r.dir.z != 0.0f) {
        if (r.inv_dir.z > 0.0f) { tmin = fmaxf(tmin, t1s.z); tmax = fminf(tmax, t2s.z); }
        else                    { tmin = fmaxf(tmin, t2s.z); tmax = fminf(tmax, t1s.z); }
    }
    if (tmin < tmax) { *otmin = tmin; *otmax = tmax; return true; }
    return false;
}

struct Hit { bool some; float t; V3 p; int face; uint32_t idx; };

static Hit leaf_min_time(const std::vector<uint32_t>& objtris, const std::vector<Tri>& tris, const Ray& r,
                         Counters& c) {                                                        /* rs:1013-1050 */
    Hit acc; acc.some = false; acc.t = 0; acc.p = mk(0, 0, 0); acc.face = 0; acc.idx = 0;
    for (uint32_t tnum : objtris) {
        float t; V3 p;
        c.tri_tests++;
        int face = tri_intersects(tris[tnum], r, &t, &p);
        if (face != F_NONE) {
            if (acc.some) {
                if (t < acc.t) { acc.t = t; acc.p = p; acc.face = face; acc.idx = tnum; }
            } else {
                acc.some = true; acc.t = t; acc.p = p; acc.face = face; acc.idx = tnum;
            }
        }
    }
    return acc;
}

static Hit bb_intersect(const BBox& b, const std::vector<Tri>& tris, const Ray& r, Counters& c) { /* rs:910-1010 */
    c.node_visits++;
    if (b.is_leaf) return leaf_min_time(b.tris, tris, r, c);

    const float FMAX = std::numeric_limits<float>::max();
    struct Ent { float tmin, tmax; const BBox* bb; };
    Ent boxmap[8];
    for (int i = 0; i < 8; i++) { boxmap[i].tmin = FMAX; boxmap[i].tmax = FMAX; boxmap[i].bb = nullptr; }
    for (size_t i = 0; i < b.boxes.size(); i++) {
        float tmin, tmax;
        c.box_tests++;
        if (bb_collides(*b.boxes[i], r, &tmin, &tmax)) { boxmap[i].tmin = tmin; boxmap[i].tmax = tmax; boxmap[i].bb = b.boxes[i]; }
    }
    for (int idx = 1; idx < 8; idx++) {                                                       /* rs:941-947 */
        int jdx = idx;
        while (jdx > 0 && boxmap[jdx - 1].tmin > boxmap[jdx].tmin) { std::swap(boxmap[jdx - 1], boxmap[jdx]); jdx--; }
    }
    float bboxtmax = 0.0f;
    Hit acc; acc.some = false; acc.t = 0; acc.p = mk(0, 0, 0); acc.face = 0; acc.idx = 0;
    for (int i = 0; i < 8; i++) {                                                             /* rs:949-1007 */
        const Ent& e = boxmap[i];
        if (!e.bb) continue;
        if (acc.some) {
            if (e.tmin < bboxtmax) {
                Hit sub = bb_intersect(*e.bb, tris, r, c);
                if (sub.some && sub.t < acc.t) { bboxtmax = sub.t; acc = sub; }
            }
        } else {
            if (e.tmin != FMAX) {
                Hit sub = bb_intersect(*e.bb, tris, r, c);
                if (sub.some) { bboxtmax = sub.t; acc = sub; }
                else { bboxtmax = 0.0f; }
            }
        }
    }
    return acc;
}

/* ------------------------------------------------------------------ */
/* oracle-only BVH (OR_ACCEL_BVH): exact closest hit, lowest index on   */
/* exact-t ties, for scenes where the octree build is impractical.      */
/* Culling is done in double with padded boxes so it can only ever be   */
/* conservative w.r.t. the f32 acceptance region of tri_intersects.     */
/* ------------------------------------------------------------------ */
struct BNode { double lo[3], hi[3]; int32_t left, right; uint32_t first, count; };

struct OBvh {
    std::vector<BNode> nodes;
    std::vector<uint32_t> order;

    void bounds(const std::vector<Tri>& tris, uint32_t a, uint32_t b, double lo[3], double hi[3], double pad) {
        for (int k = 0; k < 3; k++) { lo[k] = 1e300; hi[k] = -1e300; }
        for (uint32_t i = a; i < b; i++) {
            const Tri& t = tris[order[i]];
            for (int c = 0; c < 3; c++) {
                double v[3] = {t.corners[c].x, t.corners[c].y, t.corners[c].z};
                for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], v[k] - pad); hi[k] = std::max(hi[k], v[k] + pad); }
            }
        }
    }
    int32_t build(const std::vector<Tri>& tris, uint32_t a, uint32_t b, double pad) {
        int32_t id = (int32_t)nodes.size();
        nodes.push_back(BNode());
        BNode n;
        bounds(tris, a, b, n.lo, n.hi, pad);
        n.left = n.right = -1; n.first = a; n.count = b - a;
        if (b - a > 4) {
            int ax = 0; double ext = -1;
            double clo[3] = {1e300, 1e300, 1e300}, chi[3] = {-1e300, -1e300, -1e300};
            for (uint32_t i = a; i < b; i++) {
                const Tri& t = tris[order[i]];
                double cx[3] = {t.incenter.x, t.incenter.y, t.incenter.z};
                for (int k = 0; k < 3; k++) { clo[k] = std::min(clo[k], cx[k]); chi[k] = std::max(chi[k], cx[k]); }
            }
            for (int k = 0; k < 3; k++) if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; ax = k; }
            uint32_t mid = (a + b) / 2;
            std::nth_element(order.begin() + a, order.begin() + mid, order.begin() + b,
                             [&](uint32_t p, uint32_t q) {
                                 float cp = comp(tris[p].incenter, ax), cq = comp(tris[q].incenter, ax);
                                 return cp < cq || (cp == cq && p < q);
                             });
            n.count = 0;
            int32_t l = build(tris, a, mid, pad);
            int32_t r = build(tris, mid, b, pad);
            n.left = l; n.right = r;
        }
        nodes[id] = n;
        return id;
    }
    void init(const std::vector<Tri>& tris) {
        order.clear();
        double m = 0;
        for (uint32_t i = 1; i < tris.size(); i++) {
            order.push_back(i);
            for (int c = 0; c < 3; c++)
                m = std::max(m, (double)std::max(fabsf(tris[i].corners[c].x), std::max(fabsf(tris[i].corners[c].y), fabsf(tris[i].corners[c].z))));
        }
        nodes.clear();
        if (!order.empty()) build(tris, 0, (uint32_t)order.size(), 1e-3 * std::max(1.0, m / 16.0));
    }
    /* does any triangle other than `exclude` intersect the ray (any t >= 0)?  (shadow rays) */
    bool any_hit(const std::vector<Tri>& tris, const Ray& r, uint32_t exclude) const {
        if (nodes.empty()) return false;
        int32_t stack[128]; int sp = 0; stack[sp++] = 0;
        double o[3] = {r.orig.x, r.orig.y, r.orig.z}, d[3] = {r.dir.x, r.dir.y, r.dir.z};
        while (sp) {
            const BNode& n = nodes[stack[--sp]];
            double t0 = 0.0, t1 = 1e300;
            bool miss = false;
            for (int k = 0; k < 3 && !miss; k++) {
                if (d[k] == 0.0) { if (o[k] < n.lo[k] || o[k] > n.hi[k]) miss = true; continue; }
                double a = (n.lo[k] - o[k]) / d[k], b = (n.hi[k] - o[k]) / d[k];
                if (a > b) std::swap(a, b);
                t0 = std::max(t0, a - 1e-6); t1 = std::min(t1, b + 1e-6);
                if (t0 > t1) miss = true;
            }
            if (miss) continue;
            if (n.left < 0) {
                for (uint32_t i = n.first; i < n.first + n.count; i++) {
                    uint32_t tnum = order[i];
                    float t; V3 p;
                    if (tnum != exclude && tri_intersects(tris[tnum], r, &t, &p) != F_NONE) return true;
                }
            } else if (sp + 2 <= 128) { stack[sp++] = n.left; stack[sp++] = n.right; }
        }
        return false;
    }
    Hit intersect(const std::vector<Tri>& tris, const Ray& r, Counters& c) const {
        Hit acc; acc.some = false; acc.t = 0; acc.p = mk(0, 0, 0); acc.face = 0; acc.idx = 0;
        if (nodes.empty()) return acc;
        int32_t stack[128]; int sp = 0; stack[sp++] = 0;
        double o[3] = {r.orig.x, r.orig.y, r.orig.z}, d[3] = {r.dir.x, r.dir.y, r.dir.z};
        while (sp) {
            const BNode& n = nodes[stack[--sp]];
            c.box_tests++;
            double t0 = 0.0, t1 = acc.some ? (double)acc.t * (1.0 + 1e-6) + 1e-6 : 1e300;
            bool miss = false;
            for (int k = 0; k < 3 && !miss; k++) {
                if (d[k] == 0.0) { if (o[k] < n.lo[k] || o[k] > n.hi[k]) miss = true; continue; }
                double a = (n.lo[k] - o[k]) / d[k], b = (n.hi[k] - o[k]) / d[k];
                if (a > b) std::swap(a, b);
                t0 = std::max(t0, a - 1e-6); t1 = std::min(t1, b + 1e-6);
                if (t0 > t1) miss = true;
            }
            if (miss) continue;
            c.node_visits++;
            if (n.left < 0) {
                for (uint32_t i = n.first; i < n.first + n.count; i++) {
                    uint32_t tnum = order[i];
                    float t; V3 p;
                    c.tri_tests++;
                    int face = tri_intersects(tris[tnum], r, &t, &p);
                    if (face != F_NONE) {
                        if (!acc.some || t < acc.t || (t == acc.t && tnum < acc.idx)) {
                            acc.some = true; acc.t = t; acc.p = p; acc.face = face; acc.idx = tnum;
                        }
                    }
                }
            } else {
                if (sp + 2 <= 128) { stack[sp++] = n.left; stack[sp++] = n.right; }
            }
        }
        return acc;
    }
};

} // namespace

/* ------------------------------------------------------------------ */
/* Scene + integrator — rs:1199-1303, 1396-1440                         */
/* ------------------------------------------------------------------ */
struct OrScene {
    std::vector<Tri> tris;
    std::vector<Sph> spheres;     /* EXTENSION: primitive ids tris.size() + j */
    bool has_light = false;       /* Scene.lights: Option<LightSource> (rs:594-597, use commented out at rs:1204-1224) */
    V3 light_orig = {0, 0, 0};
    float light_len2 = 0;
    int accel;
    BBox* root;       /* OCTREE / TRIVIAL */
    OBvh bvh;         /* BVH */
    ~OrScene() { delete root; }
};

namespace {

static inline Hit scene_hit(const OrScene& s, const Ray& r, Counters& c) {
    Hit h;
    if (s.accel == OR_ACCEL_BVH) h = s.bvh.intersect(s.tris, r, c);
    else h = bb_intersect(*s.root, s.tris, r, c);
    /* EXTENSION: analytic spheres, ids after the triangles; strict < keeps the lowest id on exact-t ties */
    for (uint32_t j = 0; j < s.spheres.size(); j++) {
        float t; V3 p;
        int face = sph_intersects(s.spheres[j], r, &t, &p);
        if (face != F_NONE && (!h.some || t < h.t)) {
            h.some = true; h.t = t; h.p = p; h.face = face; h.idx = (uint32_t)s.tris.size() + j;
        }
    }
    if (h.some && !(h.t < std::numeric_limits<float>::infinity())) c.nan_t++;
    return h;
}

/* The shadow test of the commented-out block rs:1204-1224: does ANY object other than the one that was hit
 * intersect the light ray (anywhere along it: `intersects(..).is_some()`, no distance limit)?  The old code
 * collected candidates with get_all_objects_for_ray, whose BTreeMap drops leaves with equal tmin;
 * the oracle tests every object, which is what that code means to do. */
static bool scene_any_hit(const OrScene& s, const Ray& r, uint32_t exclude) {
    if (s.accel == OR_ACCEL_BVH) { if (s.bvh.any_hit(s.tris, r, exclude)) return true; }
    else {
        for (uint32_t i = 1; i < s.tris.size(); i++) {
            float t; V3 p;
            if (i != exclude && tri_intersects(s.tris[i], r, &t, &p) != F_NONE) return true;
        }
    }
    for (uint32_t j = 0; j < s.spheres.size(); j++) {
        float t; V3 p;
        if ((uint32_t)s.tris.size() + j != exclude && sph_intersects(s.spheres[j], r, &t, &p) != F_NONE) return true;
    }
    return false;
}

static V3 project_ray(const Ray& r, const OrScene& s, uint32_t depth, Rng& g, Counters& c,
                      uint32_t* prim, float* tt);

static V3 color_ray(const Ray& r, const OrScene& s, uint32_t objidx, V3 point, int face, uint32_t depth,
                    Rng& g, Counters& c) {                                                     /* rs:1199-1254 */
    const bool is_sph = objidx >= s.tris.size();
    uint32_t kind; V3 color; float alpha, scattering; V3 normal;
    if (is_sph) {
        const Sph& sp = s.spheres[objidx - s.tris.size()];
        kind = sp.kind; color = sp.color; alpha = sp.alpha; scattering = sp.scattering;
        normal = sph_normal_out(sp, point);
        if (face == F_BACK) normal = vmul(normal, -1.0f);
    } else {
        const Tri& tr = s.tris[objidx];
        kind = tr.kind; color = tr.color; alpha = tr.alpha; scattering = tr.scattering;
        normal = tri_normal(tr, face);
    }
    /* `shadowed` is evaluated first, for every hit (rs:1203-1224, LightSource::get_shadow_ray rs:600-610): a point of
     * the light cube [orig, orig + len2)^3, the ray starts 0.005..0.01 off the surface along the facing normal */
    bool shadowed = false;
    if (s.has_light) {
        float rx = g.next_f32(), ry = g.next_f32(), rz = g.next_f32();
        V3 adj = mk(s.light_orig.x + rx * s.light_len2, s.light_orig.y + ry * s.light_len2, s.light_orig.z + rz * s.light_len2);
        V3 dir = vunit(vsub(adj, point));
        V3 smudge = vmul(normal, 0.005f * (g.next_f32() + 1.0f));
        Ray light_ray = make_ray(vadd(point, smudge), dir);
        shadowed = scene_any_hit(s, light_ray, objidx);
    }
    V3 black = make_color(0, 0, 0);
    if (face == F_EDGEFRONT || face == F_EDGEBACK) return black;                               /* rs:450-459 */
    V3 base = shadowed ? black : color;
    switch (kind) {
    case OR_SOLID:
        return base;
    case OR_MATTE: {
        Ray nr = lambertian_ray(point, normal, g);
        V3 sub = project_ray(nr, s, depth - 1, g, c, nullptr, nullptr);
        return mix_color(base, sub, alpha);
    }
    default: {
        Ray nr = reflect_ray(point, normal, r.dir, scattering, g);
        V3 sub = project_ray(nr, s, depth - 1, g, c, nullptr, nullptr);
        return mix_color(base, sub, alpha);
    }
    }
}

static V3 project_ray(const Ray& r, const OrScene& s, uint32_t depth, Rng& g, Counters& c,
                      uint32_t* prim, float* tt) {                                             /* rs:1256-1295 */
    if (depth == 0) return make_color(0, 0, 0);
    V3 blue = make_color(128, 180, 255);
    Hit hit = scene_hit(s, r, c);
    c.rays++;
    if (!hit.some) { if (prim) *prim = 0; if (tt) *tt = 0.0f; return blue; }
    if (prim) *prim = hit.idx;
    if (tt) *tt = hit.t;
    return color_ray(r, s, hit.idx, hit.p, hit.face, depth, g, c);
}

static void tri_from_c(const OrTriangle& o, Tri* t) {
    t->incenter = mk(o.incenter[0], o.incenter[1], o.incenter[2]);
    t->norm = mk(o.norm[0], o.norm[1], o.norm[2]);
    t->bounding_r2 = o.bounding_r2;
    for (int i = 0; i < 3; i++) {
        t->sides[i] = mk(o.sides[3 * i], o.sides[3 * i + 1], o.sides[3 * i + 2]);
        t->side_lens[i] = o.side_lens[i];
        t->corners[i] = mk(o.corners[3 * i], o.corners[3 * i + 1], o.corners[3 * i + 2]);
    }
    t->edge_thickness = o.edge_thickness;
    t->kind = o.kind;
    t->color = mk(o.color[0], o.color[1], o.color[2]);
    t->alpha = o.alpha; t->scattering = o.scattering;
}
static void tri_to_c(const Tri& t, OrTriangle* o) {
    o->incenter[0] = t.incenter.x; o->incenter[1] = t.incenter.y; o->incenter[2] = t.incenter.z;
    o->norm[0] = t.norm.x; o->norm[1] = t.norm.y; o->norm[2] = t.norm.z;
    o->bounding_r2 = t.bounding_r2;
    for (int i = 0; i < 3; i++) {
        o->sides[3 * i] = t.sides[i].x; o->sides[3 * i + 1] = t.sides[i].y; o->sides[3 * i + 2] = t.sides[i].z;
        o->side_lens[i] = t.side_lens[i];
        o->corners[3 * i] = t.corners[i].x; o->corners[3 * i + 1] = t.corners[i].y; o->corners[3 * i + 2] = t.corners[i].z;
    }
    o->edge_thickness = t.edge_thickness;
    o->kind = t.kind;
    o->color[0] = t.color.x; o->color[1] = t.color.y; o->color[2] = t.color.z;
    o->alpha = t.alpha; o->scattering = t.scattering;
}
static View view_from_c(const OrView& v) {
    View w;
    w.width = v.width; w.height = v.height;
    w.orig = mk(v.orig[0], v.orig[1], v.orig[2]);
    w.cam = mk(v.cam[0], v.cam[1], v.cam[2]);
    w.vu = mk(v.vu[0], v.vu[1], v.vu[2]);
    w.vv = mk(v.vv[0], v.vv[1], v.vv[2]);
    w.maxdepth = v.maxdepth; w.spp = v.spp;
    return w;
}
static void tree_stats(const BBox* b, OrTreeStats* st, uint32_t maxdepth_seen) {
    st->nodes++;
    st->max_depth = std::max<uint64_t>(st->max_depth, b->depth);
    if (b->is_leaf) {
        st->leaves++;
        st->leaf_refs += b->tris.size();
        st->max_leaf = std::max<uint64_t>(st->max_leaf, b->tris.size());
    } else {
        for (const BBox* k : b->boxes) tree_stats(k, st, maxdepth_seen);
    }
}
static void count_depth(const BBox* b, uint64_t depth, uint64_t* n) {
    if (b->is_leaf) { if (b->depth == depth) (*n)++; }
    else for (const BBox* k : b->boxes) count_depth(k, depth, n);
}

} // namespace

/* ================================================================== */
/* C interface                                                          */
/* ================================================================== */
extern "C" {

int or_make_triangle(const float pts[9], uint32_t kind, const float color[3], float alpha,
                     float scattering, float edge_thickness, OrTriangle* out) {
    V3 p[3] = {mk(pts[0], pts[1], pts[2]), mk(pts[3], pts[4], pts[5]), mk(pts[6], pts[7], pts[8])};
    Tri t;
    if (!make_triangle(p, kind, mk(color[0], color[1], color[2]), alpha, scattering, edge_thickness, &t)) return -1;
    tri_to_c(t, out);
    return 0;
}

void or_make_dummy_triangle(OrTriangle* out) {                                                 /* rs:385-391 */
    V3 p[3] = {mk(1, 0, 0), mk(0, 1, 0), mk(0, 0, 1)};
    Tri t;
    make_triangle(p, OR_SOLID, make_color(255, 0, 0), 0.0f, 0.0f, 0.0f, &t);
    tri_to_c(t, out);
}

int or_make_disk(const float orig[3], const float norm[3], float r, float d, uint32_t num_tris,
                 uint32_t kind, const float color[3], float alpha, float scattering,
                 uint32_t side_kind, const float side_color[3], float side_alpha, float side_scattering,
                 float edge_thickness, OrTriangle* out) {
    std::vector<Tri> v;
    Surf s = {kind, mk(color[0], color[1], color[2]), alpha, scattering};
    Surf ss = {side_kind, mk(side_color[0], side_color[1], side_color[2]), side_alpha, side_scattering};
    int n = make_disk(mk(orig[0], orig[1], orig[2]), mk(norm[0], norm[1], norm[2]), r, d, num_tris, s, ss,
                      edge_thickness, v);
    if (n < 0) return -1;
    for (int i = 0; i < n; i++) tri_to_c(v[i], &out[i]);
    return n;
}

int or_make_sphere(const float orig[3], float r, uint32_t lat, uint32_t lon,
                   uint32_t kind, const float color[3], float alpha, float scattering,
                   float edge_thickness, OrTriangle* out, uint32_t cap) {
    std::vector<Tri> v;
    Surf s = {kind, mk(color[0], color[1], color[2]), alpha, scattering};
    int n = make_sphere(mk(orig[0], orig[1], orig[2]), r, lat, lon, s, edge_thickness, v);
    if (n < 0 || (uint32_t)n > cap) return -1;
    for (int i = 0; i < n; i++) tri_to_c(v[i], &out[i]);
    return n;
}

void or_create_transform(const float dir[3], float d_roll, float out[9]) {
    Basis b = create_transform(mk(dir[0], dir[1], dir[2]), d_roll);
    out[0] = b.b0.x; out[1] = b.b0.y; out[2] = b.b0.z;
    out[3] = b.b1.x; out[4] = b.b1.y; out[5] = b.b1.z;
    out[6] = b.b2.x; out[7] = b.b2.y; out[8] = b.b2.z;
}

void or_unit(const float v[3], float out[3]) {
    V3 u = vunit(mk(v[0], v[1], v[2]));
    out[0] = u.x; out[1] = u.y; out[2] = u.z;
}

float or_to_radians(float deg) { return to_radians(deg); }

void or_make_color(uint8_t r, uint8_t g, uint8_t b, float out[3]) {
    V3 c = make_color(r, g, b);
    out[0] = c.x; out[1] = c.y; out[2] = c.z;
}

void or_create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const float pos[3],
                        const float dir[3], float fov_deg, float c_roll, uint32_t maxdepth,
                        uint32_t samples, OrView* out) {
    View v = create_viewport(px_w, px_h, size0, size1, mk(pos[0], pos[1], pos[2]), mk(dir[0], dir[1], dir[2]),
                             fov_deg, c_roll, maxdepth, samples);
    out->width = v.width; out->height = v.height;
    out->orig[0] = v.orig.x; out->orig[1] = v.orig.y; out->orig[2] = v.orig.z;
    out->cam[0] = v.cam.x; out->cam[1] = v.cam.y; out->cam[2] = v.cam.z;
    out->vu[0] = v.vu.x; out->vu[1] = v.vu.y; out->vu[2] = v.vu.z;
    out->vv[0] = v.vv.x; out->vv[1] = v.vv.y; out->vv[2] = v.vv.z;
    out->maxdepth = v.maxdepth; out->spp = v.spp;
}

int or_mesh_to_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                         const float offset[3], float scale, const float transform[9],
                         uint32_t kind, const float color[3], float alpha, float scattering,
                         float edge_thickness, OrTriangle* out) {                              /* obj_parser.rs:59-72 */
    Basis b;
    b.b0 = mk(transform[0], transform[1], transform[2]);
    b.b1 = mk(transform[3], transform[4], transform[5]);
    b.b2 = mk(transform[6], transform[7], transform[8]);
    V3 off = mk(offset[0], offset[1], offset[2]);
    for (uint32_t f = 0; f < nfaces; f++) {
        V3 p[3];
        for (int k = 0; k < 3; k++) {
            uint32_t vi = faces[3 * f + k];
            if (vi < 1 || vi > nverts) return -2;
            V3 v = mk(verts[3 * (vi - 1)], verts[3 * (vi - 1) + 1], verts[3 * (vi - 1) + 2]);
            p[k] = vadd(change_basis(vmul(v, scale), b), off);
        }
        Tri t;
        if (!make_triangle(p, kind, mk(color[0], color[1], color[2]), alpha, scattering, edge_thickness, &t)) return -1;
        tri_to_c(t, &out[f]);
    }
    return (int)nfaces;
}

int or_parse_obj_file(const char* path, float* verts, uint32_t vcap, uint32_t* nverts,
                      uint32_t* faces, uint32_t fcap, uint32_t* nfaces) {                      /* obj_parser.rs:20-57 */
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    char line[1024];
    uint32_t nv = 0, nf = 0;
    while (fgets(line, sizeof line, f)) {
        if (line[0] == 'v' && line[1] == ' ') {
            if (nv >= vcap) { fclose(f); return -2; }
            char* p = line + 2;
            for (int k = 0; k < 3; k++) { char* e; verts[3 * nv + k] = strtof(p, &e); if (e == p) { fclose(f); return -3; } p = e; }
            nv++;
        } else if (line[0] == 'f' && line[1] == ' ') {
            if (nf >= fcap) { fclose(f); return -2; }
            char* p = line + 2;
            for (int k = 0; k < 3; k++) {
                while (*p == ' ' || *p == '\t') p++;
                char* e; unsigned long idx = strtoul(p, &e, 10);
                if (e == p) { fclose(f); return -3; }
                faces[3 * nf + k] = (uint32_t)idx;
                while (*e && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r') e++;   /* skip /vt/vn */
                p = e;
            }
            nf++;
        }
    }
    fclose(f);
    *nverts = nv; *nfaces = nf;
    return 0;
}

void or_pixel_ray(const OrView* v, uint32_t row, uint32_t col, float out[9]) {
    View w = view_from_c(*v);
    w.spp = 1;
    Ray r = pixel_ray(w, row, col, nullptr);
    out[0] = r.orig.x; out[1] = r.orig.y; out[2] = r.orig.z;
    out[3] = r.dir.x; out[4] = r.dir.y; out[5] = r.dir.z;
    out[6] = r.inv_dir.x; out[7] = r.inv_dir.y; out[8] = r.inv_dir.z;
}

int or_triangle_intersects(const OrTriangle* t, const float orig[3], const float dir[3],
                           float* t_out, float p_out[3]) {
    Tri tr; tri_from_c(*t, &tr);
    Ray r; r.orig = mk(orig[0], orig[1], orig[2]); r.dir = mk(dir[0], dir[1], dir[2]);
    r.inv_dir = mk(1.0f / r.dir.x, 1.0f / r.dir.y, 1.0f / r.dir.z);
    float tt = 0; V3 p = mk(0, 0, 0);
    int face = tri_intersects(tr, r, &tt, &p);
    if (face) { *t_out = tt; p_out[0] = p.x; p_out[1] = p.y; p_out[2] = p.z; }
    return face;
}

OrScene* or_scene_create(const OrTriangle* tris, uint32_t n, int accel, const float root_orig[3],
                         float root_len2, uint32_t maxdepth, uint32_t minobjs, int build_threads) {
    OrScene* s = new OrScene();
    s->tris.resize(n);
    for (uint32_t i = 0; i < n; i++) tri_from_c(tris[i], &s->tris[i]);
    s->accel = accel;
    s->root = nullptr;
    V3 ro = mk(root_orig[0], root_orig[1], root_orig[2]);
    if (accel == OR_ACCEL_OCTREE) {                                                            /* rs:790-793 */
        std::vector<uint32_t> refvec;
        for (uint32_t i = 1; i < n; i++) refvec.push_back(i);
        s->root = build_bb_helper(s->tris, refvec, ro, root_len2, 0, maxdepth, minobjs, build_threads > 1 ? 2 : 0);
        if (!s->root) { /* reference: unwrap() panic on an empty scene; oracle: empty leaf */
            s->root = new BBox(); s->root->orig = ro; s->root->len2 = root_len2; s->root->depth = 0; s->root->is_leaf = true;
        }
    } else if (accel == OR_ACCEL_TRIVIAL) {                                                    /* rs:847-856 */
        s->root = new BBox(); s->root->orig = ro; s->root->len2 = root_len2; s->root->depth = 0; s->root->is_leaf = true;
        for (uint32_t i = 1; i < n; i++) s->root->tris.push_back(i);
    } else {
        s->bvh.init(s->tris);
    }
    return s;
}

void or_scene_destroy(OrScene* s) { delete s; }

/* EXTENSION: analytic spheres (primitive ids n_tris + j) and the light of the reference's
 * commented-out shadow code. */
void or_scene_add_spheres(OrScene* s, const OrSphere* sp, uint32_t n) {
    for (uint32_t j = 0; j < n; j++) {
        Sph q;
        q.c = mk(sp[j].center[0], sp[j].center[1], sp[j].center[2]); q.r = sp[j].radius;
        q.kind = sp[j].kind; q.color = mk(sp[j].color[0], sp[j].color[1], sp[j].color[2]);
        q.alpha = sp[j].alpha; q.scattering = sp[j].scattering;
        s->spheres.push_back(q);
    }
}
void or_scene_set_light(OrScene* s, const float orig[3], float len2) {
    s->has_light = orig != nullptr;
    if (orig) { s->light_orig = mk(orig[0], orig[1], orig[2]); s->light_len2 = len2; }
}

void or_scene_tree_stats(const OrScene* s, OrTreeStats* out) {
    memset(out, 0, sizeof *out);
    if (s->root) {
        tree_stats(s->root, out, 0);
        count_depth(s->root, out->max_depth, &out->leaves_at_maxdepth);
    } else {
        out->nodes = s->bvh.nodes.size();
        for (const BNode& n : s->bvh.nodes) if (n.left < 0) { out->leaves++; out->leaf_refs += n.count; out->max_leaf = std::max<uint64_t>(out->max_leaf, n.count); }
    }
}

uint32_t or_scene_closest_hit(const OrScene* s, const float orig[3], const float dir[3], float* t_out) {
    Ray r = make_ray(mk(orig[0], orig[1], orig[2]), mk(dir[0], dir[1], dir[2]));
    Counters c;
    Hit h = scene_hit(*s, r, c);
    if (!h.some) return 0;
    if (t_out) *t_out = h.t;
    return h.idx;
}

int or_render(const OrScene* s, const OrView* vc, uint64_t seed, int threads, uint32_t row0, uint32_t row1,
              float* rgba, uint32_t* prim, float* tbuf, OrStats* stats) {
    return or_render_samples(s, vc, seed, threads, row0, row1, 0, vc->spp, 0, rgba, prim, tbuf, stats);
}

/* The sample loop of rs:1418-1426 restricted to samples [s0, s1) of v.spp (the jitter rule still looks at v.spp);
 * sum_only = 1 skips the final `* (1/spp)` so that partial sums of different sample ranges can be added up. */
int or_render_samples(const OrScene* s, const OrView* vc, uint64_t seed, int threads, uint32_t row0, uint32_t row1,
                      uint32_t s0, uint32_t s1, int sum_only,
                      float* rgba, uint32_t* prim, float* tbuf, OrStats* stats) {
    View v = view_from_c(*vc);
    if (row1 > v.height) row1 = v.height;
    if (threads < 1) threads = 1;
    std::atomic<uint32_t> next_row(row0);   /* the row work-queue of rs:1181-1194, :1402-1408 */
    std::vector<Counters> cs(threads);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        Counters& c = cs[tid];
        for (;;) {
            uint32_t row = next_row.fetch_add(1);
            if (row >= row1) break;
            for (uint32_t col = 0; col < v.width; col++) {                                     /* rs:1413-1427 */
                V3 acc = mk(0, 0, 0);
                uint64_t pix = (uint64_t)row * v.width + col;
                for (uint32_t smp = s0; smp < s1; smp++) {
                    Rng g; g.seed(seed, pix, smp);
                    Ray ray = pixel_ray(v, row, col, &g);
                    uint32_t pr = 0; float tt = 0;
                    V3 col3 = project_ray(ray, *s, v.maxdepth, g, c, &pr, &tt);
                    acc = vadd(acc, col3);
                    if (smp == 0) { if (prim) prim[pix] = pr; if (tbuf) tbuf[pix] = tt; }
                }
                V3 o = sum_only ? acc : vmul(acc, 1.0f / (float)v.spp);
                rgba[4 * pix + 0] = o.x; rgba[4 * pix + 1] = o.y; rgba[4 * pix + 2] = o.z; rgba[4 * pix + 3] = 0.0f;
            }
        }
    };
    if (threads == 1) worker(0);
    else {
        std::vector<std::thread> th;
        for (int i = 0; i < threads; i++) th.emplace_back(worker, i);
        for (auto& t : th) t.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        memset(stats, 0, sizeof *stats);
        for (const Counters& c : cs) {
            stats->rays += c.rays; stats->box_tests += c.box_tests; stats->tri_tests += c.tri_tests;
            stats->node_visits += c.node_visits; stats->nan_t_hits += c.nan_t;
        }
        stats->seconds = std::chrono::duration<double>(t1 - t0).count();
    }
    return 0;
}

void or_quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb) {                        /* rs:1468-1473 */
    for (uint64_t i = 0; i < npix; i++)
        for (int k = 0; k < 3; k++) {
            float c = rgba[4 * i + k] * 255.0f;
            /* Rust `as u8`: truncate toward zero, saturate, NaN -> 0 */
            uint8_t q;
            if (!(c == c)) q = 0; else if (c <= 0.0f) q = 0; else if (c >= 255.0f) q = 255; else q = (uint8_t)c;
            rgb[3 * i + k] = q;
        }
}

int or_selftest_face_collision(void) {                                                         /* rs:735-750 */
    V3 orig = mk(2.0f, 2.0f, 2.0f);
    V3 norm = mk(0.0f, 0.0f, -1.0f);
    float len2 = 2.0f;
    V3 p[3] = {mk(1.0f, 0.4f, 0.2f), mk(1.0f, 0.2f, -0.3f), mk(0.6f, 0.6f, -0.5f)};
    Tri t;
    if (!make_triangle(p, OR_SOLID, make_color(0, 0, 0), 0.0f, 0.0f, 0.0f, &t)) return -1;
    return face_contains_triangle(orig, norm, len2, t) ? 1 : 0;
}

float or_rng_f32(uint64_t seed, uint64_t pixel, uint32_t sample, uint32_t n) {
    Rng g; g.seed(seed, pixel, sample);
    float f = 0;
    for (uint32_t i = 0; i <= n; i++) f = g.next_f32();
    return f;
}

} /* extern "C" */
