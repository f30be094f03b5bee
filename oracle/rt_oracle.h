/*
 * rt_oracle.h — C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference
 * renderer's hot path (gerikkub/rust_raytrace, raytrace_lib/src/raytrace.rs),
 * used as the checker for the CUDA path.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (rust_raytrace_b200/, include/rtb.h) never links, imports or calls it.
 *
 * PARITY UNPINNED by the reference's own tests: the reference is nightly Rust
 * with un-vendored crates and no toolchain exists in the build image, and its
 * only hot-path-adjacent known-answer test is `face_collision`
 * (raytrace.rs:735-750), which or_selftest_face_collision() reproduces.  The
 * checked-in PNGs predate the current code (different sky constant and disk
 * geometry) and are not golden vectors.  Everything else is pinned by this
 * restatement plus the structural invariants in tests/test_oracle.py.
 *
 * All citations `raytrace.rs:N` are to /root/reference/raytrace_lib/src/.
 */
#ifndef RT_ORACLE_H
#define RT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* SurfaceKind discriminants (raytrace.rs:303-308). */
enum { OR_SOLID = 0, OR_MATTE = 1, OR_REFLECTIVE = 2 };

/* Triangle in the reference's field order (raytrace.rs:326-337), flattened. */
typedef struct OrTriangle {
    float incenter[3];
    float norm[3];
    float bounding_r2;
    float sides[9];      /* sides[i] = floats 3i..3i+2 */
    float side_lens[3];
    float corners[9];
    float edge_thickness;
    uint32_t kind;       /* OR_SOLID / OR_MATTE / OR_REFLECTIVE */
    float color[3];
    float alpha;
    float scattering;
} OrTriangle;            /* 35 x 4 bytes = 140 bytes */

/* Viewport (raytrace.rs:1305-1318). */
typedef struct OrView {
    uint32_t width, height;
    float orig[3];
    float cam[3];
    float vu[3];
    float vv[3];
    uint32_t maxdepth;
    uint32_t spp;
} OrView;

typedef struct OrStats {
    uint64_t rays;        /* project_ray calls with depth>0 (raytrace.rs:1278) */
    uint64_t box_tests;   /* BoundingBox::collides calls / BVH slab tests       */
    uint64_t tri_tests;   /* Triangle::intersects calls                        */
    uint64_t node_visits; /* get_object_intersection_for_ray calls             */
    uint64_t nan_t_hits;  /* accepted hits whose t is NaN or +inf (pathological)*/
    double   seconds;     /* wall time of the row loop                         */
} OrStats;

typedef struct OrTreeStats {
    uint64_t nodes, leaves, leaf_refs, max_leaf, max_depth, leaves_at_maxdepth;
} OrTreeStats;

/* Acceleration structure used by the oracle. */
enum {
    OR_ACCEL_OCTREE = 0,  /* build_bounding_box, raytrace.rs:790-845 (reference algorithm)   */
    OR_ACCEL_TRIVIAL = 1, /* build_trivial_bounding_box, raytrace.rs:847-856 (brute force)   */
    OR_ACCEL_BVH = 2      /* oracle-only median-split BVH with the exact per-triangle test;   */
                          /* ties -> lowest triangle index (for scenes the octree can't build)*/
};

typedef struct OrScene OrScene;

/* ---- scene preparation (host side of the reference) ---- */
/* make_triangle raytrace.rs:340-383; returns 0 on success, -1 where the reference would panic. */
int  or_make_triangle(const float pts[9], uint32_t kind, const float color[3], float alpha,
                      float scattering, float edge_thickness, OrTriangle* out);
void or_make_dummy_triangle(OrTriangle* out);                                  /* :385-391 */
/* make_disk :531-592; writes 4*num_tris triangles; returns count or -1. */
int  or_make_disk(const float orig[3], const float norm[3], float r, float d, uint32_t num_tris,
                  uint32_t kind, const float color[3], float alpha, float scattering,
                  uint32_t side_kind, const float side_color[3], float side_alpha, float side_scattering,
                  float edge_thickness, OrTriangle* out);
/* make_sphere :464-529; writes up to 2*lat*lon triangles; returns count or -1. */
int  or_make_sphere(const float orig[3], float r, uint32_t lat, uint32_t lon,
                    uint32_t kind, const float color[3], float alpha, float scattering,
                    float edge_thickness, OrTriangle* out, uint32_t cap);
/* create_transform :1320-1341; out = 3 rows x 3. */
void or_create_transform(const float dir[3], float d_roll, float out[9]);
void or_unit(const float v[3], float out[3]);                                  /* :93-96 */
float or_to_radians(float deg);
void or_make_color(uint8_t r, uint8_t g, uint8_t b, float out[3]);             /* :176-180 */
/* create_viewport :1343-1370. */
void or_create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const float pos[3],
                        const float dir[3], float fov_deg, float c_roll, uint32_t maxdepth,
                        uint32_t samples, OrView* out);
/* parse_obj obj_parser.rs:47-73 given already-split vertex/face arrays (face indices 1-based). */
int  or_mesh_to_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                          const float offset[3], float scale, const float transform[9],
                          uint32_t kind, const float color[3], float alpha, float scattering,
                          float edge_thickness, OrTriangle* out);
/* OBJ text -> vertex/face arrays (obj_parser.rs:20-45; strtof == Rust's correctly-rounded parse). */
int  or_parse_obj_file(const char* path, float* verts, uint32_t vcap, uint32_t* nverts,
                       uint32_t* faces, uint32_t fcap, uint32_t* nfaces);
/* pixel_ray :1374-1394 with the centre sample; out = orig3, dir3, inv_dir3. */
void or_pixel_ray(const OrView* v, uint32_t row, uint32_t col, float out[9]);

/* ---- hot path ---- */
/* Triangle::intersects :400-439. Returns 0 miss, else 1 Front, 2 Back, 3 EdgeFront, 4 EdgeBack. */
int  or_triangle_intersects(const OrTriangle* t, const float orig[3], const float dir[3],
                            float* t_out, float p_out[3]);
OrScene* or_scene_create(const OrTriangle* tris, uint32_t n, int accel,
                         const float root_orig[3], float root_len2,
                         uint32_t maxdepth, uint32_t minobjs, int build_threads);
void or_scene_destroy(OrScene* s);
/* EXTENSION (not in the mounted reference, see rt_oracle.cpp): analytic spheres, primitive ids
 * n_tris + j, and the light source of the reference's commented-out shadow code (raytrace.rs:594-610, :1203-1224);
 * orig == NULL removes the light. */
typedef struct OrSphere { float center[3]; float radius; uint32_t kind; float color[3]; float alpha; float scattering; } OrSphere;
void or_scene_add_spheres(OrScene* s, const OrSphere* spheres, uint32_t n);
void or_scene_set_light(OrScene* s, const float orig[3], float len2);
void or_scene_tree_stats(const OrScene* s, OrTreeStats* out);
/* Closest hit for one explicit ray (dir is normalised by make_ray); returns prim id, 0 = miss. */
uint32_t or_scene_closest_hit(const OrScene* s, const float orig[3], const float dir[3], float* t_out);
/* Full frame: DefaultRayCaster::walk_rays_internal :1175-1195 + walk_ray_set :1396-1440.
 * rgba: W*H*4 f32 (lane 3 = 0); prim/t (nullable): primary-ray primitive id (0 = miss) and t.
 * rows [row0,row1) only are rendered (others untouched); pass 0,height for the whole frame. */
int  or_render(const OrScene* s, const OrView* v, uint64_t seed, int threads,
               uint32_t row0, uint32_t row1,
               float* rgba, uint32_t* prim, float* t, OrStats* stats);
/* Same, for samples [s0, s1) of v->spp only; sum_only != 0 returns the un-normalised sample sum (the checker for
 * the sample-partitioned multi-GPU mode, where partial sums are reduced before the 1/spp scale). */
int  or_render_samples(const OrScene* s, const OrView* v, uint64_t seed, int threads,
                       uint32_t row0, uint32_t row1, uint32_t s0, uint32_t s1, int sum_only,
                       float* rgba, uint32_t* prim, float* t, OrStats* stats);
/* write_png quantiser :1468-1473: (c*255.) as u8, rgb only. */
void or_quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb);

/* ---- known-answer tests of the reference ---- */
int  or_selftest_face_collision(void);   /* raytrace.rs:735-750 -> 1 when the assertion holds */
/* RNG draw (shared spec with the CUDA path): n-th f32 of stream (seed,pixel,sample). */
float or_rng_f32(uint64_t seed, uint64_t pixel, uint32_t sample, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif
