/*
 * rtb_host.h — C ABI of the host-side scene/camera preparation that sits in
 * front of the GPU path.  These are the product's own C++ implementations of
 * the reference's host-only helpers (no GPU needed), exported so that a
 * non-Rust host (the C++ driver, the Python tests and bench.py) can build
 * exactly the `Vec<Triangle>` / `Viewport` the reference's main.rs builds.
 * In the Rust drop-in these are NOT used: raytrace_lib itself produces the
 * triangles and the shim only flattens them into RtbTriangle.
 *
 * Mirrors (reference file:line):
 *   rtbh_make_triangle      make_triangle            raytrace.rs:340-383
 *   rtbh_make_dummy_triangle make_dummy_triangle     raytrace.rs:385-391
 *   rtbh_make_disk          make_disk                raytrace.rs:531-592
 *   rtbh_make_sphere        make_sphere              raytrace.rs:464-529
 *   rtbh_create_transform   create_transform         raytrace.rs:1320-1341
 *   rtbh_create_viewport    create_viewport          raytrace.rs:1343-1370
 *   rtbh_parse_obj          parse_obj                obj_parser.rs:47-73
 *   rtbh_box_contains_polygon box_contains_polygon   raytrace.rs:753-779
 *   rtbh_write_ppm          write_png (quantiser)    raytrace.rs:1460-1478
 *   rtbh_write_png[_rgb8]   write_png                raytrace.rs:1460-1478
 */
#ifndef RTB_HOST_H
#define RTB_HOST_H

#include "rtb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A surface description = the reference's SurfaceKind flattened. */
typedef struct RtbSurface {
    uint32_t kind;      /* RTB_SOLID / RTB_MATTE / RTB_REFLECTIVE */
    float color[3];
    float alpha;
    float scattering;
} RtbSurface;

void  rtbh_make_color(uint8_t r, uint8_t g, uint8_t b, float out[3]);      /* raytrace.rs:176-180 */
void  rtbh_unit(const float v[3], float out[3]);                            /* raytrace.rs:93-96   */
float rtbh_to_radians(float deg);                                           /* f32::to_radians     */

/* Returns RTB_OK, or RTB_ERR_INVALID where the reference panics (degenerate triangle, raytrace.rs:357). */
int rtbh_make_triangle(const float pts[9], const RtbSurface* surface, float edge_thickness, RtbTriangle* out);
int rtbh_make_dummy_triangle(RtbTriangle* out);
/* Writes 4*num_tris triangles to out (top, bottom, 2 side per segment); returns the count or <0. */
int rtbh_make_disk(const float orig[3], const float norm[3], float r, float d, uint32_t num_tris,
                   const RtbSurface* surface, const RtbSurface* side_surface, float edge_thickness,
                   RtbTriangle* out, uint32_t cap);
/* lat must be even (reference asserts); returns the count (<= 2*lat*lon) or <0. */
int rtbh_make_sphere(const float orig[3], float r, uint32_t lat, uint32_t lon, const RtbSurface* surface,
                     float edge_thickness, RtbTriangle* out, uint32_t cap);
void rtbh_create_transform(const float dir[3], float d_roll, float out_rows[9]);
void rtbh_create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const float pos[3],
                          const float dir[3], float fov_deg, float c_roll, uint32_t maxdepth, uint32_t samples,
                          RtbView* out);

/* OBJ text ("v x y z" and "f a[/..] b[/..] c[/..]" lines only, like the reference) -> triangles.
 * First call with out == NULL to get the face count. */
int rtbh_parse_obj(const char* path, const float offset[3], float scale, const float transform_rows[9],
                   const RtbSurface* surface, float edge_thickness, RtbTriangle* out, uint32_t cap);
/* Same transform/triangle pipeline from in-memory arrays (faces are 1-based vertex indices, 3 per face). */
int rtbh_mesh_to_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                           const float offset[3], float scale, const float transform_rows[9],
                           const RtbSurface* surface, float edge_thickness, RtbTriangle* out);
/* Binary mesh container used for the bundled teapot: "RTBM", u32 nverts, u32 nfaces, f32 verts[3*nv], u32 faces[3*nf]. */
int rtbh_load_mesh_bin(const char* path, float* verts, uint32_t vcap, uint32_t* nverts,
                       uint32_t* faces, uint32_t fcap, uint32_t* nfaces);

/* 1 / 0: would the reference's octree root keep this triangle? */
int rtbh_box_contains_polygon(const float orig[3], float len2, const RtbTriangle* t);

/* Quantise with `(c*255.) as u8` and write a binary PPM (the PNG encoder itself is out of scope). */
int rtbh_write_ppm(const char* path, uint32_t width, uint32_t height, const float* rgba);

/* write_png (raytrace.rs:1460-1478): 8-bit RGB PNG of the frame.  rtbh_write_png quantises f32 RGBA with the
 * reference's `(c*255.) as u8`; rtbh_write_png_rgb8 takes pixels that are already quantised (rtb_render_rgb8,
 * rtb_quantize_rgb8).  The encoder is self-contained (zlib "stored" blocks, CRC-32, Adler-32): the file holds exactly
 * the bytes the reference's `png` crate would decode to, at no compression. */
int rtbh_write_png(const char* path, uint32_t width, uint32_t height, const float* rgba);
int rtbh_write_png_rgb8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb);

#ifdef __cplusplus
}
#endif
#endif /* RTB_HOST_H */
