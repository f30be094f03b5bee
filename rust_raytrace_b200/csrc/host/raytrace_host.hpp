// raytrace_host.hpp — C++ mirror of the reference's host-side API surface for the
// accelerated path: the types and constructors a `main.rs`-style program uses
// before and after `RayCaster::walk_rays` (raytrace_lib/src/raytrace.rs).
//
// Everything here is host-only f32 arithmetic in the reference's operation
// order (no FMA contraction: this file is compiled with -ffp-contract=off), so
// that the triangles and the viewport handed to the GPU are the ones the Rust
// program would have produced.  The hot path itself lives in ../rtb_*.cu.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "rtb.h"
#include "rtb_host.h"

namespace raytrace {

// Vec3 (raytrace.rs:22-122).  Lane 3 of the reference's f32x4 is always 0 and is
// represented implicitly: dot()/len2() add the +0 lane product last.
struct Vec3 {
    float v[3];

    Vec3 add(const Vec3& o) const { return {{v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}}; }
    Vec3 sub(const Vec3& o) const { return {{v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}}; }
    Vec3 mult(float a) const { return {{v[0] * a, v[1] * a, v[2] * a}}; }
    Vec3 mult_per(const Vec3& o) const { return {{v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2]}}; }
    float dot(const Vec3& o) const;
    float len2() const { return dot(*this); }
    float len() const;
    Vec3 cross(const Vec3& o) const;
    Vec3 unit() const;
    Vec3 orthogonal() const;
};
using Point = Vec3;
using Color = Vec3;

inline Vec3 make_vec(float x, float y, float z) { return {{x, y, z}}; }
Color make_color(uint8_t r, uint8_t g, uint8_t b);
float to_radians(float deg);

struct Basis { Vec3 r0, r1, r2; };
Vec3 change_basis(const Vec3& v, const Basis& b);

// SurfaceKind (raytrace.rs:303-308)
struct SurfaceKind {
    uint32_t kind;
    Color color;
    float alpha;
    float scattering;
    static SurfaceKind Solid(Color c) { return {RTB_SOLID, c, 0.f, 0.f}; }
    static SurfaceKind Matte(Color c, float alpha) { return {RTB_MATTE, c, alpha, 0.f}; }
    static SurfaceKind Reflective(float scattering, Color c, float alpha) { return {RTB_REFLECTIVE, c, alpha, scattering}; }
};

// Triangle (raytrace.rs:326-337) is carried in its flattened C form.
using Triangle = RtbTriangle;

// Throws std::runtime_error where the reference panics.
Triangle make_triangle(const Vec3 pts[3], const SurfaceKind& surface, float edge_thickness);
bool try_make_triangle(const Vec3 pts[3], const SurfaceKind& surface, float edge_thickness, Triangle* out);
Triangle make_dummy_triangle();
inline void populate_triangle_numbers(std::vector<Triangle>&) {}  // `num` is the array index in this ABI

std::vector<Triangle> make_sphere(const Point& orig, float r, uint32_t lat, uint32_t lon,
                                  const SurfaceKind& surface, float edge_thickness);
std::vector<Triangle> make_disk(const Point& orig, const Vec3& norm, float r, float d, uint32_t num_tris,
                                const SurfaceKind& surface, const SurfaceKind& side_surface, float edge_thickness);

Basis create_transform(const Vec3& dir_in, float d_roll);

// Viewport (raytrace.rs:1305-1318) == RtbView (seed/sample range/flags are additions).
using Viewport = RtbView;
Viewport create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const Point& pos, const Vec3& dir,
                         float fov, float c_roll, uint32_t maxdepth, uint32_t samples);

namespace obj_parser {
struct Mesh { std::vector<float> verts; std::vector<uint32_t> faces; };
Mesh read_obj(const std::string& path);
Mesh read_mesh_bin(const std::string& path);
std::vector<Triangle> mesh_to_triangles(const Mesh& m, const Vec3& offset, float scale, const Basis& transform,
                                        const SurfaceKind& surface, float edge_thickness);
// parse_obj (obj_parser.rs:47-73)
std::vector<Triangle> parse_obj(const std::string& path, const Vec3& offset, float scale, const Basis& transform,
                                const SurfaceKind& surface, float edge_thickness);
}  // namespace obj_parser

bool box_contains_polygon(const Point& orig, float len2, const Triangle& t);

// write_png's quantiser (raytrace.rs:1468-1473) with a PPM container.
void quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb);
bool write_ppm(const std::string& path, uint32_t width, uint32_t height, const float* rgba);
// write_png (raytrace.rs:1460-1478): 8-bit RGB, no interlace; own encoder (stored deflate blocks).
bool write_png_rgb8(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgb);
bool write_png(const std::string& path, uint32_t width, uint32_t height, const float* rgba);

}  // namespace raytrace
