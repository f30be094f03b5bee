// raytrace_host.cpp — host-side scene / camera preparation (see raytrace_host.hpp).
// Compiled with g++ -ffp-contract=off: every expression below is evaluated with
// separate f32 multiplies and adds in the order the reference writes them.
#include "raytrace_host.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <stdexcept>

namespace raytrace {

static constexpr float kPi = 3.14159265358979323846f;      // std::f32::consts::PI
static constexpr float kFracPi2 = 1.57079632679489661923f;  // FRAC_PI_2

// ---- Vec3 ------------------------------------------------------------------
float Vec3::dot(const Vec3& o) const {  // raytrace.rs:75-77: lane-wise product, ordered reduce_sum
    float lane0 = v[0] * o.v[0];
    float lane1 = v[1] * o.v[1];
    float lane2 = v[2] * o.v[2];
    float lane3 = 0.0f;  // 0 * 0
    float acc = lane0 + lane1;
    acc = acc + lane2;
    acc = acc + lane3;
    return acc;
}
float Vec3::len() const { return std::sqrt(len2()); }
Vec3 Vec3::cross(const Vec3& o) const {  // raytrace.rs:80-90: (yzx * o.zxy) - (zxy * o.yzx)
    const float a1[3] = {v[1], v[2], v[0]}, a2[3] = {v[2], v[0], v[1]};
    const float b1[3] = {o.v[1], o.v[2], o.v[0]}, b2[3] = {o.v[2], o.v[0], o.v[1]};
    Vec3 r;
    for (int k = 0; k < 3; ++k) r.v[k] = a1[k] * b2[k] - a2[k] * b1[k];
    return r;
}
Vec3 Vec3::unit() const { return mult(1.0f / len()); }  // raytrace.rs:93-96 (reciprocal, then multiply)
Vec3 Vec3::orthogonal() const {                         // raytrace.rs:98-108
    Vec3 cur = *this;
    for (int attempt = 0; attempt < 4; ++attempt) {
        if (std::fabs(cur.v[0]) > 0.1f) return make_vec(-1.0f * (cur.v[1] + cur.v[2]) / cur.v[0], 1.0f, 1.0f).unit();
        if (std::fabs(cur.v[1]) > 0.1f) return make_vec(1.0f, -1.0f * (cur.v[0] + cur.v[2]) / cur.v[1], 1.0f).unit();
        if (std::fabs(cur.v[2]) > 0.1f) return make_vec(1.0f, 1.0f, -1.0f * (cur.v[0] + cur.v[1]) / cur.v[2]).unit();
        cur = cur.unit();
    }
    throw std::runtime_error("orthogonal(): zero vector");
}
Color make_color(uint8_t r, uint8_t g, uint8_t b) {
    return make_vec(float(r) / 255.0f, float(g) / 255.0f, float(b) / 255.0f);
}
float to_radians(float deg) {
    const float rads_per_deg = kPi / 180.0f;
    return deg * rads_per_deg;
}
Vec3 change_basis(const Vec3& v, const Basis& b) {  // raytrace.rs:117-121
    return make_vec(b.r0.dot(v), b.r1.dot(v), b.r2.dot(v));
}

// ---- Ray helpers used only by make_triangle (raytrace.rs:194-267) ------------
namespace {
struct HostRay { Point orig; Vec3 dir; };
HostRay host_make_ray(const Point& o, const Vec3& d) { return {o, d.unit()}; }
Point ray_at(const HostRay& r, float t) { return r.dir.mult(t).add(r.orig); }

// Solve the two ray parameters in the plane spanned by components (i, j).
bool solve_2d(const HostRay& s, const HostRay& r, int i, int j, float* ts, float* tr) {
    const float det = r.dir.v[i] * s.dir.v[j] - r.dir.v[j] * s.dir.v[i];
    if (std::fabs(det) < 0.0001f) return false;
    const float dx = r.orig.v[i] - s.orig.v[i];
    const float dy = r.orig.v[j] - s.orig.v[j];
    *ts = (dy * r.dir.v[i] - dx * r.dir.v[j]) / det;
    *tr = (dy * s.dir.v[i] - dx * s.dir.v[j]) / det;
    return true;
}
// Ray::intersect: try xy, then xz, then yz projections; accept if the two points are < 0.1 apart.
bool rays_meet(const HostRay& s, const HostRay& r, Point* where) {
    float ts = 0.f, tr = 0.f;
    if (!solve_2d(s, r, 0, 1, &ts, &tr) && !solve_2d(s, r, 0, 2, &ts, &tr) && !solve_2d(s, r, 1, 2, &ts, &tr))
        return false;
    const Point ps = ray_at(s, ts), pr = ray_at(r, tr);
    if (pr.sub(ps).len2() < 0.01f) { *where = ps; return true; }
    return false;
}
void store3(float* dst, const Vec3& s) { dst[0] = s.v[0]; dst[1] = s.v[1]; dst[2] = s.v[2]; }
Vec3 load3(const float* p) { return make_vec(p[0], p[1], p[2]); }
}  // namespace

// ---- make_triangle (raytrace.rs:340-383) -------------------------------------
bool try_make_triangle(const Vec3 pts[3], const SurfaceKind& surface, float edge_thickness, Triangle* out) {
    const Vec3 &A = pts[0], &B = pts[1], &Cc = pts[2];
    const Vec3 ab = B.sub(A), ac = Cc.sub(A), bc = Cc.sub(B);
    // two medians (the reference calls them bisectors); their meeting point is the centroid
    const HostRay from_a = host_make_ray(A, ac.add(ab));
    const HostRay from_b = host_make_ray(B, bc.add(ab.mult(-1.0f)));
    Point centre;
    if (!rays_meet(from_a, from_b, &centre)) return false;

    Triangle t;
    std::memset(&t, 0, sizeof t);
    Vec3 side_dirs[3];
    for (int e = 0; e < 3; ++e) {
        const Vec3 edge = pts[(e + 1) % 3].sub(pts[e]);
        const Vec3 to_centre = centre.sub(pts[e]);
        const Vec3 foot = edge.mult(edge.dot(to_centre) / edge.len2());
        const Vec3 out_vec = foot.sub(to_centre);
        side_dirs[e] = out_vec.unit();
        t.side_lens[e] = out_vec.len();
        store3(&t.sides[3 * e], side_dirs[e]);
        store3(&t.corners[3 * e], pts[e]);
    }
    store3(t.norm, side_dirs[0].cross(side_dirs[1]).unit());
    store3(t.incenter, centre);
    float r2 = 0.0f;
    for (int e = 0; e < 3; ++e) r2 = std::fmax(r2, pts[e].sub(centre).len2());
    t.bounding_r2 = r2;
    t.edge_thickness = edge_thickness;
    t.kind = surface.kind;
    store3(t.color, surface.color);
    t.alpha = surface.alpha;
    t.scattering = surface.scattering;
    *out = t;
    return true;
}
Triangle make_triangle(const Vec3 pts[3], const SurfaceKind& surface, float edge_thickness) {
    Triangle t;
    if (!try_make_triangle(pts, surface, edge_thickness, &t))
        throw std::runtime_error("make_triangle: medians do not meet (the reference panics here, raytrace.rs:357)");
    return t;
}
Triangle make_dummy_triangle() {  // raytrace.rs:385-391
    const Vec3 pts[3] = {make_vec(1, 0, 0), make_vec(0, 1, 0), make_vec(0, 0, 1)};
    return make_triangle(pts, SurfaceKind::Solid(make_color(255, 0, 0)), 0.0f);
}

// ---- make_sphere (raytrace.rs:464-529) ---------------------------------------
std::vector<Triangle> make_sphere(const Point& orig, float r, uint32_t lat, uint32_t lon,
                                  const SurfaceKind& surface, float edge_thickness) {
    if (lat % 2 != 0) throw std::runtime_error("make_sphere: num_lat must be even");
    std::vector<Triangle> tris;
    const float nlat = float(lat), nlon = float(lon);
    auto on_sphere = [&](float phi, float theta) {
        const float s = std::sin(phi), c = std::cos(phi);
        return orig.add(make_vec(r * s, r * c * std::cos(theta), r * c * std::sin(theta)));
    };
    for (uint32_t i = 0; i < lat; ++i) {
        const bool even = (i % 2 == 0);
        const float lo_band = float(i) / nlat * kPi, hi_band = float(i + 1) / nlat * kPi;
        const float phi1 = ((even ? lo_band : hi_band) - kFracPi2) * -1.0f;
        const float phi23 = ((even ? hi_band : lo_band) - kFracPi2) * -1.0f;
        const float smudge = even ? 0.0f : 0.5f;
        for (uint32_t j = 0; j < lon; ++j) {
            const float fj = float(j);
            const float theta1 = (fj + smudge) / nlon * 2.0f * kPi;
            const float theta2 = (fj + 0.5f + smudge) / nlon * 2.0f * kPi;
            const float theta3 = (fj - 0.5f + smudge) / nlon * 2.0f * kPi;
            const float theta4 = (fj + 1.0f + smudge) / nlon * 2.0f * kPi;
            const Point p1 = on_sphere(phi1, theta1), p4 = on_sphere(phi1, theta4);
            const Point p2 = on_sphere(phi23, theta2), p3 = on_sphere(phi23, theta3);
            const Vec3 cap[3] = {p1, p2, p3};
            tris.push_back(make_triangle(cap, surface, edge_thickness));
            if (i != 0 && i != lat - 1) {
                const Vec3 fill[3] = {p1, p2, p4};
                tris.push_back(make_triangle(fill, surface, edge_thickness));
            }
        }
    }
    return tris;
}

// ---- make_disk (raytrace.rs:531-592) -----------------------------------------
std::vector<Triangle> make_disk(const Point& orig, const Vec3& norm, float r, float d, uint32_t num_tris,
                                const SurfaceKind& surface, const SurfaceKind& side_surface, float edge_thickness) {
    std::vector<Triangle> tris;
    tris.reserve(size_t(num_tris) * 4);
    const Vec3 u = norm.orthogonal().unit().mult(r);
    const Vec3 w = norm.cross(u).unit().mult(r);
    const float n = float(num_tris);
    const float smudge = 0.0f;
    const Vec3 up = norm.mult(d), down = norm.mult(-1.0f * d);
    auto rim = [&](const Vec3& lift, float theta) {
        return orig.add(lift).add(u.mult(std::sin(theta))).add(w.mult(std::cos(theta)));
    };
    for (uint32_t k = 0; k < num_tris; ++k) {
        const float fk = float(k);
        const float theta1 = fk / n * 2.0f * kPi - smudge;
        const float theta2 = (fk + 1.0f) / n * 2.0f * kPi + smudge;
        const float theta3 = (fk + 0.5f) / n * 2.0f * kPi - smudge;
        const float theta4 = (fk + 1.5f) / n * 2.0f * kPi + smudge;
        const Point top_c = orig.add(up), top_a = rim(up, theta1), top_b = rim(up, theta2);
        const Point bot_c = orig.add(down), bot_a = rim(down, theta3), bot_b = rim(down, theta4);
        const Vec3 top[3] = {top_c, top_a, top_b};
        const Vec3 bottom[3] = {bot_c, bot_a, bot_b};
        const Vec3 wall0[3] = {top_a, top_b, bot_a};
        const Vec3 wall1[3] = {bot_a, bot_b, top_b};
        tris.push_back(make_triangle(top, surface, edge_thickness));
        tris.push_back(make_triangle(bottom, surface, edge_thickness));
        tris.push_back(make_triangle(wall0, side_surface, edge_thickness));
        tris.push_back(make_triangle(wall1, side_surface, edge_thickness));
    }
    return tris;
}

// ---- camera (raytrace.rs:1320-1370) --------------------------------------------
Basis create_transform(const Vec3& dir_in, float d_roll) {
    const Vec3 dir = dir_in.unit();
    const float roll = -1.0f * std::atan2(-1.0f * dir.v[1], dir.v[2]);
    const float pitch = -1.0f * std::asin(dir.v[0]);
    const float yaw = -1.0f * d_roll;
    const float cy = std::cos(yaw), sy = std::sin(yaw);
    const float cp = std::cos(pitch), sp = std::sin(pitch);
    const float cr = std::cos(roll), sr = std::sin(roll);
    Basis b;
    b.r0 = make_vec(cy * cp, sy * cp, -1.0f * sp);
    b.r1 = make_vec(cy * sp * sr - sy * cr, sy * sp * sr + cy * cr, cp * sr);
    b.r2 = make_vec(cy * sp * cr + sy * sr, sy * sp * cr - cy * sr, cp * cr);
    return b;
}

Viewport create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const Point& pos, const Vec3& dir,
                         float fov, float c_roll, uint32_t maxdepth, uint32_t samples) {
    const float dist = size0 / (2.0f * std::tan(to_radians(fov) / 2.0f));
    const Basis rot = create_transform(dir, c_roll);
    const Point corner = pos.add(make_vec(1.0f * size1 / 2.0f, -1.0f * size0 / 2.0f, 0.0f));  // not rotated
    const Point eye = pos.sub(change_basis(make_vec(0.0f, 0.0f, dist), rot));
    const Vec3 across = change_basis(make_vec(0.0f, size0, 0.0f), rot);
    const Vec3 down = change_basis(make_vec(-1.0f * size1, 0.0f, 0.0f), rot);
    Viewport v;
    std::memset(&v, 0, sizeof v);
    v.width = px_w;
    v.height = px_h;
    store3(v.orig, corner);
    store3(v.cam, eye);
    store3(v.vu, across);
    store3(v.vv, down);
    v.maxdepth = maxdepth;
    v.spp = samples;
    return v;
}

// ---- OBJ / mesh (obj_parser.rs:20-73) ------------------------------------------
namespace obj_parser {
Mesh read_obj(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw std::runtime_error("read_obj: cannot open " + path);
    Mesh m;
    std::string line;
    while (std::getline(in, line)) {
        if (line.rfind("v ", 0) == 0) {
            const char* p = line.c_str() + 2;
            for (int k = 0; k < 3; ++k) {
                char* end = nullptr;
                const float x = std::strtof(p, &end);  // correctly rounded, as Rust's str::parse::<f32>
                if (end == p) throw std::runtime_error("read_obj: bad vertex line: " + line);
                m.verts.push_back(x);
                p = end;
            }
        } else if (line.rfind("f ", 0) == 0) {
            const char* p = line.c_str() + 2;
            for (int k = 0; k < 3; ++k) {  // only the first three corners are used (obj_parser.rs:64-66)
                while (*p == ' ' || *p == '\t') ++p;
                char* end = nullptr;
                const unsigned long idx = std::strtoul(p, &end, 10);  // text before the first '/'
                if (end == p) throw std::runtime_error("read_obj: bad face line: " + line);
                m.faces.push_back(uint32_t(idx));
                p = end;
                while (*p && *p != ' ' && *p != '\t' && *p != '\r') ++p;
            }
        }
    }
    return m;
}

Mesh read_mesh_bin(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("read_mesh_bin: cannot open " + path);
    char magic[4];
    uint32_t nv = 0, nf = 0;
    in.read(magic, 4);
    in.read(reinterpret_cast<char*>(&nv), 4);
    in.read(reinterpret_cast<char*>(&nf), 4);
    if (!in || std::memcmp(magic, "RTBM", 4) != 0) throw std::runtime_error("read_mesh_bin: bad header in " + path);
    Mesh m;
    m.verts.resize(size_t(nv) * 3);
    m.faces.resize(size_t(nf) * 3);
    in.read(reinterpret_cast<char*>(m.verts.data()), std::streamsize(m.verts.size() * 4));
    in.read(reinterpret_cast<char*>(m.faces.data()), std::streamsize(m.faces.size() * 4));
    if (!in) throw std::runtime_error("read_mesh_bin: truncated " + path);
    return m;
}

std::vector<Triangle> mesh_to_triangles(const Mesh& m, const Vec3& offset, float scale, const Basis& transform,
                                        const SurfaceKind& surface, float edge_thickness) {
    const uint32_t nv = uint32_t(m.verts.size() / 3), nf = uint32_t(m.faces.size() / 3);
    std::vector<Triangle> tris;
    tris.reserve(nf);
    for (uint32_t f = 0; f < nf; ++f) {
        Vec3 pts[3];
        for (int k = 0; k < 3; ++k) {
            const uint32_t vi = m.faces[3 * f + k];
            if (vi < 1 || vi > nv) throw std::runtime_error("mesh_to_triangles: face index out of range");
            const Vec3 raw = load3(&m.verts[3 * size_t(vi - 1)]);
            pts[k] = change_basis(raw.mult(scale), transform).add(offset);
        }
        tris.push_back(make_triangle(pts, surface, edge_thickness));
    }
    return tris;
}

std::vector<Triangle> parse_obj(const std::string& path, const Vec3& offset, float scale, const Basis& transform,
                                const SurfaceKind& surface, float edge_thickness) {
    return mesh_to_triangles(read_obj(path), offset, scale, transform, surface, edge_thickness);
}
}  // namespace obj_parser

// ---- octree-root membership (raytrace.rs:636-779), used for the upload cull ------
namespace {
bool point_in_cube(const Point& c, float h, const Point& p) {
    const Vec3 d = p.sub(c);
    return std::fabs(d.v[0]) < h && std::fabs(d.v[1]) < h && std::fabs(d.v[2]) < h;
}
struct Line { Point orig; Vec3 dir; Vec3 inv; };
Line make_line(const Point& o, const Vec3& d) {
    const Vec3 u = d.unit();
    return {o, u, make_vec(1.0f / u.v[0], 1.0f / u.v[1], 1.0f / u.v[2])};
}
// face_contains_triangle (raytrace.rs:645-729): does the line where the cube-face plane meets the
// triangle plane cross both the face and the triangle?
bool face_cuts_triangle(const Point& c, const Vec3& n1, float h, const Triangle& t) {
    const Vec3 n2 = load3(t.norm);
    const Vec3 tri_c = load3(t.incenter);
    const float h1 = n1.dot(c.add(n1.mult(h)));
    const float h2 = n2.dot(tri_c);
    const float k1 = (h1 - h2 * (n1.dot(n2))) / (1.0f - (n1.dot(n2)) * (n1.dot(n2)));
    const float k2 = (h2 - h1 * (n1.dot(n2))) / (1.0f - (n1.dot(n2)) * (n1.dot(n2)));
    const Line first = make_line(n1.mult(k1).add(n2.mult(k2)), n1.cross(n2));

    const float fmax = std::numeric_limits<float>::max();
    float tmin = fmax;
    for (int ax = 0; ax < 3; ++ax) {
        if (n1.v[ax] != 0.0f) continue;
        const float a = (c.v[ax] - h - first.orig.v[ax]) * first.inv.v[ax];
        const float b = (c.v[ax] + h - first.orig.v[ax]) * first.inv.v[ax];
        tmin = std::fmin(tmin, std::fmin(a, b));
    }
    const Line line = (tmin > 0.0f) ? first : make_line(first.dir.mult(tmin * 2.0f).add(first.orig), first.dir);

    tmin = -fmax;
    float tmax = fmax;
    for (int ax = 0; ax < 3; ++ax) {
        if (n1.v[ax] != 0.0f) continue;
        const float a = (c.v[ax] - h - line.orig.v[ax]) * line.inv.v[ax];
        const float b = (c.v[ax] + h - line.orig.v[ax]) * line.inv.v[ax];
        tmin = std::fmax(tmin, std::fmin(a, b));
        tmax = std::fmin(tmax, std::fmax(a, b));
    }
    if (tmax < tmin) return false;

    Vec3 off[3];
    for (int k = 0; k < 3; ++k) {
        const Vec3 corner = load3(&t.corners[3 * k]);
        const float s = corner.sub(line.orig).dot(line.dir) / line.dir.len2();
        off[k] = line.dir.mult(s).add(line.orig).sub(corner);
    }
    return off[0].dot(off[1]) < 0.0f || off[0].dot(off[2]) < 0.0f || off[1].dot(off[2]) < 0.0f;
}
}  // namespace

bool box_contains_polygon(const Point& orig, float len2, const Triangle& t) {
    if (point_in_cube(orig, len2, load3(t.incenter))) return true;
    for (int k = 0; k < 3; ++k)
        if (point_in_cube(orig, len2, load3(&t.corners[3 * k]))) return true;
    const Vec3 faces[6] = {make_vec(1, 0, 0), make_vec(-1, 0, 0), make_vec(0, 1, 0),
                           make_vec(0, -1, 0), make_vec(0, 0, 1), make_vec(0, 0, -1)};
    for (const Vec3& n : faces)
        if (face_cuts_triangle(orig, n, len2, t)) return true;
    return false;
}

// ---- output quantiser (raytrace.rs:1468-1473) ------------------------------------
void quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb) {
    for (uint64_t i = 0; i < npix; ++i)
        for (int k = 0; k < 3; ++k) {
            const float s = rgba[4 * i + k] * 255.0f;
            uint8_t q = 0;                       // Rust `as u8`: NaN -> 0, saturating, truncating
            if (s >= 255.0f) q = 255;
            else if (s > 0.0f) q = uint8_t(s);
            rgb[3 * i + k] = q;
        }
}
bool write_ppm(const std::string& path, uint32_t width, uint32_t height, const float* rgba) {
    std::vector<uint8_t> rgb(size_t(width) * height * 3);
    quantize_rgb8(rgba, uint64_t(width) * height, rgb.data());
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    std::fprintf(f, "P6\n%u %u\n255\n", width, height);
    const bool ok = std::fwrite(rgb.data(), 1, rgb.size(), f) == rgb.size();
    std::fclose(f);
    return ok;
}


// ---- PNG container (write_png, raytrace.rs:1460-1478) ------------------------------
namespace {
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return crc;
}
void put_be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}
bool write_chunk(FILE* f, const char type[4], const std::vector<uint8_t>& body) {
    std::vector<uint8_t> head;
    put_be32(head, uint32_t(body.size()));
    head.insert(head.end(), type, type + 4);
    uint32_t crc = crc32_update(0xffffffffu, head.data() + 4, 4);
    crc = crc32_update(crc, body.data(), body.size()) ^ 0xffffffffu;
    std::vector<uint8_t> tail;
    put_be32(tail, crc);
    return std::fwrite(head.data(), 1, head.size(), f) == head.size() &&
           (body.empty() || std::fwrite(body.data(), 1, body.size(), f) == body.size()) &&
           std::fwrite(tail.data(), 1, 4, f) == 4;
}
}  // namespace

bool write_png_rgb8(const std::string& path, uint32_t width, uint32_t height, const uint8_t* rgb) {
    if (width == 0 || height == 0) return false;
    // scanlines: filter byte 0 + 3 bytes per pixel
    const size_t stride = size_t(width) * 3, raw_len = (stride + 1) * height;
    std::vector<uint8_t> idat;
    idat.reserve(raw_len + raw_len / 65535 * 5 + 16);
    idat.push_back(0x78); idat.push_back(0x01);                    // zlib header: deflate, 32 K window, no preset
    uint32_t a = 1, b = 0;                                         // Adler-32 of the raw scanlines
    std::vector<uint8_t> raw(raw_len);
    for (uint32_t y = 0; y < height; ++y) {
        raw[(stride + 1) * y] = 0;
        std::memcpy(&raw[(stride + 1) * y + 1], rgb + stride * y, stride);
    }
    for (size_t pos = 0; pos < raw_len;) {
        const size_t n = std::min<size_t>(65535, raw_len - pos);
        idat.push_back(pos + n == raw_len ? 1 : 0);                // BFINAL, BTYPE = 00 (stored)
        idat.push_back(uint8_t(n)); idat.push_back(uint8_t(n >> 8));
        idat.push_back(uint8_t(~n)); idat.push_back(uint8_t((~n) >> 8));
        idat.insert(idat.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n;) {                               // Adler-32, deferred modulo
            const size_t m = std::min<size_t>(5552, n - i);
            for (size_t k = 0; k < m; ++k) { a += raw[pos + i + k]; b += a; }
            a %= 65521u; b %= 65521u;
            i += m;
        }
        pos += n;
    }
    put_be32(idat, (b << 16) | a);
    FILE* f = std::fopen(path.c_str(), "wb");
    if (!f) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width); put_be32(ihdr, height);
    ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);   // 8-bit RGB, no interlace
    const bool ok = std::fwrite(sig, 1, 8, f) == 8 && write_chunk(f, "IHDR", ihdr) && write_chunk(f, "IDAT", idat) &&
                    write_chunk(f, "IEND", {});
    return std::fclose(f) == 0 && ok;
}
bool write_png(const std::string& path, uint32_t width, uint32_t height, const float* rgba) {
    std::vector<uint8_t> rgb(size_t(width) * height * 3);
    quantize_rgb8(rgba, uint64_t(width) * height, rgb.data());
    return write_png_rgb8(path, width, height, rgb.data());
}

}  // namespace raytrace

// ============================================================================
// C ABI (include/rtb_host.h)
// ============================================================================
using namespace raytrace;

namespace {
SurfaceKind surf_from_c(const RtbSurface* s) {
    return {s->kind, make_vec(s->color[0], s->color[1], s->color[2]), s->alpha, s->scattering};
}
Basis basis_from_rows(const float r[9]) {
    return {make_vec(r[0], r[1], r[2]), make_vec(r[3], r[4], r[5]), make_vec(r[6], r[7], r[8])};
}
}  // namespace

extern "C" {

void rtbh_make_color(uint8_t r, uint8_t g, uint8_t b, float out[3]) { store3(out, make_color(r, g, b)); }
void rtbh_unit(const float v[3], float out[3]) { store3(out, load3(v).unit()); }
float rtbh_to_radians(float deg) { return to_radians(deg); }

int rtbh_make_triangle(const float pts[9], const RtbSurface* surface, float edge_thickness, RtbTriangle* out) {
    const Vec3 p[3] = {load3(pts), load3(pts + 3), load3(pts + 6)};
    return try_make_triangle(p, surf_from_c(surface), edge_thickness, out) ? RTB_OK : RTB_ERR_INVALID;
}
int rtbh_make_dummy_triangle(RtbTriangle* out) {
    *out = make_dummy_triangle();
    return RTB_OK;
}
int rtbh_make_disk(const float orig[3], const float norm[3], float r, float d, uint32_t num_tris,
                   const RtbSurface* surface, const RtbSurface* side_surface, float edge_thickness,
                   RtbTriangle* out, uint32_t cap) {
    try {
        const auto v = make_disk(load3(orig), load3(norm), r, d, num_tris, surf_from_c(surface),
                                 surf_from_c(side_surface), edge_thickness);
        if (v.size() > cap) return RTB_ERR_INVALID;
        std::memcpy(out, v.data(), v.size() * sizeof(RtbTriangle));
        return int(v.size());
    } catch (const std::exception&) { return RTB_ERR_INVALID; }
}
int rtbh_make_sphere(const float orig[3], float r, uint32_t lat, uint32_t lon, const RtbSurface* surface,
                     float edge_thickness, RtbTriangle* out, uint32_t cap) {
    try {
        const auto v = make_sphere(load3(orig), r, lat, lon, surf_from_c(surface), edge_thickness);
        if (v.size() > cap) return RTB_ERR_INVALID;
        std::memcpy(out, v.data(), v.size() * sizeof(RtbTriangle));
        return int(v.size());
    } catch (const std::exception&) { return RTB_ERR_INVALID; }
}
void rtbh_create_transform(const float dir[3], float d_roll, float out_rows[9]) {
    const Basis b = create_transform(load3(dir), d_roll);
    store3(out_rows, b.r0);
    store3(out_rows + 3, b.r1);
    store3(out_rows + 6, b.r2);
}
void rtbh_create_viewport(uint32_t px_w, uint32_t px_h, float size0, float size1, const float pos[3],
                          const float dir[3], float fov_deg, float c_roll, uint32_t maxdepth, uint32_t samples,
                          RtbView* out) {
    *out = create_viewport(px_w, px_h, size0, size1, load3(pos), load3(dir), fov_deg, c_roll, maxdepth, samples);
}
int rtbh_parse_obj(const char* path, const float offset[3], float scale, const float transform_rows[9],
                   const RtbSurface* surface, float edge_thickness, RtbTriangle* out, uint32_t cap) {
    try {
        const obj_parser::Mesh m = obj_parser::read_obj(path);
        const uint32_t nf = uint32_t(m.faces.size() / 3);
        if (!out) return int(nf);
        if (nf > cap) return RTB_ERR_INVALID;
        const auto v = obj_parser::mesh_to_triangles(m, load3(offset), scale, basis_from_rows(transform_rows),
                                                     surf_from_c(surface), edge_thickness);
        std::memcpy(out, v.data(), v.size() * sizeof(RtbTriangle));
        return int(v.size());
    } catch (const std::exception&) { return RTB_ERR_INVALID; }
}
int rtbh_mesh_to_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                           const float offset[3], float scale, const float transform_rows[9],
                           const RtbSurface* surface, float edge_thickness, RtbTriangle* out) {
    try {
        obj_parser::Mesh m;
        m.verts.assign(verts, verts + size_t(nverts) * 3);
        m.faces.assign(faces, faces + size_t(nfaces) * 3);
        const auto v = obj_parser::mesh_to_triangles(m, load3(offset), scale, basis_from_rows(transform_rows),
                                                     surf_from_c(surface), edge_thickness);
        std::memcpy(out, v.data(), v.size() * sizeof(RtbTriangle));
        return int(v.size());
    } catch (const std::exception&) { return RTB_ERR_INVALID; }
}
int rtbh_load_mesh_bin(const char* path, float* verts, uint32_t vcap, uint32_t* nverts,
                       uint32_t* faces, uint32_t fcap, uint32_t* nfaces) {
    try {
        const obj_parser::Mesh m = obj_parser::read_mesh_bin(path);
        *nverts = uint32_t(m.verts.size() / 3);
        *nfaces = uint32_t(m.faces.size() / 3);
        if (!verts || !faces) return RTB_OK;
        if (*nverts > vcap || *nfaces > fcap) return RTB_ERR_INVALID;
        std::memcpy(verts, m.verts.data(), m.verts.size() * 4);
        std::memcpy(faces, m.faces.data(), m.faces.size() * 4);
        return RTB_OK;
    } catch (const std::exception&) { return RTB_ERR_INVALID; }
}
int rtbh_box_contains_polygon(const float orig[3], float len2, const RtbTriangle* t) {
    return box_contains_polygon(load3(orig), len2, *t) ? 1 : 0;
}
int rtbh_write_ppm(const char* path, uint32_t width, uint32_t height, const float* rgba) {
    return write_ppm(path, width, height, rgba) ? RTB_OK : RTB_ERR_INVALID;
}
int rtbh_write_png(const char* path, uint32_t width, uint32_t height, const float* rgba) {
    return (path && rgba && write_png(path, width, height, rgba)) ? RTB_OK : RTB_ERR_INVALID;
}
int rtbh_write_png_rgb8(const char* path, uint32_t width, uint32_t height, const uint8_t* rgb) {
    return (path && rgb && write_png_rgb8(path, width, height, rgb)) ? RTB_OK : RTB_ERR_INVALID;
}

}  // extern "C"
