"""ctypes binding of the CPU oracle (oracle/rt_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
product package `rust_raytrace_b200`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "liboracle.so")

OR_SOLID, OR_MATTE, OR_REFLECTIVE = 0, 1, 2
ACCEL_OCTREE, ACCEL_TRIVIAL, ACCEL_BVH = 0, 1, 2

# numpy mirror of OrTriangle (35 x 4 bytes, reference field order raytrace.rs:326-337)
TRI_DTYPE = np.dtype(
    [
        ("incenter", "<f4", (3,)),
        ("norm", "<f4", (3,)),
        ("bounding_r2", "<f4"),
        ("sides", "<f4", (9,)),
        ("side_lens", "<f4", (3,)),
        ("corners", "<f4", (9,)),
        ("edge_thickness", "<f4"),
        ("kind", "<u4"),
        ("color", "<f4", (3,)),
        ("alpha", "<f4"),
        ("scattering", "<f4"),
    ]
)
assert TRI_DTYPE.itemsize == 140
# EXTENSION: analytic sphere, mirror of OrSphere / RtbSphere (40 bytes)
SPH_DTYPE = np.dtype([("center", "<f4", (3,)), ("radius", "<f4"), ("kind", "<u4"), ("color", "<f4", (3,)),
                      ("alpha", "<f4"), ("scattering", "<f4")])
assert SPH_DTYPE.itemsize == 40


class OrView(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("orig", C.c_float * 3),
        ("cam", C.c_float * 3),
        ("vu", C.c_float * 3),
        ("vv", C.c_float * 3),
        ("maxdepth", C.c_uint32),
        ("spp", C.c_uint32),
    ]


class OrStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("box_tests", C.c_uint64),
        ("tri_tests", C.c_uint64),
        ("node_visits", C.c_uint64),
        ("nan_t_hits", C.c_uint64),
        ("seconds", C.c_double),
    ]


class OrTreeStats(C.Structure):
    _fields_ = [
        ("nodes", C.c_uint64),
        ("leaves", C.c_uint64),
        ("leaf_refs", C.c_uint64),
        ("max_leaf", C.c_uint64),
        ("max_depth", C.c_uint64),
        ("leaves_at_maxdepth", C.c_uint64),
    ]


def build(force: bool = False) -> str:
    """Compile oracle/build/liboracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "rt_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    f3 = C.POINTER(C.c_float)
    vp = C.c_void_p
    L.or_make_triangle.argtypes = [f3, C.c_uint32, f3, C.c_float, C.c_float, C.c_float, vp]
    L.or_make_triangle.restype = C.c_int
    L.or_make_dummy_triangle.argtypes = [vp]
    L.or_make_disk.argtypes = [f3, f3, C.c_float, C.c_float, C.c_uint32,
                               C.c_uint32, f3, C.c_float, C.c_float,
                               C.c_uint32, f3, C.c_float, C.c_float, C.c_float, vp]
    L.or_make_disk.restype = C.c_int
    L.or_make_sphere.argtypes = [f3, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32, f3, C.c_float,
                                 C.c_float, C.c_float, vp, C.c_uint32]
    L.or_make_sphere.restype = C.c_int
    L.or_create_transform.argtypes = [f3, C.c_float, f3]
    L.or_unit.argtypes = [f3, f3]
    L.or_to_radians.argtypes = [C.c_float]
    L.or_to_radians.restype = C.c_float
    L.or_make_color.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, f3]
    L.or_create_viewport.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, f3, f3, C.c_float,
                                     C.c_float, C.c_uint32, C.c_uint32, C.POINTER(OrView)]
    L.or_mesh_to_triangles.argtypes = [vp, C.c_uint32, vp, C.c_uint32, f3, C.c_float, f3, C.c_uint32, f3,
                                       C.c_float, C.c_float, C.c_float, vp]
    L.or_mesh_to_triangles.restype = C.c_int
    L.or_parse_obj_file.argtypes = [C.c_char_p, vp, C.c_uint32, C.POINTER(C.c_uint32), vp, C.c_uint32,
                                    C.POINTER(C.c_uint32)]
    L.or_parse_obj_file.restype = C.c_int
    L.or_pixel_ray.argtypes = [C.POINTER(OrView), C.c_uint32, C.c_uint32, f3]
    L.or_triangle_intersects.argtypes = [vp, f3, f3, C.POINTER(C.c_float), f3]
    L.or_triangle_intersects.restype = C.c_int
    L.or_scene_create.argtypes = [vp, C.c_uint32, C.c_int, f3, C.c_float, C.c_uint32, C.c_uint32, C.c_int]
    L.or_scene_create.restype = vp
    L.or_scene_destroy.argtypes = [vp]
    L.or_scene_add_spheres.argtypes = [vp, vp, C.c_uint32]
    L.or_scene_add_spheres.restype = None
    L.or_scene_set_light.argtypes = [vp, f3, C.c_float]
    L.or_scene_set_light.restype = None
    L.or_scene_tree_stats.argtypes = [vp, C.POINTER(OrTreeStats)]
    L.or_scene_closest_hit.argtypes = [vp, f3, f3, C.POINTER(C.c_float)]
    L.or_scene_closest_hit.restype = C.c_uint32
    L.or_render.argtypes = [vp, C.POINTER(OrView), C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp,
                            C.POINTER(OrStats)]
    L.or_render.restype = C.c_int
    L.or_render_samples.argtypes = [vp, C.POINTER(OrView), C.c_uint64, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_int, vp, vp, vp, C.POINTER(OrStats)]
    L.or_render_samples.restype = C.c_int
    L.or_quantize_rgb8.argtypes = [vp, C.c_uint64, vp]
    L.or_selftest_face_collision.restype = C.c_int
    L.or_rng_f32.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    L.or_rng_f32.restype = C.c_float
    _lib = L
    return L


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


def _fn(v, n):
    return (C.c_float * n)(*[float(x) for x in v])


def make_color(r, g, b):
    out = (C.c_float * 3)()
    lib().or_make_color(r, g, b, out)
    return np.array(out[:], dtype=np.float32)


def unit(v):
    out = (C.c_float * 3)()
    lib().or_unit(_f3(v), out)
    return np.array(out[:], dtype=np.float32)


def to_radians(deg):
    return float(lib().or_to_radians(deg))


def create_transform(dir3, roll):
    out = (C.c_float * 9)()
    lib().or_create_transform(_f3(dir3), roll, out)
    return np.array(out[:], dtype=np.float32)


class Surface:
    def __init__(self, kind, color, alpha=0.0, scattering=0.0):
        self.kind, self.color, self.alpha, self.scattering = kind, np.asarray(color, np.float32), alpha, scattering


def make_triangle(pts, surf: Surface, edge):
    out = np.zeros(1, TRI_DTYPE)
    rc = lib().or_make_triangle(_fn(np.asarray(pts, np.float32).ravel(), 9), surf.kind, _f3(surf.color),
                                surf.alpha, surf.scattering, edge, out.ctypes.data)
    if rc != 0:
        raise ValueError("make_triangle: degenerate triangle (reference would panic, raytrace.rs:357)")
    return out


def make_dummy_triangle():
    out = np.zeros(1, TRI_DTYPE)
    lib().or_make_dummy_triangle(out.ctypes.data)
    return out


def make_disk(orig, norm, r, d, n, surf: Surface, side: Surface, edge):
    out = np.zeros(4 * n, TRI_DTYPE)
    rc = lib().or_make_disk(_f3(orig), _f3(norm), r, d, n, surf.kind, _f3(surf.color), surf.alpha, surf.scattering,
                            side.kind, _f3(side.color), side.alpha, side.scattering, edge, out.ctypes.data)
    assert rc == 4 * n, rc
    return out


def make_sphere(orig, r, lat, lon, surf: Surface, edge):
    out = np.zeros(2 * lat * lon, TRI_DTYPE)
    rc = lib().or_make_sphere(_f3(orig), r, lat, lon, surf.kind, _f3(surf.color), surf.alpha, surf.scattering,
                              edge, out.ctypes.data, len(out))
    assert rc >= 0, rc
    return out[:rc].copy()


def parse_obj_file(path):
    cap = 1 << 20
    verts = np.zeros((cap, 3), np.float32)
    faces = np.zeros((cap, 3), np.uint32)
    nv, nf = C.c_uint32(), C.c_uint32()
    rc = lib().or_parse_obj_file(path.encode(), verts.ctypes.data, cap, C.byref(nv), faces.ctypes.data, cap,
                                 C.byref(nf))
    assert rc == 0, rc
    return verts[: nv.value].copy(), faces[: nf.value].copy()


def mesh_to_triangles(verts, faces, offset, scale, transform, surf: Surface, edge):
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    out = np.zeros(len(faces), TRI_DTYPE)
    rc = lib().or_mesh_to_triangles(verts.ctypes.data, len(verts), faces.ctypes.data, len(faces), _f3(offset),
                                    scale, _fn(transform, 9), surf.kind, _f3(surf.color), surf.alpha,
                                    surf.scattering, edge, out.ctypes.data)
    assert rc == len(faces), rc
    return out


def create_viewport(px, size, pos, dir3, fov, c_roll, maxdepth, samples) -> OrView:
    v = OrView()
    lib().or_create_viewport(px[0], px[1], size[0], size[1], _f3(pos), _f3(dir3), fov, c_roll, maxdepth, samples,
                             C.byref(v))
    return v


def pixel_ray(v: OrView, row, col):
    out = (C.c_float * 9)()
    lib().or_pixel_ray(C.byref(v), row, col, out)
    return np.array(out[:], dtype=np.float32)


class Scene:
    def __init__(self, tris, accel=ACCEL_OCTREE, root_orig=(0.0, 0.0, 20.1), root_len2=20.0, maxdepth=10,
                 minobjs=19, build_threads=8):
        self.tris = np.ascontiguousarray(tris, TRI_DTYPE)
        self.h = lib().or_scene_create(self.tris.ctypes.data, len(self.tris), accel, _f3(root_orig), root_len2,
                                       maxdepth, minobjs, build_threads)
        self.accel = accel

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_scene_destroy(self.h)
            self.h = None

    def add_spheres(self, spheres):
        """EXTENSION: analytic spheres, primitive ids len(tris) + j."""
        spheres = np.ascontiguousarray(spheres, SPH_DTYPE)
        lib().or_scene_add_spheres(self.h, spheres.ctypes.data, len(spheres))
        return self

    def set_light(self, orig, len2):
        """EXTENSION: Scene.lights = Some(LightSource{orig, len2}) (raytrace.rs:594-610, :1203-1224); None removes it."""
        lib().or_scene_set_light(self.h, None if orig is None else _f3(orig), float(len2))
        return self

    def tree_stats(self) -> OrTreeStats:
        st = OrTreeStats()
        lib().or_scene_tree_stats(self.h, C.byref(st))
        return st

    def closest_hit(self, orig, dir3):
        t = C.c_float()
        idx = lib().or_scene_closest_hit(self.h, _f3(orig), _f3(dir3), C.byref(t))
        return idx, t.value

    def render_samples(self, v: OrView, s0, s1, seed=0, threads=None, sum_only=True):
        """Samples [s0, s1) of v.spp only; with sum_only the un-normalised sum (multi-GPU sample partition checker)."""
        threads = threads or os.cpu_count() or 1
        H, W = v.height, v.width
        rgba = np.zeros((H, W, 4), np.float32)
        st = OrStats()
        lib().or_render_samples(self.h, C.byref(v), seed, threads, 0, H, int(s0), int(s1), 1 if sum_only else 0,
                                rgba.ctypes.data, None, None, C.byref(st))
        return rgba, st

    def render(self, v: OrView, seed=0, threads=None, rows=None, want_ids=True):
        """DefaultRayCaster.walk_rays equivalent -> (rgba[H,W,4], prim[H,W], t[H,W], OrStats)."""
        threads = threads or os.cpu_count() or 1
        H, W = v.height, v.width
        rgba = np.zeros((H, W, 4), np.float32)
        prim = np.zeros((H, W), np.uint32) if want_ids else None
        tt = np.zeros((H, W), np.float32) if want_ids else None
        st = OrStats()
        r0, r1 = rows if rows else (0, H)
        lib().or_render(self.h, C.byref(v), seed, threads, r0, r1, rgba.ctypes.data,
                        prim.ctypes.data if want_ids else None, tt.ctypes.data if want_ids else None, C.byref(st))
        return rgba, prim, tt, st


def quantize_rgb8(rgba):
    rgba = np.ascontiguousarray(rgba, np.float32)
    n = rgba.size // 4
    out = np.zeros((n, 3), np.uint8)
    lib().or_quantize_rgb8(rgba.ctypes.data, n, out.ctypes.data)
    return out.reshape(rgba.shape[:-1] + (3,))


# ---------------------------------------------------------------------------
# The reference's benchmark scene and camera, raytrace/src/main.rs:116-173.
# ---------------------------------------------------------------------------
def main_scene_tris(verts, faces, deterministic=False):
    """tris of main.rs:116-152.  deterministic=True swaps in the materials of the
    deterministic parity mode: teapot Solid(252,119,0) (main.rs:123),
    disks Reflective with scattering 0, disk sides Solid."""
    orange = make_color(252, 119, 0)
    grey = make_color(230, 230, 230)
    dark = make_color(40, 40, 40)
    if deterministic:
        teapot = Surface(OR_SOLID, orange)
        d1 = Surface(OR_REFLECTIVE, grey, 0.7, 0.0)
        d2 = Surface(OR_REFLECTIVE, grey, 0.7, 0.0)
        side = Surface(OR_SOLID, dark)
    else:
        teapot = Surface(OR_MATTE, orange, 0.2)
        d1 = Surface(OR_REFLECTIVE, grey, 0.7, 0.0002)
        d2 = Surface(OR_REFLECTIVE, grey, 0.7, 0.002)
        side = Surface(OR_MATTE, dark, 0.2)
    parts = [make_dummy_triangle()]
    tf = create_transform(unit([0.0, 0.3, 1.0]), to_radians(270.0))
    parts.append(mesh_to_triangles(verts, faces, [0.0, 0.5, 5.0], 1.0, tf, teapot, 0.05))
    parts.append(make_disk([4.0, 4.0, 7.0], unit([-0.3, -0.55, -0.5]), 2.0, 0.1, 50, d1, side, -1.0))
    parts.append(make_disk([4.0, -3.0, 5.0], unit([-0.5, 2.0, -0.5]), 1.0, 0.04, 50, d2, side, -1.0))
    return np.concatenate(parts)


def load_mesh_bin(path=None):
    """Reader of the committed mesh fixture (rust_raytrace_b200/data/teapot_mesh.bin = raytrace/teapot_tri.obj parsed by
    tests/golden/make_fixtures.py): b"RTBM", u32 n_verts, u32 n_faces, f32 verts[n][3], u32 faces[n][3] (1-based)."""
    import struct
    if path is None:
        path = os.path.join(os.path.dirname(_HERE), "rust_raytrace_b200", "data", "teapot_mesh.bin")
    with open(path, "rb") as fh:
        raw = fh.read()
    assert raw[:4] == b"RTBM"
    nv, nf = struct.unpack_from("<II", raw, 4)
    verts = np.frombuffer(raw, "<f4", nv * 3, 12).reshape(nv, 3).copy()
    faces = np.frombuffer(raw, "<u4", nf * 3, 12 + nv * 12).reshape(nf, 3).copy()
    return verts, faces


def teapot_field_tris(verts, faces, nz=12, ny=13, seed=1, scale=0.3):
    """BASELINE config 4: nz x ny teapots on a grid receding from main.rs's camera, every instance with
    its own roll angle from a Numerical-Recipes LCG, `Reflective{scattering: 0}` surfaces.  Same definition as the product
    package's teapot_field_scene (tests/test_host.py compares the bytes)."""
    surf = Surface(OR_REFLECTIVE, make_color(252, 119, 0), 0.5, 0.0)
    parts = [make_dummy_triangle()]
    state = int(seed) & 0xFFFFFFFF
    for iz in range(nz):
        for iy in range(ny):
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF
            roll = 270.0 + 360.0 * (state >> 8) / float(1 << 24)
            tf = create_transform(unit([0.0, 0.3, 1.0]), to_radians(roll))
            off = [-1.2, (iy - (ny - 1) / 2.0) * 2.2, 3.0 + 2.0 * iz]
            parts.append(mesh_to_triangles(verts, faces, off, scale, tf, surf, 0.05))
    return np.concatenate(parts)


def circles_scene_parts(n=64, seed=1, light=((6.0, -1.0, 2.0), 0.6)):
    """BASELINE config 1 as this build defines it (EXTENSION): (tris, spheres, light) — n analytic
    spheres of random colour (one in eight a mirror, one in eight matte) over a matte ground disk, one cube light.  Same
    definition as the product package's circles_scene (tests/test_host.py compares the bytes)."""
    rng = np.random.RandomState(seed)
    sph = np.zeros(n, SPH_DTYPE)
    for k in range(n):
        c = [float(rng.uniform(-0.5, 3.5)), float(rng.uniform(-5.0, 5.0)), float(rng.uniform(4.0, 14.0))]
        col = make_color(*[int(x) for x in rng.randint(30, 255, 3)])
        kind, alpha = (OR_REFLECTIVE, 0.6) if k % 8 == 0 else ((OR_MATTE, 0.3) if k % 8 == 1 else (OR_SOLID, 0.0))
        sph[k]["center"] = c
        sph[k]["radius"] = float(rng.uniform(0.25, 0.7))
        sph[k]["kind"], sph[k]["alpha"], sph[k]["scattering"] = kind, alpha, 0.0
        sph[k]["color"] = col
    ground = make_disk([-1.5, 0.0, 9.0], unit([1.0, 0.0, 0.05]), 9.0, 0.1, 40,
                       Surface(OR_MATTE, make_color(150, 150, 150), 0.25), Surface(OR_SOLID, [0.1, 0.1, 0.1]), -1.0)
    return np.concatenate([make_dummy_triangle(), ground]), sph, light


def main_viewport(width, height, maxdepth=5, spp=1) -> OrView:
    """main.rs:166-173 with aspect = height/width (main.rs:96-110)."""
    aspect = np.float32(height) / np.float32(width)
    return create_viewport((width, height), (1.0, float(np.float32(1.0) * aspect)), [2.0, 0.0, 0.0],
                           unit([0.0, 0.0, 1.0]), 90.0, to_radians(0.0), maxdepth, spp)
