"""Host-side mirror of the reference's API surface for the accelerated path.

Names, argument order and meaning follow raytrace_lib/src/raytrace.rs so that code
written against the reference (`main.rs`) reads the same here:

    tris  = [make_dummy_triangle()] + obj_parser.parse_obj(...) + make_disk(...)
    scene = Scene(tris, boxes=(root_orig, root_len2))
    v     = create_viewport((w, h), (1., aspect), pos, dir, fov, roll, maxdepth, spp)
    data  = new_image(v)
    ctx   = B200RayCaster().walk_rays(v, scene, data, threads=n_gpus, show_progress=False)
    ctx.print_stats()

Everything numeric is done by librtb.so (host helpers in C++, hot path in CUDA);
this module only marshals.  There is no CPU rendering path here.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import (RTB_FLAG_STATS, RTB_FLAG_SUM_ONLY, RTB_MATTE, RTB_REFLECTIVE, RTB_SOLID, SPH_DTYPE, TRI_DTYPE, RtbMeshInstance,
                   RtbSceneInfo, RtbStats, RtbSurface, RtbView, check, lib)

_PKG = os.path.dirname(os.path.abspath(__file__))
TEAPOT_MESH = os.path.join(_PKG, "data", "teapot_mesh.bin")


def _f(v, n=3):
    return (C.c_float * n)(*[float(x) for x in v])


# ---------------------------------------------------------------------------
# Vec3 / colours (raytrace.rs:22-192)
# ---------------------------------------------------------------------------
def make_vec(v):
    return np.asarray(v, np.float32).reshape(3)


def make_color(color):
    out = (C.c_float * 3)()
    lib().rtbh_make_color(int(color[0]), int(color[1]), int(color[2]), out)
    return np.array(out[:], np.float32)


def unit(v):
    out = (C.c_float * 3)()
    lib().rtbh_unit(_f(v), out)
    return np.array(out[:], np.float32)


def to_radians(deg):
    return float(lib().rtbh_to_radians(float(deg)))


# ---------------------------------------------------------------------------
# SurfaceKind (raytrace.rs:303-308)
# ---------------------------------------------------------------------------
@dataclass
class SurfaceKind:
    kind: int
    color: np.ndarray
    alpha: float = 0.0
    scattering: float = 0.0

    @staticmethod
    def Solid(color):
        return SurfaceKind(RTB_SOLID, make_vec(color))

    @staticmethod
    def Matte(color, alpha):
        return SurfaceKind(RTB_MATTE, make_vec(color), float(alpha))

    @staticmethod
    def Reflective(scattering, color, alpha):
        return SurfaceKind(RTB_REFLECTIVE, make_vec(color), float(alpha), float(scattering))

    def _c(self):
        return RtbSurface(self.kind, _f(self.color), self.alpha, self.scattering)


# ---------------------------------------------------------------------------
# Triangles and generators (raytrace.rs:326-592); arrays of TRI_DTYPE
# ---------------------------------------------------------------------------
def make_triangle(points, surface: SurfaceKind, edge_thickness: float):
    out = np.zeros(1, TRI_DTYPE)
    s = surface._c()
    rc = lib().rtbh_make_triangle(_f(np.asarray(points, np.float32).ravel(), 9), C.byref(s), edge_thickness,
                                  out.ctypes.data)
    if rc != 0:
        raise ValueError("make_triangle: degenerate triangle (the reference panics here, raytrace.rs:357)")
    return out


def make_dummy_triangle():
    out = np.zeros(1, TRI_DTYPE)
    lib().rtbh_make_dummy_triangle(out.ctypes.data)
    return out


def populate_triangle_numbers(tris):
    """`num` is the array index in this ABI (raytrace.rs:393-397); kept for source compatibility."""
    return tris


def make_disk(orig, norm, r, d, num_tris, surface: SurfaceKind, side_surface: SurfaceKind, edge_thickness):
    out = np.zeros(4 * num_tris, TRI_DTYPE)
    s, ss = surface._c(), side_surface._c()
    rc = lib().rtbh_make_disk(_f(orig), _f(norm), r, d, num_tris, C.byref(s), C.byref(ss), edge_thickness,
                              out.ctypes.data, len(out))
    if rc < 0:
        raise ValueError("make_disk failed")
    return out[:rc]


def make_sphere(orig, r, lat_lon, surface: SurfaceKind, edge_thickness):
    lat, lon = lat_lon
    out = np.zeros(2 * lat * lon, TRI_DTYPE)
    s = surface._c()
    rc = lib().rtbh_make_sphere(_f(orig), r, lat, lon, C.byref(s), edge_thickness, out.ctypes.data, len(out))
    if rc < 0:
        raise ValueError("make_sphere failed (lat must be even; triangles must not be degenerate)")
    return out[:rc].copy()


def create_transform(dir_in, d_roll):
    out = (C.c_float * 9)()
    lib().rtbh_create_transform(_f(dir_in), float(d_roll), out)
    return np.array(out[:], np.float32)


class obj_parser:
    """obj_parser.rs:47-73"""

    @staticmethod
    def parse_obj(path, offset, scale, transform, surface: SurfaceKind, edge_thickness):
        s = surface._c()
        n = lib().rtbh_parse_obj(path.encode(), _f(offset), scale, _f(transform, 9), C.byref(s), edge_thickness,
                                 None, 0)
        if n < 0:
            raise ValueError(f"parse_obj: cannot read {path}")
        out = np.zeros(n, TRI_DTYPE)
        rc = lib().rtbh_parse_obj(path.encode(), _f(offset), scale, _f(transform, 9), C.byref(s), edge_thickness,
                                  out.ctypes.data, n)
        if rc != n:
            raise ValueError("parse_obj failed (degenerate face?)")
        return out

    @staticmethod
    def load_mesh_bin(path=TEAPOT_MESH):
        nv, nf = C.c_uint32(), C.c_uint32()
        check(lib().rtbh_load_mesh_bin(path.encode(), None, 0, C.byref(nv), None, 0, C.byref(nf)), "load_mesh_bin")
        verts = np.zeros((nv.value, 3), np.float32)
        faces = np.zeros((nf.value, 3), np.uint32)
        check(lib().rtbh_load_mesh_bin(path.encode(), verts.ctypes.data, nv.value, C.byref(nv), faces.ctypes.data,
                                       nf.value, C.byref(nf)), "load_mesh_bin")
        return verts, faces

    @staticmethod
    def mesh_to_triangles(verts, faces, offset, scale, transform, surface: SurfaceKind, edge_thickness):
        verts = np.ascontiguousarray(verts, np.float32)
        faces = np.ascontiguousarray(faces, np.uint32)
        out = np.zeros(len(faces), TRI_DTYPE)
        s = surface._c()
        rc = lib().rtbh_mesh_to_triangles(verts.ctypes.data, len(verts), faces.ctypes.data, len(faces), _f(offset),
                                          scale, _f(transform, 9), C.byref(s), edge_thickness, out.ctypes.data)
        if rc != len(faces):
            raise ValueError("mesh_to_triangles failed (degenerate face or bad index)")
        return out


# ---------------------------------------------------------------------------
# Viewport (raytrace.rs:1305-1370)
# ---------------------------------------------------------------------------
Viewport = RtbView


def create_viewport(px, size, pos, dir3, fov, c_roll, maxdepth, samples) -> RtbView:
    v = RtbView()
    lib().rtbh_create_viewport(int(px[0]), int(px[1]), float(size[0]), float(size[1]), _f(pos), _f(dir3), float(fov),
                               float(c_roll), int(maxdepth), int(samples), C.byref(v))
    return v


def new_image(v: RtbView):
    """`vec![make_vec(&[0.,0.,0.]); width*height]` (main.rs:190): H x W x 4 f32, lane 3 = 0."""
    return np.zeros((v.height, v.width, 4), np.float32)


# ---------------------------------------------------------------------------
# Scene (raytrace.rs:1297-1303)
# ---------------------------------------------------------------------------
class Scene:
    """`Scene{tris, boxes, ..}`.  `boxes` keeps only what the GPU path needs of the octree:
    the root cube (orig, len2) used for the reference's visibility cull; None disables it."""

    def __init__(self, tris, boxes=((0.0, 0.0, 20.1), 20.0), spheres=None, light=None):
        self.tris = np.ascontiguousarray(tris, TRI_DTYPE)
        self.boxes = boxes
        self._h = None
        # EXTENSION (include/rtb.h): analytic spheres (ids len(tris) + j) and `lights: Option<LightSource>` = (orig, len2)
        self.spheres = None if spheres is None else np.ascontiguousarray(spheres, SPH_DTYPE)
        self.light = light

    # device residency -----------------------------------------------------
    def upload(self):
        if self._h is None:
            h = C.c_void_p()
            if self.boxes is None:
                ro, rl = None, 0.0
            else:
                ro, rl = _f(self.boxes[0]), float(self.boxes[1])
            if self.spheres is not None and len(self.spheres):
                check(lib().rtb_scene_create_ext(self.tris.ctypes.data, len(self.tris), self.spheres.ctypes.data,
                                                 len(self.spheres), ro, rl, C.byref(h)), "rtb_scene_create_ext")
            else:
                check(lib().rtb_scene_create(self.tris.ctypes.data, len(self.tris), ro, rl, C.byref(h)), "rtb_scene_create")
            self._h = h
            if self.light is not None:
                self.set_light(*self.light)
        return self._h

    def set_light(self, orig, len2=0.0):
        """`lights = Some(LightSource{orig, len2})` (raytrace.rs:594-597); orig None removes it."""
        self.light = None if orig is None else (tuple(float(x) for x in orig), float(len2))
        if self._h is not None:
            check(lib().rtb_scene_set_light(self._h, None if orig is None else _f(orig), float(len2)), "rtb_scene_set_light")

    def info(self) -> RtbSceneInfo:
        out = RtbSceneInfo()
        check(lib().rtb_scene_info(self.upload(), C.byref(out)), "rtb_scene_info")
        return out

    def download_bvh(self):
        inf = self.info()
        nodes = np.zeros((inf.n_nodes, 8), np.float32)
        order = np.zeros(inf.n_refs, np.uint32)
        check(lib().rtb_scene_download_bvh(self._h, nodes.ctypes.data, order.ctypes.data), "rtb_scene_download_bvh")
        return nodes, order

    def release(self):
        if self._h is not None:
            lib().rtb_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def analytic_sphere(orig, r, surface: SurfaceKind):
    """EXTENSION: one analytic sphere record (SPH_DTYPE) for Scene(spheres=...)."""
    c = surface._c()
    out = np.zeros(1, SPH_DTYPE)
    out["center"][0] = [float(x) for x in orig]
    out["radius"], out["kind"], out["alpha"], out["scattering"] = float(r), c.kind, c.alpha, c.scattering
    out["color"][0] = c.color[:]
    return out


def circles_scene(n=64, seed=1, light=((6.0, -1.0, 2.0), 0.6)) -> "Scene":
    """BASELINE config 1 ("circles scene at 2K, primary + shadow rays"), defined by this build because the mounted
    reference has none: n analytic spheres of random colour (one in eight a mirror, one in eight matte)
    over a matte ground disk, one cube light."""
    rng = np.random.RandomState(seed)
    parts = []
    for k in range(n):
        c = [float(rng.uniform(-0.5, 3.5)), float(rng.uniform(-5.0, 5.0)), float(rng.uniform(4.0, 14.0))]
        col = make_color(tuple(int(x) for x in rng.randint(30, 255, 3)))
        surf = (SurfaceKind.Reflective(0.0, col, 0.6) if k % 8 == 0 else
                SurfaceKind.Matte(col, 0.3) if k % 8 == 1 else SurfaceKind.Solid(col))
        parts.append(analytic_sphere(c, float(rng.uniform(0.25, 0.7)), surf))
    ground = make_disk([-1.5, 0.0, 9.0], unit([1.0, 0.0, 0.05]), 9.0, 0.1, 40,
                       SurfaceKind.Matte(make_color((150, 150, 150)), 0.25), SurfaceKind.Solid([0.1, 0.1, 0.1]), -1.0)
    return Scene(np.concatenate([make_dummy_triangle(), ground]), boxes=((0.0, 0.0, 20.1), 20.0),
                 spheres=np.concatenate(parts), light=light)


# ---------------------------------------------------------------------------
# Scene assembly on the GPU: parse_obj's per-face work + make_triangle as a kernel
# ---------------------------------------------------------------------------
def mesh_instance(offset, scale, transform, surface: SurfaceKind, edge_thickness) -> RtbMeshInstance:
    """One `parse_obj(path, offset, scale, transform, surface, edge_thickness)` call (obj_parser.rs:47-52) minus the file."""
    c = surface._c()
    m = RtbMeshInstance()
    m.transform_rows[:] = [float(x) for x in np.asarray(transform, np.float32).reshape(9)]
    m.offset[:] = [float(x) for x in offset]
    m.scale, m.edge_thickness = float(scale), float(edge_thickness)
    m.kind, m.alpha, m.scattering = c.kind, c.alpha, c.scattering
    m.color[:] = c.color[:]
    return m


def _inst_array(instances):
    arr = (RtbMeshInstance * max(len(instances), 1))()
    for k, m in enumerate(instances):
        arr[k] = m
    return arr


def assemble_triangles(verts, faces, instances):
    """`Triangle` records of every (instance, face), computed on the GPU (rtb_assemble_triangles)."""
    verts = np.ascontiguousarray(verts, np.float32)
    faces = np.ascontiguousarray(faces, np.uint32)
    out = np.zeros(len(faces) * len(instances), TRI_DTYPE)
    arr = _inst_array(instances)       # keep the ctypes array alive across the call
    check(lib().rtb_assemble_triangles(verts.ctypes.data, len(verts), faces.ctypes.data, len(faces),
                                       C.addressof(arr), len(instances), out.ctypes.data), "rtb_assemble_triangles")
    return out


def cull_triangles(tris, boxes):
    """Indices of the triangles the octree root cube keeps (box_contains_polygon, raytrace.rs:753-779), on the GPU."""
    tris = np.ascontiguousarray(tris, TRI_DTYPE)
    keep = np.zeros(len(tris), np.uint32)
    n = C.c_uint32()
    ro, rl = (None, 0.0) if boxes is None else (_f(boxes[0]), float(boxes[1]))
    check(lib().rtb_cull_triangles(tris.ctypes.data, len(tris), ro, rl, keep.ctypes.data, C.byref(n)), "rtb_cull_triangles")
    return keep[:n.value]


class InstancedScene(Scene):
    """A scene given as one indexed mesh + instances (+ ready-made extra triangles): the `Triangle` array
    [dummy, instance 0, instance 1, ..., extra] is built on the GPU and never exists on the host unless `.tris` is read."""

    def __init__(self, verts, faces, instances, extra=None, boxes=((0.0, 0.0, 20.1), 20.0)):
        self.verts = np.ascontiguousarray(verts, np.float32)
        self.faces = np.ascontiguousarray(faces, np.uint32)
        self.instances = list(instances)
        self.extra = np.zeros(0, TRI_DTYPE) if extra is None else np.ascontiguousarray(extra, TRI_DTYPE)
        self.boxes = boxes
        self._h = None
        self._tris = None

    @property
    def tris(self):
        if self._tris is None:
            self._tris = np.concatenate([make_dummy_triangle(), assemble_triangles(self.verts, self.faces, self.instances),
                                         self.extra])
        return self._tris

    def upload(self):
        if self._h is None:
            h = C.c_void_p()
            ro, rl = (None, 0.0) if self.boxes is None else (_f(self.boxes[0]), float(self.boxes[1]))
            arr = _inst_array(self.instances)      # keep the ctypes array alive across the call
            check(lib().rtb_scene_create_instanced(self.verts.ctypes.data, len(self.verts), self.faces.ctypes.data,
                                                   len(self.faces), C.addressof(arr),
                                                   len(self.instances), self.extra.ctypes.data if len(self.extra) else None,
                                                   len(self.extra), ro, rl, C.byref(h)), "rtb_scene_create_instanced")
            self._h = h
        return self._h


# ---------------------------------------------------------------------------
# progress.rs:9-185 (the part a caster must feed)
# ---------------------------------------------------------------------------
@dataclass
class ProgressCtx:
    width: int
    height: int
    num_threads: int
    start_time: float = field(default_factory=time.perf_counter)
    stop_time: float = 0.0
    total_rays: int = 0
    finished_pixels: int = 0
    runtimes: dict = field(default_factory=dict)

    def update(self, tnum, row, pixels, runstats):
        self.finished_pixels += pixels
        for k, v in runstats.items():
            if k == "Rays":
                self.total_rays += v
            self.runtimes[k] = self.runtimes.get(k, 0) + v

    def finish(self):
        self.stop_time = time.perf_counter()

    def seconds(self):
        return self.stop_time - self.start_time

    def print_stats(self):
        secs = self.seconds()
        print("Processed {:.3f} million rays in {:.3f} seconds. {:.3f} million rays/s".format(
            self.total_rays / 1e6, secs, self.total_rays / secs / 1e6 if secs > 0 else 0.0))
        for k in sorted(self.runtimes):
            print(f"{k}: {self.runtimes[k]}")


# ---------------------------------------------------------------------------
# RayCaster (raytrace.rs:1128-1165) — the plugin seam
# ---------------------------------------------------------------------------
class B200RayCaster:
    """Sibling of DefaultRayCaster / CudaRayCaster (main.rs:183-200).

    walk_rays(v, s, data, threads, show_progress): `threads` is reinterpreted as the number of GPUs
    (0 = all visible).  `data` is the H*W*4 f32 image (`&mut [Color]`) and is filled in place.
    Extra outputs of the last call (primitive ids, hit times, RtbStats) are kept on the instance.
    """

    def __init__(self, want_ids=False, seed=0, stats=False):
        self.want_ids = want_ids
        self.seed = seed
        self.stats_flag = stats
        self.prim = None
        self.t = None
        self.stats = None
        self._n_gpus = None

    def _init(self, threads):
        """`threads` (main.rs:83 passes 16) as a GPU count: 0 = all visible devices, more than exist is clamped."""
        visible = check(lib().rtb_visible_device_count(), "rtb_visible_device_count")
        n = min(int(threads) if int(threads) > 0 else visible, visible, _lib.RTB_MAX_GPUS)
        if self._n_gpus != n:
            check(lib().rtb_init(n, None), "rtb_init")
            self._n_gpus = n

    def walk_rays_internal(self, v: RtbView, s: Scene, data, threads, progress_tx):
        self._init(threads)
        h = s.upload()
        assert data.dtype == np.float32 and data.size == v.width * v.height * 4 and data.flags["C_CONTIGUOUS"]
        vv = RtbView.from_buffer_copy(v)
        vv.seed = self.seed
        if self.stats_flag:
            vv.flags |= RTB_FLAG_STATS
        prim_p = t_p = None
        if self.want_ids:
            self.prim = np.zeros((v.height, v.width), np.uint32)
            self.t = np.zeros((v.height, v.width), np.float32)
            prim_p, t_p = self.prim.ctypes.data, self.t.ctypes.data
        st = RtbStats()
        check(lib().rtb_render(h, C.byref(vv), data.ctypes.data, prim_p, t_p, C.byref(st)), "rtb_render")
        self.stats = st
        progress_tx((0, v.height - 1, v.width * v.height,
                     {"Rays": int(st.rays), "GPU ms": float(st.ms_render), "GPU launches": int(st.kernel_launches)}))

    def walk_rays(self, v: RtbView, s: Scene, data, threads=1, show_progress=False) -> ProgressCtx:
        ctx = ProgressCtx(v.width, v.height, threads)
        self.walk_rays_internal(v, s, data, threads, lambda msg: ctx.update(*msg))
        ctx.finish()
        return ctx

    def walk_rays_rgb8(self, v: RtbView, s: Scene, rgb, threads=1) -> ProgressCtx:
        """walk_rays followed by write_png's quantiser, fused on the device (rtb_render_rgb8): `rgb` is H x W x 3 uint8."""
        assert rgb.dtype == np.uint8 and rgb.size == v.width * v.height * 3 and rgb.flags["C_CONTIGUOUS"]
        ctx = ProgressCtx(v.width, v.height, threads)
        self._init(threads)
        h = s.upload()
        vv = RtbView.from_buffer_copy(v)
        vv.seed = self.seed
        st = RtbStats()
        check(lib().rtb_render_rgb8(h, C.byref(vv), rgb.ctypes.data, C.byref(st)), "rtb_render_rgb8")
        self.stats = st
        ctx.update(0, v.height - 1, v.width * v.height, {"Rays": int(st.rays), "GPU ms": float(st.ms_render)})
        ctx.finish()
        return ctx

    def walk_rays_progressive(self, v: RtbView, s: Scene, data, threads=0) -> ProgressCtx:
        """Multi-sample frame with samples partitioned over the GPUs (rtb_render_progressive)."""
        ctx = ProgressCtx(v.width, v.height, threads)
        self._init(threads)
        h = s.upload()
        vv = RtbView.from_buffer_copy(v)
        vv.seed = self.seed
        st = RtbStats()
        check(lib().rtb_render_progressive(h, C.byref(vv), data.ctypes.data, C.byref(st)), "rtb_render_progressive")
        self.stats = st
        ctx.update(0, v.height - 1, v.width * v.height, {"Rays": int(st.rays), "GPU ms": float(st.ms_render)})
        ctx.finish()
        return ctx


def write_png(path, img_size, data):
    """write_png (raytrace.rs:1460-1478): 8-bit RGB PNG.  `data` is the f32 RGBA frame (quantised with the reference's
    `(c*255.) as u8`) or an already quantised H x W x 3 uint8 frame (walk_rays_rgb8)."""
    w, h = (img_size.width, img_size.height) if hasattr(img_size, "width") else img_size
    data = np.asarray(data)
    if data.dtype == np.uint8:
        data = np.ascontiguousarray(data)
        assert data.size == w * h * 3
        check(lib().rtbh_write_png_rgb8(os.fsencode(path), w, h, data.ctypes.data), "rtbh_write_png_rgb8")
    else:
        data = np.ascontiguousarray(data, np.float32)
        assert data.size == w * h * 4
        check(lib().rtbh_write_png(os.fsencode(path), w, h, data.ctypes.data), "rtbh_write_png")


def write_ppm(path, v_or_size, data):
    """write_png's quantiser (raytrace.rs:1460-1478) with a PPM container."""
    w, h = (v_or_size.width, v_or_size.height) if hasattr(v_or_size, "width") else v_or_size
    data = np.ascontiguousarray(data, np.float32)
    check(lib().rtbh_write_ppm(path.encode(), w, h, data.ctypes.data), "rtbh_write_ppm")


def quantize_rgb8(data):
    """(c*255.) as u8 on the GPU (rtb_quantize_rgb8)."""
    data = np.ascontiguousarray(data, np.float32)
    n = data.size // 4
    out = np.zeros((n, 3), np.uint8)
    check(lib().rtb_quantize_rgb8(data.ctypes.data, n, out.ctypes.data), "rtb_quantize_rgb8")
    return out.reshape(data.shape[:-1] + (3,))


# ---------------------------------------------------------------------------
# The reference's benchmark scene and camera (raytrace/src/main.rs:116-173)
# ---------------------------------------------------------------------------
def main_scene(deterministic=False, mesh_path=TEAPOT_MESH, instanced=False) -> Scene:
    """Scene of main.rs:116-164.  deterministic=True uses the materials of the deterministic
    parity mode: teapot Solid(252,119,0) (the commented line main.rs:123), disks Reflective with
    scattering 0, disk sides Solid."""
    orange, grey, dark = make_color((252, 119, 0)), make_color((230, 230, 230)), make_color((40, 40, 40))
    if deterministic:
        teapot = SurfaceKind.Solid(orange)
        d1 = d2 = SurfaceKind.Reflective(0.0, grey, 0.7)
        side = SurfaceKind.Solid(dark)
    else:
        teapot = SurfaceKind.Matte(orange, 0.2)
        d1 = SurfaceKind.Reflective(0.0002, grey, 0.7)
        d2 = SurfaceKind.Reflective(0.002, grey, 0.7)
        side = SurfaceKind.Matte(dark, 0.2)
    tf = create_transform(unit([0.0, 0.3, 1.0]), to_radians(270.0))
    if instanced:      # the teapot's `Triangle`s are computed on the GPU (InstancedScene)
        verts, faces = obj_parser.load_mesh_bin(mesh_path)
        disks = np.concatenate([make_disk([4.0, 4.0, 7.0], unit([-0.3, -0.55, -0.5]), 2.0, 0.1, 50, d1, side, -1.0),
                                make_disk([4.0, -3.0, 5.0], unit([-0.5, 2.0, -0.5]), 1.0, 0.04, 50, d2, side, -1.0)])
        return InstancedScene(verts, faces, [mesh_instance([0.0, 0.5, 5.0], 1.0, tf, teapot, 0.05)], disks,
                              boxes=((0.0, 0.0, 20.1), 20.0))
    if mesh_path.endswith(".obj"):
        pot = obj_parser.parse_obj(mesh_path, [0.0, 0.5, 5.0], 1.0, tf, teapot, 0.05)
    else:
        verts, faces = obj_parser.load_mesh_bin(mesh_path)
        pot = obj_parser.mesh_to_triangles(verts, faces, [0.0, 0.5, 5.0], 1.0, tf, teapot, 0.05)
    tris = np.concatenate([
        make_dummy_triangle(),
        pot,
        make_disk([4.0, 4.0, 7.0], unit([-0.3, -0.55, -0.5]), 2.0, 0.1, 50, d1, side, -1.0),
        make_disk([4.0, -3.0, 5.0], unit([-0.5, 2.0, -0.5]), 1.0, 0.04, 50, d2, side, -1.0),
    ])
    return Scene(tris, boxes=((0.0, 0.0, 20.1), 20.0))


def teapot_field_scene(nz=12, ny=13, seed=1, surface=None, scale=0.3, mesh_path=TEAPOT_MESH, instanced=False) -> Scene:
    """BASELINE config 4: nz x ny instanced teapots (12 x 13 = 156 -> 985,920 triangles + the dummy) standing on a
    grid that recedes from main.rs's camera, every instance with its own roll angle from an LCG (fixed seed) and
    mirror-like `Reflective{scattering: 0}` surfaces so that bounce rays are incoherent.  All instances lie inside
    main.rs's octree root cube ((0,0,20.1), 20), so the reference's visibility cull keeps every triangle."""
    if surface is None:
        surface = SurfaceKind.Reflective(0.0, make_color((252, 119, 0)), 0.5)
    verts, faces = obj_parser.load_mesh_bin(mesh_path)
    parts = [make_dummy_triangle()]
    insts = []
    state = int(seed) & 0xFFFFFFFF
    for iz in range(nz):
        for iy in range(ny):
            state = (state * 1664525 + 1013904223) & 0xFFFFFFFF          # Numerical Recipes LCG
            roll = 270.0 + 360.0 * (state >> 8) / float(1 << 24)
            tf = create_transform(unit([0.0, 0.3, 1.0]), to_radians(roll))
            off = [-1.2, (iy - (ny - 1) / 2.0) * 2.2, 3.0 + 2.0 * iz]
            if instanced:
                insts.append(mesh_instance(off, scale, tf, surface, 0.05))
            else:
                parts.append(obj_parser.mesh_to_triangles(verts, faces, off, scale, tf, surface, 0.05))
    if instanced:      # 76 KB of mesh + 156 instance records instead of 138 MB of finished triangles
        return InstancedScene(verts, faces, insts, None, boxes=((0.0, 0.0, 20.1), 20.0))
    return Scene(np.concatenate(parts), boxes=((0.0, 0.0, 20.1), 20.0))


def main_viewport(width, height, maxdepth=5, spp=1) -> RtbView:
    """main.rs:166-173 with aspect = height/width as in main.rs:96-110."""
    aspect = np.float32(height) / np.float32(width)
    return create_viewport((width, height), (1.0, float(np.float32(1.0) * aspect)), [2.0, 0.0, 0.0],
                           unit([0.0, 0.0, 1.0]), 90.0, to_radians(0.0), maxdepth, spp)
