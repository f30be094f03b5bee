"""ctypes loader for librtb.so (the C ABI of include/rtb.h + include/rtb_host.h).

There is no fallback: if the shared library is missing this raises, and every
entry point that needs a GPU returns RTB_ERR_NO_DEVICE without one.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RTB_LIB", os.path.join(_PKG, "librtb.so"))   # RTB_LIB: A/B builds of the same ABI

RTB_OK, RTB_ERR_NO_DEVICE, RTB_ERR_CUDA, RTB_ERR_INVALID, RTB_ERR_NOMEM = 0, -1, -2, -3, -4
RTB_SOLID, RTB_MATTE, RTB_REFLECTIVE = 0, 1, 2
RTB_FLAG_SUM_ONLY, RTB_FLAG_STATS, RTB_FLAG_BRUTE, RTB_FLAG_MEGAKERNEL, RTB_FLAG_TIMING, RTB_FLAG_FUSED, RTB_FLAG_BVH8, RTB_FLAG_COPY_ONLY = 1, 2, 4, 8, 16, 64, 128, 256
RTB_STAGES = ("raygen", "trace", "shade", "bounce")
RTB_MAX_DEPTH = 16
RTB_MAX_GPUS = 8

# numpy mirror of RtbTriangle (35 x 4 bytes; reference field order raytrace.rs:326-337)
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


class RtbView(C.Structure):
    _fields_ = [
        ("width", C.c_uint32),
        ("height", C.c_uint32),
        ("orig", C.c_float * 3),
        ("cam", C.c_float * 3),
        ("vu", C.c_float * 3),
        ("vv", C.c_float * 3),
        ("maxdepth", C.c_uint32),
        ("spp", C.c_uint32),
        ("seed", C.c_uint64),
        ("sample_begin", C.c_uint32),
        ("sample_end", C.c_uint32),
        ("flags", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class RtbStats(C.Structure):
    _fields_ = [
        ("rays", C.c_uint64),
        ("node_tests", C.c_uint64),
        ("tri_tests", C.c_uint64),
        ("ms_render", C.c_double),
        ("ms_total", C.c_double),
        ("kernel_launches", C.c_uint32),
        ("n_gpus", C.c_uint32),
        ("bounce_rays", C.c_uint64),
        ("node_tests_bounce", C.c_uint64),
        ("tri_tests_bounce", C.c_uint64),
        ("ms_stage", C.c_double * 4),
        ("ms_reduce", C.c_double),
    ]


class RtbSceneInfo(C.Structure):
    _fields_ = [
        ("n_tris", C.c_uint32),
        ("n_prims", C.c_uint32),
        ("n_nodes", C.c_uint32),
        ("n_leaves", C.c_uint32),
        ("max_leaf", C.c_uint32),
        ("tree_height", C.c_uint32),
        ("scene_lo", C.c_float * 3),
        ("scene_hi", C.c_float * 3),
        ("ms_upload", C.c_double),
        ("ms_build", C.c_double),
        ("build_launches", C.c_uint32),
        ("n_gpus", C.c_uint32),
        ("n_refs", C.c_uint32),
        ("reserved", C.c_uint32),
    ]


class RtbSurface(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("color", C.c_float * 3), ("alpha", C.c_float), ("scattering", C.c_float)]


# EXTENSION (include/rtb.h): analytic sphere, 40 bytes
SPH_DTYPE = np.dtype([("center", "<f4", (3,)), ("radius", "<f4"), ("kind", "<u4"), ("color", "<f4", (3,)),
                      ("alpha", "<f4"), ("scattering", "<f4")])


class RtbMeshInstance(C.Structure):
    _fields_ = [("transform_rows", C.c_float * 9), ("offset", C.c_float * 3), ("scale", C.c_float),
                ("edge_thickness", C.c_float), ("kind", C.c_uint32), ("color", C.c_float * 3), ("alpha", C.c_float),
                ("scattering", C.c_float)]


# Every symbol include/rtb.h and include/rtb_host.h declare (checked by tests/test_host.py::test_library_exports_every_declared_symbol).
RTB_SYMBOLS = [
    "rtb_init", "rtb_device_count", "rtb_visible_device_count", "rtb_shutdown", "rtb_last_error", "rtb_scene_create", "rtb_scene_info",
    "rtb_scene_destroy", "rtb_scene_download_bvh", "rtb_render_rgb8", "rtb_scene_create_instanced", "rtb_assemble_triangles",
    "rtb_cull_triangles", "rtb_scene_create_ext", "rtb_scene_set_light", "rtb_render", "rtb_render_device", "rtb_render_progressive",
    "rtb_quantize_rgb8", "rtb_scale_device", "rtb_selftest_sort", "rtb_selftest_udiv", "rtb_partition_rows", "rtb_host_register", "rtb_host_unregister",
    "rtb_device_alloc", "rtb_device_free", "rtb_ipc_export", "rtb_ipc_open", "rtb_ipc_close",
]
RTBH_SYMBOLS = [
    "rtbh_make_color", "rtbh_unit", "rtbh_to_radians", "rtbh_make_triangle", "rtbh_make_dummy_triangle",
    "rtbh_make_disk", "rtbh_make_sphere", "rtbh_create_transform", "rtbh_create_viewport", "rtbh_parse_obj",
    "rtbh_mesh_to_triangles", "rtbh_load_mesh_bin", "rtbh_box_contains_polygon", "rtbh_write_ppm", "rtbh_write_png", "rtbh_write_png_rgb8",
]


def build(verbose: bool = False) -> str:
    """Compile librtb.so for sm_100a with the committed Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j4"]
    subprocess.run(cmd, check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the GPU path)"
        )
    L = C.CDLL(LIB_PATH)
    f = C.POINTER(C.c_float)
    vp = C.c_void_p
    u32 = C.c_uint32
    L.rtb_init.argtypes = [C.c_int, C.POINTER(C.c_int)]
    L.rtb_device_count.argtypes = []
    L.rtb_visible_device_count.argtypes = []
    L.rtb_last_error.restype = C.c_char_p
    L.rtb_scene_create.argtypes = [vp, u32, f, C.c_float, C.POINTER(vp)]
    L.rtb_scene_create_instanced.argtypes = [vp, u32, vp, u32, vp, u32, vp, u32, f, C.c_float, C.POINTER(vp)]
    L.rtb_assemble_triangles.argtypes = [vp, u32, vp, u32, vp, u32, vp]
    L.rtb_cull_triangles.argtypes = [vp, u32, f, C.c_float, vp, C.POINTER(u32)]
    L.rtb_scene_create_ext.argtypes = [vp, u32, vp, u32, f, C.c_float, C.POINTER(vp)]
    L.rtb_scene_set_light.argtypes = [vp, f, C.c_float]
    L.rtb_scene_info.argtypes = [vp, C.POINTER(RtbSceneInfo)]
    L.rtb_scene_destroy.argtypes = [vp]
    L.rtb_scene_destroy.restype = None
    L.rtb_scene_download_bvh.argtypes = [vp, vp, vp]
    L.rtb_render.argtypes = [vp, C.POINTER(RtbView), vp, vp, vp, C.POINTER(RtbStats)]
    L.rtb_render_rgb8.argtypes = [vp, C.POINTER(RtbView), vp, C.POINTER(RtbStats)]
    L.rtb_render_device.argtypes = [vp, C.POINTER(RtbView), C.c_int, u32, u32, vp, vp, vp, vp, C.POINTER(RtbStats)]
    L.rtb_render_progressive.argtypes = [vp, C.POINTER(RtbView), vp, C.POINTER(RtbStats)]
    L.rtb_quantize_rgb8.argtypes = [vp, C.c_uint64, vp]
    L.rtb_scale_device.argtypes = [vp, C.c_uint64, u32, C.c_int, vp]
    L.rtb_selftest_sort.argtypes = [u32, C.c_int, C.c_uint64]
    L.rtb_selftest_udiv.argtypes = [u32, u32, C.c_uint64]
    L.rtb_partition_rows.argtypes = [u32, u32, u32, vp, u32]
    L.rtb_host_register.argtypes = [vp, C.c_size_t]
    L.rtb_host_unregister.argtypes = [vp]
    L.rtb_device_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp)]
    L.rtb_device_free.argtypes = [C.c_int, vp]
    L.rtb_ipc_export.argtypes = [vp, C.c_char_p]
    L.rtb_ipc_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.rtb_ipc_close.argtypes = [C.c_int, vp]
    # host helpers
    S = C.POINTER(RtbSurface)
    L.rtbh_make_color.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, f]
    L.rtbh_make_color.restype = None
    L.rtbh_unit.argtypes = [f, f]
    L.rtbh_unit.restype = None
    L.rtbh_to_radians.argtypes = [C.c_float]
    L.rtbh_to_radians.restype = C.c_float
    L.rtbh_make_triangle.argtypes = [f, S, C.c_float, vp]
    L.rtbh_make_dummy_triangle.argtypes = [vp]
    L.rtbh_make_disk.argtypes = [f, f, C.c_float, C.c_float, u32, S, S, C.c_float, vp, u32]
    L.rtbh_make_sphere.argtypes = [f, C.c_float, u32, u32, S, C.c_float, vp, u32]
    L.rtbh_create_transform.argtypes = [f, C.c_float, f]
    L.rtbh_create_transform.restype = None
    L.rtbh_create_viewport.argtypes = [u32, u32, C.c_float, C.c_float, f, f, C.c_float, C.c_float, u32, u32,
                                       C.POINTER(RtbView)]
    L.rtbh_create_viewport.restype = None
    L.rtbh_parse_obj.argtypes = [C.c_char_p, f, C.c_float, f, S, C.c_float, vp, u32]
    L.rtbh_mesh_to_triangles.argtypes = [vp, u32, vp, u32, f, C.c_float, f, S, C.c_float, vp]
    L.rtbh_load_mesh_bin.argtypes = [C.c_char_p, vp, u32, C.POINTER(u32), vp, u32, C.POINTER(u32)]
    L.rtbh_box_contains_polygon.argtypes = [f, C.c_float, vp]
    L.rtbh_write_ppm.argtypes = [C.c_char_p, u32, u32, vp]
    L.rtbh_write_png.argtypes = [C.c_char_p, u32, u32, vp]
    L.rtbh_write_png_rgb8.argtypes = [C.c_char_p, u32, u32, vp]
    _lib = L
    return L


class RtbError(RuntimeError):
    def __init__(self, code: int, where: str):
        msg = lib().rtb_last_error()
        super().__init__(f"{where} failed with status {code}: {msg.decode() if msg else ''}")
        self.code = code


def check(code: int, where: str) -> int:
    if code < 0:
        raise RtbError(code, where)
    return code
