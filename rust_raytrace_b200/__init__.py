"""rust_raytrace_b200 — B200-native renderer core behind gerikkub/rust_raytrace's RayCaster API.

Only what the hot path needs: `csrc/` (CUDA kernels + the C ABI of include/rtb.h, host-side
scene preparation in C++) and `raytrace.py`, the host-side mirror of the reference interface.
"""
from . import _lib  # noqa: F401
from .raytrace import (B200RayCaster, InstancedScene, analytic_sphere, circles_scene, ProgressCtx, Scene, SurfaceKind, Viewport, create_transform, create_viewport,  # noqa: F401
                       main_scene, main_viewport, teapot_field_scene, make_color, make_disk, make_dummy_triangle, make_sphere,
                       make_triangle, make_vec, new_image, obj_parser, populate_triangle_numbers, quantize_rgb8,
                       to_radians, unit, write_png, write_ppm)

__all__ = [
    "B200RayCaster", "InstancedScene", "analytic_sphere", "circles_scene", "ProgressCtx", "Scene", "SurfaceKind", "Viewport", "create_transform", "create_viewport",
    "main_scene", "main_viewport", "teapot_field_scene", "make_color", "make_disk", "make_dummy_triangle", "make_sphere", "make_triangle",
    "make_vec", "new_image", "obj_parser", "populate_triangle_numbers", "quantize_rgb8", "to_radians", "unit",
    "write_png", "write_ppm",
]
