"""Generates the committed fixtures from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  rust_raytrace_b200/data/teapot_mesh.bin   vertices + faces of raytrace/teapot_tri.obj, parsed with the
                                             oracle's OBJ reader (strtof == Rust's correctly-rounded parse)
  tests/golden/main_scene_64.npz             oracle (reference-algorithm: octree) render of main.rs's scene
                                             at main.rs's own 64x64 size: prim ids, t, rgba
  tests/golden/main_scene_stats.json         structural numbers of the reference scene/octree

Usage: python tests/golden/make_fixtures.py
"""
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/raytrace"


def main():
    v1, f1 = O.parse_obj_file(os.path.join(REF, "teapot_tri.obj"))
    v2, f2 = O.parse_obj_file(os.path.join(REF, "teapot.obj"))
    assert np.array_equal(v1, v2) and np.array_equal(f1, f2), "teapot.obj and teapot_tri.obj differ"
    path = os.path.join(ROOT, "rust_raytrace_b200", "data", "teapot_mesh.bin")
    with open(path, "wb") as fh:
        fh.write(b"RTBM")
        fh.write(struct.pack("<II", len(v1), len(f1)))
        fh.write(v1.astype("<f4").tobytes())
        fh.write(f1.astype("<u4").tobytes())
    print("wrote", path, len(v1), "verts", len(f1), "faces")

    stats = {}
    out = {}
    for det in (False, True):
        tris = O.main_scene_tris(v1, f1, deterministic=det)
        sc = O.Scene(tris, O.ACCEL_OCTREE)
        if not det:
            st = sc.tree_stats()
            stats = {"n_tris": int(len(tris)), **{k: int(getattr(st, k)) for k, _ in st._fields_}}
        v = O.main_viewport(64, 64, maxdepth=5, spp=1)
        rgba, prim, t, rs = sc.render(v, seed=0, threads=1)
        tag = "det" if det else "shipped"
        out[f"{tag}_rgba"] = rgba
        out[f"{tag}_prim"] = prim
        out[f"{tag}_t"] = t
        stats[f"{tag}_rays_64"] = int(rs.rays)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "main_scene_64.npz"), **out)
    with open(os.path.join(ROOT, "tests", "golden", "main_scene_stats.json"), "w") as fh:
        json.dump(stats, fh, indent=1)
    print(stats)


if __name__ == "__main__":
    main()
