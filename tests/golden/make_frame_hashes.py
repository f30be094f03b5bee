"""Golden frame hashes of bench.py's workloads, rendered by the CPU oracle (runs anywhere: needs no reference tree).

    python tests/golden/make_frame_hashes.py            # writes tests/golden/frame_hashes.json

sha256 over the little-endian f32 RGBA frame (H*W*16 bytes, lane 3 = 0) of the oracle in BVH mode, seed 7 — which
tests/test_oracle.py pins to the reference-algorithm (octree) mode.  bench.py compares every timed configuration's frame
with these (`parity.equals_golden`), at every GPU count."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

SEED = 7


def main():
    verts, faces = O.load_mesh_bin()
    out = {}
    tris = O.main_scene_tris(verts, faces, False)
    sc = O.Scene(tris, O.ACCEL_BVH)
    rgba, _, _, st = sc.render(O.main_viewport(3840, 2160, 5, 1), seed=SEED, want_ids=False)
    out["teapot4k"] = {"sha256": hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest(), "rays": int(st.rays)}
    print("teapot4k", out["teapot4k"], flush=True)
    ft = O.teapot_field_tris(verts, faces)
    sc = O.Scene(ft, O.ACCEL_BVH)
    rgba, _, _, st = sc.render(O.main_viewport(2560, 1440, 5, 1), seed=SEED, want_ids=False)
    out["field1m"] = {"sha256": hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest(), "rays": int(st.rays)}
    print("field1m", out["field1m"], flush=True)
    t, sph, light = O.circles_scene_parts()
    sc = O.Scene(t, O.ACCEL_BVH).add_spheres(sph).set_light(*light)
    rgba, _, _, st = sc.render(O.main_viewport(2560, 1440, 2, 1), seed=SEED, want_ids=False)
    out["circles2k"] = {"sha256": hashlib.sha256(np.ascontiguousarray(rgba).tobytes()).hexdigest(), "rays": int(st.rays)}
    print("circles2k", out["circles2k"], flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "frame_hashes.json"), "w") as fh:
        json.dump({"seed": SEED, "frames": out}, fh, indent=1)


if __name__ == "__main__":
    main()
