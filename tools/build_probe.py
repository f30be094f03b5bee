"""Probe: scene upload + BVH build times and shape for the two scenes (best of 3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
for name, mk in (("teapot (6,721 tris)", lambda: R.main_scene(False)), ("field (985,921 tris)", R.teapot_field_scene)):
    best = None
    for _ in range(3):
        s = mk(); t0 = time.perf_counter(); i = s.info(); wall = (time.perf_counter() - t0) * 1e3
        if best is None or i.ms_build < best[0]: best = (i.ms_build, i.ms_upload, wall, i.n_nodes, i.n_leaves, i.tree_height, i.build_launches)
        s.release()
    print(f"{name}: build {best[0]:.3f} ms  upload {best[1]:.3f} ms  create wall {best[2]:.1f} ms  nodes {best[3]} leaves {best[4]} height {best[5]} launches {best[6]}", flush=True)
