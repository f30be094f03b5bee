"""Writes the corners (9 f32 per triangle, dummy excluded) of main.rs's scene for tools/experiments/bvh_quality.cpp."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
out = sys.argv[1] if len(sys.argv) > 1 else "/tmp/bq/corners.bin"
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
tris = R.main_scene().tris
np.ascontiguousarray(tris["corners"][1:], np.float32).tofile(out)
print(len(tris) - 1, "triangles ->", out)
