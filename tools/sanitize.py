"""Small workload for compute-sanitizer (memcheck / racecheck / initcheck): scene build (sort, PLOC, emit), one frame with
every renderer variant, the progressive path, the quantiser and the sort self-test."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
s = R.main_scene(False)
for flags in (0, _lib.RTB_FLAG_STATS, _lib.RTB_FLAG_FUSED, _lib.RTB_FLAG_MEGAKERNEL, _lib.RTB_FLAG_BRUTE):
    for spp in (1, 2):
        v = R.main_viewport(48, 40, 5, spp)
        v.flags = flags
        if flags == _lib.RTB_FLAG_BRUTE and spp == 2:
            continue
        data = R.new_image(v)
        ctx = R.B200RayCaster(want_ids=True, seed=3).walk_rays(v, s, data, threads=1)
        print("flags", flags, "spp", spp, "rays", ctx.total_rays, flush=True)
v = R.main_viewport(40, 24, 5, 4)
data = R.new_image(v)
R.B200RayCaster(seed=1).walk_rays_progressive(v, s, data, threads=1)
print("progressive ok", float(data.mean()))
print("quantise", R.quantize_rgb8(data).shape)
_lib.check(_lib.lib().rtb_selftest_sort(5000, 63, 9), "sort")
print("sanitize workload done")
# second half of round 1: scene assembly + cull kernels, reference splitting, rgb8 output, extension renderer
from rust_raytrace_b200 import raytrace as rt
si = R.main_scene(False, instanced=True)
v = R.main_viewport(48, 40, 5, 1)
rgb = np.zeros((40, 48, 3), np.uint8)
R.B200RayCaster(seed=3).walk_rays_rgb8(v, si, rgb)
print("instanced + rgb8 ok", int(rgb.sum()), "refs", si.info().n_refs, flush=True)
print("cull", len(rt.cull_triangles(s.tris, ((0.0, 0.5, 5.0), 1.0))))
c = R.circles_scene(n=8, seed=2)
for spp in (1, 2):
    vv = R.main_viewport(48, 40, 3, spp)
    d = R.new_image(vv)
    print("circles spp", spp, R.B200RayCaster(want_ids=True, seed=3).walk_rays(vv, c, d, threads=1).total_rays, flush=True)
s.set_light((6.0, -2.0, 0.0), 0.5)
d = R.new_image(v)
print("teapot + light", R.B200RayCaster(seed=3).walk_rays(v, s, d, threads=1).total_rays)
print("sanitize workload (part 2) done")
