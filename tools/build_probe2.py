"""Build-time probe: scene creation repeated, ms_upload / ms_build per repetition (teapot scene and the 1M field, host array and instanced)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_raytrace_b200 as R
for name, mk in (("teapot", lambda: R.main_scene(False)), ("field host", lambda: R.teapot_field_scene()), ("field instanced", lambda: R.teapot_field_scene(instanced=True))):
    sc = mk()
    out = []
    for rep in range(5):
        sc.release(); sc.upload(); i = sc.info()
        out.append(f"{i.ms_upload:.2f}/{i.ms_build:.2f}")
    print(name, "upload/build ms:", " ".join(out), "launches", i.build_launches, "height", i.tree_height, flush=True)
    sc.release()
