#!/usr/bin/env python
"""compare_golden.py — checks a dump of the REFERENCE's own renderer against this repository's goldens.

    python tools/compare_golden.py golden_out [--gpu]

`golden_out/` is what `cargo +nightly run --release --bin dump_golden` (rust_raytrace_b200/rust/b200_raytrace_lib) wrote
on a machine with the reference's nightly toolchain: per frame the reference's debug CSV (Pixel_x = row, Pixel_y = col,
tri_hit, hit_t; debug.rs:118-140) and the raw f32 RGBA frame.  This script compares

  * golden_64x64_{shipped,det}     with tests/golden/main_scene_64.npz (the committed vectors every GPU test uses),
  * every dumped frame             with the CPU oracle (oracle/rt_oracle.cpp, reference-algorithm mode) rendered here,
  * with --gpu, also               with the CUDA path through the C ABI (needs a B200),

bit for bit: primitive ids, hit times, and — for the deterministic material set — RGBA.  (With the shipped Matte / fuzzy
materials the reference draws from an unseeded ThreadRng: only the primary hits are comparable.)  Mismatches are listed in
the style of DebugCtx::compare_to (debug.rs:150-221).  A clean run upgrades the oracle's status from "parity unpinned" to
"pinned by the reference" (DESIGN.md section 2): exit status 0 and the line `Found 0 errors`.
"""
import argparse
import glob
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def read_dump(csv_path):
    m = re.search(r"golden_(\d+)x(\d+)_(shipped|det)\.csv$", csv_path)
    w, h, tag = int(m.group(1)), int(m.group(2)), m.group(3)
    prim = np.zeros((h, w), np.uint32)
    t = np.zeros((h, w), np.float32)
    seen = np.zeros((h, w), bool)
    with open(csv_path) as fh:
        next(fh)                                     # Pixel_x;Pixel_y;ray_p;ray_v;tri_hit;hit_t;check_tris
        for ln in fh:
            f = ln.rstrip("\n").split(";")
            if len(f) < 6:
                continue
            row, col = int(f[0]), int(f[1])
            prim[row, col] = int(f[4])
            t[row, col] = np.float32(f[5])           # shortest round-trip digits -> the exact f32
            seen[row, col] = True
    rgba = np.fromfile(csv_path[:-4] + ".rgba", "<f4").reshape(h, w, 4)
    return w, h, tag, prim, t, rgba, seen


def report(name, w, prim, t, rgba, oprim, ot, orgba, det, out):
    bad = 0
    for r, c in np.argwhere(prim != oprim):
        out.append(f"{name} ({r},{c}): Hit Mismatch {prim[r, c]} vs {oprim[r, c]}")
        bad += 1
    same = prim == oprim
    tb = (t.view(np.uint32) != ot.view(np.uint32)) & same
    for r, c in np.argwhere(tb):
        out.append(f"{name} ({r},{c}): Hit times differ {t[r, c]!r} vs {ot[r, c]!r}")
        bad += 1
    if det:
        cb = (rgba.view(np.uint32) != orgba.view(np.uint32)).any(-1)
        for r, c in np.argwhere(cb):
            out.append(f"{name} ({r},{c}): Colour Mismatch {rgba[r, c].tolist()} vs {orgba[r, c].tolist()}")
            bad += 1
    return bad


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("dump_dir")
    ap.add_argument("--gpu", action="store_true", help="also compare with the CUDA path (needs a B200)")
    args = ap.parse_args()
    from oracle import oracle as O
    verts, faces = O.load_mesh_bin()
    npz = np.load(os.path.join(ROOT, "tests", "golden", "main_scene_64.npz"))
    lines, errors, frames = [], 0, 0
    for csv_path in sorted(glob.glob(os.path.join(args.dump_dir, "golden_*.csv"))):
        w, h, tag, prim, t, rgba, seen = read_dump(csv_path)
        det = tag == "det"
        frames += 1
        if not seen.all():
            lines.append(f"{tag} {w}x{h}: {int((~seen).sum())} pixels have no entry")
            errors += int((~seen).sum())
        if (w, h) == (64, 64):
            errors += report(f"npz {tag} 64x64", w, prim, t, rgba, npz[f"{tag}_prim"], npz[f"{tag}_t"], npz[f"{tag}_rgba"], det, lines)
        osc = O.Scene(O.main_scene_tris(verts, faces, det), O.ACCEL_OCTREE)
        orgba, oprim, ot, _ = osc.render(O.main_viewport(w, h, 5, 1), seed=0)
        errors += report(f"oracle {tag} {w}x{h}", w, prim, t, rgba, oprim, ot, orgba, det, lines)
        if args.gpu:
            import rust_raytrace_b200 as R
            v = R.main_viewport(w, h, 5, 1)
            data = R.new_image(v)
            caster = R.B200RayCaster(want_ids=True, seed=0)
            caster.walk_rays(v, R.main_scene(deterministic=det), data, threads=1)
            errors += report(f"gpu {tag} {w}x{h}", w, prim, t, rgba, caster.prim, caster.t, data, det, lines)
    for ln in lines[:200]:
        print(ln)
    if frames == 0:
        print(f"no golden_*.csv in {args.dump_dir}")
        return 2
    print(f"{frames} frames compared")
    print(f"Found {errors} errors")
    return 0 if errors == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
