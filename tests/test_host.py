"""CPU tests of the product's host-side scene preparation (C++ in librtb.so) against the oracle's
independent restatement, byte for byte; the C ABI surface; the band partition."""
import ctypes as C
import os
import re

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(R):
    from rust_raytrace_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    declared = set()
    for hdr in ("rtb.h", "rtb_host.h"):
        text = open(os.path.join(root, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b(rtbh?_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.RTB_SYMBOLS + _lib.RTBH_SYMBOLS)
    L = _lib.lib()
    for name in sorted(declared):
        assert getattr(L, name) is not None


def test_struct_sizes_match_the_header(R, tmp_path):
    """The ctypes mirrors against the real header: a C program compiled with gcc from include/rtb.h prints sizeof /
    offsetof of every ABI struct."""
    import os
    import subprocess
    from rust_raytrace_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sizes.c"
    src.write_text('''#include <stdio.h>
#include <stddef.h>
#include "rtb.h"
#include "rtb_host.h"
int main(void) {
    printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu ", sizeof(RtbTriangle), sizeof(RtbView), sizeof(RtbStats), sizeof(RtbSceneInfo),
           sizeof(RtbSurface), offsetof(RtbView, seed), offsetof(RtbStats, ms_stage), offsetof(RtbSceneInfo, ms_upload),
           sizeof(RtbMeshInstance), offsetof(RtbMeshInstance, kind), offsetof(RtbSceneInfo, n_refs));
    printf("%zu %zu\\n", sizeof(RtbSphere), offsetof(RtbSphere, kind));
    return 0;
}''')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = [_lib.TRI_DTYPE.itemsize, C.sizeof(_lib.RtbView), C.sizeof(_lib.RtbStats), C.sizeof(_lib.RtbSceneInfo),
            C.sizeof(_lib.RtbSurface), _lib.RtbView.seed.offset, _lib.RtbStats.ms_stage.offset, _lib.RtbSceneInfo.ms_upload.offset,
            C.sizeof(_lib.RtbMeshInstance), _lib.RtbMeshInstance.kind.offset, _lib.RtbSceneInfo.n_refs.offset,
            _lib.SPH_DTYPE.itemsize, _lib.SPH_DTYPE.fields["kind"][1]]
    assert got == want, (got, want)
    assert got[0] == 140 and got[1] == 88


def test_main_scene_bytes_equal_oracle(R, O, teapot_mesh):
    verts, faces = teapot_mesh
    for det in (False, True):
        assert R.main_scene(det).tris.tobytes() == O.main_scene_tris(verts, faces, det).tobytes()


def test_bench_scene_builders_equal_the_oracles(R, O, teapot_mesh):
    """bench.py's CPU arms build their scenes with the oracle's own generators (no product import there): the bytes must be
    the product's — teapot field (config 4), circles scene (config 1 extension), the mesh fixture reader."""
    verts, faces = teapot_mesh
    ov, of = O.load_mesh_bin()
    assert np.array_equal(ov, verts) and np.array_equal(of, faces)
    assert R.teapot_field_scene(nz=2, ny=3).tris.tobytes() == O.teapot_field_tris(verts, faces, nz=2, ny=3).tobytes()
    c = R.circles_scene()
    t, sph, light = O.circles_scene_parts()
    assert c.tris.tobytes() == t.tobytes() and c.spheres.tobytes() == sph.tobytes() and c.light == light


def test_reference_arm_of_the_bench_does_not_load_the_product():
    """`bench.py --impl reference` must time the oracle alone: nothing of rust_raytrace_b200 (hence not librtb.so) may be
    imported, and its `config` must be the GPU arm's."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']; "
            "sys.path.insert(0, %r); import bench; bench.main(); "
            "assert not [m for m in sys.modules if m.startswith('rust_raytrace_b200')], 'product imported'; "
            "assert not [l for l in open('/proc/self/maps') if 'librtb' in l], 'librtb.so mapped'; "
            "print('CONFIG ' + json.dumps(bench.workload_config('teapot4k')))" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["cpu_baseline"]["kind"] == "port" and line["value"] > 0
    assert line["config"] == json.loads(lines[1][len("CONFIG "):])


@pytest.mark.parametrize("wh", [(64, 64), (640, 480), (2560, 1440), (3840, 2160), (7680, 4320), (333, 217), (1, 1)])
def test_viewport_bytes_equal_oracle(R, O, wh):
    a, b = R.main_viewport(*wh, maxdepth=5, spp=1), O.main_viewport(*wh, maxdepth=5, spp=1)
    assert bytes(a)[:64] == bytes(b)[:64]


def test_rotated_camera_and_transform(R, O):
    d = [0.3, -0.2, 0.9]
    assert R.create_transform(R.unit(d), 0.7).tobytes() == O.create_transform(O.unit(d), 0.7).tobytes()
    a = R.create_viewport((100, 50), (1.0, 0.5), [1, 2, 3], R.unit(d), 70.0, 0.3, 4, 2)
    b = O.create_viewport((100, 50), (1.0, 0.5), [1, 2, 3], O.unit(d), 70.0, 0.3, 4, 2)
    assert bytes(a)[:64] == bytes(b)[:64]


def test_sphere_and_disk_generators(R, O):
    s = R.make_sphere([1, 2, 3], 1.5, (8, 12), R.SurfaceKind.Matte(R.make_color((10, 20, 30)), 0.3), 0.02)
    so = O.make_sphere([1, 2, 3], 1.5, 8, 12, O.Surface(O.OR_MATTE, O.make_color(10, 20, 30), 0.3), 0.02)
    assert len(s) == 8 * 12 * 2 - 2 * 12 and s.tobytes() == so.tobytes()
    with pytest.raises(ValueError):
        R.make_sphere([0, 0, 0], 1.0, (7, 8), R.SurfaceKind.Solid([1, 1, 1]), 0.0)   # odd lat: reference asserts
    d = R.make_disk([0, 1, 2], R.unit([0.2, 0.1, -1]), 2.0, 0.1, 17, R.SurfaceKind.Reflective(0.01, [1, 1, 1], 0.5),
                    R.SurfaceKind.Solid([0, 0, 0]), -1.0)
    do = O.make_disk([0, 1, 2], O.unit([0.2, 0.1, -1]), 2.0, 0.1, 17, O.Surface(O.OR_REFLECTIVE, [1, 1, 1], 0.5, 0.01),
                     O.Surface(O.OR_SOLID, [0, 0, 0]), -1.0)
    assert len(d) == 68 and d.tobytes() == do.tobytes()


def test_obj_text_roundtrip(R, O, teapot_mesh, tmp_path):
    verts, faces = teapot_mesh
    p = tmp_path / "pot.obj"
    with open(p, "w") as fh:
        fh.write("# comment\nvn 0 0 1\n")
        for v in verts[:400]:
            fh.write("v %.9g %.9g %.9g\n" % tuple(v))
        for f in faces:
            if f.max() <= 400:
                fh.write("f %d//1 %d//2 %d//3\n" % tuple(f))   # the a//n form of teapot_tri.obj
    sel = faces[(faces.max(axis=1) <= 400)]
    surf = R.SurfaceKind.Solid([1, 0, 0])
    tf = R.create_transform(R.unit([0, 0.3, 1]), R.to_radians(270.0))
    a = R.obj_parser.parse_obj(str(p), [0, 0.5, 5], 1.0, tf, surf, 0.05)
    b = R.obj_parser.mesh_to_triangles(verts[:400], sel, [0, 0.5, 5], 1.0, tf, surf, 0.05)
    assert len(a) == len(sel) > 100 and a.tobytes() == b.tobytes()
    ov, of = O.parse_obj_file(str(p))
    assert np.array_equal(ov, verts[:400]) and np.array_equal(of, sel)
    with pytest.raises(ValueError):
        R.obj_parser.parse_obj(str(tmp_path / "missing.obj"), [0, 0, 0], 1.0, tf, surf, 0.0)


def test_degenerate_triangle_is_an_error_not_a_crash(R):
    with pytest.raises(ValueError):
        R.make_triangle([[0, 0, 0], [1, 1, 1], [2, 2, 2]], R.SurfaceKind.Solid([0, 0, 0]), 0.0)


def test_root_cube_membership(R):
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    f3 = lambda v: (C.c_float * 3)(*v)  # noqa: E731
    inside = R.make_triangle([[0, 0, 5], [1, 0, 5], [0, 1, 5]], R.SurfaceKind.Solid([1, 1, 1]), 0.0)
    outside = R.make_triangle([[0, 0, 50], [1, 0, 50], [0, 1, 50]], R.SurfaceKind.Solid([1, 1, 1]), 0.0)
    assert L.rtbh_box_contains_polygon(f3([0, 0, 20.1]), 20.0, inside.ctypes.data) == 1
    assert L.rtbh_box_contains_polygon(f3([0, 0, 20.1]), 20.0, outside.ctypes.data) == 0


def test_band_partition_covers_every_row_once(R):
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    for H in (1, 7, 8, 9, 64, 1440, 2160, 2161):
        for world in (1, 2, 3, 4, 8):
            seen = np.zeros(H, np.int32)
            for rank in range(world):
                rows = np.zeros(H, np.uint32)
                n = L.rtb_partition_rows(H, rank, world, rows.ctypes.data, H)
                assert n >= 0
                seen[rows[:n]] += 1
                assert np.all((rows[:n] // 8) % world == rank)
            assert np.all(seen == 1)
    assert L.rtb_partition_rows(10, 3, 2, None, 0) < 0


def test_slot_to_pixel_division_is_exact(R):
    """slot_to_pixel divides a warp-tile index by the tiles of a band with a multiplier computed once per frame on the host
    (rtb_udiv_make / rtb_udiv, Granlund & Montgomery): it must equal the machine's division for every divisor an image width
    can produce (1 .. 2 * ceil(65536 / 8)) and for awkward ones, over the whole 32-bit range of numerators."""
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    divisors = list(range(1, 1200)) + [2 * ((w + 7) // 8) for w in (3840, 7680, 2560, 1283, 65535)] + \
        [2**k for k in range(1, 32)] + [2**k - 1 for k in range(2, 32)] + [2**k + 1 for k in range(1, 31)] + [0xffffffff, 0x80000001]
    for d in divisors:
        assert L.rtb_selftest_udiv(d, 400, d) == 0, d
    assert L.rtb_selftest_udiv(0, 1, 0) == _lib.RTB_ERR_INVALID


def test_gpu_path_fails_loudly_without_a_device(R):
    """No CPU fallback: on a machine without CUDA the render call must raise, not produce pixels."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from rust_raytrace_b200._lib import RtbError
    v = R.main_viewport(8, 8)
    with pytest.raises(RtbError) as e:
        R.B200RayCaster().walk_rays(v, R.main_scene(), R.new_image(v), threads=1)
    assert e.value.code == -1
    # the scene-assembly entry points have no host fallback either
    from rust_raytrace_b200 import raytrace as rt
    sc = R.main_scene(instanced=True)
    for call in (sc.upload, lambda: sc.tris, lambda: rt.cull_triangles(R.main_scene().tris, ((0.0, 0.0, 20.1), 20.0))):
        with pytest.raises(RtbError) as e:
            call()
        assert e.value.code == -1


@pytest.mark.parametrize("wh", [(1, 1), (7, 5), (640, 360), (300, 73)])
def test_write_png_holds_the_reference_quantisation(R, O, wh, tmp_path):
    """write_png (raytrace.rs:1460-1478): the file must decode (zlib, CRC-checked) to exactly `(c*255.) as u8` of the
    frame — the oracle's quantiser — for f32 input and for already quantised input; > 64 KiB frames span several
    stored deflate blocks."""
    w, h = wh
    rs = np.random.RandomState(w * 1000 + h)
    data = rs.uniform(-0.1, 1.1, (h, w, 4)).astype(np.float32)
    data[0, 0, :3] = [np.nan, 1.0, 0.999999]
    want = O.quantize_rgb8(data.reshape(-1, 4)).reshape(h, w, 3)
    p1, p2 = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    R.write_png(p1, (w, h), data)
    R.write_png(p2, (w, h), want)
    from png_util import decode_png
    for p in (p1, p2):
        gw, gh, px = decode_png(p)
        assert (gw, gh) == (w, h) and np.array_equal(px, want)
    assert open(p1, "rb").read() == open(p2, "rb").read()


def test_driver_binary_fails_loudly_without_a_device(R):
    """raytrace_b200, the main.rs-equivalent driver: built by the Makefile, exits non-zero with the library's message
    when there is no GPU (no CPU rendering path)."""
    import os
    import subprocess
    import torch
    from rust_raytrace_b200 import _lib
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "raytrace_b200")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "--mesh", R.raytrace.TEAPOT_MESH], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr
