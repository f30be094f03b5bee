"""GPU parity tests: the CUDA path (through the C ABI, librtb.so) against the CPU oracle.

Bar (BASELINE.json north_star): hit mask, primitive ids and hit times BIT-EXACT; RGB within 1/255 on
>= 99.9 % of pixels in the deterministic mode (we assert the stronger bit-exact RGBA), PSNR >= 40 dB
in the stochastic mode.  Because the oracle and the CUDA path share one counter-based RNG spec,
stochastic materials are ALSO compared bit for bit; the PSNR test covers the case where the
summation order legitimately differs (samples partitioned over GPUs).
"""
import ctypes as C
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def gpu_render(R, scene, v, seed=0, want_ids=True, threads=1, stats=False):
    data = R.new_image(v)
    caster = R.B200RayCaster(want_ids=want_ids, seed=seed, stats=stats)
    ctx = caster.walk_rays(v, scene, data, threads=threads, show_progress=False)
    return data, caster.prim, caster.t, ctx, caster


def assert_bit_exact(gpu, ora, what=""):
    data, prim, t, ctx, _ = gpu
    rgba, oprim, ot, st = ora
    bad = int((prim != oprim).sum())
    assert bad == 0, f"{what}: {bad} primitive ids differ, first at {np.argwhere(prim != oprim)[:5].tolist()}"
    assert np.array_equal(bits(t), bits(ot)), f"{what}: hit times differ"
    nb = int((bits(data) != bits(rgba)).any(-1).sum())
    assert nb == 0, f"{what}: {nb} pixels differ in RGBA"
    assert ctx.total_rays == st.rays, f"{what}: ray count {ctx.total_rays} != oracle {st.rays}"


@pytest.fixture(scope="module")
def scenes(R, O):
    out = {}
    for det in (False, True):
        s = R.main_scene(deterministic=det)
        out[det] = (s, O.Scene(s.tris.view(O.TRI_DTYPE), O.ACCEL_OCTREE), O.Scene(s.tris.view(O.TRI_DTYPE), O.ACCEL_BVH))
    return out


def test_golden_64(R, scenes, golden):
    """main.rs's own 64x64 frame against the committed golden vectors (oracle, reference octree)."""
    stats, g = golden
    for det, tag in ((False, "shipped"), (True, "det")):
        data, prim, t, ctx, _ = gpu_render(R, scenes[det][0], R.main_viewport(64, 64, 5, 1))
        assert np.array_equal(prim, g[f"{tag}_prim"])
        assert np.array_equal(bits(t), bits(g[f"{tag}_t"]))
        assert np.array_equal(bits(data), bits(g[f"{tag}_rgba"]))
        assert ctx.total_rays == stats[f"{tag}_rays_64"]


@pytest.mark.parametrize("det", [True, False])
def test_vs_reference_octree_640x480(R, O, scenes, det):
    """Config 0/2 at the reference's 640x480 size: octree (reference algorithm) oracle, all bounces."""
    s, oct_sc, _ = scenes[det]
    assert_bit_exact(gpu_render(R, s, R.main_viewport(640, 480, 5, 1), seed=7),
                     oct_sc.render(O.main_viewport(640, 480, 5, 1), seed=7), f"640x480 det={det}")


def test_deterministic_rgb_within_1_255(R, O, scenes):
    """The north_star's stated RGB criterion, evaluated in write_png's u8 domain."""
    s, oct_sc, _ = scenes[True]
    data = gpu_render(R, s, R.main_viewport(640, 480, 5, 1))[0]
    rgba = oct_sc.render(O.main_viewport(640, 480, 5, 1))[0]
    q, oq = O.quantize_rgb8(data).astype(np.int32), O.quantize_rgb8(rgba).astype(np.int32)
    ok = (np.abs(q - oq).max(-1) <= 1).mean()
    assert ok >= 0.999


@pytest.mark.parametrize("wh", [(1, 1), (7, 5), (16, 8), (17, 9), (333, 217), (1000, 3)])
def test_ragged_sizes(R, O, scenes, wh):
    s, _, bvh = scenes[False]
    assert_bit_exact(gpu_render(R, s, R.main_viewport(*wh, 5, 1), seed=2), bvh.render(O.main_viewport(*wh, 5, 1), seed=2),
                     f"{wh}")


@pytest.mark.parametrize("maxdepth", [1, 2, 3, 8, 16])
def test_maxdepth(R, O, scenes, maxdepth):
    s, _, bvh = scenes[False]
    assert_bit_exact(gpu_render(R, s, R.main_viewport(320, 200, maxdepth, 1), seed=5),
                     bvh.render(O.main_viewport(320, 200, maxdepth, 1), seed=5), f"maxdepth {maxdepth}")


def test_maxdepth_above_cap_is_an_error(R, scenes):
    from rust_raytrace_b200._lib import RtbError
    v = R.main_viewport(8, 8, 17, 1)
    with pytest.raises(RtbError):
        gpu_render(R, scenes[True][0], v)


def test_multisample_jitter_bit_exact(R, O, scenes):
    """spp > 1: jittered primary rays (raytrace.rs:1382-1386) and in-order accumulation (:1422-1426)."""
    s, _, bvh = scenes[False]
    assert_bit_exact(gpu_render(R, s, R.main_viewport(240, 160, 5, 5), seed=9),
                     bvh.render(O.main_viewport(240, 160, 5, 5), seed=9), "spp=5")


def test_teapot_2k_prim_ids(R, O, scenes):
    """BASELINE config 2: teapot_tri.obj at 2560x1440, 1 spp, deterministic materials — bit-exact ids on
    all 3,686,400 pixels (oracle in BVH mode, which tests/test_oracle.py pins to the octree)."""
    s, _, bvh = scenes[True]
    assert_bit_exact(gpu_render(R, s, R.main_viewport(2560, 1440, 5, 1)), bvh.render(O.main_viewport(2560, 1440, 5, 1)),
                     "2K deterministic")


def test_teapot_4k_multibounce(R, O, scenes):
    """BASELINE config 3: 3840x2160, maxdepth 5, shipped materials (Matte teapot, fuzzy mirrors)."""
    s, oct_sc, bvh = scenes[False]
    gpu = gpu_render(R, s, R.main_viewport(3840, 2160, 5, 1), seed=7)
    assert_bit_exact(gpu, bvh.render(O.main_viewport(3840, 2160, 5, 1), seed=7), "4K shipped")
    assert gpu[3].total_rays == 14259831          # rays of the reference octree algorithm, seed 7
    # a band of the same frame against the reference's own octree traversal
    rows = (1000, 1064)
    rgba, prim, t, _ = oct_sc.render(O.main_viewport(3840, 2160, 5, 1), seed=7, rows=rows)
    sl = slice(*rows)
    assert np.array_equal(gpu[1][sl], prim[sl]) and np.array_equal(bits(gpu[2][sl]), bits(t[sl]))
    assert np.array_equal(bits(gpu[0][sl]), bits(rgba[sl]))


def test_bvh_equals_gpu_bruteforce(R, scenes):
    """Size-independent property: the LBVH traversal returns exactly what a linear scan over every
    primitive returns (same kernel, RTB_FLAG_BRUTE) — ids, t and colour, all bounces."""
    from rust_raytrace_b200 import _lib
    s = scenes[False][0]
    v = R.main_viewport(1280, 720, 5, 1)
    a = gpu_render(R, s, v, seed=3)
    vb = _lib.RtbView.from_buffer_copy(v)
    vb.flags |= _lib.RTB_FLAG_BRUTE
    b = gpu_render(R, s, vb, seed=3)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
    assert a[3].total_rays == b[3].total_rays


@pytest.mark.parametrize("spp", [1, 3])
def test_megakernel_equals_wavefront(R, O, scenes, spp):
    """The two renderers in the library (default wavefront pipeline, RTB_FLAG_MEGAKERNEL) are the same
    function: ids, t, colour and ray count, including multi-sample accumulation order."""
    from rust_raytrace_b200 import _lib
    s, _, bvh = scenes[False]
    v = R.main_viewport(801, 453, 5, spp)
    a = gpu_render(R, s, v, seed=13)
    vm = _lib.RtbView.from_buffer_copy(v)
    vm.flags |= _lib.RTB_FLAG_MEGAKERNEL
    b = gpu_render(R, s, vm, seed=13)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
    assert a[3].total_rays == b[3].total_rays
    assert_bit_exact(b, bvh.render(O.main_viewport(801, 453, 5, spp), seed=13), f"megakernel spp={spp}")


@pytest.mark.parametrize("spp,maxdepth,wh", [(1, 5, (1283, 721)), (3, 16, (640, 363)), (1, 1, (333, 217)), (2, 2, (17, 9))])
def test_fused_launch_equals_the_two_phase_launches(R, O, scenes, spp, maxdepth, wh):
    """The path kernel runs its primary phase and its bounce phase as two launches per sample.  RTB_FLAG_FUSED runs both in
    ONE launch whose bounce phase consumes the queue its primary phase is still filling (tagged entries + tickets, no
    kernel boundary): ids, t, colour, ray and test counts must be identical, and both must match the oracle."""
    from rust_raytrace_b200 import _lib
    s, _, bvh = scenes[False]
    v = R.main_viewport(*wh, maxdepth, spp)
    a = gpu_render(R, s, v, seed=23, stats=True)
    vs = _lib.RtbView.from_buffer_copy(v)
    vs.flags |= _lib.RTB_FLAG_FUSED
    b = gpu_render(R, s, vs, seed=23, stats=True)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
    assert a[3].total_rays == b[3].total_rays
    # the bounce rays do identical work in both; the primary rays of the two-launch renderer push their far children
    # half-sorted (three compare-exchanges instead of five), so their box / triangle test counts differ a little
    assert a[4].stats.node_tests_bounce == b[4].stats.node_tests_bounce and a[4].stats.tri_tests_bounce == b[4].stats.tri_tests_bounce
    assert abs(a[4].stats.node_tests - b[4].stats.node_tests) <= 0.01 * b[4].stats.node_tests
    assert abs(a[4].stats.tri_tests - b[4].stats.tri_tests) <= 0.01 * b[4].stats.tri_tests
    assert_bit_exact(b, bvh.render(O.main_viewport(*wh, maxdepth, spp), seed=23), f"fused {wh} spp={spp} maxdepth={maxdepth}")


@pytest.mark.parametrize("det,spp,maxdepth,wh", [(False, 1, 5, (1283, 721)), (True, 1, 5, (640, 363)), (False, 3, 16, (333, 217))])
def test_bvh8_equals_bvh4(R, O, det, spp, maxdepth, wh, monkeypatch):
    """A/B accelerator: the 8-wide compressed BVH (80-byte nodes, quantised child boxes, octant-ordered hit masks, built
    only with RTB_BVH8=1) against the default 4-wide uncompressed BVH of the same binary tree.  The closest hit does not
    depend on the accelerator: ids, t, colour and ray count identical, both equal to the oracle."""
    from rust_raytrace_b200 import _lib
    monkeypatch.setenv("RTB_BVH8", "1")
    s = R.main_scene(deterministic=det)
    s.upload()
    monkeypatch.delenv("RTB_BVH8")
    v = R.main_viewport(*wh, maxdepth, spp)
    a = gpu_render(R, s, v, seed=29, stats=True)
    v8 = _lib.RtbView.from_buffer_copy(v)
    v8.flags |= _lib.RTB_FLAG_BVH8
    b = gpu_render(R, s, v8, seed=29, stats=True)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
    assert a[3].total_rays == b[3].total_rays
    assert b[4].stats.node_tests != a[4].stats.node_tests          # it really was another tree
    want = O.Scene(s.tris.view(O.TRI_DTYPE), O.ACCEL_BVH).render(O.main_viewport(*wh, maxdepth, spp), seed=29)
    assert_bit_exact(a, want, f"bvh4 {wh}")
    assert_bit_exact(b, want, f"bvh8 {wh}")
    s.release()
    with pytest.raises(_lib.RtbError):                               # a scene built without it refuses the flag
        gpu_render(R, R.main_scene(deterministic=det), v8, seed=29)


def test_fused_kernel_is_repeatable(R, scenes):
    """Twenty frames through the fused launch on one scene handle (the queue is never cleared: every launch has its own
    entry tag) — every frame identical, and identical again after a differently sized frame reused the workspace."""
    from rust_raytrace_b200 import _lib
    s = scenes[False][0]
    v = R.main_viewport(960, 541, 5, 1)
    v.flags |= _lib.RTB_FLAG_FUSED
    first = gpu_render(R, s, v, seed=4)
    for k in range(20):
        if k == 10:
            v2 = R.main_viewport(400, 300, 5, 2)
            v2.flags |= _lib.RTB_FLAG_FUSED
            gpu_render(R, s, v2, seed=1)
        again = gpu_render(R, s, v, seed=4)
        assert np.array_equal(bits(first[0]), bits(again[0])) and np.array_equal(first[1], again[1])
        assert first[3].total_rays == again[3].total_rays


def test_rotated_camera(R, O, scenes):
    s, _, bvh = scenes[True]
    args = ((400, 300), (1.0, 0.75), [1.0, -1.0, 0.5], None, 60.0, 0.4, 5, 1)
    rv = R.create_viewport(*args[:3], R.unit([0.2, 0.15, 1.0]), *args[4:])
    ov = O.create_viewport(*args[:3], O.unit([0.2, 0.15, 1.0]), *args[4:])
    assert_bit_exact(gpu_render(R, s, rv), bvh.render(ov), "rotated camera")


def test_far_camera_near_the_range_limit(R, O, scenes):
    """The conservative slab test of the BVH4 traversal (centre / half-extent boxes, three FMA roundings per plane) is
    covered by the builder's box padding only while the viewport stays within 32x the scene's largest coordinate
    (rtb_api.cu check_camera).  Close to that limit — a viewport 260 units away from a scene whose largest coordinate is
    8.6 (the limit is 276), looking at the teapot through a 1 degree lens — the traversal must still return exactly what the linear scan over every
    primitive (RTB_FLAG_BRUTE) and the CPU oracle return; one step beyond the limit the call is refused."""
    from rust_raytrace_b200 import _lib
    s, _, bvh = scenes[True]
    args = ((400, 300), (1.0, 0.75), [2.0, 0.0, -260.0], None, 1.0, 0.0, 5, 1)
    rv = R.create_viewport(*args[:3], R.unit([0.0, 0.0, 1.0]), *args[4:])
    ov = O.create_viewport(*args[:3], O.unit([0.0, 0.0, 1.0]), *args[4:])
    a = gpu_render(R, s, rv, seed=5)
    assert (a[1] != 0).mean() > 0.15                                  # the teapot and the disks are in view
    vb = _lib.RtbView.from_buffer_copy(rv)
    vb.flags |= _lib.RTB_FLAG_BRUTE
    b = gpu_render(R, s, vb, seed=5)
    assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
    assert a[3].total_rays == b[3].total_rays
    assert_bit_exact(a, bvh.render(ov, seed=5), "far camera")
    too_far = R.create_viewport(args[0], args[1], [2.0, 0.0, -400.0], R.unit([0.0, 0.0, 1.0]), *args[4:])
    with pytest.raises(Exception):
        gpu_render(R, s, too_far)


def test_empty_and_tiny_scenes(R, O):
    v, ov = R.main_viewport(64, 48, 5, 1), O.main_viewport(64, 48, 5, 1)
    only_dummy = R.Scene(R.make_dummy_triangle())
    data, prim, t, ctx, _ = gpu_render(R, only_dummy, v)
    assert not prim.any() and ctx.total_rays == 64 * 48
    sky = np.array([128 / 255, 180 / 255, 1.0, 0.0], np.float32)           # raytrace.rs:1264
    assert np.array_equal(bits(data), np.broadcast_to(bits(sky), data.shape))
    # 1, 2, 5 triangles (root leaf / one split / two levels)
    surf = R.SurfaceKind.Reflective(0.0, R.make_color((200, 50, 50)), 0.5)
    tris = [R.make_dummy_triangle()]
    for k in range(5):
        z = 3.0 + k
        tris.append(R.make_triangle([[1 + k * 0.1, -1, z], [3, 0.2 * k, z], [1.5, 1, z + 0.5]], surf, 0.05))
        arr = np.concatenate(tris)
        sc = R.Scene(arr)
        osc = O.Scene(arr.view(O.TRI_DTYPE), O.ACCEL_TRIVIAL)
        assert_bit_exact(gpu_render(R, sc, v), osc.render(ov), f"{k + 1} triangles")


def test_root_cube_cull(R, O):
    """Triangles outside the reference's octree root are invisible there (raytrace.rs:795-805)."""
    surf = R.SurfaceKind.Solid(R.make_color((10, 200, 10)))
    arr = np.concatenate([R.make_dummy_triangle(),
                          R.make_triangle([[1, -1, 50], [3, 0, 50], [1.5, 1, 50]], surf, 0.0),    # z=50: outside
                          R.make_triangle([[1, -1, 6], [3, 0, 6], [1.5, 1, 6]], surf, 0.0)])
    v, ov = R.main_viewport(96, 96, 2, 1), O.main_viewport(96, 96, 2, 1)
    got = gpu_render(R, R.Scene(arr), v)
    assert got[4].stats.rays == 96 * 96 and set(np.unique(got[1])) == {0, 2}
    assert_bit_exact(got, O.Scene(arr.view(O.TRI_DTYPE), O.ACCEL_OCTREE).render(ov), "cull")
    assert R.Scene(arr).info().n_prims == 1 and R.Scene(arr, boxes=None).info().n_prims == 2


def _random_scene(R, rng, n_tris, inside_root_cube=True):
    """A triangle soup in main.rs's octree root cube ((0, 0, 20.1) +- 20): sizes from slivers to triangles spanning a third of
    the cube, every SurfaceKind, wire-frame edges on some.  inside_root_cube=False lets triangles stick out of the cube: the
    reference's octree then sees their outer parts only from some directions (tests/test_oracle.py)."""
    parts = [R.make_dummy_triangle()]
    while len(parts) <= n_tris:
        c = np.array([rng.uniform(-6, 8), rng.uniform(-9, 9), rng.uniform(3, 30)], np.float32)
        size = float(10 ** rng.uniform(-1.3, 0.9))
        p = [c + (rng.uniform(-1, 1, 3) * size * (0.05 if (k == 2 and rng.rand() < 0.15) else 1.0)).astype(np.float32) for k in range(3)]
        if inside_root_cube and any(abs(q[0]) > 19.5 or abs(q[1]) > 19.5 or not (0.6 <= q[2] <= 39.6) for q in p):
            continue
        col = R.make_color(tuple(int(x) for x in rng.randint(0, 256, 3)))
        kind = rng.randint(0, 3)
        surf = (R.SurfaceKind.Solid(col) if kind == 0 else R.SurfaceKind.Matte(col, float(rng.uniform(0.05, 0.9))) if kind == 1
                else R.SurfaceKind.Reflective(float(rng.choice([0.0, 0.001, 0.05])), col, float(rng.uniform(0.1, 0.9))))
        try:
            parts.append(R.make_triangle(p, surf, float(rng.choice([-1.0, 0.0, 0.03, 0.2]))))
        except ValueError:
            pass                                   # a degenerate draw: the reference would panic, skip it
    return np.concatenate(parts)


@pytest.mark.parametrize("seed", range(12))
def test_random_scenes_differential(R, O, seed):
    """Randomised differential test against the oracle: random triangle soups (2 .. 600 triangles, all three surface
    kinds, slivers and cube-spanning triangles), random cameras (position, direction, roll, field of view), odd image sizes
    (the centre pixel's ray runs exactly along the view axis: zero direction components in the slab test), maxdepth 1..8,
    1..3 samples — ids, t, RGBA and ray counts bit-exact against the oracle BVH AND the reference-algorithm octree (the soups
    lie inside the root cube; what happens when triangles stick out of it: tests/test_oracle.py)."""
    rng = np.random.RandomState(1000 + seed)
    n_tris = int(rng.choice([2, 5, 17, 60, 200, 600]))
    tris = _random_scene(R, rng, n_tris)
    w, h = int(rng.choice([33, 101, 161, 257])), int(rng.choice([31, 75, 121]))
    maxdepth, spp = int(rng.randint(1, 9)), int(rng.choice([1, 1, 2, 3]))
    if seed % 3 == 0:          # looking straight down +z from inside the cube: axis-parallel centre ray
        pos, dirv, roll = [float(rng.uniform(-1, 3)), float(rng.uniform(-2, 2)), 0.0], [0.0, 0.0, 1.0], 0.0
    else:
        pos = [float(rng.uniform(-4, 6)), float(rng.uniform(-6, 6)), float(rng.uniform(-3, 2))]
        dirv = [float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.5, 0.5)), 1.0]
        roll = float(rng.uniform(-0.6, 0.6))
    fov = float(rng.choice([40.0, 90.0, 120.0]))
    args = ((w, h), (1.0, h / w), pos)
    v = R.create_viewport(*args, R.unit(dirv), fov, roll, maxdepth, spp)
    ov = O.create_viewport(*args, O.unit(dirv), fov, roll, maxdepth, spp)
    sc = R.Scene(tris)
    got = gpu_render(R, sc, v, seed=seed)
    assert_bit_exact(got, O.Scene(tris.view(O.TRI_DTYPE), O.ACCEL_BVH).render(ov, seed=seed), f"random scene {seed} ({n_tris} tris, {w}x{h}, depth {maxdepth}, spp {spp})")
    assert_bit_exact(got, O.Scene(tris.view(O.TRI_DTYPE), O.ACCEL_OCTREE).render(ov, seed=seed), f"random scene {seed} vs octree")
    sc.release()


def test_cell_spanning_triangles_vs_the_reference_octree(R, O):
    """Triangles that span most of the octree root cube (tests/test_oracle.py::large_triangle_scene): the GPU path against
    the reference-algorithm (octree) oracle, mirror bounces included."""
    from test_oracle import large_triangle_scene
    tris = large_triangle_scene(O)
    v, ov = R.main_viewport(320, 240, 4, 1), O.main_viewport(320, 240, 4, 1)
    assert_bit_exact(gpu_render(R, R.Scene(tris.view(R.raytrace.TRI_DTYPE)), v, seed=2),
                     O.Scene(tris, O.ACCEL_OCTREE).render(ov, seed=2), "cell-spanning triangles")


def test_sphere_scene_config1(R, O):
    """BASELINE config 1 ('circles'): tessellated spheres (make_sphere, raytrace.rs:464-529) over a disk,
    primary + one bounce.  The analytic-sphere scene of circles_2k.png no longer exists in the reference."""
    rng = np.random.RandomState(1)
    parts = [R.make_dummy_triangle()]
    for k in range(6):
        c = [float(rng.uniform(-1, 3)), float(rng.uniform(-3, 3)), float(rng.uniform(5, 9))]
        col = R.make_color(tuple(int(x) for x in rng.randint(30, 255, 3)))
        surf = R.SurfaceKind.Reflective(0.0, col, 0.6) if k == 0 else R.SurfaceKind.Solid(col)
        parts.append(R.make_sphere(c, 0.8, (8, 16), surf, 0.0))
    parts.append(R.make_disk([-2.0, 0.0, 8.0], R.unit([1.0, 0.0, 0.05]), 8.0, 0.1, 40,
                             R.SurfaceKind.Matte(R.make_color((90, 90, 90)), 0.3), R.SurfaceKind.Solid([0.1, 0.1, 0.1]), -1.0))
    arr = np.concatenate(parts)
    v, ov = R.main_viewport(512, 288, 2, 1), O.main_viewport(512, 288, 2, 1)
    assert_bit_exact(gpu_render(R, R.Scene(arr), v, seed=4), O.Scene(arr.view(O.TRI_DTYPE), O.ACCEL_BVH).render(ov, seed=4),
                     "spheres")


def test_teapot_field_1m_triangles(R, O):
    """BASELINE config 4: 156 instanced teapots = 985,920 triangles, mirror surfaces (incoherent bounces), GPU LBVH
    build + traversal against the oracle's own BVH (the reference octree cannot be built at this size)."""
    s = R.teapot_field_scene()
    inf = s.info()
    assert inf.n_tris == 985921 and inf.n_prims == 985920 and inf.max_leaf <= 4
    _, order = s.download_bvh()
    assert inf.n_refs >= inf.n_prims and np.array_equal(np.unique(order), np.arange(1, 985921, dtype=np.uint32))
    osc = O.Scene(s.tris.view(O.TRI_DTYPE), O.ACCEL_BVH)
    v, ov = R.main_viewport(2560, 1440, 5, 1), O.main_viewport(2560, 1440, 5, 1)      # the config's own size
    assert_bit_exact(gpu_render(R, s, v, seed=3), osc.render(ov, seed=3), "teapot field 2K")
    s.release()


def test_stats_variant_and_bvh_shape(R, scenes):
    s = scenes[False][0]
    inf = s.info()
    assert inf.n_tris == 6721 and inf.n_prims == 6720 and inf.max_leaf <= 4 and inf.tree_height < 62
    nodes, order = s.download_bvh()
    assert inf.n_refs >= 6720 and len(order) == inf.n_refs and sorted(set(order.tolist())) == list(range(1, 6721))
    a = gpu_render(R, s, R.main_viewport(640, 360, 5, 1), seed=1)
    b = gpu_render(R, s, R.main_viewport(640, 360, 5, 1), seed=1, stats=True)
    assert np.array_equal(bits(a[0]), bits(b[0]))
    st = b[4].stats
    assert st.node_tests > st.rays and st.tri_tests > 0 and st.node_tests / st.rays < 200


@pytest.mark.parametrize("n,bits", [(1, 8), (31, 3), (4096, 64), (4097, 63), (100000, 5), (1 << 20, 63), (3000001, 40)])
def test_builder_sort_and_scan(R, n, bits):
    """The LBVH builder's hand-written stable radix sort and exclusive scans against std::stable_sort / a host sum."""
    from rust_raytrace_b200 import _lib
    _lib.check(_lib.lib().rtb_selftest_sort(n, bits, 1234 + n), "rtb_selftest_sort")


def test_quantiser_on_gpu(R, O):
    px = np.random.RandomState(0).uniform(-0.2, 1.2, (1000, 4)).astype(np.float32)
    px[0, :3] = [np.nan, 1.0, 0.0]
    assert np.array_equal(R.quantize_rgb8(px), O.quantize_rgb8(px))


def test_multi_gpu_tiles_identical(R, scenes):
    """Image bands split over all visible GPUs must reproduce the single-GPU frame bit for bit."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    s = R.main_scene(False)
    v = R.main_viewport(1920, 1083, 5, 1)
    one = gpu_render(R, s, v, seed=6, threads=1)
    s2 = R.main_scene(False)
    many = gpu_render(R, s2, v, seed=6, threads=n)
    assert np.array_equal(one[1], many[1]) and np.array_equal(bits(one[0]), bits(many[0]))
    assert one[3].total_rays == many[3].total_rays


@pytest.mark.parametrize("world", [2, 3, 8])
def test_band_partition_assembles_the_single_gpu_frame(R, scenes, world):
    """The multi-GPU band partition on ONE GPU: ranks 0..world-1 of `world` rendered one after the other into one device
    buffer (rtb_render_device's full-frame indexing) must leave exactly the frame a single rank renders — every pixel
    written once, bit for bit, same ray total.  Ragged height (1083 = 135 bands + 3 rows)."""
    import torch
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    _lib.check(L.rtb_init(1, None), "rtb_init")
    s = scenes[False][0]
    h = s.upload()
    W, H = 1920, 1083
    v = R.main_viewport(W, H, 5, 1)
    v.seed = 6

    def render(rank, wld, rgba, prim):
        st = _lib.RtbStats()
        _lib.check(L.rtb_render_device(h, C.byref(v), 0, rank, wld, rgba.data_ptr(), prim.data_ptr(), None, None, C.byref(st)), "render")
        return int(st.rays)

    one = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    one_prim = torch.zeros((H, W), dtype=torch.int32, device="cuda")
    rays_one = render(0, 1, one, one_prim)
    parts = torch.full((H, W, 4), float("nan"), dtype=torch.float32, device="cuda")
    parts_prim = torch.full((H, W), -1, dtype=torch.int32, device="cuda")
    rays = sum(render(r, world, parts, parts_prim) for r in range(world))
    torch.cuda.synchronize()
    assert torch.equal(parts.view(torch.int32), one.view(torch.int32))
    assert torch.equal(parts_prim, one_prim) and rays == rays_one
    # and each rank alone touches only its own rows (rtb_partition_rows)
    alone = torch.full((H, W, 4), float("nan"), dtype=torch.float32, device="cuda")
    render(1, world, alone, parts_prim)
    rows = np.zeros(H, np.uint32)
    n = L.rtb_partition_rows(H, 1, world, rows.ctypes.data, H)
    written = (~torch.isnan(alone[..., 0])).any(dim=1).cpu().numpy()
    assert sorted(np.flatnonzero(written).tolist()) == rows[:n].tolist()


def test_two_processes_render_into_one_frame_over_ipc(R, scenes, tmp_path):
    """The one-process-per-GPU shape of bench.py on ONE GPU: the parent allocates the frame (rtb_device_alloc), exports it
    (rtb_ipc_export) and renders band rank 0 of 2; a child PROCESS maps it (rtb_ipc_open) and renders rank 1 of 2 straight
    into the parent's memory.  The assembled frame must be the single-rank frame bit for bit."""
    import subprocess
    import sys
    import torch
    from rust_raytrace_b200 import _lib
    L = _lib.lib()
    _lib.check(L.rtb_init(1, None), "rtb_init")
    s = scenes[False][0]
    h = s.upload()
    W, H = 1283, 721
    v = R.main_viewport(W, H, 5, 1)
    v.seed = 8
    frame = C.c_void_p()
    _lib.check(L.rtb_device_alloc(0, W * H * 16, C.byref(frame)), "rtb_device_alloc")
    hbuf = C.create_string_buffer(64)
    _lib.check(L.rtb_ipc_export(frame, hbuf), "rtb_ipc_export")
    st = _lib.RtbStats()
    _lib.check(L.rtb_render_device(h, C.byref(v), 0, 0, 2, frame, None, None, None, C.byref(st)), "render rank 0")
    child = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
L = _lib.lib(); _lib.check(L.rtb_init(1, None), "init")
s = R.main_scene(False); h = s.upload()
v = R.main_viewport(%d, %d, 5, 1); v.seed = 8
p = C.c_void_p()
_lib.check(L.rtb_ipc_open(0, bytes.fromhex(sys.argv[1]), C.byref(p)), "rtb_ipc_open")
st = _lib.RtbStats()
_lib.check(L.rtb_render_device(h, C.byref(v), 0, 1, 2, p, None, None, None, C.byref(st)), "render rank 1")
_lib.check(L.rtb_ipc_close(0, p), "rtb_ipc_close")
print("RAYS", st.rays)
""" % (os.path.dirname(os.path.dirname(_lib.LIB_PATH)), W, H)
    r = subprocess.run([sys.executable, "-c", child, hbuf.raw.hex()], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    rays = int(st.rays) + int([ln for ln in r.stdout.splitlines() if ln.startswith("RAYS")][0].split()[1])
    one = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    st1 = _lib.RtbStats()
    _lib.check(L.rtb_render_device(h, C.byref(v), 0, 0, 1, one.data_ptr(), None, None, None, C.byref(st1)), "render whole")

    class Dev:
        __cuda_array_interface__ = {"shape": (H, W, 4), "typestr": "<f4", "data": (frame.value, False), "version": 2}
    both = torch.as_tensor(Dev(), device="cuda")
    torch.cuda.synchronize()
    assert torch.equal(both.view(torch.int32), one.view(torch.int32)) and rays == int(st1.rays)
    del both
    _lib.check(L.rtb_device_free(0, frame), "rtb_device_free")


def test_config5_8k_band_64spp_partitioned_psnr(R, O, scenes):
    """BASELINE config 5 at its own width: one 8-row band of the 7680x4320 frame (the band partition with tile_world = 540
    selects it), 64 spp partitioned over 8 sample shares with RTB_FLAG_SUM_ONLY, summed and scaled (raytrace.rs:1426) —
    PSNR >= 40 dB (the north_star's stochastic criterion) against a 4,096-spp oracle render of 4 of those rows."""
    import torch
    from rust_raytrace_b200 import _lib, dist as RD
    L = _lib.lib()
    _lib.check(L.rtb_init(1, None), "rtb_init")
    s, _, bvh = scenes[False]
    h = s.upload()
    W, H, spp, shares, band = 7680, 4320, 64, 8, 269
    v = R.main_viewport(W, H, 5, spp)
    v.seed = 1
    total = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    buf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    rays = 0
    for rank in range(shares):
        sv = RD.sample_view(v, rank, shares)
        assert sv.sample_end - sv.sample_begin == spp // shares
        st = _lib.RtbStats()
        _lib.check(L.rtb_render_device(h, C.byref(sv), 0, band, H // 8, buf.data_ptr(), None, None, None, C.byref(st)), "render")
        total += buf
        rays += int(st.rays)
    out = RD.reduce_samples(total, spp)
    torch.cuda.synchronize()
    rows = slice(band * 8, band * 8 + 4)
    got = out[rows].cpu().numpy()[..., :3]
    assert float(out[: band * 8].abs().max()) == 0.0 and rays >= 8 * W * spp       # only the band was rendered
    ref = bvh.render(O.main_viewport(W, H, 5, 4096), seed=99, rows=(rows.start, rows.stop), want_ids=False)[0][rows][..., :3]
    mse = float(np.mean((np.clip(got, 0, 1) - np.clip(ref, 0, 1)) ** 2))
    psnr = 10 * np.log10(1.0 / mse)
    assert psnr >= 40.0, psnr
    del total, buf


def test_two_threads_share_a_scene_handle(R, O, scenes):
    """thread-safe per scene handle.  Two host threads render different views through ONE handle at the same
    time (ctypes releases the GIL); a handle serialises its frames, so every frame must equal its single-threaded result."""
    import threading
    s, _, bvh = scenes[False]
    jobs = [(R.main_viewport(640, 360, 5, 1), 3), (R.main_viewport(801, 453, 5, 2), 4)]
    want = [gpu_render(R, s, v, seed=seed) for v, seed in jobs]
    errors = []

    def worker(k):
        try:
            v, seed = jobs[k]
            for _ in range(6):
                got = gpu_render(R, s, v, seed=seed)
                assert np.array_equal(bits(got[0]), bits(want[k][0])) and np.array_equal(got[1], want[k][1])
                assert got[3].total_rays == want[k][3].total_rays
        except Exception as e:       # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_copy_only_flag_and_camera_range(R, scenes):
    """RTB_FLAG_COPY_ONLY (bench.py's device-to-host floor) launches no kernel: the frame of the previous call comes home
    again.  A viewport farther out than 32x the scene's extent is refused (the box padding no longer covers the ray-origin
    rounding of the conservative slab test) instead of risking a missed hit."""
    from rust_raytrace_b200 import _lib
    s = scenes[True][0]
    v = R.main_viewport(320, 200, 5, 1)
    a = gpu_render(R, s, v, seed=2, want_ids=False)
    vc = _lib.RtbView.from_buffer_copy(v)
    vc.flags |= _lib.RTB_FLAG_COPY_ONLY
    b = gpu_render(R, s, vc, seed=2, want_ids=False)
    assert np.array_equal(bits(a[0]), bits(b[0])) and b[4].stats.kernel_launches == 0
    far = R.create_viewport((64, 64), (1.0, 1.0), [2.0, 0.0, -5000.0], R.unit([0.0, 0.0, 1.0]), 90.0, 0.0, 5, 1)
    with pytest.raises(_lib.RtbError) as e:
        gpu_render(R, s, far)
    assert e.value.code == -3


def test_sample_range_sum_only_bit_exact(R, O, scenes):
    """The per-rank piece of the sample-partitioned mode (torchrun): samples [b, e) of spp with RTB_FLAG_SUM_ONLY
    into a device buffer; must equal the oracle's partial sum bit for bit, and reduce + rtb_scale_device must
    reproduce the single-process frame to rounding."""
    import torch
    from rust_raytrace_b200 import _lib, dist as RD
    s, _, bvh = scenes[False]
    W, H, spp, world = 200, 120, 6, 4
    L = _lib.lib()
    _lib.check(L.rtb_init(1, None), "rtb_init")
    h = s.upload()
    v = R.main_viewport(W, H, 5, spp)
    v.seed = 21
    ov = O.main_viewport(W, H, 5, spp)
    total = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    rays = 0
    for rank in range(world):
        sv = RD.sample_view(v, rank, world)
        buf = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        st = _lib.RtbStats()
        _lib.check(L.rtb_render_device(h, C.byref(sv), 0, 0, 1, buf.data_ptr(), None, None, None, C.byref(st)), "render")
        ref, ost = bvh.render_samples(ov, sv.sample_begin, sv.sample_end, seed=21)
        assert np.array_equal(bits(buf.cpu().numpy()), bits(ref)), f"partial sum of rank {rank} differs"
        assert st.rays == ost.rays
        rays += st.rays
        total += buf
    out = RD.reduce_samples(total, spp)
    torch.cuda.synchronize()
    full = bvh.render(ov, seed=21)
    assert rays == full[3].rays
    got = out.cpu().numpy()
    assert np.all(got[..., 3] == 0) and np.max(np.abs(got - full[0])) < 1e-6


def test_progressive_psnr(R, O, scenes):
    """Stochastic mode: samples partitioned + peer reduce; PSNR >= 40 dB against a 4,096-spp oracle render."""
    s, _, bvh = scenes[False]
    W, H = 160, 120
    ref = bvh.render(O.main_viewport(W, H, 5, 4096), seed=99)[0][..., :3]
    v = R.main_viewport(W, H, 5, 256)
    data = R.new_image(v)
    R.B200RayCaster(seed=1).walk_rays_progressive(v, R.main_scene(False), data, threads=0)
    mse = float(np.mean((np.clip(data[..., :3], 0, 1) - np.clip(ref, 0, 1)) ** 2))
    psnr = 10 * np.log10(1.0 / mse)
    assert psnr >= 40.0, psnr


# ---------------------------------------------------------------------------
# scene assembly on the GPU: make_triangle / parse_obj's transform / root-cube cull as kernels
# ---------------------------------------------------------------------------
def test_gpu_make_triangle_bit_exact(R, O, teapot_mesh):
    """rtb_assemble_triangles (k_assemble) against the oracle's make_triangle (raytrace.rs:340-383) and the host mirror:
    every one of the 35 fields of every record, bit for bit — main.rs's teapot and rolled / scaled / shifted instances."""
    from rust_raytrace_b200 import raytrace as rt
    verts, faces = teapot_mesh
    cases = [([0.0, 0.5, 5.0], 1.0, 270.0, R.SurfaceKind.Matte(R.make_color((252, 119, 0)), 0.2), 0.05),
             ([-1.2, 3.3, 9.0], 0.3, 301.7, R.SurfaceKind.Reflective(0.002, R.make_color((1, 2, 3)), 0.5), -1.0),
             ([7.0, -2.0, 30.0], 2.5, 12.0, R.SurfaceKind.Solid(R.make_color((9, 8, 7))), 0.0)]
    insts, want, want_o = [], [], []
    for off, scale, roll, surf, edge in cases:
        tf = R.create_transform(R.unit([0.0, 0.3, 1.0]), R.to_radians(roll))
        insts.append(rt.mesh_instance(off, scale, tf, surf, edge))
        want.append(R.obj_parser.mesh_to_triangles(verts, faces, off, scale, tf, surf, edge))
        otf = O.create_transform(O.unit([0.0, 0.3, 1.0]), O.to_radians(roll))
        c = surf._c()
        want_o.append(O.mesh_to_triangles(verts, faces, off, scale, otf, O.Surface(c.kind, list(c.color), c.alpha, c.scattering), edge))
    got = rt.assemble_triangles(verts, faces, insts)
    assert len(got) == 3 * len(faces)
    assert got.tobytes() == np.concatenate(want_o).tobytes(), "GPU make_triangle differs from the oracle"
    assert got.tobytes() == np.concatenate(want).tobytes()


def test_gpu_make_triangle_rejects_what_the_reference_panics_on(R):
    from rust_raytrace_b200 import raytrace as rt
    from rust_raytrace_b200._lib import RtbError
    tf = R.create_transform(R.unit([0.0, 0.0, 1.0]), 0.0)
    inst = [rt.mesh_instance([0, 0, 5], 1.0, tf, R.SurfaceKind.Solid([1, 1, 1]), 0.0)]
    good = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1], [2, 2, 2]], np.float32)
    assert len(rt.assemble_triangles(good, np.array([[1, 2, 3]], np.uint32), inst)) == 1
    for faces in ([[1, 4, 5]], [[1, 2, 3], [1, 2, 9]], [[0, 1, 2]]):      # collinear corners; index out of range; 0-based
        with pytest.raises(RtbError) as e:
            rt.assemble_triangles(good, np.array(faces, np.uint32), inst)
        assert e.value.code == -3


def test_gpu_root_cube_cull_equals_host(R, scenes):
    """k_cull_flags + compaction against box_contains_polygon of the host mirror (raytrace.rs:753-779) on cubes that
    cut through the scene, so that all three outcomes (corner inside, face crossing, outside) occur."""
    from rust_raytrace_b200 import _lib, raytrace as rt
    L = _lib.lib()
    tris = scenes[False][0].tris
    for boxes in (((0.0, 0.0, 20.1), 20.0), ((0.0, 0.5, 5.0), 1.0), ((4.0, 4.0, 7.0), 0.7), ((2.0, 0.0, 4.0), 0.3), ((50.0, 0.0, 0.0), 1.0)):
        f3 = (C.c_float * 3)(*boxes[0])
        want = [i for i in range(1, len(tris)) if L.rtbh_box_contains_polygon(f3, boxes[1], tris[i:i + 1].ctypes.data) == 1]
        got = rt.cull_triangles(tris, boxes)
        assert got.tolist() == want, boxes
    assert rt.cull_triangles(tris, None).tolist() == list(range(1, len(tris)))
    assert 0 < len(rt.cull_triangles(tris, ((0.0, 0.5, 5.0), 1.0))) < 6720


def test_instanced_scene_equals_uploaded_scene(R, scenes):
    """rtb_scene_create_instanced (triangles computed on the GPU) must give the same tree and the same frame as
    rtb_scene_create with the host-made triangle array."""
    s_host = scenes[False][0]
    s_inst = R.main_scene(deterministic=False, instanced=True)
    a, b = s_host.info(), s_inst.info()
    assert (a.n_tris, a.n_prims, a.n_refs, a.n_nodes, a.n_leaves) == (b.n_tris, b.n_prims, b.n_refs, b.n_nodes, b.n_leaves)
    assert s_inst.tris.tobytes() == s_host.tris.tobytes()
    v = R.main_viewport(640, 360, 5, 1)
    x, y = gpu_render(R, s_host, v, seed=2), gpu_render(R, s_inst, v, seed=2)
    assert np.array_equal(x[1], y[1]) and np.array_equal(bits(x[2]), bits(y[2])) and np.array_equal(bits(x[0]), bits(y[0]))
    assert x[3].total_rays == y[3].total_rays
    s_inst.release()


def test_instanced_teapot_field_1m(R):
    """BASELINE config 4 built from 76 KB of mesh + 156 instance records: same counts and the same frame as the
    138 MB host-array upload (which test_teapot_field_1m_triangles compares with the oracle)."""
    s_inst, s_host = R.teapot_field_scene(instanced=True), R.teapot_field_scene()
    a, b = s_host.info(), s_inst.info()
    assert (a.n_tris, a.n_prims, a.n_refs, a.n_nodes) == (b.n_tris, b.n_prims, b.n_refs, b.n_nodes) and b.n_prims == 985920
    v = R.main_viewport(480, 270, 5, 1)
    x, y = gpu_render(R, s_host, v, seed=3), gpu_render(R, s_inst, v, seed=3)
    assert np.array_equal(x[1], y[1]) and np.array_equal(bits(x[0]), bits(y[0]))
    s_host.release()
    s_inst.release()


def test_render_rgb8_is_the_quantised_frame(R, O, scenes, tmp_path):
    """rtb_render_rgb8 (render + write_png's quantiser fused on the device, 3 B/px home) against the oracle's frame
    put through the oracle's quantiser; ragged height, and the PNG written from it decodes to the same pixels."""
    from png_util import decode_png
    s, _, obvh = scenes[False]
    for (w, h) in ((640, 363), (1283, 721)):
        v, ov = R.main_viewport(w, h, 5, 1), O.main_viewport(w, h, 5, 1)
        rgb = np.zeros((h, w, 3), np.uint8)
        caster = R.B200RayCaster(seed=5)
        ctx = caster.walk_rays_rgb8(v, s, rgb)
        rgba, _, _, st = obvh.render(ov, seed=5)
        want = O.quantize_rgb8(rgba.reshape(-1, 4)).reshape(h, w, 3)
        assert np.array_equal(rgb, want) and ctx.total_rays == st.rays
    path = str(tmp_path / "frame.png")
    R.write_png(path, (w, h), rgb)
    assert np.array_equal(decode_png(path)[2], want)


# ---------------------------------------------------------------------------
# EXTENSION: shadow rays and analytic spheres, oracle-vs-GPU
# ---------------------------------------------------------------------------
def _oracle_ext(O, scene, accel):
    osc = O.Scene(scene.tris.view(O.TRI_DTYPE), accel)
    if scene.spheres is not None:
        osc.add_spheres(scene.spheres.view(O.SPH_DTYPE))
    if scene.light is not None:
        osc.set_light(*scene.light)
    return osc


@pytest.mark.parametrize("det", [True, False])
def test_shadow_rays_on_the_teapot_scene(R, O, det):
    """main.rs's scene with `lights = Some(..)`: the shadow block of color_ray (commented out at raytrace.rs:1203-1224)
    live on the GPU and in the oracle — ids, t, RGBA and the ray count (shadow rays are not `Rays`) bit-exact; the
    reference-algorithm oracle (octree) and the oracle BVH agree with each other as well."""
    s = R.main_scene(deterministic=det)
    s.set_light((6.0, -2.0, 0.0), 0.5)
    v, ov = R.main_viewport(480, 270, 5, 1), O.main_viewport(480, 270, 5, 1)
    got = gpu_render(R, s, v, seed=6)
    want = _oracle_ext(O, s, O.ACCEL_BVH).render(ov, seed=6)
    assert_bit_exact(got, want, "teapot + light")
    unlit = gpu_render(R, R.main_scene(deterministic=det), v, seed=6)
    assert (bits(unlit[0]) != bits(got[0])).any(-1).sum() > 1000       # the light does something
    assert np.array_equal(unlit[1], got[1])                              # ... but not to the primary hits
    v2, ov2 = R.main_viewport(160, 90, 5, 1), O.main_viewport(160, 90, 5, 1)
    assert_bit_exact(gpu_render(R, s, v2, seed=6), _oracle_ext(O, s, O.ACCEL_OCTREE).render(ov2, seed=6), "teapot + light, octree oracle")
    s.set_light(None)                                                    # back to the reference's live integrator
    again = gpu_render(R, s, v, seed=6)
    assert np.array_equal(bits(again[0]), bits(unlit[0]))
    s.release()


@pytest.mark.parametrize("spp,light", [(1, True), (1, False), (3, True)])
def test_circles_scene_analytic_spheres(R, O, spp, light):
    """BASELINE config 1: analytic spheres (Solid / Matte / mirror) over a ground disk, primary + shadow rays + bounces,
    against the oracle; sphere ids follow the triangles'."""
    s = R.circles_scene(n=24, seed=3)
    if not light:
        s.light = None
    w, h = (512, 288) if spp == 1 else (200, 113)
    v, ov = R.main_viewport(w, h, 3, spp), O.main_viewport(w, h, 3, spp)
    got = gpu_render(R, s, v, seed=9)
    assert_bit_exact(got, _oracle_ext(O, s, O.ACCEL_BVH).render(ov, seed=9), "circles")
    ids = np.unique(got[1])
    assert ids.max() >= len(s.tris) and ids.max() < len(s.tris) + 24 and (ids >= len(s.tris)).sum() >= 10
    s.release()


@pytest.mark.parametrize("spp", [1, 2])
def test_extension_wavefront_equals_the_one_kernel_renderer(R, O, spp, monkeypatch):
    """Extension scenes (analytic spheres, shadow rays) run on the wavefront renderer's EXT variants: spheres as a leaf
    record kind, the shadow query as an any-hit ray of the lane's path (scenes of fewer than 4,096 references default to
    the one-kernel renderer, so the threshold is lowered here).  RTB_FLAG_MEGAKERNEL selects round 1's one-thread-per-pixel
    renderer (rtb_ext.cu): same ids, t, colours and ray count — also with RTB_FLAG_BRUTE (no BVH) — and all equal to the
    oracle; the teapot scene with a light likewise."""
    from rust_raytrace_b200 import _lib
    monkeypatch.setenv("RTB_EXT_WAVEFRONT_MIN", "0")
    for s, wh, depth in ((R.circles_scene(n=40, seed=5), (640, 360), 4), (R.main_scene(deterministic=False), (400, 225), 5)):
        if s.spheres is None:
            s.set_light((6.0, -2.0, 0.0), 0.5)
        v = R.main_viewport(*wh, depth, spp)
        a = gpu_render(R, s, v, seed=12, stats=True)
        vm = _lib.RtbView.from_buffer_copy(v)
        vm.flags |= _lib.RTB_FLAG_MEGAKERNEL
        b = gpu_render(R, s, vm, seed=12)
        assert np.array_equal(a[1], b[1]) and np.array_equal(bits(a[2]), bits(b[2])) and np.array_equal(bits(a[0]), bits(b[0]))
        assert a[3].total_rays == b[3].total_rays
        if s.spheres is not None:
            vb = _lib.RtbView.from_buffer_copy(R.main_viewport(160, 90, depth, spp))
            vb.flags |= _lib.RTB_FLAG_BRUTE
            c = gpu_render(R, s, vb, seed=12)
            d = gpu_render(R, s, R.main_viewport(160, 90, depth, spp), seed=12)
            assert np.array_equal(c[1], d[1]) and np.array_equal(bits(c[0]), bits(d[0])) and c[3].total_rays == d[3].total_rays
        assert_bit_exact(a, _oracle_ext(O, s, O.ACCEL_BVH).render(O.main_viewport(*wh, depth, spp), seed=12), "ext wavefront")
        s.release()


def test_sphere_edge_cases(R, O):
    """Camera inside a sphere (far root, back face), a sphere behind the camera (both roots negative), touching
    spheres, a huge sphere around everything; 2K-wide band to cover the config's resolution."""
    col = R.make_color((200, 100, 50))
    sph = np.concatenate([
        R.analytic_sphere([2.0, 0.0, 0.0], 1.5, R.SurfaceKind.Matte(col, 0.4)),          # contains the camera plane
        R.analytic_sphere([2.0, 0.0, -5.0], 1.0, R.SurfaceKind.Solid(col)),              # behind
        R.analytic_sphere([2.0, -1.0, 6.0], 1.0, R.SurfaceKind.Reflective(0.0, col, 0.5)),
        R.analytic_sphere([2.0, 1.0, 6.0], 1.0, R.SurfaceKind.Reflective(0.01, col, 0.5)),   # touches the previous one
        R.analytic_sphere([0.0, 0.0, 0.0], 19.0, R.SurfaceKind.Solid(R.make_color((10, 20, 30)))),
    ])
    s = R.Scene(R.make_dummy_triangle(), boxes=None, spheres=sph, light=((3.0, 0.0, 3.0), 0.2))
    v, ov = R.main_viewport(2560, 64, 4, 1), O.main_viewport(2560, 64, 4, 1)
    assert_bit_exact(gpu_render(R, s, v, seed=2), _oracle_ext(O, s, O.ACCEL_BVH).render(ov, seed=2), "sphere edge cases")
    s.release()


@pytest.mark.parametrize("extra", [[], ["--instanced", "--rgb8"]])
def test_main_rs_equivalent_driver(R, O, scenes, tmp_path, extra):
    """raytrace_b200 (csrc/host/raytrace_main.cpp) does what main.rs:116-227 does — scene, camera, walk_rays, print_stats,
    write_png — over the C ABI: its PNG must hold the oracle's frame through the oracle's quantiser, its "Rays" line the
    oracle's ray count."""
    import os
    import subprocess
    from png_util import decode_png
    from rust_raytrace_b200 import _lib
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "raytrace_b200")
    out = str(tmp_path / "test.png")
    r = subprocess.run([exe, "--mesh", R.raytrace.TEAPOT_MESH, "--size", "320x180", "--out", out, "--seed", "11"] + extra,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rgba, _, _, st = scenes[False][2].render(O.main_viewport(320, 180, 5, 1), seed=11)
    w, h, px = decode_png(out)
    assert (w, h) == (320, 180) and np.array_equal(px, O.quantize_rgb8(rgba.reshape(-1, 4)).reshape(180, 320, 3))
    assert f"Rays: {st.rays}" in r.stdout and "million rays/s" in r.stdout


def test_circles_2k_full_size(R, O):
    """BASELINE config 1 at its own size: 64 analytic spheres + ground disk + light, 2560x1440, maxdepth 2 (primary + one
    bounce) + one shadow ray per hit — every pixel against the oracle (bench.py's `circles2k` workload)."""
    s = R.circles_scene()
    v, ov = R.main_viewport(2560, 1440, 2, 1), O.main_viewport(2560, 1440, 2, 1)
    got = gpu_render(R, s, v, seed=7)
    assert_bit_exact(got, _oracle_ext(O, s, O.ACCEL_BVH).render(ov, seed=7), "circles 2K")
    assert (got[1] >= len(s.tris)).mean() > 0.05          # spheres cover a good part of the frame
    s.release()


def test_all_three_bvh_builders_give_the_same_frame(R, tmp_path):
    """Size-independent property: the closest hit does not depend on the accelerator.  The driver binary is run once per
    topology builder (binned SAH = default, PLOC, Karras radix tree; the knob is read once per process) and once without
    reference splitting; every PNG must hold the same bytes and every run must count the same rays."""
    import os
    import subprocess
    from rust_raytrace_b200 import _lib
    exe = os.path.join(os.path.dirname(_lib.LIB_PATH), "raytrace_b200")
    outs = []
    for k, env in enumerate(({}, {"RTB_BUILDER": "ploc"}, {"RTB_BUILDER": "karras"}, {"RTB_SPLIT_DIV": "0"})):
        out = str(tmp_path / f"b{k}.png")
        r = subprocess.run([exe, "--mesh", R.raytrace.TEAPOT_MESH, "--size", "800x450", "--out", out, "--seed", "5"],
                           capture_output=True, text=True, env={**os.environ, **env})
        assert r.returncode == 0, r.stderr
        rays = [ln for ln in r.stdout.splitlines() if ln.startswith("Rays:")]
        outs.append((open(out, "rb").read(), rays))
    assert all(o == outs[0] for o in outs[1:])



def test_debug_build_runs_clean(R, O, scenes, tmp_path):
    """librtb_debug.so (make debug: -DRTB_DEBUG, device-side bounds asserts on node / reference / stack / queue / slot
    indices — the stand-in for compute-sanitizer, which the GPU pool refuses) renders a frame through every renderer
    variant without tripping an assert, and the frames are the release library's."""
    import os
    import subprocess
    import sys
    from rust_raytrace_b200 import _lib
    dbg = os.path.join(os.path.dirname(_lib.LIB_PATH), "librtb_debug.so")
    assert os.path.exists(dbg), "librtb_debug.so missing: __graft_entry__.build() makes it"
    code = r"""
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
import rust_raytrace_b200 as R
from rust_raytrace_b200 import _lib
assert _lib.LIB_PATH.endswith("librtb_debug.so")
s = R.main_scene(False)
out = []
for fl in (0, _lib.RTB_FLAG_FUSED, _lib.RTB_FLAG_BVH8, _lib.RTB_FLAG_MEGAKERNEL, _lib.RTB_FLAG_STATS):
    v = R.main_viewport(333, 217, 5, 2); v.flags = fl
    data = R.new_image(v)
    c = R.B200RayCaster(want_ids=True, seed=3); c.walk_rays(v, s, data)
    out.append(hashlib.sha256(data.tobytes() + c.prim.tobytes()).hexdigest())
e = R.circles_scene(n=30, seed=2)
for fl in (0, _lib.RTB_FLAG_MEGAKERNEL):
    v = R.main_viewport(200, 120, 3, 1); v.flags = fl
    data = R.new_image(v); R.B200RayCaster(seed=3).walk_rays(v, e, data)
    out.append(hashlib.sha256(data.tobytes()).hexdigest())
print("HASHES", " ".join(out))
""" % os.path.dirname(os.path.dirname(_lib.LIB_PATH))
    res = {}
    for tag, lib in (("release", _lib.LIB_PATH), ("debug", dbg)):
        env = dict(os.environ, RTB_LIB=lib, RTB_BVH8="1", RTB_EXT_WAVEFRONT_MIN="0")
        r = subprocess.run([sys.executable, "-c", code.replace('assert _lib.LIB_PATH.endswith("librtb_debug.so")', "" if tag == "release" else 'assert _lib.LIB_PATH.endswith("librtb_debug.so")')],
                           capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-3000:]
        res[tag] = [ln for ln in r.stdout.splitlines() if ln.startswith("HASHES")][0]
    assert res["debug"] == res["release"]
    h = res["release"].split()[1:]
    assert len(set(h[:5])) == 1 and h[5] == h[6]          # every renderer variant: the same frame
