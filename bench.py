#!/usr/bin/env python
"""bench.py — Mrays/s and ms/frame of the hot path on B200, with the CPU reference beside it.

Workload (BASELINE.json configs[2], the one the metric is quoted on): the main.rs scene
(teapot mesh 6,320 triangles + two 200-triangle mirror disks + dummy = 6,721 `Triangle`s) at
3840x2160, maxdepth 5, 1 spp, shipped materials (Matte teapot, fuzzy Reflective disks).
A "step" is one full frame: primary-ray generation, BVH traversal, exact ray/triangle tests,
bounce shading and accumulation for all 8,294,400 pixels (~14.26 M rays).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # one JSON line (rank 0)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                            # CPU reference arm (oracle port)

value   : Mrays/s of a COMPLETE frame resident on GPU 0.  At N > 1 (one process per GPU) every rank renders its
          8-row bands straight into rank 0's frame buffer (CUDA IPC mapping, stores over NVLink peer access); the
          timed region of a rank is its own stream, CUDA events, max over ranks; rays = project_ray calls with
          depth>0 (the reference's own counter, raytrace.rs:1278), summed over ranks.
parity  : outside the timed region, the frame the timed steps produced is hashed on rank 0 and compared with the
          committed golden hash (the CPU oracle's frame, tests/golden/frame_hashes.json) and, at N > 1, with a
          single-GPU render of the same frame; a mismatch makes the run fail.
e2e     : same metric through the public API B200RayCaster.walk_rays (-> rtb_render) with HOST
          buffers: per step the view goes H2D as kernel parameters and the whole W*H*16-byte image
          comes back D2H into pinned host memory, inside the timed region.  One process drives all N
          GPUs (the drop-in shape); under torchrun rank 0 does it while the other ranks wait.
"""
from __future__ import annotations

import argparse
import ctypes as C
import glob
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 7
FP32_LANES_PER_SM, N_SM = 128, 148

# name -> (description, width, height, maxdepth, spp, partition)
WORKLOADS = {
    # BASELINE.json configs[2]: the configuration the metric is quoted on (default)
    "teapot4k": ("teapot 4K multi-bounce (main.rs scene, 6721 tris, 3840x2160, maxdepth 5, 1 spp, shipped materials)",
                 3840, 2160, 5, 1, "bands"),
    # configs[0]: the "circles" scene — not in the mounted reference; analytic spheres + shadow rays are
    # this build's extension (EXT variants of the wavefront kernel; this 224-primitive scene runs on the one-kernel
    # renderer, which is faster for scenes of a few thousand references or fewer), checked against the oracle's restatement of the same
    "circles2k": ("circles 2K (extension): 64 analytic spheres + ground disk + one cube light, 2560x1440, maxdepth 2 "
                  "(primary + 1 bounce) + one shadow ray per hit, 1 spp", 2560, 1440, 2, 1, "bands"),
    # configs[3]: ~1M triangles, GPU LBVH build + incoherent (mirror) bounces
    "field1m": ("teapot field, 156 instanced teapots = 985,921 tris, Reflective{0}, 2560x1440, maxdepth 5, 1 spp",
                2560, 1440, 5, 1, "bands"),
    # configs[4]: progressive render, samples partitioned over the GPUs + reduce of the accumulation buffers
    "progressive8k": ("8K progressive, main.rs scene, 7680x4320, maxdepth 5, 64 spp, samples partitioned over ranks + "
                      "NCCL reduce(sum) of the f32 accumulation buffers + 1/spp", 7680, 4320, 5, 64, "samples"),
}


def workload_config(name):
    """The `config` object of the JSON line — identical in the GPU arm and the reference arm."""
    desc, W, H, maxdepth, spp, partition = WORKLOADS[name]
    return {"workload": desc, "name": name, "width": W, "height": H, "maxdepth": maxdepth, "spp": spp, "seed": SEED}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), float(p.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


def profiled_dram_traffic(name, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from the newest `ncu --set full`
    summary under profiles/ (written by tools/summarize_profiles.sh for the teapot4k frame at N=1); None for every other
    configuration.  Returns (bytes, source file)."""
    if name != "teapot4k" or world != 1:
        return None, None
    # the newest capture: profiles/CURRENT names its tag (file times do not survive the snapshot that travels to the GPU box)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_k_wf_path_bounce_raw.csv")))
    cur = os.path.join(ROOT, "profiles", "CURRENT")
    if os.path.exists(cur):
        tagged = os.path.join(ROOT, "profiles", open(cur).read().strip() + "_k_wf_path_bounce_raw.csv")
        if os.path.exists(tagged):
            files = [tagged]
    if not files:
        return None, None
    tot, seen = 0.0, 0
    for ln in open(files[-1]):
        f = ln.strip().split(",")
        if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(f[1], None)
            if scale is not None:
                tot += float(f[2]) * scale
                seen += 1
    return (tot, os.path.relpath(files[-1], ROOT)) if seen == 2 else (None, None)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self):
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, gpu_index=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9 or not f[0].isdigit() or int(f[0]) != gpu_index:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------
# CPU reference arm: the oracle restatement running the reference's octree algorithm.  Nothing of the product package
# is imported here: the scenes come from oracle.py's own generators (tests/test_host.py compares their bytes with the
# product's), so the only native code this arm loads is oracle/build/liboracle.so.
# --------------------------------------------------------------------------------------------
def oracle_scene(O, name, cores):
    """Reference-algorithm mode (octree) wherever the reference's builder can handle the scene; the 1M-triangle
    scene cannot be built as an octree in reasonable time, there the oracle uses its own BVH."""
    verts, faces = O.load_mesh_bin()
    accel = O.ACCEL_BVH if name == "field1m" else O.ACCEL_OCTREE
    if name == "circles2k":
        tris, spheres, light = O.circles_scene_parts()
        osc = O.Scene(tris, accel, build_threads=cores).add_spheres(spheres).set_light(*light)
    elif name == "field1m":
        osc = O.Scene(O.teapot_field_tris(verts, faces), accel, build_threads=cores)
    else:
        osc = O.Scene(O.main_scene_tris(verts, faces, False), accel, build_threads=cores)
    return osc, ("bvh" if accel == O.ACCEL_BVH else "octree")


def accel_text(accel):
    return ("reference-algorithm mode (octree 204,894 nodes, row work-queue)" if accel == "octree"
            else "BVH mode (the reference octree cannot be built for 1M triangles; row work-queue)")


def cpu_reference(name, rows, threads=None, scene=None):
    """Times the oracle on image rows [rows[0], rows[1]) of the workload's frame (all samples).
    Returns (Mrays/s, rays, seconds, cores, accel, scene)."""
    from oracle import oracle as O
    desc, W, H, maxdepth, spp, _ = WORKLOADS[name]
    cores = threads or os.cpu_count() or 1
    osc, accel = scene if scene else oracle_scene(O, name, os.cpu_count() or 1)   # build excluded, as main.rs:160 vs :191
    ov = O.main_viewport(W, H, maxdepth, spp)
    _, _, _, st = osc.render(ov, seed=SEED, threads=cores, rows=rows, want_ids=False)
    return st.rays / st.seconds / 1e6, int(st.rays), float(st.seconds), cores, accel, (osc, accel)


def cpu_sample_rows(name):
    """A bounded band through the middle of the frame (teapot + both disks), ~10-30 s of CPU work."""
    desc, W, H, maxdepth, spp, _ = WORKLOADS[name]
    if name == "teapot4k":
        return (0, H)                 # the whole frame: 14.26 M rays, ~9 s on 16 host threads
    if name == "field1m":
        return (H // 2 + 100, H // 2 + 164)
    if name == "circles2k":
        return (0, H)                 # 64 spheres by brute force + a 160-triangle octree: a few seconds
    return (H // 2, H // 2 + 4)       # 64 spp: 4 rows x 7680 px x 64 samples


def cpu_one_thread_rows(name):
    """The 1-thread figure (the reference's `threads` argument, main.rs:191) on a band 1/16 the size."""
    r0, r1 = cpu_sample_rows(name)
    mid, n = (r0 + r1) // 2, max(1, (r1 - r0) // 16)
    return (mid - n // 2, mid - n // 2 + n)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    name = args.workload
    desc, W, H, maxdepth, spp, _ = WORKLOADS[name]
    cores = os.cpu_count() or 1
    osc, accel = oracle_scene(O, name, cores)
    ov = O.main_viewport(W, H, maxdepth, spp)
    # bounded sample per step, sized so that warmup+steps stay within a few minutes whatever the host:
    # ~2 s of work at the rate of a small probe band through the middle of the frame
    mid = H // 2 + (100 if name == "field1m" else 0)      # the field's teapots fill the lower half of the frame
    probe_rows = 16 if spp == 1 else 1
    _, _, _, probe = osc.render(ov, seed=SEED, threads=cores, rows=(mid, mid + probe_rows), want_ids=False)
    n_rows = int(min(H - mid, max(probe_rows, probe_rows * 2.0 / max(probe.seconds, 1e-3))))
    if spp == 1:
        n_rows = max(8, n_rows // 8 * 8)
    rows = (max(0, mid - n_rows // 2), min(H, max(0, mid - n_rows // 2) + n_rows)) if spp == 1 else (mid, min(H, mid + n_rows))
    rates, secs, rays = [], [], 0
    for i in range(args.warmup + args.steps):
        _, _, _, st = osc.render(ov, seed=SEED, threads=cores, rows=rows, want_ids=False)
        rays = int(st.rays)
        if i >= args.warmup:
            rates.append(st.rays / st.seconds / 1e6); secs.append(st.seconds)
    value = float(np.mean(rates))
    sample = f"image rows {rows[0]}..{rows[1]} of {H} ({rays} rays per step), oracle in {accel_text(accel)}"
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(secs) * 1e3),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name),
        "details": {"sample": f"rows {rows[0]}..{rows[1]} of {H} per step", "rays_per_step": rays},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def build_scene(R, name):
    if name == "circles2k":
        return R.circles_scene()
    return R.teapot_field_scene() if name == "field1m" else R.main_scene(deterministic=False)


class DevFrame:
    """A raw device allocation (rtb_device_alloc or an IPC mapping of one) seen by torch without a copy."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def run_gpu(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # CPU-side barrier for the e2e leg: an NCCL barrier parks a spinning kernel on every waiting rank's GPU,
        # which would compete with the kernels rank 0 launches on those GPUs while it drives all N of them
        host_group = dist.new_group(backend="gloo")

    import rust_raytrace_b200 as R
    from rust_raytrace_b200 import _lib, dist as RD
    L = _lib.lib()

    name = args.workload
    desc, W, H, maxdepth, spp, partition = WORKLOADS[name]
    by_samples = partition == "samples"
    dev_ids = (C.c_int * 1)(local_rank)
    _lib.check(L.rtb_init(1, dev_ids), "rtb_init")
    scene = build_scene(R, name)
    h = scene.upload()
    scene.release()                 # the first build of a process pays CUDA's lazy kernel loading: report the second
    h = scene.upload()
    info = scene.info()
    view = R.main_viewport(W, H, maxdepth, spp)
    view.seed = SEED
    npix = W * H
    if by_samples:
        my_view = RD.sample_view(view, rank, world)           # samples [b, e) of spp, RTB_FLAG_SUM_ONLY
        tile_rank, tile_world = 0, 1
    else:
        my_view = view
        tile_rank, tile_world = rank, world

    # The frame.  Bands over N > 1 processes: ONE frame buffer on rank 0's GPU, mapped into every other rank with CUDA IPC;
    # each rank's kernels store their bands into it over NVLink.  Samples over ranks: every rank owns a full sum buffer
    # and NCCL reduces them to rank 0.
    shared_frame = world > 1 and not by_samples
    gathered_frame = False          # fallback without peer access: per-rank buffers, gathered outside the timed region
    frame_ptr = C.c_void_p()
    if shared_frame:
        handle = torch.zeros(64, dtype=torch.uint8, device="cuda")
        if rank == 0:
            _lib.check(L.rtb_device_alloc(0, npix * 16, C.byref(frame_ptr)), "rtb_device_alloc")
            hbuf = C.create_string_buffer(64)
            _lib.check(L.rtb_ipc_export(frame_ptr, hbuf), "rtb_ipc_export")
            handle.copy_(torch.tensor(list(hbuf.raw), dtype=torch.uint8))
        dist.broadcast(handle, 0)
        opened = torch.ones(1, dtype=torch.int32, device="cuda")
        if os.environ.get("RTB_BENCH_NO_IPC"):          # test knob: behave as on a box without peer access
            opened.zero_()
        elif rank != 0 and L.rtb_ipc_open(0, bytes(handle.cpu().tolist()), C.byref(frame_ptr)) != 0:
            print(f"bench.py: rank {rank}: {L.rtb_last_error().decode()}", file=sys.stderr, flush=True)
            frame_ptr = C.c_void_p()
            opened.zero_()
        dist.all_reduce(opened, op=dist.ReduceOp.MIN)
        if int(opened.item()) == 0:     # some pair of GPUs has no peer access: every rank keeps its bands in its own buffer
            if rank != 0 and frame_ptr.value:
                L.rtb_ipc_close(0, frame_ptr)
            shared_frame, gathered_frame = False, True
    if shared_frame:
        d_rgba = torch.as_tensor(DevFrame(frame_ptr.value, (H, W, 4)), device="cuda") if rank == 0 else None
        d_rgba_ptr = frame_ptr.value
    else:
        d_rgba = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        d_rgba_ptr = d_rgba.data_ptr()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    # an explicit (non-null) stream: handle 0 would mean "the library's own stream" to rtb_render_device,
    # and torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    sp = C.c_void_p(stream.cuda_stream)

    def render(v, stats=None, ptr=None, tr=None, tw=None):
        if v is not None:
            _lib.check(L.rtb_render_device(h, C.byref(v), 0, tile_rank if tr is None else tr, tile_world if tw is None else tw,
                                           d_rgba_ptr if ptr is None else ptr, None, None, sp, stats), "rtb_render_device")

    def step():
        render(my_view)
        if by_samples:
            if my_view is None:
                d_rgba.zero_()
            RD.reduce_samples(d_rgba, spp)                     # NCCL reduce(sum) to rank 0 + 1/spp (rtb_scale_device)

    def with_flags(v, flags):
        if v is None:
            return None
        c = _lib.RtbView.from_buffer_copy(v)
        c.flags |= flags
        return c

    # one counted frame: rays, then (STATS kernel variant) node / triangle tests per ray, per phase
    # (counting / timing renders go to a scratch buffer: the frame buffer keeps what the timed steps produced)
    scratch = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    sptr = scratch.data_ptr()
    st = _lib.RtbStats()
    render(my_view, C.byref(st), ptr=sptr)
    my_rays = int(st.rays)
    st2 = _lib.RtbStats()
    render(with_flags(my_view, _lib.RTB_FLAG_STATS), C.byref(st2), ptr=sptr)

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler()
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    t_wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(0)              # L2 flush between timed iterations (outside the events)
        a.record(stream)
        step()
        b.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_mine = float(np.mean(ms_steps))

    # per-phase device times of the same step, live (CUDA events on the launching stream between the two launches of the
    # path kernel), still under the clock sampler; used for the roofline of the dominant launch
    stage_ms = np.zeros(4)
    n_timing = 0
    if my_view is not None:
        stt = _lib.RtbStats()
        tv = with_flags(my_view, _lib.RTB_FLAG_TIMING)
        for _ in range(5):
            flush.fill_(0)
            render(tv, C.byref(stt), ptr=sptr)
            stage_ms += np.array(stt.ms_stage[:])
            n_timing += 1
        stage_ms /= n_timing
    clocks = sampler.stop(local_rank) if rank == 0 else None

    tens = torch.tensor([ms_mine, float(my_rays)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tens.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tens.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_step, total_rays = float(mx[0]), float(sm[1])
    else:
        ms_step, total_rays = ms_mine, float(my_rays)
    value = total_rays / (ms_step * 1e-3) / 1e6

    # ---- parity of the frame the timed steps left on rank 0's GPU (outside the timed region) ----
    parity = None
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if gathered_frame:                  # disjoint bands, zeros elsewhere: an integer sum of the bit patterns is a gather
        dist.reduce(d_rgba.view(torch.int32), dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    if rank == 0:
        frame_host = d_rgba.cpu().numpy()
        parity = {"checked": True, "frame_sha": hashlib.sha256(np.ascontiguousarray(frame_host).tobytes()).hexdigest(),
                  "frame_on": "GPU 0" + (" (written by all ranks over NVLink peer mappings)" if shared_frame else
                                          " (NO peer access on this box: per-rank band buffers, gathered outside the timed region)"
                                          if gathered_frame else "")}
        try:
            with open(os.path.join(ROOT, "tests", "golden", "frame_hashes.json")) as fh:
                gold = json.load(fh)["frames"].get(name)
        except Exception:
            gold = None
        if gold is not None and not by_samples:
            parity["golden_sha"] = gold["sha256"]
            parity["equals_golden"] = parity["frame_sha"] == gold["sha256"] and int(total_rays) == int(gold["rays"])
            parity["golden"] = "CPU oracle frame, tests/golden/frame_hashes.json (make_frame_hashes.py)"
        if world > 1:
            # the same frame rendered by this GPU alone
            one = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
            st1 = _lib.RtbStats()
            render(view, C.byref(st1), ptr=one.data_ptr(), tr=0, tw=1)
            torch.cuda.synchronize()
            if by_samples:        # the summation order differs (sum of per-rank partial sums): rounding-level differences
                diff = (one - d_rgba).abs().max().item()
                mse = ((one - d_rgba) ** 2).mean().item()
                parity["max_abs_diff_vs_n1"] = diff
                parity["psnr_vs_n1_db"] = float(10 * np.log10(1.0 / max(mse, 1e-30)))
                parity["equals_n1"] = bool(diff <= 2e-6)
            else:
                parity["equals_n1"] = bool(torch.equal(one.view(torch.int32), d_rgba.view(torch.int32)))
            parity["rays_equal_n1"] = int(st1.rays) == int(total_rays)
            del one
    if world > 1:
        dist.barrier()

    # ---- e2e through the public API with host buffers (rank 0 drives all N GPUs) ----
    e2e = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
    if rank == 0:
        ids = (C.c_int * args.gpus)(*range(args.gpus))
        scene.release()
        _lib.check(L.rtb_init(args.gpus, ids), "rtb_init(all)")
        sc_all = build_scene(R, name)
        caster = R.B200RayCaster(seed=SEED)
        caster._n_gpus = args.gpus
        host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()
        data = host.numpy()
        vv = R.main_viewport(W, H, maxdepth, spp)
        call = (lambda: caster.walk_rays_progressive(vv, sc_all, data, threads=args.gpus)) if by_samples else \
               (lambda: caster.walk_rays(vv, sc_all, data, threads=args.gpus))
        e_steps = args.steps if not by_samples else max(1, min(args.steps, 3))
        for _ in range(max(args.warmup, 3) if not by_samples else 1):
            call()
        t0 = time.perf_counter()
        rays_e2e = 0
        for _ in range(e_steps):
            rays_e2e += call().total_rays
        dt = time.perf_counter() - t0
        e2e = {"value": rays_e2e / dt / 1e6, "unit": "Mrays/s", "ms_per_frame": dt / e_steps * 1e3,
               "h2d_bytes_per_step": C.sizeof(_lib.RtbView), "d2h_bytes_per_step": npix * 16,
               "api": ("B200RayCaster.walk_rays_progressive -> rtb_render_progressive" if by_samples else
                       "B200RayCaster.walk_rays -> rtb_render") + " (pinned host image, scene resident)"}
        assert int(caster.stats.rays) == int(total_rays), (caster.stats.rays, total_rays)
        if parity is not None and not by_samples:
            parity["e2e_equals_device_frame"] = bool(np.array_equal(data.view(np.uint32), frame_host.view(np.uint32)))
        if by_samples:
            e2e["ms_reduce"] = float(caster.stats.ms_reduce)     # rtb_render_progressive: cross-GPU reduce + copy home, max over GPUs
        if not by_samples:
            # the device-to-host floor of the same call: every copy and event of the frame, no kernel (RTB_FLAG_COPY_ONLY)
            vc = _lib.RtbView.from_buffer_copy(vv)
            vc.flags |= _lib.RTB_FLAG_COPY_ONLY
            scratch_host = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory().numpy()
            for _ in range(3):
                caster.walk_rays(vc, sc_all, scratch_host, threads=args.gpus)
            t0 = time.perf_counter()
            for _ in range(e_steps):
                caster.walk_rays(vc, sc_all, scratch_host, threads=args.gpus)
            e2e["d2h_floor_ms"] = (time.perf_counter() - t0) / e_steps * 1e3
            del scratch_host
            # the same frame the way main.rs consumes it (walk_rays -> write_png): quantised on the device, 3 B/px home
            host8 = torch.zeros((H, W, 3), dtype=torch.uint8).pin_memory()
            rgb = host8.numpy()
            for _ in range(3):
                caster.walk_rays_rgb8(vv, sc_all, rgb, threads=args.gpus)
            t0 = time.perf_counter()
            rays8 = 0
            for _ in range(e_steps):
                rays8 += caster.walk_rays_rgb8(vv, sc_all, rgb, threads=args.gpus).total_rays
            dt8 = time.perf_counter() - t0
            e2e["rgb8"] = {"value": rays8 / dt8 / 1e6, "unit": "Mrays/s", "ms_per_frame": dt8 / e_steps * 1e3,
                           "d2h_bytes_per_step": npix * 3,
                           "api": "B200RayCaster.walk_rays_rgb8 -> rtb_render_rgb8 (write_png's quantiser fused on the device)"}
        # scene assembly (the step in front of the path): finished host triangles vs mesh + instances assembled on the GPU
        if name != "circles2k":
            inf_host = sc_all.info()
            sc_inst = R.teapot_field_scene(instanced=True) if name == "field1m" else R.main_scene(deterministic=False, instanced=True)
            inf_inst = sc_inst.info()
            e2e["scene_assembly"] = {
                "host_triangle_array": {"ms_upload_and_cull": inf_host.ms_upload, "h2d_bytes": int(inf_host.n_tris) * 140},
                "instanced_on_gpu": {"ms_assemble_and_cull": inf_inst.ms_upload,
                                     "h2d_bytes": int(sc_inst.verts.nbytes + sc_inst.faces.nbytes + 80 * len(sc_inst.instances)
                                                      + sc_inst.extra.nbytes + 140)},
                "n_tris": int(inf_inst.n_tris), "n_refs": int(inf_inst.n_refs), "ms_build": inf_inst.ms_build}
            sc_inst.release()
        sc_all.release()
    if world > 1:
        dist.barrier(group=host_group)

    ok = True
    if rank == 0:
        hbm_peak, sm_max_mhz, peak_src = measured_peaks()
        # algorithmic FP32 lane-ops / bytes per ray = 24 ops, 32 B per AABB test; 54 ops, 80 B per exact
        # triangle test; 45 ops per generated ray, 16 B per output pixel.  Dominant launch = the bounce phase of
        # k_wf_path on rank 0.
        n_node, n_tri = st2.node_tests / max(st2.rays, 1), st2.tri_tests / max(st2.rays, 1)
        # (extension scenes run the EXT variant of the same kernel: its tests include the shadow rays', its rays do not)
        dom_kernel, b_rays = "k_wf_path<bounce phase>", max(int(st2.bounce_rays), 1)
        nb_node, nb_tri = st2.node_tests_bounce / b_rays, st2.tri_tests_bounce / b_rays
        if name == "circles2k":      # 224 primitives (~1,500 references): below the wavefront threshold, one kernel per frame (rtb_ext.cu)
            dom_kernel, b_rays, nb_node, nb_tri = "k_trace_ext", max(int(st2.rays), 1), n_node, n_tri
            stage_ms = np.array([0.0, 0.0, 0.0, ms_mine])
        elif stage_ms[3] < stage_ms[1]:           # the primary phase dominates (maxdepth 2): report that launch
            dom_kernel = dom_kernel.replace("bounce", "primary")
            b_rays = max(int(st2.rays - st2.bounce_rays), 1)
            nb_node, nb_tri = (st2.node_tests - st2.node_tests_bounce) / b_rays, (st2.tri_tests - st2.tri_tests_bounce) / b_rays
            stage_ms = np.array([stage_ms[0], stage_ms[3], stage_ms[2], stage_ms[1]])   # [3] = the dominant launch below
        launch_s = max(stage_ms[3], 1e-6) * 1e-3
        bytes_launch = b_rays * (nb_node * 32 + nb_tri * 80 + 16)
        ops_launch = b_rays * (nb_node * 24 + nb_tri * 54 + 45)
        fp32_peak = N_SM * FP32_LANES_PER_SM * sm_max_mhz * 1e6
        step_ops = my_rays * (n_node * 24 + n_tri * 54 + 45)
        traffic, traffic_src = profiled_dram_traffic(name, world)
        cpu = None
        if world == 1 and not args.no_cpu:
            rows = cpu_sample_rows(name)
            mr, rays, s, cores, accel, osc = cpu_reference(name, rows)
            rows1 = cpu_one_thread_rows(name)
            mr1, rays1, s1, _, _, _ = cpu_reference(name, rows1, threads=1, scene=osc)
            cpu = {"value": mr, "unit": "Mrays/s", "cores": cores, "kind": "port",
                   "sample": f"image rows {rows[0]}..{rows[1]} of the same frame ({rays} rays, {s:.1f} s), oracle in "
                             + accel_text(accel) + ", all host threads",
                   "one_thread": {"value": mr1, "unit": "Mrays/s", "cores": 1,
                                  "sample": f"image rows {rows1[0]}..{rows1[1]} ({rays1} rays, {s1:.1f} s), same oracle, 1 thread"}}
        launches_per_step = int(st.kernel_launches) + (1 if by_samples else 0)
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "ms_per_frame": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name),
            "details": {"rays_per_frame": int(total_rays),
                        "partition": (f"samples: rank r renders samples [64r/{world}, 64(r+1)/{world}) of the full frame, then NCCL reduce to rank 0"
                                      if by_samples else f"8-row bands, band b -> rank b % {world}; every rank stores into the one frame on GPU 0"),
                        "l2": "flushed between timed iterations (256 MiB fill); the scene is re-read from HBM each step",
                        "accelerator": os.environ.get("RTB_BVH", "4") + "-wide BVH",
                        "bvh": {"nodes": info.n_nodes, "leaves": info.n_leaves, "max_leaf": info.max_leaf,
                                "height": info.tree_height, "prims": info.n_prims, "refs": info.n_refs,
                                "ms_build": info.ms_build, "ms_upload": info.ms_upload},
                        "node_tests_per_ray": n_node, "tri_tests_per_ray": n_tri, "wall_s_timed_region": t_wall},
            "gpu_launches": args.steps * launches_per_step * world,
            "clocks": clocks,
            "parity": parity,
            "e2e": e2e,
            "stages_ms": {k: float(v) for k, v in zip(_lib.RTB_STAGES, stage_ms)},
            "roofline": {"bound": "fp32_issue", "kernel": dom_kernel, "achieved": ops_launch / launch_s / 1e12,
                         "peak": fp32_peak / 1e12, "unit": "Tlane-op/s", "frac": ops_launch / launch_s / fp32_peak,
                         "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": f"{N_SM} SMs x {FP32_LANES_PER_SM} FP32 lanes x {sm_max_mhz:.0f} MHz (non-FMA: the exactness contract forbids contraction; SURVEY 8d)",
                         "launch_ms": float(stage_ms[3]), "share_of_step": float(stage_ms[3] / max(stage_ms.sum(), 1e-9)),
                         "algorithmic_lane_ops_per_launch": ops_launch,
                         "lane_ops_per_ray": nb_node * 24 + nb_tri * 54 + 45,
                         "per_ray": {"aabb_tests": nb_node, "tri_tests": nb_tri, "rays": b_rays},
                         "whole_step_frac": step_ops / (ms_mine * 1e-3) / fp32_peak},
            "roofline_hbm": {"bound": "hbm", "kernel": dom_kernel, "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": bytes_launch,
                             "algorithmic_gbs": bytes_launch / launch_s / 1e9,
                             "measured_dram_bytes_per_launch": traffic, "measured_source": traffic_src,
                             "achieved": (traffic / launch_s / 1e9) if traffic else None,
                             "frac": (traffic / launch_s / 1e9 / hbm_peak) if traffic else None,
                             "note": "the scene is cache resident: algorithmic bytes are served by L1/L2, DRAM sees the "
                                     "workspace streaming through; HBM is not the bound of this path (SURVEY 8d)"},
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
        if parity is not None:
            bad = [k for k in ("equals_golden", "equals_n1", "rays_equal_n1", "e2e_equals_device_frame") if parity.get(k) is False]
            if bad:
                print(f"bench.py: PARITY FAILURE: {bad}", file=sys.stderr, flush=True)
                ok = False
    if shared_frame and rank != 0:
        L.rtb_ipc_close(0, frame_ptr)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="teapot4k", choices=sorted(WORKLOADS),
                    help="teapot4k = the configuration BASELINE.json's metric is quoted on (default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
