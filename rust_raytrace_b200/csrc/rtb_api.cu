// rtb_api.cu — the C ABI of include/rtb.h: device selection, scene upload + LBVH build,
// frame rendering into host or device buffers, progressive multi-GPU rendering.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>

#include "rtb_internal.cuh"
#include "host/raytrace_host.hpp"

namespace {

thread_local std::string g_err;
std::mutex g_mu;
std::vector<int> g_devices;   // devices selected by rtb_init
bool g_p2p[RTB_MAX_GPUS][RTB_MAX_GPUS] = {};   // [a][b]: GPU slot a can load from / store to GPU slot b's memory (rtb_init)

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// One persistent host thread per GPU slot for the multi-GPU entry points: a frame is ~15 runtime calls per GPU, and
// issuing them for 8 GPUs from one thread (or from threads created per call) cost more than the frame itself
// (8 GPUs, 4K teapot frame: 1.88 ms per rtb_render call for 0.45 ms of device time).
class Worker {
public:
    Worker() : th_([this] { loop(); }) {}
    ~Worker() {
        { std::lock_guard<std::mutex> lk(mu_); quit_ = true; }
        cv_.notify_all();
        th_.join();
    }
    void start(std::function<void()> job) {
        { std::lock_guard<std::mutex> lk(mu_); job_ = std::move(job); busy_ = true; }
        cv_.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return !busy_; });
    }
private:
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [this] { return quit_ || (busy_ && job_); });
                if (quit_) return;
                job = std::move(job_);
                job_ = nullptr;
            }
            job();
            { std::lock_guard<std::mutex> lk(mu_); busy_ = false; }
            done_cv_.notify_all();
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    std::function<void()> job_;
    bool busy_ = false, quit_ = false;
    std::thread th_;
};
std::vector<std::unique_ptr<Worker>> g_workers;
std::mutex g_render_mu;     // one multi-GPU frame at a time uses the workers

void ensure_workers(size_t n) {
    while (g_workers.size() < n) g_workers.emplace_back(new Worker());
}

}  // namespace

void rtb_set_error(const std::string& msg) { g_err = msg; }

int rtb_cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    g_err = std::string("CUDA error ") + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ") in " + what + " at " +
            file + ":" + std::to_string(line);
    return (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? RTB_ERR_NO_DEVICE
           : (e == cudaErrorMemoryAllocation)                            ? RTB_ERR_NOMEM
                                                                          : RTB_ERR_CUDA;
}

namespace {

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

int ensure_init() {
    if (!g_devices.empty()) return RTB_OK;
    return rtb_init(1, nullptr);
}

void free_gpu_scene(GpuScene& g) {
    if (g.device < 0) return;
    cudaSetDevice(g.device);
    if (g.stream) cudaStreamSynchronize(g.stream);
    cudaFree(g.d_nodes8); cudaFree(g.d_tri8); cudaFree(g.d_shade8);
    cudaFree(g.d_nodes); cudaFree(g.d_nodes4); cudaFree(g.d_tri); cudaFree(g.d_shade); cudaFree(g.d_prim_order); cudaFree(g.d_counters);
    cudaFree(g.d_rgba); cudaFree(g.d_prim); cudaFree(g.d_t); cudaFree(g.d_rgb8);
    cudaFree(g.d_stage);
    if (g.prog_done) cudaEventDestroy(g.prog_done);
    if (g.red0) cudaEventDestroy(g.red0);
    if (g.red1) cudaEventDestroy(g.red1);
    for (auto& l : g.lanes) {
        cudaFree(l.d_ws);
        if (l.busy) cudaEventDestroy(l.busy);
        if (l.done) cudaEventDestroy(l.done);
        for (auto& e : l.stage_ev) if (e) cudaEventDestroy(e);
        if (l.st) cudaStreamDestroy(l.st);
    }
    if (g.fork_ev) cudaEventDestroy(g.fork_ev);
    if (g.ev0) cudaEventDestroy(g.ev0);
    if (g.ev1) cudaEventDestroy(g.ev1);
    for (auto& e : g.chunk_ev) if (e) cudaEventDestroy(e);
    if (g.copy_stream) cudaStreamDestroy(g.copy_stream);
    if (g.stream) cudaStreamDestroy(g.stream);
    g = GpuScene();
}

int ensure_framebuffer(GpuScene& g, size_t pixels, bool want_prim, bool want_t) {
    if (g.fb_pixels < pixels) {
        cudaFree(g.d_rgba); cudaFree(g.d_prim); cudaFree(g.d_t);
        g.d_rgba = nullptr; g.d_prim = nullptr; g.d_t = nullptr; g.fb_pixels = 0;
        RTB_CUDA(cudaMalloc(&g.d_rgba, pixels * sizeof(float4)));
        g.fb_pixels = pixels;
    }
    if (want_prim && !g.d_prim) RTB_CUDA(cudaMalloc(&g.d_prim, g.fb_pixels * sizeof(uint32_t)));
    if (want_t && !g.d_t) RTB_CUDA(cudaMalloc(&g.d_t, g.fb_pixels * sizeof(float)));
    return RTB_OK;
}

int check_view(const RtbView* v) {
    if (!v) return fail(RTB_ERR_INVALID, "view is NULL");
    if (v->width == 0 || v->height == 0) return fail(RTB_ERR_INVALID, "viewport has zero width or height");
    if (v->maxdepth == 0) return fail(RTB_ERR_INVALID, "maxdepth must be >= 1");
    if (v->maxdepth > RTB_MAX_DEPTH)
        return fail(RTB_ERR_INVALID, "maxdepth " + std::to_string(v->maxdepth) + " exceeds RTB_MAX_DEPTH");
    if (v->spp == 0) return fail(RTB_ERR_INVALID, "samples_per_pixel must be >= 1");
    if (v->sample_end > v->spp || v->sample_begin > v->sample_end)
        return fail(RTB_ERR_INVALID, "sample range outside [0, spp]");
    return RTB_OK;
}

// The conservative slab tests compute t = fma(plane, 1/d, -o/d): the rounding of o/d moves a plane by up to |o| * 2^-23,
// which the builder's padding (2^-16 of the scene's largest |coordinate| for the BVH4, 2^-17 for the BVH2) covers only while the ray origins stay within
// 32x the scene's extent.  Bounce rays start on surfaces; primary rays start on the viewport plane — checked here.
int check_camera(const rtb_scene* s, const RtbView* v) {
    float max_abs = 0.f;
    for (int k = 0; k < 3; ++k) max_abs = std::max(max_abs, std::max(std::fabs(s->info.scene_lo[k]), std::fabs(s->info.scene_hi[k])));
    if (s->info.n_prims == 0 || !(max_abs > 0.f)) return RTB_OK;
    float far = 0.f;
    for (int k = 0; k < 3; ++k) {
        // the four corners of the viewport: orig, orig + vu, orig + vv, orig + vu + vv
        const float c[4] = {v->orig[k], v->orig[k] + v->vu[k], v->orig[k] + v->vv[k], v->orig[k] + v->vu[k] + v->vv[k]};
        for (float x : c) far = std::max(far, std::fabs(x));
    }
    if (!(far <= 32.0f * max_abs))
        return fail(RTB_ERR_INVALID, "viewport farther from the origin than 32x the scene's largest coordinate (" + std::to_string(far) +
                                         " vs " + std::to_string(max_abs) + "): beyond the range in which the BVH's box padding keeps the "
                                         "traversal exact; move the camera closer or translate the scene (INTEGRATION.md, limits)");
    return RTB_OK;
}

ViewDev make_view(const RtbView& v, uint32_t rank, uint32_t world, bool compact) {
    ViewDev d;
    std::memset(&d, 0, sizeof d);
    d.width = v.width; d.height = v.height;
    for (int k = 0; k < 3; ++k) { d.orig[k] = v.orig[k]; d.cam[k] = v.cam[k]; d.vu[k] = v.vu[k]; d.vv[k] = v.vv[k]; }
    d.maxdepth = v.maxdepth; d.spp = v.spp; d.seed = v.seed;
    d.s_begin = v.sample_begin; d.s_end = v.sample_end;
    if (d.s_begin == 0 && d.s_end == 0) d.s_end = v.spp;
    d.flags = v.flags;
    d.tile_rank = rank; d.tile_world = world;
    d.tiles_x = (v.width + RTB_TILE_W - 1) / RTB_TILE_W;
    const uint32_t tiles_y = (v.height + RTB_TILE_H - 1) / RTB_TILE_H;
    d.my_tile_rows = tiles_y > rank ? (tiles_y - rank + world - 1) / world : 0;
    d.compact = compact ? 1u : 0u;
    // the per-frame constants of pixel_ray, the RNG seeding and the slot -> pixel map (host code is compiled without
    // contraction or fast-math: the same IEEE division and products the device would do per ray)
    const float inv_w = 1.0f / (float)v.width, inv_h = 1.0f / (float)v.height;
    for (int k = 0; k < 3; ++k) { d.vu_delta[k] = v.vu[k] * inv_w; d.vv_delta[k] = v.vv[k] * inv_h; }
    uint64_t z = v.seed + 0x9E3779B97F4A7C15ull;                          // splitmix64, as rtb_device.cuh
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    d.seed_mixed = z ^ (z >> 31);
    const uint32_t tiles_x8 = (v.width + 7u) / 8u;
    d.div_band = rtb_udiv_make(2u * tiles_x8);
    d.div_tx8 = rtb_udiv_make(tiles_x8);
    return d;
}

SceneDev scene_dev(const GpuScene& g, uint32_t n_prims) {
    SceneDev s;
    s.nodes = g.d_nodes; s.tri = g.d_tri; s.shade = g.d_shade; s.n_prims = n_prims; s.n_nodes = g.n_nodes;
    s.height = g.height;
    s.nodes4 = g.d_nodes4;
    s.stack4 = g.stack4_need ? g.stack4_need + 1u : 3u * (g.depth4 + 1u) + 2u;    // exact bound from the builder, else 3 per level
    s.nodes8 = g.d_nodes8; s.tri8 = g.d_tri8; s.shade8 = g.d_shade8; s.depth8 = g.depth8;
    s.n_nodes4 = g.n_nodes4; s.n_nodes8 = g.n_nodes8;
    return s;
}

// Pixels of the image this rank renders (rows of its 8-row bands that lie inside the image).
uint64_t owned_pixels(const ViewDev& vd) {
    uint64_t rows = 0;
    for (uint32_t b = vd.band_begin; b < vd.band_begin + vd.my_tile_rows; ++b) {
        const uint32_t row0 = (b * vd.tile_world + vd.tile_rank) * RTB_TILE_H;
        if (row0 < vd.height) rows += std::min<uint32_t>(RTB_TILE_H, vd.height - row0);
    }
    return rows * vd.width;
}

// One frame (all samples, all bounces) of this rank's bands on stream `st`.  Default: the wavefront
// pipeline; RTB_FLAG_MEGAKERNEL selects the one-kernel renderer.  *primary_rays is what must be added to
// counters->rays afterwards (the wavefront counts bounce rays on the device, primaries on the host).
int launch_frame(GpuScene& g, GpuLane& lane, uint32_t n_prims, const ViewDev& vd, float4* d_rgba, uint32_t* d_prim,
                 float* d_t, cudaStream_t st, uint32_t* launches, uint64_t* primary_rays) {
    const bool is_ext = g.has_spheres || g.ext.has_light;      // EXTENSION scenes: analytic spheres / shadow rays
    // Their default renderer is the wavefront kernel's EXT variant, except for small scenes: with a few hundred primitives a
    // ray costs a handful of node visits and the per-path queue traffic of the wavefront design outweighs what it saves
    // (B200, circles scene, 224 primitives = ~1,500 references, 2K, maxdepth 2 / 3 / 5: one kernel 0.58 / 0.83 / 1.35 ms, wavefront 0.79 /
    // 1.06 / 1.50 ms; teapot scene + light, 7,120 references, 4K: one kernel 5.34 ms, wavefront 3.93 ms).
    // RTB_EXT_WAVEFRONT_MIN (references, read per call) moves the threshold; RTB_FLAG_MEGAKERNEL forces the one-kernel renderer.
    if (is_ext) {
        const char* e = getenv("RTB_EXT_WAVEFRONT_MIN");
        const uint32_t min_refs = e ? (uint32_t)std::max(0, atoi(e)) : 4096u;
        if ((vd.flags & RTB_FLAG_MEGAKERNEL) || n_prims < min_refs)
            return rtb_launch_trace_ext(scene_dev(g, n_prims), vd, g.ext, d_rgba, d_prim, d_t, g.d_counters, st, launches);
    }
    if (vd.flags & RTB_FLAG_MEGAKERNEL)
        return rtb_launch_trace(scene_dev(g, n_prims), vd, d_rgba, d_prim, d_t, g.d_counters, st, launches);
    const uint32_t n_slots = vd.my_tile_rows * 2u * ((vd.width + 7u) / 8u) * 32u;
    const size_t need = rtb_wf_workspace_bytes(n_slots, vd.maxdepth ? vd.maxdepth : 1, (vd.s_end - vd.s_begin) > 1);
    if (lane.ws_bytes < need) {
        RTB_CUDA(cudaDeviceSynchronize());
        if (lane.d_ws) RTB_CUDA(cudaFree(lane.d_ws));
        lane.d_ws = nullptr; lane.ws_bytes = 0; lane.epoch = 0;
        RTB_CUDA(cudaMalloc(&lane.d_ws, need + need / 16));     // slack: pieces differ by a band
        // the bounce queue's tag words start as "no launch"; on the LAUNCHING stream: a null-stream memset is not ordered
        // before kernels on a non-blocking stream and would wipe entries of the running frame
        RTB_CUDA(cudaMemsetAsync(lane.d_ws, 0, need + need / 16, st));
        lane.ws_bytes = need + need / 16;
    }
    *primary_rays += owned_pixels(vd) * (uint64_t)(vd.s_end - vd.s_begin);
    cudaEvent_t* sev = nullptr;
    if (vd.flags & RTB_FLAG_TIMING) {
        for (auto& e : lane.stage_ev) if (!e) RTB_CUDA(cudaEventCreate(&e));
        sev = lane.stage_ev;
    }
    return rtb_launch_wavefront(scene_dev(g, n_prims), vd, is_ext ? &g.ext : nullptr, lane.d_ws, &lane.epoch, d_rgba, d_prim, d_t,
                                g.d_counters, st, launches, sev, lane.stage_ms);
}

int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? std::max(1, atoi(e)) : dflt;
}

// Renders the bands described by `whole` in `pieces` pieces, round-robin over `n_lanes` compute streams that are
// forked from and joined back into `st`.  Why: every stage of the wavefront pipeline ends with a drain in which a
// few long rays keep a few warps alive (each bounce level costs >= ~0.15 ms however few rays it has); with
// several independent pieces in flight on different streams the hardware fills one piece's drain with another
// piece's blocks.  after_piece(c, piece_view, lane_stream) runs after a piece's kernels are queued.
template <class F>
int render_pieces(GpuScene& g, uint32_t n_prims, const ViewDev& whole, float4* d_rgba, uint32_t* d_prim, float* d_t,
                  cudaStream_t st, uint32_t pieces, uint32_t n_lanes, uint32_t* launches, uint64_t* primary, F after_piece,
                  bool taper = false) {
    pieces = std::max(1u, std::min<uint32_t>(std::min<uint32_t>(pieces, RTB_MAX_CHUNKS), whole.my_tile_rows));
    n_lanes = std::max(1u, std::min<uint32_t>(std::min<uint32_t>(n_lanes, RTB_MAX_LANES), pieces));
    if (whole.flags & RTB_FLAG_MEGAKERNEL) n_lanes = 1;
    for (uint32_t l = 0; l < n_lanes; ++l) {
        GpuLane& lane = g.lanes[l];
        if (!lane.done) RTB_CUDA(cudaEventCreateWithFlags(&lane.done, cudaEventDisableTiming));
        if (l > 0 && !lane.st) {
            // later lanes get lower priority: earlier pieces then finish first and their D2H copies start while
            // the later pieces still render (equal priorities made all pieces finish together)
            int lo = 0, hi = 0;
            RTB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));     // lo = least (numerically largest)
            RTB_CUDA(cudaStreamCreateWithPriority(&lane.st, cudaStreamNonBlocking, std::min(lo, hi + (int)l)));
        }
    }
    if (n_lanes > 1) {
        if (!g.fork_ev) RTB_CUDA(cudaEventCreateWithFlags(&g.fork_ev, cudaEventDisableTiming));
        RTB_CUDA(cudaEventRecord(g.fork_ev, st));
        for (uint32_t l = 1; l < n_lanes; ++l) RTB_CUDA(cudaStreamWaitEvent(g.lanes[l].st, g.fork_ev, 0));
    }
    // Piece boundaries.  With copies to overlap (taper) the pieces shrink linearly, first : last = (pieces+1) : 2, so the
    // copy of the last piece — the one nothing can hide — is short.
    static const bool taper_on = getenv("RTB_PIECE_TAPER") ? atoi(getenv("RTB_PIECE_TAPER")) != 0 : true;
    // experiment knob: explicit relative piece sizes, "RTB_PIECE_WEIGHTS=1,3,4,4,2,1" (also sets the piece count)
    static const std::vector<double> weights = [] {
        std::vector<double> w;
        if (const char* e = getenv("RTB_PIECE_WEIGHTS"))
            for (const char* p = e; *p;) { char* end; const double x = strtod(p, &end); if (end == p) break; if (x > 0) w.push_back(x); p = *end ? end + 1 : end; }
        return w;
    }();
    if (taper && weights.size() >= 2) pieces = std::min<uint32_t>((uint32_t)std::min<size_t>(weights.size(), RTB_MAX_CHUNKS), whole.my_tile_rows);
    auto bound = [&](uint32_t c) -> uint32_t {        // first band of piece c (relative), bound(pieces) = all bands
        if (taper && weights.size() >= 2) {
            double tot = 0, acc = 0;
            for (uint32_t i = 0; i < pieces; ++i) { tot += weights[i]; if (i < c) acc += weights[i]; }
            return c >= pieces ? whole.my_tile_rows : (uint32_t)((double)whole.my_tile_rows * acc / tot);
        }
        if (!taper || !taper_on || pieces < 3) return (uint32_t)((uint64_t)whole.my_tile_rows * c / pieces);
        const uint64_t total_w = (uint64_t)pieces * (pieces + 3) / 2;                 // sum of (pieces + 1 - i), i < pieces
        const uint64_t w = (uint64_t)c * (2 * pieces + 3 - c) / 2;                    // sum over i < c
        return (uint32_t)((uint64_t)whole.my_tile_rows * w / total_w);
    };
    for (uint32_t c = 0; c < pieces; ++c) {
        ViewDev vd = whole;
        vd.band_begin = whole.band_begin + bound(c);
        vd.my_tile_rows = whole.band_begin + bound(c + 1) - vd.band_begin;
        if (vd.my_tile_rows == 0) continue;
        const uint32_t l = c % n_lanes;
        cudaStream_t ls = l == 0 ? st : g.lanes[l].st;     // lane 0 is the caller's stream itself
        GpuLane& lane = g.lanes[l];
        // the lane's workspace may still be in use by a frame issued earlier on ANOTHER stream (asynchronous
        // rtb_render_device calls with different streams): that frame's kernels first
        if (lane.used && lane.busy_stream != ls) RTB_CUDA(cudaStreamWaitEvent(ls, lane.busy, 0));
        int rc = (whole.flags & RTB_FLAG_COPY_ONLY) ? (int)RTB_OK
                                                    : launch_frame(g, lane, n_prims, vd, d_rgba, d_prim, d_t, ls, launches, primary);
        if (rc != RTB_OK) return rc;
        if (!lane.busy) RTB_CUDA(cudaEventCreateWithFlags(&lane.busy, cudaEventDisableTiming));
        RTB_CUDA(cudaEventRecord(lane.busy, ls));
        lane.busy_stream = ls; lane.used = true;
        rc = after_piece(c, vd, ls);
        if (rc != RTB_OK) return rc;
    }
    for (uint32_t l = 1; l < n_lanes; ++l) {
        RTB_CUDA(cudaEventRecord(g.lanes[l].done, g.lanes[l].st));
        RTB_CUDA(cudaStreamWaitEvent(st, g.lanes[l].done, 0));
    }
    return RTB_OK;
}

}  // namespace

extern "C" {

const char* rtb_last_error(void) { return g_err.c_str(); }

int rtb_init(int n_gpus, const int* device_ids) {
    std::lock_guard<std::mutex> lk(g_mu);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(RTB_ERR_NO_DEVICE, std::string("no CUDA device available (") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                           "); this library has no CPU fallback");
    }
    if (n_gpus <= 0) n_gpus = count;
    if (n_gpus > RTB_MAX_GPUS) return fail(RTB_ERR_INVALID, "more than RTB_MAX_GPUS devices requested");
    std::vector<int> devs;
    for (int i = 0; i < n_gpus; ++i) {
        int d = device_ids ? device_ids[i] : i;
        if (d < 0 || d >= count) return fail(RTB_ERR_INVALID, "device id " + std::to_string(d) + " out of range");
        devs.push_back(d);
    }
    // peer access between every pair (rtb_render_progressive reads its peers' sample sums directly where it can and
    // stages them with cudaMemcpyPeerAsync where it cannot); what works is recorded per pair
    for (size_t a = 0; a < devs.size(); ++a)
        for (size_t b = 0; b < devs.size(); ++b) {
            g_p2p[a][b] = (a == b);
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devs[a], devs[b]) != cudaSuccess) { cudaGetLastError(); can = 0; }
            if (can) {
                cudaSetDevice(devs[a]);
                const cudaError_t pe = cudaDeviceEnablePeerAccess(devs[b], 0);
                if (pe != cudaSuccess) cudaGetLastError();
                g_p2p[a][b] = (pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled);
            }
        }
    // scene-build scratch is stream-ordered (cudaMallocAsync): keep freed blocks in the pool instead of returning them
    for (int d : devs) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) {
            uint64_t keep_all = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep_all);
        } else {
            cudaGetLastError();
        }
    }
    RTB_CUDA(cudaSetDevice(devs[0]));
    g_devices = devs;
    return RTB_OK;
}

int rtb_device_count(void) { return g_devices.empty() ? RTB_ERR_INVALID : (int)g_devices.size(); }

int rtb_visible_device_count(void) {
    int count = 0;
    const cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(RTB_ERR_NO_DEVICE, "no CUDA device available; this library has no CPU fallback");
    }
    return count;
}

void rtb_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d : g_devices) {            // hand the cached scene-build scratch back to the system
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        else cudaGetLastError();
    }
    g_devices.clear();
    std::lock_guard<std::mutex> lr(g_render_mu);
    g_workers.clear();
}

namespace {

// Where a scene's triangle array comes from: finished `Triangle`s of the host (rtb_scene_create), or one mesh + per-
// instance records assembled on the device (rtb_scene_create_instanced) followed by ready-made extras.
struct TriSource {
    const RtbTriangle* tris = nullptr; uint32_t n = 0;            // host triangles (index 0 = dummy)
    const float* verts = nullptr; uint32_t nverts = 0;            // instanced mesh
    const uint32_t* faces = nullptr; uint32_t nfaces = 0;
    const RtbMeshInstance* inst = nullptr; uint32_t n_inst = 0;
    const RtbTriangle* extra = nullptr; uint32_t n_extra = 0;
    bool instanced = false;
    const RtbSphere* spheres = nullptr; uint32_t n_spheres = 0;   // EXTENSION: appended as pseudo-triangles
    uint32_t n_triangles() const { return instanced ? 1u + nfaces * n_inst + n_extra : n; }
    uint32_t total() const { return n_triangles() + n_spheres; }
};

cudaError_t upload(void** d, const void* h, size_t bytes, cudaStream_t st) {
    cudaError_t e = cudaMalloc(d, bytes ? bytes : 4);
    if (e == cudaSuccess && bytes) e = cudaMemcpyAsync(*d, h, bytes, cudaMemcpyHostToDevice, st);
    return e;
}

// EXTENSION: a sphere as the pseudo-`Triangle` the builder and the extension renderer understand (rtb_ext.cu):
// norm = 0 marks it, incenter = centre, bounding_r2 = r*r, corners span its AABB, kind carries RTB_PRIM_SPHERE.
RtbTriangle sphere_record(const RtbSphere& sp) {
    RtbTriangle t;
    std::memset(&t, 0, sizeof t);
    for (int k = 0; k < 3; ++k) {
        t.incenter[k] = sp.center[k];
        t.corners[k] = sp.center[k] - sp.radius; t.corners[3 + k] = sp.center[k] + sp.radius; t.corners[6 + k] = t.corners[k];
        t.color[k] = sp.color[k];
    }
    t.bounding_r2 = sp.radius * sp.radius;
    t.edge_thickness = -1.0f;
    t.kind = (sp.kind & 0xffu) | RTB_PRIM_SPHERE;
    t.alpha = sp.alpha;
    t.scattering = sp.scattering;
    return t;
}

int produce_base_triangles(const TriSource& src, RtbTriangle* d_tris, cudaStream_t st);

// Fills d_tris[0 .. src.total()) on the current device.
int produce_triangles(const TriSource& src, RtbTriangle* d_tris, cudaStream_t st) {
    int rc = produce_base_triangles(src, d_tris, st);
    if (rc != RTB_OK || src.n_spheres == 0) return rc;
    std::vector<RtbTriangle> rec(src.n_spheres);
    for (uint32_t j = 0; j < src.n_spheres; ++j) rec[j] = sphere_record(src.spheres[j]);
    RTB_CUDA(cudaMemcpyAsync(d_tris + src.n_triangles(), rec.data(), sizeof(RtbTriangle) * rec.size(), cudaMemcpyHostToDevice, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    return RTB_OK;
}

int produce_base_triangles(const TriSource& src, RtbTriangle* d_tris, cudaStream_t st) {
    if (!src.instanced) {
        if (src.n) RTB_CUDA(cudaMemcpyAsync(d_tris, src.tris, sizeof(RtbTriangle) * src.n, cudaMemcpyHostToDevice, st));
        RTB_CUDA(cudaStreamSynchronize(st));
        return RTB_OK;
    }
    const RtbTriangle dummy = raytrace::make_dummy_triangle();     // triangle 0 (main.rs:117), never tested
    RTB_CUDA(cudaMemcpyAsync(d_tris, &dummy, sizeof dummy, cudaMemcpyHostToDevice, st));
    float* d_verts = nullptr; uint32_t* d_faces = nullptr; RtbMeshInstance* d_inst = nullptr;
    auto cleanup = [&] { cudaFree(d_verts); cudaFree(d_faces); cudaFree(d_inst); };
    cudaError_t e = upload((void**)&d_verts, src.verts, sizeof(float) * 3 * (size_t)src.nverts, st);
    if (e == cudaSuccess) e = upload((void**)&d_faces, src.faces, sizeof(uint32_t) * 3 * (size_t)src.nfaces, st);
    if (e == cudaSuccess) e = upload((void**)&d_inst, src.inst, sizeof(RtbMeshInstance) * (size_t)src.n_inst, st);
    if (e != cudaSuccess) { cleanup(); return rtb_cuda_fail(e, "mesh upload", __FILE__, __LINE__); }
    uint32_t bad = 0xffffffffu;
    int rc = rtb_launch_assemble(d_verts, src.nverts, d_faces, src.nfaces, d_inst, src.n_inst, d_tris + 1, st, &bad);
    cleanup();
    if (rc != RTB_OK) return rc;
    if (bad != 0xffffffffu)
        return fail(RTB_ERR_INVALID, "make_triangle fails for face " + std::to_string(bad % src.nfaces) + " of instance " +
                                         std::to_string(bad / src.nfaces) +
                                         " (vertex index out of range, or a degenerate triangle: the reference panics, raytrace.rs:357)");
    if (src.n_extra)
        RTB_CUDA(cudaMemcpyAsync(d_tris + 1 + (size_t)src.nfaces * src.n_inst, src.extra, sizeof(RtbTriangle) * src.n_extra,
                                 cudaMemcpyHostToDevice, st));
    RTB_CUDA(cudaStreamSynchronize(st));
    return RTB_OK;
}

int scene_create_common(const TriSource& src, const float root_orig[3], float root_len2, rtb_scene** out) {
    *out = nullptr;
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    const uint32_t n = src.total();

    rtb_scene* s = new rtb_scene();
    std::memset(&s->info, 0, sizeof s->info);
    s->info.n_tris = n;
    s->info.n_gpus = (uint32_t)g_devices.size();
    s->gpu.resize(g_devices.size());

    for (size_t gi = 0; gi < g_devices.size(); ++gi) {
        GpuScene& g = s->gpu[gi];
        g.device = g_devices[gi];
        auto bail = [&](int code) { for (auto& x : s->gpu) free_gpu_scene(x); delete s; return code; };
        cudaError_t e;
        if ((e = cudaSetDevice(g.device)) != cudaSuccess) return bail(rtb_cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__));
        {
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            if ((e = cudaStreamCreateWithPriority(&g.stream, cudaStreamNonBlocking, hi)) != cudaSuccess)
                return bail(rtb_cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__));
        }
        cudaEventCreate(&g.ev0);
        cudaEventCreate(&g.ev1);
        if ((e = cudaMalloc(&g.d_counters, sizeof(TraceCounters))) != cudaSuccess)
            return bail(rtb_cuda_fail(e, "cudaMalloc(counters)", __FILE__, __LINE__));

        // triangles on the device (H2D, or assembled there), then the root-cube cull (raytrace.rs:795-805; triangle 0
        // is never in the tree, :791) as a kernel + ordered compaction
        const double t0 = now_ms();
        RtbTriangle* d_tris = nullptr;
        uint32_t* d_keep = nullptr;
        uint32_t n_prims = 0;
        if ((e = cudaMallocAsync(&d_tris, sizeof(RtbTriangle) * (n ? n : 1), g.stream)) != cudaSuccess)
            return bail(rtb_cuda_fail(e, "cudaMallocAsync(tris)", __FILE__, __LINE__));
        rc = produce_triangles(src, d_tris, g.stream);
        if (rc == RTB_OK) rc = rtb_launch_cull(d_tris, n, root_orig, root_len2, g.stream, &d_keep, &n_prims);
        if (rc != RTB_OK) { cudaFreeAsync(d_tris, g.stream); cudaFreeAsync(d_keep, g.stream); return bail(rc); }
        const double t1 = now_ms();
        if (gi == 0) s->info.n_prims = n_prims;

        BuildResult br;
        rc = rtb_build_lbvh(d_tris, d_keep, n_prims, g.stream, &br);
        cudaFreeAsync(d_tris, g.stream);
        cudaFreeAsync(d_keep, g.stream);
        g.d_nodes = br.d_nodes; g.d_tri = br.d_tri; g.d_shade = br.d_shade; g.d_prim_order = br.d_prim_order;
        g.n_nodes = br.n_nodes;
        g.height = br.tree_height;
        g.d_nodes4 = br.d_nodes4; g.n_nodes4 = br.n_nodes4; g.depth4 = br.depth4; g.stack4_need = br.stack4_need;
        g.d_nodes8 = br.d_nodes8; g.d_tri8 = br.d_tri8; g.d_shade8 = br.d_shade8; g.n_nodes8 = br.n_nodes8; g.depth8 = br.depth8;
        g.has_spheres = src.n_spheres > 0;
        if (rc != RTB_OK) return bail(rc);
        if (gi == 0) {
            s->info.n_nodes = br.n_nodes; s->info.n_leaves = br.n_leaves; s->info.max_leaf = br.max_leaf;
            s->info.tree_height = br.tree_height;
            for (int k = 0; k < 3; ++k) { s->info.scene_lo[k] = br.lo[k]; s->info.scene_hi[k] = br.hi[k]; }
            s->info.build_launches = br.launches;
            s->info.n_refs = br.n_refs;
        }
        s->info.ms_upload = std::max(s->info.ms_upload, t1 - t0);
        s->info.ms_build = std::max(s->info.ms_build, (double)br.ms_build);
    }
    *out = s;
    return RTB_OK;
}

}  // namespace

int rtb_scene_create(const RtbTriangle* tris, uint32_t n, const float root_orig[3], float root_len2,
                     rtb_scene** out) {
    if (!out) return fail(RTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n > 0 && !tris) return fail(RTB_ERR_INVALID, "tris is NULL");
    TriSource src;
    src.tris = tris; src.n = n;
    return scene_create_common(src, root_orig, root_len2, out);
}

namespace {
int check_mesh_args(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                    const RtbMeshInstance* inst, uint32_t n_inst) {
    if ((nverts && !verts) || (nfaces && !faces) || (n_inst && !inst)) return fail(RTB_ERR_INVALID, "NULL mesh argument");
    if ((uint64_t)nfaces * n_inst > 0x7fffff00ull) return fail(RTB_ERR_INVALID, "too many triangles");
    return RTB_OK;
}
}  // namespace

int rtb_scene_create_ext(const RtbTriangle* tris, uint32_t n, const RtbSphere* spheres, uint32_t n_spheres,
                         const float root_orig[3], float root_len2, rtb_scene** out) {
    if (!out) return fail(RTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if ((n > 0 && !tris) || (n_spheres > 0 && !spheres)) return fail(RTB_ERR_INVALID, "NULL primitive array");
    if (n == 0 && n_spheres > 0) return fail(RTB_ERR_INVALID, "primitive 0 must be the dummy triangle (ids: triangles 1..n-1, spheres n..)");
    for (uint32_t j = 0; j < n_spheres; ++j)
        if (!(spheres[j].radius > 0.0f) || (spheres[j].kind & ~0xffu)) return fail(RTB_ERR_INVALID, "bad sphere " + std::to_string(j));
    TriSource src;
    src.tris = tris; src.n = n; src.spheres = spheres; src.n_spheres = n_spheres;
    return scene_create_common(src, root_orig, root_len2, out);
}

int rtb_scene_set_light(rtb_scene* s, const float orig[3], float len2) {
    if (!s) return fail(RTB_ERR_INVALID, "scene is NULL");
    for (auto& g : s->gpu) {
        g.ext.has_light = orig ? 1u : 0u;
        for (int k = 0; k < 3; ++k) g.ext.light[k] = orig ? orig[k] : 0.0f;
        g.ext.light[3] = orig ? len2 : 0.0f;
    }
    return RTB_OK;
}

int rtb_scene_create_instanced(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                               const RtbMeshInstance* inst, uint32_t n_inst, const RtbTriangle* extra,
                               uint32_t n_extra, const float root_orig[3], float root_len2, rtb_scene** out) {
    if (!out) return fail(RTB_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int rc = check_mesh_args(verts, nverts, faces, nfaces, inst, n_inst);
    if (rc != RTB_OK) return rc;
    if (n_extra && !extra) return fail(RTB_ERR_INVALID, "extra is NULL");
    TriSource src;
    src.instanced = true;
    src.verts = verts; src.nverts = nverts; src.faces = faces; src.nfaces = nfaces;
    src.inst = inst; src.n_inst = n_inst; src.extra = extra; src.n_extra = n_extra;
    return scene_create_common(src, root_orig, root_len2, out);
}

int rtb_assemble_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                           const RtbMeshInstance* inst, uint32_t n_inst, RtbTriangle* out) {
    int rc = check_mesh_args(verts, nverts, faces, nfaces, inst, n_inst);
    if (rc != RTB_OK) return rc;
    if (!out) return fail(RTB_ERR_INVALID, "out is NULL");
    rc = ensure_init();
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(g_devices[0]));
    TriSource src;
    src.instanced = true;
    src.verts = verts; src.nverts = nverts; src.faces = faces; src.nfaces = nfaces; src.inst = inst; src.n_inst = n_inst;
    const uint32_t n = src.total();
    RtbTriangle* d_tris = nullptr;
    RTB_CUDA(cudaMalloc(&d_tris, sizeof(RtbTriangle) * n));
    rc = produce_triangles(src, d_tris, nullptr);
    if (rc == RTB_OK) {
        cudaError_t e = cudaMemcpy(out, d_tris + 1, sizeof(RtbTriangle) * (n - 1), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = rtb_cuda_fail(e, "cudaMemcpy(triangles)", __FILE__, __LINE__);
    }
    cudaFree(d_tris);
    return rc;
}

int rtb_cull_triangles(const RtbTriangle* tris, uint32_t n, const float root_orig[3], float root_len2,
                       uint32_t* keep_out, uint32_t* n_keep) {
    if ((n && !tris) || !n_keep) return fail(RTB_ERR_INVALID, "NULL argument");
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(g_devices[0]));
    RtbTriangle* d_tris = nullptr;
    RTB_CUDA(cudaMalloc(&d_tris, sizeof(RtbTriangle) * (n ? n : 1)));
    cudaError_t e = n ? cudaMemcpy(d_tris, tris, sizeof(RtbTriangle) * n, cudaMemcpyHostToDevice) : cudaSuccess;
    uint32_t* d_keep = nullptr;
    rc = e == cudaSuccess ? rtb_launch_cull(d_tris, n, root_orig, root_len2, nullptr, &d_keep, n_keep)
                          : rtb_cuda_fail(e, "cudaMemcpy(triangles)", __FILE__, __LINE__);
    if (rc == RTB_OK && keep_out && *n_keep) {
        e = cudaMemcpy(keep_out, d_keep, sizeof(uint32_t) * *n_keep, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = rtb_cuda_fail(e, "cudaMemcpy(keep)", __FILE__, __LINE__);
    }
    cudaFree(d_tris); cudaFree(d_keep);
    return rc;
}

int rtb_scene_info(const rtb_scene* s, RtbSceneInfo* out) {
    if (!s || !out) return fail(RTB_ERR_INVALID, "NULL argument");
    *out = s->info;
    return RTB_OK;
}

void rtb_scene_destroy(rtb_scene* s) {
    if (!s) return;
    for (auto& g : s->gpu) free_gpu_scene(g);
    delete s;
}

int rtb_scene_download_bvh(const rtb_scene* s, float* nodes, uint32_t* prim_order) {
    if (!s) return fail(RTB_ERR_INVALID, "scene is NULL");
    const GpuScene& g = s->gpu[0];
    RTB_CUDA(cudaSetDevice(g.device));
    if (nodes) RTB_CUDA(cudaMemcpy(nodes, g.d_nodes, sizeof(float4) * 2 * g.n_nodes, cudaMemcpyDeviceToHost));
    if (prim_order && s->info.n_refs)
        RTB_CUDA(cudaMemcpy(prim_order, g.d_prim_order, sizeof(uint32_t) * s->info.n_refs, cudaMemcpyDeviceToHost));
    return RTB_OK;
}

int rtb_render_device(rtb_scene* s, const RtbView* view, int gpu, uint32_t tile_rank, uint32_t tile_world,
                      float* d_rgba, uint32_t* d_prim, float* d_t, void* stream, RtbStats* stats) {
    if (!s) return fail(RTB_ERR_INVALID, "scene is NULL");
    int rc = check_view(view);
    if (rc != RTB_OK) return rc;
    if (gpu < 0 || gpu >= (int)s->gpu.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    if (tile_world == 0 || tile_rank >= tile_world) return fail(RTB_ERR_INVALID, "bad tile_rank/tile_world");
    if (!d_rgba) return fail(RTB_ERR_INVALID, "d_rgba is NULL");
    rc = check_camera(s, view);
    if (rc != RTB_OK) return rc;
    std::lock_guard<std::mutex> scene_lock(s->mu);
    GpuScene& g = s->gpu[gpu];
    RTB_CUDA(cudaSetDevice(g.device));
    cudaStream_t st = stream ? (cudaStream_t)stream : g.stream;
    const ViewDev vd = make_view(*view, tile_rank, tile_world, false);
    const double t0 = now_ms();
    uint32_t launches = 0;
    if (stats) {
        RTB_CUDA(cudaMemsetAsync(g.d_counters, 0, sizeof(TraceCounters), st));
        RTB_CUDA(cudaEventRecord(g.ev0, st));
    }
    uint64_t primary = 0;
    const size_t px = (size_t)vd.my_tile_rows * RTB_TILE_H * vd.width;
    uint32_t pieces = (uint32_t)env_int("RTB_PIECES", (int)std::min<size_t>(RTB_DEFAULT_PIECES_DEVICE, std::max<size_t>(1, px / (1u << 20))));
    uint32_t n_lanes = (uint32_t)env_int("RTB_LANES", RTB_DEFAULT_LANES);
    const bool timing = (view->flags & RTB_FLAG_TIMING) != 0;
    if (timing) {
        if (!stats) return fail(RTB_ERR_INVALID, "RTB_FLAG_TIMING needs a stats pointer");
        pieces = 1; n_lanes = 1;
        for (float& m : g.lanes[0].stage_ms) m = 0.f;
    }
    rc = render_pieces(g, s->info.n_refs, vd, (float4*)d_rgba, d_prim, d_t, st, pieces, n_lanes, &launches, &primary,
                       [](uint32_t, const ViewDev&, cudaStream_t) { return (int)RTB_OK; });
    if (rc != RTB_OK) return rc;
    if (stats) {
        RTB_CUDA(cudaEventRecord(g.ev1, st));
        TraceCounters c;
        RTB_CUDA(cudaMemcpyAsync(&c, g.d_counters, sizeof c, cudaMemcpyDeviceToHost, st));
        RTB_CUDA(cudaStreamSynchronize(st));
        float ms = 0.f;
        RTB_CUDA(cudaEventElapsedTime(&ms, g.ev0, g.ev1));
        if (c.stalled) return fail(RTB_ERR_CUDA, "path kernel watchdog: a reserved bounce-queue entry never arrived");
        std::memset(stats, 0, sizeof *stats);
        stats->rays = c.rays + primary; stats->node_tests = c.node_tests; stats->tri_tests = c.tri_tests;
        stats->bounce_rays = c.rays; stats->node_tests_bounce = c.node_tests_bounce; stats->tri_tests_bounce = c.tri_tests_bounce;
        if (timing) for (int k = 0; k < RTB_N_STAGES; ++k) stats->ms_stage[k] = g.lanes[0].stage_ms[k];
        stats->ms_render = ms; stats->ms_total = now_ms() - t0; stats->kernel_launches = launches; stats->n_gpus = 1;
    }
    return RTB_OK;
}

}  // extern "C"

namespace {
// rtb_render (f32 RGBA home) and rtb_render_rgb8 (quantised on the device, 3 bytes per pixel home) share one body.
int render_host(rtb_scene* s, const RtbView* view, float* rgba_out, uint8_t* rgb8_out, uint32_t* prim_out, float* t_out,
                RtbStats* stats) {
    if (!s) return fail(RTB_ERR_INVALID, "scene is NULL");
    int rc = check_view(view);
    if (rc != RTB_OK) return rc;
    if (!rgba_out && !rgb8_out) return fail(RTB_ERR_INVALID, "output buffer is NULL");
    rc = check_camera(s, view);
    if (rc != RTB_OK) return rc;
    std::lock_guard<std::mutex> scene_lock(s->mu);     // one frame at a time per handle: two threads may share a handle
    const double t0 = now_ms();
    const uint32_t world = (uint32_t)s->gpu.size();
    const uint32_t W = view->width, H = view->height;
    const uint32_t tiles_y = (H + RTB_TILE_H - 1) / RTB_TILE_H;
    uint32_t launches = 0;
    uint64_t primary_total = 0;

    // Every GPU renders its 8-row bands (band b -> GPU b % world) into a compact device buffer and copies
    // them home with one strided cudaMemcpy2DAsync per chunk over its own PCIe link.  The bands of a GPU are
    // processed in `chunks` pieces so that the D2H copy of piece i overlaps the kernels of piece i+1
    // (compute stream + copy stream, one event per piece).
    // Tried instead (tools/experiments/streaming_d2h_completion_flags.patch): ONE render of all bands with per-group
    // finished-pixel counters and copies started as groups complete (cuStreamWaitValue32, then host-polled flags in
    // mapped memory).  The mechanism works, but with the persistent bounce kernel every group of bands still has a
    // few long paths in flight until ~0.3 ms before the end of the frame, so nothing overlapped: 4.65 ms vs 3.47.
    std::vector<uint32_t> launches_r(world, 0);
    std::vector<uint64_t> primary_r(world, 0);
    std::vector<std::string> err_r(world);
    auto issue = [&](uint32_t r) -> int {
        int rc = RTB_OK;
        uint32_t& launches = launches_r[r];
        uint64_t& primary_total = primary_r[r];
        GpuScene& g = s->gpu[r];
        RTB_CUDA(cudaSetDevice(g.device));
        const ViewDev whole = make_view(*view, r, world, true);
        if (whole.my_tile_rows == 0) return RTB_OK;
        const size_t pixels = (size_t)whole.my_tile_rows * RTB_TILE_H * W;
        rc = ensure_framebuffer(g, pixels, prim_out != nullptr, t_out != nullptr);
        if (rc != RTB_OK) return rc;
        if (rgb8_out && g.rgb8_pixels < pixels) {
            cudaFree(g.d_rgb8); g.d_rgb8 = nullptr; g.rgb8_pixels = 0;
            RTB_CUDA(cudaMalloc(&g.d_rgb8, pixels * 3 + 16));
            g.rgb8_pixels = pixels;
        }
        if (!g.copy_stream) {
            RTB_CUDA(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
            for (auto& e : g.chunk_ev) RTB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        // pieces exist to overlap the D2H copy with the kernels: 16 B/px need 8 of them, the 3 B/px of the quantised frame
        // only 3 (4K frame through rtb_render_rgb8: 2.37 ms with 3 pieces, 2.42 with 4, 2.51 with 8, 2.57 with 2)
        const size_t dflt_pieces = rgb8_out ? RTB_DEFAULT_PIECES_RGB8 : RTB_DEFAULT_PIECES;
        // a piece of the f32 frame should not be much smaller than 256 K pixels (its kernels' fixed tails), a piece of the
        // quantised frame not smaller than 1 M (the copy it overlaps is short).  With several GPUs the copies share the host's
        // memory path (8 GPUs: 133 MB come home in 1.24 ms however they are issued), so what counts is how soon the FIRST
        // copy starts: 4 pieces per GPU at N = 8 instead of one.
        const size_t per_piece = rgb8_out ? (1u << 20) : (1u << 18);
        const uint32_t pieces = (uint32_t)env_int("RTB_PIECES", (int)std::min<size_t>(dflt_pieces, std::max<size_t>(1, pixels / per_piece)));
        const uint32_t n_lanes = (uint32_t)env_int("RTB_LANES", RTB_DEFAULT_LANES);
        RTB_CUDA(cudaMemsetAsync(g.d_counters, 0, sizeof(TraceCounters), g.stream));
        RTB_CUDA(cudaEventRecord(g.ev0, g.stream));
        rc = render_pieces(g, s->info.n_refs, whole, g.d_rgba, prim_out ? g.d_prim : nullptr, t_out ? g.d_t : nullptr,
                           g.stream, pieces, n_lanes, &launches, &primary_total,
                           [&](uint32_t c, const ViewDev& vd, cudaStream_t ls) -> int {
            // the piece's bands go home on the copy stream as soon as its kernels are done
            if (rgb8_out && !(vd.flags & RTB_FLAG_COPY_ONLY)) {      // write_png's `(c * 255.) as u8` (raytrace.rs:1468-1473) on the piece's pixels, on its stream
                const size_t first_px = (size_t)vd.band_begin * RTB_TILE_H * W, n_px = (size_t)vd.my_tile_rows * RTB_TILE_H * W;
                int rcq = rtb_launch_quantize(g.d_rgba + first_px, n_px, g.d_rgb8 + 3 * first_px, ls);
                if (rcq != RTB_OK) return rcq;
                ++launches;
            }
            RTB_CUDA(cudaEventRecord(g.chunk_ev[c], ls));
            RTB_CUDA(cudaStreamWaitEvent(g.copy_stream, g.chunk_ev[c], 0));
            // local band b = image rows [(b*world + r)*8, +8); copy bands [band_begin, band_begin+n)
            const uint32_t b0 = vd.band_begin, nb = vd.my_tile_rows;
            const uint32_t last_ty = (b0 + nb - 1) * world + r;
            const bool ragged = (last_ty == tiles_y - 1) && (H % RTB_TILE_H != 0);
            const uint32_t full_bands = ragged ? nb - 1 : nb;
            auto copy_plane = [&](void* host, const void* dev, size_t px_bytes) -> int {
                const size_t band_bytes = (size_t)RTB_TILE_H * W * px_bytes;
                char* dst0 = (char*)host + ((size_t)b0 * world + r) * band_bytes;
                const char* src0 = (const char*)dev + (size_t)b0 * band_bytes;
                if (full_bands)
                    RTB_CUDA(cudaMemcpy2DAsync(dst0, band_bytes * world, src0, band_bytes, band_bytes, full_bands,
                                               cudaMemcpyDeviceToHost, g.copy_stream));
                if (ragged) {
                    const size_t rows = H - (size_t)last_ty * RTB_TILE_H;
                    RTB_CUDA(cudaMemcpyAsync((char*)host + (size_t)last_ty * band_bytes,
                                             src0 + (size_t)full_bands * band_bytes, rows * W * px_bytes,
                                             cudaMemcpyDeviceToHost, g.copy_stream));
                }
                return RTB_OK;
            };
            int rc2;
            if (rgba_out && (rc2 = copy_plane(rgba_out, g.d_rgba, sizeof(float4))) != RTB_OK) return rc2;
            if (rgb8_out && (rc2 = copy_plane(rgb8_out, g.d_rgb8, 3)) != RTB_OK) return rc2;
            if (prim_out && (rc2 = copy_plane(prim_out, g.d_prim, sizeof(uint32_t))) != RTB_OK) return rc2;
            if (t_out && (rc2 = copy_plane(t_out, g.d_t, sizeof(float))) != RTB_OK) return rc2;
            return RTB_OK;
        }, true);
        if (rc != RTB_OK) return rc;
        RTB_CUDA(cudaEventRecord(g.ev1, g.stream));
        return RTB_OK;
    };
    // issue + wait + counter read-back of one GPU; runs on that GPU's worker thread when there are several
    std::vector<TraceCounters> cnt_r(world);
    std::vector<float> ms_r(world, 0.f);
    auto frame_of = [&](uint32_t r) -> int {
        int rc = issue(r);
        if (rc != RTB_OK) return rc;
        GpuScene& g = s->gpu[r];
        std::memset(&cnt_r[r], 0, sizeof(TraceCounters));
        if (make_view(*view, r, world, true).my_tile_rows == 0) return RTB_OK;
        RTB_CUDA(cudaMemcpyAsync(&cnt_r[r], g.d_counters, sizeof(TraceCounters), cudaMemcpyDeviceToHost, g.stream));
        RTB_CUDA(cudaStreamSynchronize(g.stream));
        RTB_CUDA(cudaStreamSynchronize(g.copy_stream));
        RTB_CUDA(cudaEventElapsedTime(&ms_r[r], g.ev0, g.ev1));
        return RTB_OK;
    };
    if (world == 1) {
        rc = frame_of(0);
    } else {
        std::lock_guard<std::mutex> lr(g_render_mu);
        ensure_workers(world);
        std::vector<int> rcs(world, RTB_OK);
        for (uint32_t r = 0; r < world; ++r)
            g_workers[r]->start([&, r]() { rcs[r] = frame_of(r); if (rcs[r] != RTB_OK) err_r[r] = rtb_last_error(); });
        for (uint32_t r = 0; r < world; ++r) g_workers[r]->wait();
        for (uint32_t r = 0; r < world; ++r)
            if (rcs[r] != RTB_OK) { rc = rcs[r]; g_err = err_r[r]; break; }
    }
    if (rc != RTB_OK) return rc;
    RtbStats st;
    std::memset(&st, 0, sizeof st);
    for (uint32_t r = 0; r < world; ++r) {
        launches += launches_r[r]; primary_total += primary_r[r];
        const TraceCounters& c = cnt_r[r];
        if (c.stalled) return fail(RTB_ERR_CUDA, "path kernel watchdog: a reserved bounce-queue entry never arrived");
        st.rays += c.rays; st.node_tests += c.node_tests; st.tri_tests += c.tri_tests;
        st.bounce_rays += c.rays; st.node_tests_bounce += c.node_tests_bounce; st.tri_tests_bounce += c.tri_tests_bounce;
        st.ms_render = std::max(st.ms_render, (double)ms_r[r]);
    }
    st.rays += primary_total;
    st.ms_total = now_ms() - t0;
    st.kernel_launches = launches;
    st.n_gpus = world;
    if (stats) *stats = st;
    return RTB_OK;
}
}  // namespace

extern "C" {

int rtb_render(rtb_scene* s, const RtbView* view, float* rgba_out, uint32_t* prim_out, float* t_out, RtbStats* stats) {
    if (!rgba_out) return fail(RTB_ERR_INVALID, "rgba_out is NULL");
    return render_host(s, view, rgba_out, nullptr, prim_out, t_out, stats);
}

int rtb_render_rgb8(rtb_scene* s, const RtbView* view, uint8_t* rgb_out, RtbStats* stats) {
    if (!rgb_out) return fail(RTB_ERR_INVALID, "rgb_out is NULL");
    return render_host(s, view, nullptr, rgb_out, nullptr, nullptr, stats);
}

int rtb_render_progressive(rtb_scene* s, const RtbView* view, float* rgba_out, RtbStats* stats) {
    if (!s) return fail(RTB_ERR_INVALID, "scene is NULL");
    int rc = check_view(view);
    if (rc != RTB_OK) return rc;
    if (!rgba_out) return fail(RTB_ERR_INVALID, "rgba_out is NULL");
    rc = check_camera(s, view);
    if (rc != RTB_OK) return rc;
    std::lock_guard<std::mutex> scene_lock(s->mu);
    const double t0 = now_ms();
    const uint32_t world = (uint32_t)s->gpu.size();
    const uint32_t W = view->width, H = view->height;
    const size_t pixels = (size_t)W * H;
    const uint32_t s_lo = view->sample_begin, s_hi = (view->sample_begin == 0 && view->sample_end == 0) ? view->spp : view->sample_end;
    const uint32_t n_s = s_hi - s_lo;
    const float inv_spp = 1.0f / (float)n_s;
    std::vector<uint32_t> launches_r(world, 0);
    std::vector<uint64_t> primary_r(world, 0);
    std::vector<TraceCounters> cnt_r(world);
    std::vector<float> ms_r(world, 0.f), msred_r(world, 0.f);
    std::vector<std::string> err_r(world);
    // direct peer loads only if EVERY pair can do them; else every peer's share is staged (one code path, not N^2)
    bool all_p2p = true;
    for (uint32_t a = 0; a < world; ++a)
        for (uint32_t b = 0; b < world; ++b) all_p2p = all_p2p && g_p2p[a][b];
    auto range_of = [&](uint32_t r, uint64_t* first, uint64_t* last) { *first = (uint64_t)pixels * r / world; *last = (uint64_t)pixels * (r + 1) / world; };

    // phase 1 (per GPU, issue only): accumulate this GPU's contiguous share of the samples over the FULL frame (sum only),
    // then record "my sums are complete"
    auto phase1 = [&](uint32_t r) -> int {
        GpuScene& g = s->gpu[r];
        RTB_CUDA(cudaSetDevice(g.device));
        int rc = ensure_framebuffer(g, 2 * pixels, false, false);   // [0,pixels) = sum buffer, [pixels,2*pixels) = reduced share
        if (rc != RTB_OK) return rc;
        if (!g.prog_done) {
            RTB_CUDA(cudaEventCreateWithFlags(&g.prog_done, cudaEventDisableTiming));
            RTB_CUDA(cudaEventCreate(&g.red0));
            RTB_CUDA(cudaEventCreate(&g.red1));
        }
        RtbView v = *view;
        v.sample_begin = s_lo + (uint32_t)(((uint64_t)n_s * r) / world);
        v.sample_end = s_lo + (uint32_t)(((uint64_t)n_s * (r + 1)) / world);
        v.flags |= RTB_FLAG_SUM_ONLY;
        RTB_CUDA(cudaMemsetAsync(g.d_counters, 0, sizeof(TraceCounters), g.stream));
        RTB_CUDA(cudaEventRecord(g.ev0, g.stream));
        if (v.sample_begin == v.sample_end) {
            RTB_CUDA(cudaMemsetAsync(g.d_rgba, 0, pixels * sizeof(float4), g.stream));
        } else {
            const ViewDev vd = make_view(v, 0, 1, false);
            GpuLane& lane = g.lanes[0];
            if (lane.used && lane.busy_stream != g.stream) RTB_CUDA(cudaStreamWaitEvent(g.stream, lane.busy, 0));
            rc = launch_frame(g, lane, s->info.n_refs, vd, g.d_rgba, nullptr, nullptr, g.stream, &launches_r[r], &primary_r[r]);
            if (rc != RTB_OK) return rc;
            if (!lane.busy) RTB_CUDA(cudaEventCreateWithFlags(&lane.busy, cudaEventDisableTiming));
            RTB_CUDA(cudaEventRecord(lane.busy, g.stream));
            lane.busy_stream = g.stream; lane.used = true;
        }
        RTB_CUDA(cudaEventRecord(g.ev1, g.stream));
        RTB_CUDA(cudaEventRecord(g.prog_done, g.stream));
        return RTB_OK;
    };
    // phase 2 (per GPU): wait on the device for every peer's sums (events, no host synchronisation), reduce pixel range r
    // over NVLink peer loads fused with the 1/spp scale, copy the range home over this GPU's own PCIe link, read the counters
    auto phase2 = [&](uint32_t r) -> int {
        GpuScene& g = s->gpu[r];
        RTB_CUDA(cudaSetDevice(g.device));
        uint64_t first, last;
        range_of(r, &first, &last);
        for (uint32_t p = 0; p < world; ++p)
            if (p != r) RTB_CUDA(cudaStreamWaitEvent(g.stream, s->gpu[p].prog_done, 0));
        RTB_CUDA(cudaEventRecord(g.red0, g.stream));
        RtbPeerBufs bufs;
        for (uint32_t p = 0; p < RTB_MAX_GPUS; ++p) bufs.p[p] = p < world ? s->gpu[p].d_rgba : nullptr;
        if (!all_p2p && world > 1) {
            const size_t cnt = (size_t)(last - first);
            if (g.stage_pixels < cnt * (world - 1)) {
                RTB_CUDA(cudaStreamSynchronize(g.stream));
                cudaFree(g.d_stage); g.d_stage = nullptr; g.stage_pixels = 0;
                RTB_CUDA(cudaMalloc(&g.d_stage, cnt * (world - 1) * sizeof(float4)));
                g.stage_pixels = cnt * (world - 1);
            }
            uint32_t k = 0;
            for (uint32_t p = 0; p < world; ++p) {
                if (p == r) continue;
                float4* dst = g.d_stage + (size_t)k * cnt;
                RTB_CUDA(cudaMemcpyPeerAsync(dst, g.device, s->gpu[p].d_rgba + first, s->gpu[p].device, cnt * sizeof(float4), g.stream));
                bufs.p[p] = dst - first;          // indexed with [first + i] by the kernel
                ++k;
            }
        }
        float4* d_out = g.d_rgba + pixels;
        int rc = rtb_launch_peer_reduce(bufs, (int)world, inv_spp, first, last - first, d_out, g.stream);
        if (rc != RTB_OK) return rc;
        ++launches_r[r];
        RTB_CUDA(cudaMemcpyAsync(rgba_out + 4 * first, d_out + first, (last - first) * sizeof(float4), cudaMemcpyDeviceToHost, g.stream));
        RTB_CUDA(cudaEventRecord(g.red1, g.stream));
        RTB_CUDA(cudaMemcpyAsync(&cnt_r[r], g.d_counters, sizeof(TraceCounters), cudaMemcpyDeviceToHost, g.stream));
        RTB_CUDA(cudaStreamSynchronize(g.stream));
        RTB_CUDA(cudaEventElapsedTime(&ms_r[r], g.ev0, g.ev1));
        RTB_CUDA(cudaEventElapsedTime(&msred_r[r], g.red0, g.red1));
        return RTB_OK;
    };
    // an event must be RECORDED before another stream is told to wait for it: all GPUs finish issuing phase 1 (a host-side
    // join of the workers, nothing waits for the device) before any GPU issues phase 2
    auto run_phase = [&](const std::function<int(uint32_t)>& fn) -> int {
        if (world == 1) return fn(0);
        std::vector<int> rcs(world, RTB_OK);
        for (uint32_t r = 0; r < world; ++r)
            g_workers[r]->start([&, r]() { rcs[r] = fn(r); if (rcs[r] != RTB_OK) err_r[r] = rtb_last_error(); });
        for (uint32_t r = 0; r < world; ++r) g_workers[r]->wait();
        for (uint32_t r = 0; r < world; ++r)
            if (rcs[r] != RTB_OK) { g_err = err_r[r]; return rcs[r]; }
        return RTB_OK;
    };
    {
        std::unique_lock<std::mutex> lr(g_render_mu, std::defer_lock);
        if (world > 1) { lr.lock(); ensure_workers(world); }
        rc = run_phase(phase1);
        if (rc == RTB_OK) rc = run_phase(phase2);
        if (rc != RTB_OK) {      // leave no work in flight on an error path
            for (uint32_t r = 0; r < world; ++r) { cudaSetDevice(s->gpu[r].device); cudaStreamSynchronize(s->gpu[r].stream); }
            cudaGetLastError();
            return rc;
        }
    }
    RtbStats st;
    std::memset(&st, 0, sizeof st);
    for (uint32_t r = 0; r < world; ++r) {
        const TraceCounters& c = cnt_r[r];
        if (c.stalled) return fail(RTB_ERR_CUDA, "path kernel watchdog: a reserved bounce-queue entry never arrived");
        st.rays += c.rays + primary_r[r]; st.node_tests += c.node_tests; st.tri_tests += c.tri_tests;
        st.bounce_rays += c.rays; st.node_tests_bounce += c.node_tests_bounce; st.tri_tests_bounce += c.tri_tests_bounce;
        st.ms_render = std::max(st.ms_render, (double)ms_r[r]);
        st.ms_reduce = std::max(st.ms_reduce, (double)msred_r[r]);
        st.kernel_launches += launches_r[r];
    }
    st.ms_total = now_ms() - t0;
    st.n_gpus = world;
    if (stats) *stats = st;
    return RTB_OK;
}

int rtb_scale_device(float* d_rgba, uint64_t npix, uint32_t spp, int gpu, void* stream) {
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    if (!d_rgba || spp == 0) return fail(RTB_ERR_INVALID, "rtb_scale_device: NULL buffer or spp == 0");
    if (gpu < 0 || gpu >= (int)g_devices.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    RTB_CUDA(cudaSetDevice(g_devices[gpu]));
    return rtb_launch_scale((float4*)d_rgba, npix, 1.0f / (float)spp, (cudaStream_t)stream);
}

int rtb_quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb_out) {
    if (!rgba || !rgb_out) return fail(RTB_ERR_INVALID, "NULL argument");
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaSetDevice(g_devices[0]));
    float4* d_in = nullptr;
    uint8_t* d_out = nullptr;
    RTB_CUDA(cudaMalloc(&d_in, npix * sizeof(float4) + 16));
    if (cudaMalloc(&d_out, npix * 3 + 16) != cudaSuccess) { cudaFree(d_in); return fail(RTB_ERR_NOMEM, "cudaMalloc"); }
    cudaError_t e = cudaMemcpy(d_in, rgba, npix * sizeof(float4), cudaMemcpyHostToDevice);
    rc = e == cudaSuccess ? rtb_launch_quantize(d_in, npix, d_out, 0) : rtb_cuda_fail(e, "cudaMemcpy(rgba)", __FILE__, __LINE__);
    if (rc == RTB_OK) e = cudaMemcpy(rgb_out, d_out, npix * 3, cudaMemcpyDeviceToHost);
    cudaFree(d_in);
    cudaFree(d_out);
    if (rc != RTB_OK) return rc;
    if (e != cudaSuccess) return rtb_cuda_fail(e, "cudaMemcpy", __FILE__, __LINE__);
    return RTB_OK;
}

int rtb_selftest_sort(uint32_t n, int key_bits, uint64_t seed) {
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    if (key_bits < 1 || key_bits > 64) return fail(RTB_ERR_INVALID, "key_bits must be in [1, 64]");
    RTB_CUDA(cudaSetDevice(g_devices[0]));
    return rtb_sort_selftest(n, key_bits, seed);
}

int rtb_selftest_udiv(uint32_t d, uint32_t samples, uint64_t seed) {
    if (d == 0) return fail(RTB_ERR_INVALID, "division by zero");
    const RtbUdiv m = rtb_udiv_make(d);
    int bad = 0;
    const uint32_t edge[] = {0u, 1u, d - 1u, d, d + 1u, 2u * d - 1u, 2u * d, 0x7fffffffu, 0x80000000u, 0xfffffffeu, 0xffffffffu,
                             0xffffffffu / d * d, 0xffffffffu / d * d - 1u};
    for (uint32_t n : edge) bad += rtb_udiv(n, m) != n / d;
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
    for (uint32_t k = 0; k < samples; ++k) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;                         // xorshift64
        const uint32_t n = (uint32_t)(x >> 16);
        bad += rtb_udiv(n, m) != n / d;
    }
    return bad;
}

int rtb_partition_rows(uint32_t height, uint32_t rank, uint32_t world, uint32_t* rows_out, uint32_t cap) {
    if (world == 0 || rank >= world) return fail(RTB_ERR_INVALID, "bad rank/world");
    uint32_t n = 0;
    for (uint32_t row = 0; row < height; ++row) {
        if ((row / RTB_TILE_H) % world != rank) continue;
        if (rows_out && n < cap) rows_out[n] = row;
        ++n;
    }
    return (int)n;
}

int rtb_device_alloc(int gpu, size_t bytes, void** d_ptr) {
    if (!d_ptr) return fail(RTB_ERR_INVALID, "d_ptr is NULL");
    *d_ptr = nullptr;
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    if (gpu < 0 || gpu >= (int)g_devices.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    RTB_CUDA(cudaSetDevice(g_devices[gpu]));
    RTB_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 4));
    RTB_CUDA(cudaMemset(*d_ptr, 0, bytes ? bytes : 4));
    RTB_CUDA(cudaDeviceSynchronize());
    return RTB_OK;
}

int rtb_device_free(int gpu, void* d_ptr) {
    if (gpu < 0 || gpu >= (int)g_devices.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    RTB_CUDA(cudaSetDevice(g_devices[gpu]));
    RTB_CUDA(cudaFree(d_ptr));
    return RTB_OK;
}

int rtb_ipc_export(void* d_ptr, unsigned char handle_out[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    if (!d_ptr || !handle_out) return fail(RTB_ERR_INVALID, "NULL argument");
    cudaIpcMemHandle_t h;
    RTB_CUDA(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle_out, &h, 64);
    return RTB_OK;
}

int rtb_ipc_open(int gpu, const unsigned char handle[64], void** d_ptr_out) {
    if (!handle || !d_ptr_out) return fail(RTB_ERR_INVALID, "NULL argument");
    *d_ptr_out = nullptr;
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    if (gpu < 0 || gpu >= (int)g_devices.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    RTB_CUDA(cudaSetDevice(g_devices[gpu]));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    RTB_CUDA(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return RTB_OK;
}

int rtb_ipc_close(int gpu, void* d_ptr) {
    if (gpu < 0 || gpu >= (int)g_devices.size()) return fail(RTB_ERR_INVALID, "gpu slot out of range");
    RTB_CUDA(cudaSetDevice(g_devices[gpu]));
    RTB_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return RTB_OK;
}

int rtb_host_register(void* ptr, size_t bytes) {
    int rc = ensure_init();
    if (rc != RTB_OK) return rc;
    RTB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return RTB_OK;
}

int rtb_host_unregister(void* ptr) {
    RTB_CUDA(cudaHostUnregister(ptr));
    return RTB_OK;
}

}  // extern "C"
