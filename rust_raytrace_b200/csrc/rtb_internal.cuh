// rtb_internal.cuh — shared declarations of the CUDA library (not installed).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "rtb.h"

// ---------------------------------------------------------------------------
// Device-side scene layout (all arrays in HBM, L2-resident for the scenes of
// interest; see DESIGN.md "Data layout").
//
//  nodes : 2 x float4 per BVH2 node (32 bytes):
//            n0 = (lo.x, lo.y, lo.z, bits(a))     n1 = (hi.x, hi.y, hi.z, bits(b))
//          internal: a = index of the left child, the right child is a+1 (sibling
//                    pairs are adjacent and 64-byte aligned), b = 0
//          leaf    : a = first primitive slot, b = primitive count (> 0)
//          node 0 is the root, node 1 is padding, pairs start at node 2.
//  tri   : 5 x float4 per primitive (80 bytes) in leaf order — the 19 floats
//          Triangle::intersects reads (raytrace.rs:400-422) plus the original index:
//            q0 = (norm.xyz, bounding_r2)   q1 = (incenter.xyz, bits(orig_index))
//            q2 = (sides[0].xyz, side_lens[0])  q3, q4 likewise
//  shade : 2 x float4 per primitive in leaf order, read once per hit:
//            s0 = (color.rgb, alpha)
//            s1 = (bits(kind), scattering, edge_thickness, 0)
// ---------------------------------------------------------------------------
struct SceneDev {
    const float4* nodes;
    const float4* tri;
    const float4* shade;
    uint32_t n_prims;
    uint32_t n_nodes;
    uint32_t height;     // tree height: the traversal stack never holds more than `height` entries
    const float4* nodes4;   // 4-wide collapse of the same tree: 8 x float4 (128 B) per node, see rtb_lbvh.cu
    uint32_t stack4;     // stack entries a BVH4 traversal can need: 3 per level
    uint32_t n_nodes4, n_nodes8;   // node counts (bounds of the RTB_DEBUG checks)
    // 8-wide compressed collapse (80 B per node) with its own reference order: tri8 / shade8 are tri / shade permuted
    const uint4* nodes8;
    const float4* tri8;
    const float4* shade8;
    uint32_t depth8;     // levels of the BVH8: its traversal stack holds at most one entry per level
};

// RTB_DEBUG build (make debug -> librtb_debug.so): device-side bounds checks on everything the traversal indexes — node and
// reference indices, traversal-stack depth, queue and pixel-slot indices.  compute-sanitizer is not available on the
// B200 pool this was developed on; a tripped check ends the kernel with cudaErrorAssert, which the C ABI reports as
// RTB_ERR_CUDA.  tests/test_gpu_parity.py runs a parity frame of every renderer through the debug library.
#ifdef RTB_DEBUG
#include <cassert>
#define RTB_DASSERT(cond) assert(cond)
#else
#define RTB_DASSERT(cond) ((void)0)
#endif

#define RTB_TRI_F4 5
// BVH4 child boxes as (centre, half extent) instead of (lo, hi): the slab test is then t_c = c/d - o/d, t_near = t_c - e/|d|,
// t_far = t_c + e/|d| — nine FFMA per child on the FMA pipe instead of six FFMA + six FMNMX; min/max/compare/select run on the
// half-rate ALU pipe, which is the busiest pipe of the path kernel (ncu: 62 % / 73 % in the bounce / primary launch against
// 20 % / 24 % for the FMA pipe).  0 = (lo, hi) boxes with per-axis min / max (the A/B baseline).
#ifndef RTB_NODE_CE
#define RTB_NODE_CE 1
#endif
#define RTB_SHADE_F4 2
#ifndef RTB_LEAF_MAX
#define RTB_LEAF_MAX 4
#endif
#define RTB_STACK 64
#define RTB_MAX_CHUNKS 16
#define RTB_MAX_LANES 4
// rtb_render (host output) renders the frame's bands in pieces on a few streams so that the D2H copy of a finished
// piece (own copy stream) overlaps the kernels of the next ones.  Measured on B200, 4K teapot frame, ms per call incl.
// the 133 MB D2H (2.34 ms alone; one full-frame render 2.47 ms): pieces/lanes 4/1 3.70, 4/2 3.62, 8/1 4.00, 8/2 3.47,
// 8/4 4.07, 12/3 3.81, 16/2 3.95, 16/4 4.32.  Every piece pays its own launches and kernel tails (8 pieces on one
// stream: 3.65 ms of device time), so more pieces is not better.
#define RTB_DEFAULT_PIECES 8          /* rtb_render: more pieces = earlier D2H overlap */
#define RTB_DEFAULT_PIECES_DEVICE 1   /* rtb_render_device: no copies to overlap */
#define RTB_DEFAULT_PIECES_RGB8 3     /* rtb_render_rgb8: 3 B/px go home, the copy is short */
#define RTB_DEFAULT_LANES 2

// Image tiling: one CTA = 128 threads = 16 x 8 pixels; one warp = 8 x 4 pixels.
#define RTB_TILE_W 16
#define RTB_TILE_H 8

// Unsigned division by a divisor known on the host (Granlund & Montgomery 1994, the round-up method): exact for every 32-bit
// numerator.  q = (t + ((n - t) >> s1)) >> s2 with t = mulhi(m, n).
struct RtbUdiv { uint32_t m, s1, s2; };
static inline RtbUdiv rtb_udiv_make(uint32_t d) {
    RtbUdiv r;
    uint32_t l = 0;
    while ((1ull << l) < (unsigned long long)d) ++l;                       // ceil(log2 d)
    r.m = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1ull);
    r.s1 = l < 1u ? l : 1u;
    r.s2 = l > 0u ? l - 1u : 0u;
    return r;
}
#ifdef __CUDACC__
__host__ __device__
#endif
static inline uint32_t rtb_udiv(uint32_t n, RtbUdiv d) {
#ifdef __CUDA_ARCH__
    const uint32_t t = __umulhi(d.m, n);
#else
    const uint32_t t = (uint32_t)(((unsigned long long)d.m * n) >> 32);
#endif
    return (t + ((n - t) >> d.s1)) >> d.s2;
}

struct ViewDev {
    uint32_t width, height;
    float orig[3], cam[3], vu[3], vv[3];
    uint32_t maxdepth, spp;
    uint64_t seed;
    uint32_t s_begin, s_end;
    uint32_t flags;
    // band partition: tile row ty belongs to rank ty % world
    uint32_t tile_rank, tile_world;
    uint32_t tiles_x;        // ceil(width / 16)
    uint32_t my_tile_rows;   // number of this rank's bands rendered by this launch ...
    uint32_t band_begin;     // ... starting at its band_begin-th band (chunked rendering; 0 = from the first)
    uint32_t compact;        // 1: output rows are packed (own bands only), 0: full-frame indexing
    // derived once on the host (make_view), so that no ray pays for them:
    float vu_delta[3], vv_delta[3];   // vu * (1/width), vv * (1/height) — pixel_ray :1379-1380, the same IEEE operations
    uint64_t seed_mixed;              // splitmix64(seed), the first of the three mixing steps of Rng::seed
    RtbUdiv div_band, div_tx8;        // division of a warp-tile index by the tiles of a band / of a band half (slot_to_pixel)
};

// EXTENSION (rtb_ext.cu): analytic spheres travel as pseudo-triangles tagged in `kind`; the light of the reference's
// commented-out shadow code (raytrace.rs:594-610).
#define RTB_PRIM_SPHERE 0x100u
struct ExtParams {
    float light[4];        // LightSource: orig xyz, len2
    uint32_t has_light;
};

struct TraceCounters {
    unsigned long long rays;
    unsigned long long node_tests;
    unsigned long long tri_tests;
    unsigned long long node_tests_bounce;   // share of the two above spent on bounce rays (incl. their shadow rays)
    unsigned long long tri_tests_bounce;
    unsigned long long stalled;             // fused path kernel: watchdog trips (0 unless something is badly wrong)
};

// One compute lane of a GPU: a stream with its own path workspace (bounce queue, mix stacks, sample sums).
struct GpuLane {
    cudaStream_t st = nullptr;     // lane 0 runs on the caller's stream and leaves this null
    cudaEvent_t done = nullptr;
    void* d_ws = nullptr;
    size_t ws_bytes = 0;
    uint32_t epoch = 0;            // launches of the path kernel on this workspace (the bounce queue's entry tag)
    cudaEvent_t busy = nullptr;    // recorded after the last kernel that used this workspace ...
    cudaStream_t busy_stream = nullptr;   // ... on this stream: a frame issued on another stream waits for it first
    bool used = false;
    cudaEvent_t stage_ev[RTB_N_STAGES + 1] = {};   // RTB_FLAG_TIMING: stage boundaries of one sample
    float stage_ms[RTB_N_STAGES] = {};
};

// Per-GPU state of a scene.
struct GpuScene {
    int device = -1;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t copy_stream = nullptr;           // D2H of finished pieces overlaps the next piece's kernels
    cudaEvent_t chunk_ev[RTB_MAX_CHUNKS] = {};
    float4* d_nodes = nullptr;
    float4* d_tri = nullptr;
    float4* d_shade = nullptr;
    uint32_t* d_prim_order = nullptr;  // leaf slot -> original triangle index
    TraceCounters* d_counters = nullptr;
    // frame buffers for rtb_render (host-output mode), grown on demand
    float4* d_rgba = nullptr;
    uint32_t* d_prim = nullptr;
    float* d_t = nullptr;
    size_t fb_pixels = 0;
    uint8_t* d_rgb8 = nullptr;         // rtb_render_rgb8: quantised frame, 3 B per pixel
    size_t rgb8_pixels = 0;
    uint32_t n_nodes = 0;
    uint32_t height = 0;
    float4* d_nodes4 = nullptr;
    uint32_t n_nodes4 = 0, depth4 = 0, stack4_need = 0;
    uint4* d_nodes8 = nullptr;
    float4* d_tri8 = nullptr;
    float4* d_shade8 = nullptr;
    uint32_t n_nodes8 = 0, depth8 = 0;
    GpuLane lanes[RTB_MAX_LANES];
    cudaEvent_t fork_ev = nullptr;
    // rtb_render_progressive: "my sample sums are complete" (waited for by every peer's reduce), reduce timing, peer staging
    cudaEvent_t prog_done = nullptr, red0 = nullptr, red1 = nullptr;
    float4* d_stage = nullptr;
    size_t stage_pixels = 0;
    // scenes with analytic spheres or a light are rendered by the extension renderer (rtb_ext.cu)
    bool has_spheres = false;
    ExtParams ext = {};
};

struct rtb_scene {
    std::vector<GpuScene> gpu;
    RtbSceneInfo info;
    std::mutex mu;          // a handle serialises its render calls (workspaces, counters and frame buffers are per handle)
};

// ---- error plumbing ---------------------------------------------------------
void rtb_set_error(const std::string& msg);
int rtb_cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define RTB_CUDA(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) return rtb_cuda_fail(e_, #call, __FILE__, __LINE__); \
    } while (0)

// ---- implemented in rtb_lbvh.cu ---------------------------------------------
struct BuildResult {
    float4* d_nodes = nullptr;
    float4* d_tri = nullptr;
    float4* d_shade = nullptr;
    uint32_t* d_prim_order = nullptr;
    uint32_t n_nodes = 0, n_leaves = 0, max_leaf = 0, tree_height = 0;
    uint32_t n_refs = 0;             // primitive references in the tree (>= n_prims: long primitives are split)
    float4* d_nodes4 = nullptr;      // 4-wide collapse, 8 x float4 per node
    uint32_t n_nodes4 = 0, depth4 = 0, stack4_need = 0;   // stack4_need: exact BVH4 stack bound (0 = use 3 per level)
    uint4* d_nodes8 = nullptr;       // 8-wide compressed collapse, 5 x uint4 per node, and the references in its order
    float4* d_tri8 = nullptr;
    float4* d_shade8 = nullptr;
    uint32_t n_nodes8 = 0, depth8 = 0;
    float lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    float ms_build = 0.f;
    uint32_t launches = 0;
};
// d_tris: the caller's RtbTriangle array on the device; d_keep: original indices of the n_prims kept triangles.
int rtb_build_lbvh(const RtbTriangle* d_tris, const uint32_t* d_keep, uint32_t n_prims, cudaStream_t stream,
                   BuildResult* out);

// Self-test of the hand-written radix sort / scan (rtb_sort.cuh) against the host.
int rtb_sort_selftest(uint32_t n, int key_bits, uint64_t seed);

// ---- implemented in rtb_scene.cu ----------------------------------------------
// Scene assembly on the device: mesh instances -> `Triangle` records (make_triangle), root-cube cull + compaction.
int rtb_launch_assemble(const float* d_verts, uint32_t nverts, const uint32_t* d_faces, uint32_t nfaces,
                        const RtbMeshInstance* d_inst, uint32_t n_inst, RtbTriangle* d_out, cudaStream_t stream,
                        uint32_t* h_first_bad);
int rtb_launch_cull(const RtbTriangle* d_tris, uint32_t n, const float root_orig[3], float root_len2,
                    cudaStream_t stream, uint32_t** d_keep_out, uint32_t* n_keep);

// ---- implemented in rtb_trace.cu ----------------------------------------------
// Launches the trace kernel for the tile rows owned by (tile_rank, tile_world).
int rtb_launch_trace(const SceneDev& sc, const ViewDev& vw, float4* d_rgba, uint32_t* d_prim, float* d_t,
                     TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches);
// ---- implemented in rtb_ext.cu --------------------------------------------------
int rtb_launch_trace_ext(const SceneDev& sc, const ViewDev& vw, const ExtParams& ex, float4* d_rgba, uint32_t* d_prim,
                         float* d_t, TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches);
// ---- implemented in rtb_wavefront.cu -------------------------------------------
// The default renderer: one persistent kernel per sample (primary phase + bounce phase, rtb_wavefront.cu).
// counters->rays receives the BOUNCE rays only; the caller adds the primary rays (valid pixels x samples).
// workspace: rtb_wf_workspace_bytes, ZERO-FILLED when allocated; epoch: the workspace's launch counter (queue entry tags).
size_t rtb_wf_workspace_bytes(uint32_t n_slots, uint32_t maxdepth, bool multisample);
// stage_ev (nullable): 5 events recorded at the stage boundaries of every sample; stage_ms accumulates their gaps
// (this synchronises the stream once per sample: timing mode only).
// ext: non-null for an extension scene (analytic spheres in the leaves and / or a light): the EXT variants of the kernel.
int rtb_launch_wavefront(const SceneDev& sc, const ViewDev& vw, const ExtParams* ext, void* workspace, uint32_t* epoch,
                         float4* d_rgba, uint32_t* d_prim, float* d_t, TraceCounters* d_counters, cudaStream_t stream,
                         uint32_t* launches, cudaEvent_t* stage_ev = nullptr, float* stage_ms = nullptr);

int rtb_launch_quantize(const float4* d_rgba, uint64_t npix, uint8_t* d_rgb, cudaStream_t stream);
int rtb_launch_scale(float4* d_rgba, uint64_t npix, float inv_spp, cudaStream_t stream);
// Fused cross-GPU reduce of per-GPU sample sums over peer memory: out[i] = (sum_g bufs.p[g][i]) * inv_spp for
// pixels [first, first+count).  The pointers travel by value (kernel parameters): no pointer table in device memory.
struct RtbPeerBufs { const float4* p[RTB_MAX_GPUS]; };
int rtb_launch_peer_reduce(const RtbPeerBufs& bufs, int n_bufs, float inv_spp, uint64_t first,
                           uint64_t count, float4* d_out, cudaStream_t stream);
