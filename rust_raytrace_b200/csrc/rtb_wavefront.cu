// rtb_wavefront.cu — the hot path as a wavefront pipeline (the default renderer).
//
// The one-kernel version (rtb_trace.cu) keeps only ~9 of 32 lanes busy per issued instruction
// (ncu, profiles/r1_v1_*): rays of one warp leave the scene at different times and the surviving
// Matte/mirror paths bounce up to maxdepth times.  Here the same arithmetic is split into stages:
//
//   (raygen)      Viewport::pixel_ray (raytrace.rs:1374-1394), generated in place by the two kernels below
//   k_wf_trace    PERSISTENT kernel over the coherent primary rays: warps pull pixel slots, every lane that
//                 finishes its ray is refilled (ballot + one atomic per refill), so the traversal loop runs
//                 with ~27 of 32 lanes; closest hit -> hit record
//   k_wf_shade    one thread per primary ray: Triangle::intersects classification, color_ray (:1199-1254):
//                 terminal paths add their sample to the pixel; bouncing paths push (colour, alpha) on the
//                 pixel's mix stack and append the bounce ray to the bounce queue (warp-aggregated atomic)
//   k_wf_bounce   PERSISTENT kernel that takes every bounce path to its end: trace, shade, bounce again in
//                 place (project_ray's recursion, iteratively), refilling a lane from the bounce queue only
//                 when its path dies.  A first version ran one trace + one shade kernel per bounce level;
//                 every level then paid its own drain (a few 100-step rays keep a few warps alive, >= ~0.15 ms
//                 per level however few rays it holds) and ran with 9-16 of 32 lanes.
//   k_wf_tally    per-sample counter reset
//
// Per-pixel state lives in HBM between stages (hit 8 B, bounce ray 32 B, mix stack 16 B/level, RNG 8 B).
#include <algorithm>
#include <cstdlib>

#include "rtb_device.cuh"

using namespace rtbdev;

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t LEAF_FLAG = 0x80000000u;
constexpr int WF_BLOCK = 128;           // threads per CTA of the persistent kernels
constexpr int WF_OVF = 3 * (RTB_STACK / 2 + 1) + 2;   // worst-case BVH4 stack (tree height < RTB_STACK), thread-local overflow part
constexpr int WF_SMEM_STACK = 8;        // stack entries per thread kept in shared memory
constexpr uint32_t WF_SENTINEL = 0x7fffffffu;   // bottom of a whole-in-shared-memory stack (never a node index: < 2^31 nodes)
#ifndef WF_TRACE_MIN_BLOCKS
#define WF_TRACE_MIN_BLOCKS 4
#endif
#ifndef WF_BOUNCE_MIN_BLOCKS
#define WF_BOUNCE_MIN_BLOCKS 8          // 64 registers: 32 warps/SM instead of 24 at the natural 80
#endif
// Tunables (env RTB_WF_DESCEND / RTB_WF_REFILL override for experiments).  Measured on B200, 4K teapot frame, current
// pipeline with exact work fetch: (descend, refill) = (4,16) 2.82 ms, (4,20) 2.79, (4,24) 2.80, (4,28) 2.86,
// (3,20) 2.79, (6,20) 2.87, (8,20) 2.90.
constexpr uint32_t WF_DESCEND_MAX = 3;  // BVH4 node visits per lane per round before leaves are processed (binned-SAH tree: 4 -> 3
                                        // = 2.133 -> 2.102 ms on the 4K teapot frame, 1.120 -> 1.091 ms on the 1 M field)
constexpr uint32_t WF_REFILL_MIN = 20;  // bounce kernel: service (shade / refill) lanes once at least this many wait
constexpr uint32_t WF_REFILL_MIN_PRIMARY = 24;   // primary kernel: refill once at least this many lanes are done
struct WfTune { uint32_t descend_max, refill_min; int smem_depth; uint32_t pool_node_min; };

// slot -> pixel.  Slots enumerate 8x4 warp tiles inside the 8-row bands this launch renders.
struct Pixel { uint32_t row, col, out_idx; bool inside; };
__device__ __forceinline__ Pixel slot_to_pixel(const ViewDev& vw, uint32_t slot) {
    const uint32_t wt = slot >> 5, l = slot & 31u;
    const uint32_t tiles_x8 = (vw.width + 7u) >> 3;
    const uint32_t per_band = 2u * tiles_x8;
    const uint32_t band_rel = wt / per_band, rem = wt - band_rel * per_band;
    const uint32_t band_local = band_rel + vw.band_begin;
    const uint32_t half = rem / tiles_x8, tx8 = rem - half * tiles_x8;
    Pixel p;
    const uint32_t in_band = half * 4u + (l >> 3);
    p.row = (band_local * vw.tile_world + vw.tile_rank) * RTB_TILE_H + in_band;
    p.col = tx8 * 8u + (l & 7u);
    p.inside = (p.row < vw.height) && (p.col < vw.width);
    const uint32_t out_row = vw.compact ? (band_local * RTB_TILE_H + in_band) : p.row;
    p.out_idx = out_row * vw.width + p.col;
    return p;
}

// ---------------------------------------------------------------------------
// stage 0: primary rays.  Not a kernel of its own any more: the trace kernel generates a slot's ray when a lane
// takes the slot, and the shade kernel regenerates it (same arithmetic, same RNG draws, ~80 instructions) instead
// of reading it back.  A separate raygen kernel cost 0.08 ms of the 4K frame plus 64 B per pixel written to and
// read from HBM twice.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool primary_ray(const ViewDev& vw, uint32_t slot, uint32_t smp, V3* o, V3* d, Rng* g,
                                            Pixel* px_out) {
    const Pixel px = slot_to_pixel(vw, slot);
    *px_out = px;
    if (!px.inside) return false;
    g->seed(vw.seed, (uint64_t)px.row * vw.width + px.col, smp);
    float u_off = 0.5f, v_off = 0.5f;
    if (vw.spp != 1) { u_off = g->next_f32(); v_off = g->next_f32(); }   // :1382-1386
    gen_primary(vw, px.row, px.col, u_off, v_off, o, d);
    return true;
}

// ---------------------------------------------------------------------------
// traversal machinery shared by the two persistent kernels
// ---------------------------------------------------------------------------
struct TravState {
    V3 o, d;
    float ix, iy, iz, ox, oy, oz;
    float tbest;
    Hit h;
    uint32_t cur;      // LEAF_FLAG | first<<3 | count, or the index of a BVH4 node
    int sp;
};

__device__ __forceinline__ uint32_t root_code_of(const SceneDev&) {
    return 0u;   // BVH4 node 0 (an empty scene has one node whose four slots are all empty)
}

__device__ __forceinline__ void start_ray(TravState& s, V3 o, V3 d, uint32_t root_code, int sp0 = 0) {
    s.o = o; s.d = d;
    // |1/d| clamped so a zero direction component gives +-huge, not inf (inf - inf = NaN in the fused slab test)
    s.ix = fminf(fmaxf(1.0f / d.x, -1e30f), 1e30f);
    s.iy = fminf(fmaxf(1.0f / d.y, -1e30f), 1e30f);
    s.iz = fminf(fmaxf(1.0f / d.z, -1e30f), 1e30f);
    s.ox = -o.x * s.ix; s.oy = -o.y * s.iy; s.oz = -o.z * s.iz;
    s.tbest = FLT_MAX;
    s.h.t = FLT_MAX; s.h.slot = -1; s.h.orig = 0xffffffffu;
    s.sp = sp0;
    s.cur = root_code;
}

// One traversal round for the lanes with trav == true; every lane of the warp must call it.
//   phase 1: at most `k_nodes` BVH4 node visits (unbounded descent, the classic while-while, left lanes that had
//            reached a leaf waiting for the slowest lane of the warp)
//   phase 2: the leaf this lane holds, if it reached one
// trav becomes false when the lane's ray is finished (result in s.h).  The stack lives in SHARED memory, one
// 4-byte code per entry, [depth][thread]: as thread-local arrays the stacks went through L1 and evicted the BVH.
//
// Tried and dropped (B200, 4K teapot frame, bounce kernel 2.3 ms): postponing leaves into a per-lane queue so
// that node visits and triangle tests run in separate, fuller phases — 15-25% slower (more box and triangle
// tests because t_best tightens later, plus the queue bookkeeping), see DESIGN.md "What did not work".
template <bool STATS, bool OVF>
__device__ __forceinline__ void trav_round(const SceneDev& sc, TravState& s, bool& trav, uint32_t* stack,
                                           uint32_t* ovf, int smem_depth, uint32_t k_nodes, unsigned long long& n_node,
                                           unsigned long long& n_tri) {
    // OVF: the first `smem_depth` stack entries live in shared memory, deeper ones (rare) in thread-local memory
    // (!OVF: the whole worst-case stack fits in shared memory and the depth tests compile away, ~2 % faster):
    // sizing the shared stack for the worst case (3 entries per BVH4 level, 37 for the teapot scene) took 143 KB
    // of each SM's 256 KB and grows with the tree height (a 1 M-triangle scene would drop to 4 CTAs/SM).  Measured on
    // B200, 4K teapot frame: 40 entries 3.22 ms, 16: 3.21, 12: 3.20, 8: 3.18, 6: 3.15, 4: 3.16 — L1 capacity is not the limiter.
    auto pop = [&]() {
        if (!OVF) { --s.sp; s.cur = stack[s.sp * WF_BLOCK]; trav = s.cur != WF_SENTINEL; return; }
        if (s.sp == 0) { trav = false; return; }
        --s.sp;
        s.cur = (s.sp < smem_depth) ? stack[s.sp * WF_BLOCK] : ovf[s.sp - smem_depth];
    };
    auto push = [&](uint32_t code) {
        if (!OVF || s.sp < smem_depth) stack[s.sp * WF_BLOCK] = code; else ovf[s.sp - smem_depth] = code;
        ++s.sp;
    };
#pragma unroll 1
    for (uint32_t it = 0; it < k_nodes; ++it) {
        const bool go = trav && !(s.cur & LEAF_FLAG);
        if (!__any_sync(FULL, go)) break;
        if (!go) continue;
        // one BVH4 node = one 128-byte line: 4 child boxes (SoA) + 4 child codes
        const float4* np = sc.nodes4 + 8u * s.cur;
        float4 lx, hx, ly, hy, lz, hz, cdf, pad_;
        ldg256(np + 0, lx, hx); ldg256(np + 2, ly, hy); ldg256(np + 4, lz, hz); ldg256(np + 6, cdf, pad_);
        const uint4 cd = make_uint4(__float_as_uint(cdf.x), __float_as_uint(cdf.y), __float_as_uint(cdf.z), __float_as_uint(cdf.w));
        if (STATS) n_node += 4;
        const float INF = __int_as_float(0x7f800000);
        float key[4];
        uint32_t code[4] = {cd.x, cd.y, cd.z, cd.w};
#define RTB_SLAB(c, LX, HX, LY, HY, LZ, HZ)                                                             \
        {                                                                                               \
            const float x0 = fmaf(LX, s.ix, s.ox), x1 = fmaf(HX, s.ix, s.ox);                           \
            const float y0 = fmaf(LY, s.iy, s.oy), y1 = fmaf(HY, s.iy, s.oy);                           \
            const float z0 = fmaf(LZ, s.iz, s.oz), z1 = fmaf(HZ, s.iz, s.oz);                           \
            const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));    \
            const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000005f;    \
            key[c] = ((tn <= tf) && (tn <= s.tbest)) ? tn : INF;   /* an empty slot has a NaN box */     \
        }
        RTB_SLAB(0, lx.x, hx.x, ly.x, hy.x, lz.x, hz.x)
        RTB_SLAB(1, lx.y, hx.y, ly.y, hy.y, lz.y, hz.y)
        RTB_SLAB(2, lx.z, hx.z, ly.z, hy.z, lz.z, hz.z)
        RTB_SLAB(3, lx.w, hx.w, ly.w, hy.w, lz.w, hz.w)
#undef RTB_SLAB
        // sort the (up to 4) hits near-to-far: 5 compare-exchanges
#define RTB_CE(a, b)                                                                  \
        {                                                                             \
            const bool sw = key[a] > key[b];                                          \
            const float ka = sw ? key[b] : key[a], kb = sw ? key[a] : key[b];         \
            const uint32_t ca_ = sw ? code[b] : code[a], cb_ = sw ? code[a] : code[b]; \
            key[a] = ka; key[b] = kb; code[a] = ca_; code[b] = cb_;                   \
        }
        RTB_CE(0, 1) RTB_CE(2, 3) RTB_CE(0, 2) RTB_CE(1, 3) RTB_CE(1, 2)
#undef RTB_CE
        if (!OVF) {
            // whole stack in shared memory, entry 0 = WF_SENTINEL: no branches — the three far children are stored
            // unconditionally (a store above the top is harmless), the top only moves past the ones that were hit,
            // and a pop that reaches the sentinel ends the ray
            stack[s.sp * WF_BLOCK] = code[3]; s.sp += (key[3] < INF) ? 1 : 0;
            stack[s.sp * WF_BLOCK] = code[2]; s.sp += (key[2] < INF) ? 1 : 0;
            stack[s.sp * WF_BLOCK] = code[1]; s.sp += (key[1] < INF) ? 1 : 0;
            const bool miss = key[0] == INF;
            s.sp -= miss ? 1 : 0;
            s.cur = miss ? stack[s.sp * WF_BLOCK] : code[0];
            trav = s.cur != WF_SENTINEL;
        } else if (key[0] == INF) {
            pop();
        } else {
            s.cur = code[0];                                   // nearest first, the rest far-to-near on the stack
            if (key[3] < INF) push(code[3]);
            if (key[2] < INF) push(code[2]);
            if (key[1] < INF) push(code[1]);
        }
    }
    if (trav && (s.cur & LEAF_FLAG)) {
        const uint32_t first = (s.cur & ~LEAF_FLAG) >> 3, cnt = s.cur & 7u;
        // software pipeline: the first 32 bytes of triangle k+1 are in flight while triangle k is tested
        const float4* q = sc.tri + (size_t)RTB_TRI_F4 * first;
        float4 a0 = __ldg(q), a1 = __ldg(q + 1);
        for (uint32_t k = first; k < first + cnt; ++k, q += RTB_TRI_F4) {
            float t;
            if (STATS) ++n_tri;
            const float4 c0 = a0, c1 = a1;
            if (k + 1u < first + cnt) { a0 = __ldg(q + RTB_TRI_F4); a1 = __ldg(q + RTB_TRI_F4 + 1); }
            if (tri_test_pre(q, c0, c1, s.o, s.d, s.h.slot >= 0, s.h.t, &t)) {
                const uint32_t orig = __float_as_uint(c1.w);
                if (s.h.slot < 0 || t < s.h.t || (t == s.h.t && orig < s.h.orig)) {
                    s.h.t = t; s.h.slot = (int)k; s.h.orig = orig;
                    if (t < s.tbest) s.tbest = t;      // a NaN t never tightens the bound
                }
            }
        }
        pop();
    }
}

// validation mode (RTB_FLAG_BRUTE): linear scan over every primitive, no BVH
template <bool STATS>
__device__ __forceinline__ void brute_scan(const SceneDev& sc, TravState& s, unsigned long long& n_tri) {
    for (uint32_t k = 0; k < sc.n_prims; ++k) {
        float t;
        if (STATS) ++n_tri;
        const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
        if (tri_test(q, s.o, s.d, s.h.slot >= 0, s.h.t, &t)) {
            const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
            if (s.h.slot < 0 || t < s.h.t || (t == s.h.t && orig < s.h.orig)) { s.h.t = t; s.h.slot = (int)k; s.h.orig = orig; }
        }
    }
}

// Warp-level work fetch: the lanes in `need` take consecutive queue indices, ONE global atomic per refill that
// reserves exactly popc(need) rays.  Reserving a fixed portion per warp instead (128 rays per atomic at first) left
// the kernel waiting for the warps that happened to draw one portion more: on B200, 4K teapot frame, 128 rays per
// fetch 3.18 ms, 32 rays 2.99 ms, exact 2.95 ms; on a 1/8 band share of the frame 0.97 / 0.58 / 0.52 ms.
struct WorkFetch {
    bool exhausted = false;
    // returns this lane's queue index or 0xffffffff
    __device__ __forceinline__ uint32_t fetch(unsigned need, bool me, uint32_t n, uint32_t* counter, unsigned lane) {
        if (exhausted || need == 0u) return 0xffffffffu;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(counter, (uint32_t)__popc(need));
        base = __shfl_sync(FULL, base, 0);
        if (base >= n) { exhausted = true; return 0xffffffffu; }
        const uint32_t idx = base + __popc(need & ((1u << lane) - 1u));
        return (me && idx < n) ? idx : 0xffffffffu;
    }
};

// ---------------------------------------------------------------------------
// stage 1: closest hit of the primary rays, persistent threads with per-lane refill
// ---------------------------------------------------------------------------
template <bool STATS, bool OVF>
__global__ void __launch_bounds__(WF_BLOCK, WF_TRACE_MIN_BLOCKS)
k_wf_trace(const SceneDev sc, const ViewDev vw, uint32_t smp, uint32_t n,
           float2* __restrict__ hit_out, uint32_t* __restrict__ work_counter, uint32_t brute, const WfTune tune,
           TraceCounters* __restrict__ counters) {
    const uint32_t descend_max = tune.descend_max, refill_min = tune.refill_min;
    const int smem_depth = tune.smem_depth;
    extern __shared__ uint32_t smem_stack[];
    uint32_t* const stack = smem_stack + threadIdx.x;     // entry k lives at stack[k * WF_BLOCK]
    uint32_t ovf[OVF ? WF_OVF : 1];
    const unsigned lane = threadIdx.x & 31u;
    WorkFetch wf;
    wf.exhausted = (n == 0u);
    const uint32_t root_code = root_code_of(sc);
    bool trav = false;
    uint32_t ray_id = 0;
    TravState s;
    constexpr int SP0 = OVF ? 0 : 1;         // !OVF: stack entry 0 holds the sentinel
    if (!OVF) stack[0] = WF_SENTINEL;
    start_ray(s, mk(0.f, 0.f, 0.f), mk(1.f, 1.f, 1.f), root_code, SP0);
    unsigned long long n_node = 0, n_tri = 0;

    for (;;) {
        const unsigned need = __ballot_sync(FULL, !trav);
        if (!wf.exhausted && (uint32_t)__popc(need) >= refill_min) {
            const uint32_t idx = wf.fetch(need, !trav, n, work_counter, lane);
            if (idx != 0xffffffffu) {
                V3 o, d; Rng g; Pixel px;
                if (primary_ray(vw, idx, smp, &o, &d, &g, &px)) {   // slots outside the image have no ray
                    ray_id = idx;
                    start_ray(s, o, d, root_code, SP0);
                    trav = true;
                }
            }
        }
        if (__ballot_sync(FULL, trav) == 0u) {
            if (wf.exhausted) break;
            continue;
        }
        for (;;) {
            const bool was = trav;
            if (brute) { if (trav) { brute_scan<STATS>(sc, s, n_tri); trav = false; } }
            else trav_round<STATS, OVF>(sc, s, trav, stack, ovf, smem_depth, descend_max, n_node, n_tri);
            if (was && !trav) __stcs(hit_out + ray_id, make_float2(s.h.t, __int_as_float(s.h.slot)));
            const unsigned act = __ballot_sync(FULL, trav);
            if (act == 0u) break;
            if (!wf.exhausted && (32u - __popc(act)) >= refill_min) break;
        }
    }
    if (STATS) {
        for (int off = 16; off > 0; off >>= 1) {
            n_node += __shfl_xor_sync(FULL, n_node, off);
            n_tri += __shfl_xor_sync(FULL, n_tri, off);
        }
        if (lane == 0) { atomicAdd(&counters->node_tests, n_node); atomicAdd(&counters->tri_tests, n_tri); }
    }
}

// ---------------------------------------------------------------------------
// path bookkeeping shared by the shade kernel and the bounce kernel
// ---------------------------------------------------------------------------
struct PathBuffers {
    float4* stack;            // [maxdepth][n_slots]  (colour, alpha) per bounce level
    uint64_t* rng_state;      // [n_slots]
    float4* acc;              // [n_slots] running sample sum (multi-sample only)
    float4* rgba; uint32_t* prim_out; float* t_out;
    uint32_t n_slots;
};

// A path of `levels` bounces ended with colour `term`: unwind the recursion innermost-first (mix_color :299-301),
// then add the sample to the pixel and, on the last sample, scale and store it (walk_ray_set :1422-1426).
__device__ __forceinline__ void finish_path(const ViewDev& vw, const PathBuffers& pb, uint32_t slot, uint32_t smp,
                                            uint32_t levels, V3 term) {
    V3 c = term;
    for (int k = (int)levels - 1; k >= 0; --k) {
        const float4 e = pb.stack[(size_t)k * pb.n_slots + slot];
        c = mix_color(mk(e.x, e.y, e.z), c, e.w);
    }
    V3 sum = mk(0.f, 0.f, 0.f);
    if (smp != vw.s_begin) { const float4 p = pb.acc[slot]; sum = mk(p.x, p.y, p.z); }
    sum = vadd(sum, c);
    if (smp + 1u == vw.s_end) {
        if (!(vw.flags & RTB_FLAG_SUM_ONLY)) sum = vmul(sum, __fdiv_rn(1.0f, (float)vw.spp));
        const Pixel px = slot_to_pixel(vw, slot);
        __stcs(pb.rgba + px.out_idx, make_float4(sum.x, sum.y, sum.z, 0.f));
    } else {
        pb.acc[slot] = make_float4(sum.x, sum.y, sum.z, 0.f);
    }
}

// ---------------------------------------------------------------------------
// stage 2: shade the primary hits, emit the bounce queue
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_shade(const SceneDev sc, const ViewDev vw, const PathBuffers pb,
                                                  const float2* __restrict__ hit, uint32_t n, uint32_t smp,
                                                  float4* __restrict__ qo_out, float4* __restrict__ qd_out,
                                                  uint32_t* __restrict__ n_out) {
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool emit = false;
        V3 no = mk(0.f, 0.f, 0.f), nd = mk(0.f, 0.f, 0.f);
        const uint32_t slot = i;
        V3 o = mk(0.f, 0.f, 0.f), d = mk(0.f, 0.f, 0.f);
        Rng g; g.state = 0;
        Pixel px; px.inside = false;
        // a missed pixel (64 % of the 4K teapot frame) needs no ray, only its position
        float2 hr = make_float2(0.f, __int_as_float(-1));
        bool live = false;
        if (i < n) {
            px = slot_to_pixel(vw, slot);
            if (px.inside) {
                hr = __ldcs(hit + i);
                live = __float_as_int(hr.y) < 0 ? true : primary_ray(vw, slot, smp, &o, &d, &g, &px);
            }
        }
        if (live) {
            const int prim_slot = __float_as_int(hr.y);
            const float t = hr.x;
            V3 term = mk(0.f, 0.f, 0.f);
            uint32_t levels = 0;
            if (smp == 0u && (pb.prim_out || pb.t_out)) {
                if (pb.prim_out) pb.prim_out[px.out_idx] = prim_slot >= 0 ? __float_as_uint(__ldg(sc.tri + (size_t)RTB_TRI_F4 * prim_slot + 1).w) : 0u;
                if (pb.t_out) pb.t_out[px.out_idx] = prim_slot >= 0 ? t : 0.0f;
            }
            if (prim_slot < 0) {
                term = sky_color();                                            // project_ray miss :1284
            } else {
                V3 color;
                float alpha = 0.f;
                if (shade_hit(sc, prim_slot, t, o, d, g, &color, &alpha, &no, &nd) == 0) {
                    term = color;
                } else {
                    pb.stack[slot] = make_float4(color.x, color.y, color.z, alpha);   // level 0
                    levels = 1u;
                    if (1u < vw.maxdepth) { emit = true; pb.rng_state[slot] = g.state; }
                    // else: project_ray(depth 0) returns black (:1261), term stays black
                }
            }
            if (!emit) finish_path(vw, pb, slot, smp, levels, term);
        }
        // stream compaction of the surviving paths into the bounce queue
        const unsigned m = __ballot_sync(FULL, emit);
        if (m) {
            uint32_t base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(n_out, (uint32_t)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (emit) {
                const uint32_t j = base + __popc(m & ((1u << lane) - 1u));
                qo_out[j] = make_float4(no.x, no.y, no.z, __uint_as_float(slot));
                qd_out[j] = make_float4(nd.x, nd.y, nd.z, 0.f);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// stage 3: every bounce path to its end, persistent
// ---------------------------------------------------------------------------
template <bool STATS, bool OVF>
__global__ void __launch_bounds__(WF_BLOCK, WF_BOUNCE_MIN_BLOCKS)
k_wf_bounce(const SceneDev sc, const ViewDev vw, const PathBuffers pb, const float4* __restrict__ qo,
            const float4* __restrict__ qd, const uint32_t* __restrict__ n_ptr, uint32_t smp,
            uint32_t* __restrict__ work_counter, uint32_t brute, const WfTune tune,
            TraceCounters* __restrict__ counters) {
    const uint32_t descend_max = tune.descend_max, refill_min = tune.refill_min;
    const int smem_depth = tune.smem_depth;
    extern __shared__ uint32_t smem_stack[];
    uint32_t* const stack = smem_stack + threadIdx.x;
    uint32_t ovf[OVF ? WF_OVF : 1];
    const uint32_t n = *n_ptr;
    const unsigned lane = threadIdx.x & 31u;
    WorkFetch wf;
    wf.exhausted = (n == 0u);
    const uint32_t root_code = root_code_of(sc);

    bool has_path = false;   // this lane carries a path (pixel) ...
    bool trav = false;       // ... whose current ray is still being traversed
    uint32_t slot = 0, level = 0;
    Rng g; g.state = 0;
    TravState s;
    constexpr int SP0 = OVF ? 0 : 1;         // !OVF: stack entry 0 holds the sentinel
    if (!OVF) stack[0] = WF_SENTINEL;
    start_ray(s, mk(0.f, 0.f, 0.f), mk(1.f, 1.f, 1.f), root_code, SP0);
    unsigned long long n_node = 0, n_tri = 0, n_rays = 0;

    for (;;) {
        // ---- service: shade finished rays (bounce again in place or end the path), then refill dead lanes ----
        const unsigned waiting = __ballot_sync(FULL, !trav);
        if ((uint32_t)__popc(waiting) >= refill_min || __ballot_sync(FULL, trav) == 0u) {
            if (has_path && !trav) {
                // project_ray :1283-1293 for the ray that just finished; `level` entries are on the mix stack
                V3 term = mk(0.f, 0.f, 0.f);
                bool ended = true;
                if (s.h.slot < 0) {
                    term = sky_color();
                } else {
                    V3 color, no, nd;
                    float alpha = 0.f;
                    if (shade_hit(sc, s.h.slot, s.h.t, s.o, s.d, g, &color, &alpha, &no, &nd) == 0) {
                        term = color;
                    } else {
                        pb.stack[(size_t)level * pb.n_slots + slot] = make_float4(color.x, color.y, color.z, alpha);
                        ++level;
                        if (level < vw.maxdepth) {          // project_ray(depth-1) with depth-1 > 0
                            start_ray(s, no, nd, root_code, SP0);
                            trav = true; ended = false; ++n_rays;
                        }                                   // else depth 0: black (:1261), not counted
                    }
                }
                if (ended) { finish_path(vw, pb, slot, smp, level, term); has_path = false; }
            }
            const unsigned need = __ballot_sync(FULL, !has_path);
            const uint32_t idx = wf.fetch(need, !has_path, n, work_counter, lane);
            if (idx != 0xffffffffu) {
                const float4 ro = __ldcs(qo + idx), rd = __ldcs(qd + idx);
                slot = __float_as_uint(ro.w);
                level = 1u;
                g.state = pb.rng_state[slot];
                start_ray(s, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z), root_code, SP0);
                has_path = true; trav = true; ++n_rays;
            }
        }
        if (__ballot_sync(FULL, has_path) == 0u) {
            if (wf.exhausted) break;
            continue;
        }
        // ---- traversal rounds until enough lanes wait for service ----
        for (;;) {
            if (brute) { if (trav) { brute_scan<STATS>(sc, s, n_tri); trav = false; } }
            else trav_round<STATS, OVF>(sc, s, trav, stack, ovf, smem_depth, descend_max, n_node, n_tri);
            const unsigned act = __ballot_sync(FULL, trav);
            if (act == 0u || (32u - __popc(act)) >= refill_min) break;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_xor_sync(FULL, n_rays, off);
        if (STATS) {
            n_node += __shfl_xor_sync(FULL, n_node, off);
            n_tri += __shfl_xor_sync(FULL, n_tri, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&counters->rays, n_rays);
        if (STATS) {
            atomicAdd(&counters->node_tests, n_node); atomicAdd(&counters->tri_tests, n_tri);
            atomicAdd(&counters->node_tests_bounce, n_node); atomicAdd(&counters->tri_tests_bounce, n_tri);
        }
    }
}

// ---------------------------------------------------------------------------
// stage 3, pool variant: the same function as k_wf_bounce, organised for lane utilisation.
//
// k_wf_bounce keeps one ray per lane in registers, so an instruction of the traversal loop only serves the lanes
// whose ray happens to need that step: 14.5 of 32 lanes in a node visit, 8-11 in a triangle test, ~20 in shading
// (ncu r1_v5).  Here a warp owns a POOL of 64 rays whose state lives in shared memory; before every step the warp
// counts how many pooled rays want a node visit / a leaf / shading / a refill, picks the kind with the most
// candidates and deals (up to) 32 of them out to its lanes (ballot + popc ranks), so each step runs closer to full
// width.  The price is the state traffic: a node step loads 9 and stores 2-5 words of ray state per lane.
//
// Measured on B200 (ncu, 4K teapot frame): active lanes per instruction 12.8 -> 18.9 (node visits 14.5 -> 20.1,
// triangle tests 11.5 -> 19.2), warp instructions -7 %, kernel 1.85 -> 1.78 ms, frame 2.465 -> 2.41 ms; but a 1/8
// band share of the frame (the 8-GPU case) takes 0.537 ms instead of 0.453 ms, and the 27 KB of shared memory per
// CTA leave L1 with 50 % hits instead of 76 %.  Lane utilisation is evidently not what bounds this kernel (L1
// throughput 74 % and the ALU pipe 60 % are), so the register-resident kernel stays the default and this one is
// selected with RTB_FLAG_POOL (or RTB_WF_POOL=1) for A/B runs; both produce the same bits.
// ---------------------------------------------------------------------------
constexpr int POOL = 64;                 // rays per warp
#ifndef POOL_STACK_N
#define POOL_STACK_N 6
#endif
constexpr int POOL_STACK = POOL_STACK_N;   // traversal stack entries per ray in shared memory; deeper ones in HBM
enum : uint32_t { PS_DEAD = 0u, PS_NODE = 1u, PS_LEAF = 2u, PS_SHADE = 3u };
// shared-memory words per warp: field-major [field][POOL]
enum { PF_OX, PF_OY, PF_OZ, PF_DX, PF_DY, PF_DZ, PF_IX, PF_IY, PF_IZ, PF_TBEST, PF_HT,
       PF_HSLOT, PF_HORIG, PF_CUR, PF_SP, PF_PIX, PF_LEVEL, PF_RNGLO, PF_RNGHI, PF_ST, PF_STACK0,
       PF_FIELDS = PF_STACK0 + POOL_STACK };
constexpr int POOL_WORDS = PF_FIELDS * POOL + 32;     // + the deal list

#ifndef POOL_MIN_BLOCKS
#define POOL_MIN_BLOCKS 8
#endif
template <bool STATS>
__global__ void __launch_bounds__(WF_BLOCK, POOL_MIN_BLOCKS)
k_wf_bounce_pool(const SceneDev sc, const ViewDev vw, const PathBuffers pb, const float4* __restrict__ qo,
                 const float4* __restrict__ qd, const uint32_t* __restrict__ n_ptr, uint32_t smp,
                 uint32_t* __restrict__ work_counter, uint32_t* __restrict__ ovf_base, uint32_t ovf_depth,
                 const WfTune tune, TraceCounters* __restrict__ counters) {
    extern __shared__ uint32_t smem_pool[];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t* const W = smem_pool + warp * POOL_WORDS;
    float* const F = reinterpret_cast<float*>(W);
    uint32_t* const list = W + PF_FIELDS * POOL;
    uint32_t* const ovf = ovf_base + (size_t)(blockIdx.x * (WF_BLOCK / 32) + warp) * POOL * ovf_depth;
    const uint32_t n = *n_ptr;
    const uint32_t root_code = root_code_of(sc);
    const unsigned lt = (1u << lane) - 1u;
    bool exhausted = (n == 0u);
    unsigned long long n_node = 0, n_tri = 0, n_rays = 0;
    W[PF_ST * POOL + lane] = PS_DEAD; W[PF_ST * POOL + lane + 32] = PS_DEAD;
    __syncwarp();
    const uint32_t node_min = tune.pool_node_min;

    // (re)start the ray of pooled slot `p`
    auto put_ray = [&](uint32_t p, V3 o, V3 d) {
        F[PF_OX * POOL + p] = o.x; F[PF_OY * POOL + p] = o.y; F[PF_OZ * POOL + p] = o.z;
        F[PF_DX * POOL + p] = d.x; F[PF_DY * POOL + p] = d.y; F[PF_DZ * POOL + p] = d.z;
        F[PF_IX * POOL + p] = fminf(fmaxf(1.0f / d.x, -1e30f), 1e30f);
        F[PF_IY * POOL + p] = fminf(fmaxf(1.0f / d.y, -1e30f), 1e30f);
        F[PF_IZ * POOL + p] = fminf(fmaxf(1.0f / d.z, -1e30f), 1e30f);
        F[PF_TBEST * POOL + p] = FLT_MAX; F[PF_HT * POOL + p] = FLT_MAX;
        W[PF_HSLOT * POOL + p] = 0xffffffffu; W[PF_HORIG * POOL + p] = 0xffffffffu;
        W[PF_CUR * POOL + p] = root_code; W[PF_SP * POOL + p] = 0u;
        W[PF_ST * POOL + p] = PS_NODE;
    };

    for (;;) {
        const uint32_t s0 = W[PF_ST * POOL + lane], s1 = W[PF_ST * POOL + lane + 32];
        const unsigned dead0 = __ballot_sync(FULL, s0 == PS_DEAD), dead1 = __ballot_sync(FULL, s1 == PS_DEAD);
        const unsigned node0 = __ballot_sync(FULL, s0 == PS_NODE), node1 = __ballot_sync(FULL, s1 == PS_NODE);
        const unsigned leaf0 = __ballot_sync(FULL, s0 == PS_LEAF), leaf1 = __ballot_sync(FULL, s1 == PS_LEAF);
        const unsigned shd0 = __ballot_sync(FULL, s0 == PS_SHADE), shd1 = __ballot_sync(FULL, s1 == PS_SHADE);
        const uint32_t n_dead = __popc(dead0) + __popc(dead1), n_nodew = __popc(node0) + __popc(node1);
        const uint32_t n_leafw = __popc(leaf0) + __popc(leaf1), n_shade = __popc(shd0) + __popc(shd1);
        if (n_dead == (uint32_t)POOL && exhausted) break;

        // ---- refill: every dead slot takes the next path of the bounce queue ----
        if (!exhausted && (n_dead >= tune.refill_min || n_dead == (uint32_t)POOL)) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(work_counter, n_dead);
            base = __shfl_sync(FULL, base, 0);
            if (base >= n) { exhausted = true; continue; }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const bool dead = h ? (s1 == PS_DEAD) : (s0 == PS_DEAD);
                const uint32_t idx = base + (h ? __popc(dead0) + __popc(dead1 & lt) : __popc(dead0 & lt));
                if (dead && idx < n) {
                    const uint32_t p = lane + 32u * h;
                    const float4 ro = __ldcs(qo + idx), rd = __ldcs(qd + idx);
                    const uint32_t pix = __float_as_uint(ro.w);
                    const unsigned long long rs = pb.rng_state[pix];
                    W[PF_PIX * POOL + p] = pix; W[PF_LEVEL * POOL + p] = 1u;
                    W[PF_RNGLO * POOL + p] = (uint32_t)rs; W[PF_RNGHI * POOL + p] = (uint32_t)(rs >> 32);
                    put_ray(p, mk(ro.x, ro.y, ro.z), mk(rd.x, rd.y, rd.z));
                    ++n_rays;
                }
            }
            __syncwarp();
            continue;
        }

        // ---- pick the step kind with the most candidates and deal the candidates out to the lanes ----
        uint32_t kind = PS_NODE, m0 = node0, m1 = node1, cnt = n_nodew;
        if (n_leafw > cnt) { kind = PS_LEAF; m0 = leaf0; m1 = leaf1; cnt = n_leafw; }
        if (n_shade > cnt || (cnt == 0u)) { kind = PS_SHADE; m0 = shd0; m1 = shd1; cnt = n_shade; }
        if (cnt == 0u) continue;                      // cannot happen: some slot is alive
        {
            const uint32_t r0 = __popc(m0 & lt), r1 = __popc(m0) + __popc(m1 & lt);
            if (((m0 >> lane) & 1u) && r0 < 32u) list[r0] = lane;
            if (((m1 >> lane) & 1u) && r1 < 32u) list[r1] = lane + 32u;
        }
        __syncwarp();
        const bool have = lane < min(cnt, 32u);
        const uint32_t p = have ? list[lane] : 0u;
        __syncwarp();

        if (kind == PS_NODE) {
            // ---- up to descend_max BVH4 node visits for the dealt rays ----
            float ix = 0.f, iy = 0.f, iz = 0.f, ox = 0.f, oy = 0.f, oz = 0.f, tbest = 0.f;
            uint32_t cur = 0u, sp = 0u;
            bool go = have;
            if (have) {
                ix = F[PF_IX * POOL + p]; iy = F[PF_IY * POOL + p]; iz = F[PF_IZ * POOL + p];
                ox = -F[PF_OX * POOL + p] * ix; oy = -F[PF_OY * POOL + p] * iy; oz = -F[PF_OZ * POOL + p] * iz;
                tbest = F[PF_TBEST * POOL + p];
                cur = W[PF_CUR * POOL + p]; sp = W[PF_SP * POOL + p];
            }
            uint32_t st = PS_NODE;
#pragma unroll 1
            for (uint32_t it = 0; it < tune.descend_max; ++it) {
                // re-deal as soon as too few of the dealt rays still want a node (the others reached a leaf or ended)
                if ((uint32_t)__popc(__ballot_sync(FULL, go)) < (it == 0u ? 1u : node_min)) break;
                if (!go) continue;
                const float4* np = sc.nodes4 + 8u * cur;
                const float4 lx = __ldg(np + 0), hx = __ldg(np + 1), ly = __ldg(np + 2), hy = __ldg(np + 3);
                const float4 lz = __ldg(np + 4), hz = __ldg(np + 5);
                const uint4 cd = __ldg(reinterpret_cast<const uint4*>(np + 6));
                if (STATS) n_node += 4;
                const float INF = __int_as_float(0x7f800000);
                float key[4];
                uint32_t code[4] = {cd.x, cd.y, cd.z, cd.w};
#define RTB_SLAB(c, LX, HX, LY, HY, LZ, HZ)                                                             \
                {                                                                                       \
                    const float x0 = fmaf(LX, ix, ox), x1 = fmaf(HX, ix, ox);                           \
                    const float y0 = fmaf(LY, iy, oy), y1 = fmaf(HY, iy, oy);                           \
                    const float z0 = fmaf(LZ, iz, oz), z1 = fmaf(HZ, iz, oz);                           \
                    const float tn = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f)); \
                    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000005f; \
                    key[c] = ((tn <= tf) && (tn <= tbest)) ? tn : INF;                                  \
                }
                RTB_SLAB(0, lx.x, hx.x, ly.x, hy.x, lz.x, hz.x)
                RTB_SLAB(1, lx.y, hx.y, ly.y, hy.y, lz.y, hz.y)
                RTB_SLAB(2, lx.z, hx.z, ly.z, hy.z, lz.z, hz.z)
                RTB_SLAB(3, lx.w, hx.w, ly.w, hy.w, lz.w, hz.w)
#undef RTB_SLAB
#define RTB_CE(a, b)                                                                          \
                {                                                                             \
                    const bool sw = key[a] > key[b];                                          \
                    const float ka = sw ? key[b] : key[a], kb = sw ? key[a] : key[b];         \
                    const uint32_t ca_ = sw ? code[b] : code[a], cb_ = sw ? code[a] : code[b]; \
                    key[a] = ka; key[b] = kb; code[a] = ca_; code[b] = cb_;                   \
                }
                RTB_CE(0, 1) RTB_CE(2, 3) RTB_CE(0, 2) RTB_CE(1, 3) RTB_CE(1, 2)
#undef RTB_CE
                auto push = [&](uint32_t c) {
                    if (sp < (uint32_t)POOL_STACK) W[(PF_STACK0 + sp) * POOL + p] = c;
                    else ovf[(size_t)p * ovf_depth + (sp - POOL_STACK)] = c;
                    ++sp;
                };
                if (key[0] == INF) {
                    if (sp == 0u) { st = PS_SHADE; go = false; }       // the ray is finished
                    else {
                        --sp;
                        cur = sp < (uint32_t)POOL_STACK ? W[(PF_STACK0 + sp) * POOL + p] : ovf[(size_t)p * ovf_depth + (sp - POOL_STACK)];
                    }
                } else {
                    cur = code[0];
                    if (key[3] < INF) push(code[3]);
                    if (key[2] < INF) push(code[2]);
                    if (key[1] < INF) push(code[1]);
                }
                if (go && (cur & LEAF_FLAG)) { st = PS_LEAF; go = false; }
            }
            if (have) { W[PF_CUR * POOL + p] = cur; W[PF_SP * POOL + p] = sp; W[PF_ST * POOL + p] = st; }
        } else if (kind == PS_LEAF) {
            // ---- the exact tests of one leaf per dealt ray (get_box_min_time_intersection, raytrace.rs:1013-1050) ----
            if (have) {
                const V3 o = mk(F[PF_OX * POOL + p], F[PF_OY * POOL + p], F[PF_OZ * POOL + p]);
                const V3 d = mk(F[PF_DX * POOL + p], F[PF_DY * POOL + p], F[PF_DZ * POOL + p]);
                float ht = F[PF_HT * POOL + p], tbest = F[PF_TBEST * POOL + p];
                int hslot = (int)W[PF_HSLOT * POOL + p];
                uint32_t horig = W[PF_HORIG * POOL + p];
                uint32_t cur = W[PF_CUR * POOL + p], sp = W[PF_SP * POOL + p];
                const uint32_t first = (cur & ~LEAF_FLAG) >> 3, cnt_t = cur & 7u;
                const float4* q = sc.tri + (size_t)RTB_TRI_F4 * first;
                float4 a0 = __ldg(q), a1 = __ldg(q + 1);
                for (uint32_t k = first; k < first + cnt_t; ++k, q += RTB_TRI_F4) {
                    float t;
                    if (STATS) ++n_tri;
                    const float4 c0 = a0, c1 = a1;
                    if (k + 1u < first + cnt_t) { a0 = __ldg(q + RTB_TRI_F4); a1 = __ldg(q + RTB_TRI_F4 + 1); }
                    if (tri_test_pre(q, c0, c1, o, d, hslot >= 0, ht, &t)) {
                        const uint32_t orig = __float_as_uint(c1.w);
                        if (hslot < 0 || t < ht || (t == ht && orig < horig)) {
                            ht = t; hslot = (int)k; horig = orig;
                            if (t < tbest) tbest = t;      // a NaN t never tightens the bound
                        }
                    }
                }
                uint32_t st = PS_SHADE;
                if (sp != 0u) {
                    --sp;
                    cur = sp < (uint32_t)POOL_STACK ? W[(PF_STACK0 + sp) * POOL + p] : ovf[(size_t)p * ovf_depth + (sp - POOL_STACK)];
                    st = (cur & LEAF_FLAG) ? PS_LEAF : PS_NODE;
                }
                F[PF_HT * POOL + p] = ht; F[PF_TBEST * POOL + p] = tbest;
                W[PF_HSLOT * POOL + p] = (uint32_t)hslot; W[PF_HORIG * POOL + p] = horig;
                W[PF_CUR * POOL + p] = cur; W[PF_SP * POOL + p] = sp; W[PF_ST * POOL + p] = st;
            }
        } else {
            // ---- project_ray :1283-1293 for finished rays: bounce again in place or end the path ----
            if (have) {
                const V3 o = mk(F[PF_OX * POOL + p], F[PF_OY * POOL + p], F[PF_OZ * POOL + p]);
                const V3 d = mk(F[PF_DX * POOL + p], F[PF_DY * POOL + p], F[PF_DZ * POOL + p]);
                const float ht = F[PF_HT * POOL + p];
                const int hslot = (int)W[PF_HSLOT * POOL + p];
                const uint32_t pix = W[PF_PIX * POOL + p];
                uint32_t level = W[PF_LEVEL * POOL + p];
                Rng g;
                g.state = (unsigned long long)W[PF_RNGLO * POOL + p] | ((unsigned long long)W[PF_RNGHI * POOL + p] << 32);
                V3 term = mk(0.f, 0.f, 0.f);
                bool ended = true;
                if (hslot < 0) {
                    term = sky_color();
                } else {
                    V3 color, no, nd;
                    float alpha = 0.f;
                    if (shade_hit(sc, hslot, ht, o, d, g, &color, &alpha, &no, &nd) == 0) {
                        term = color;
                    } else {
                        pb.stack[(size_t)level * pb.n_slots + pix] = make_float4(color.x, color.y, color.z, alpha);
                        ++level;
                        if (level < vw.maxdepth) {          // project_ray(depth-1) with depth-1 > 0
                            W[PF_LEVEL * POOL + p] = level;
                            W[PF_RNGLO * POOL + p] = (uint32_t)g.state; W[PF_RNGHI * POOL + p] = (uint32_t)(g.state >> 32);
                            put_ray(p, no, nd);
                            ended = false; ++n_rays;
                        }                                   // else depth 0: black (:1261), not counted
                    }
                }
                if (ended) { finish_path(vw, pb, pix, smp, level, term); W[PF_ST * POOL + p] = PS_DEAD; }
            }
        }
        __syncwarp();
    }
    for (int off = 16; off > 0; off >>= 1) {
        n_rays += __shfl_xor_sync(FULL, n_rays, off);
        if (STATS) {
            n_node += __shfl_xor_sync(FULL, n_node, off);
            n_tri += __shfl_xor_sync(FULL, n_tri, off);
        }
    }
    if (lane == 0) {
        atomicAdd(&counters->rays, n_rays);
        if (STATS) {
            atomicAdd(&counters->node_tests, n_node); atomicAdd(&counters->tri_tests, n_tri);
            atomicAdd(&counters->node_tests_bounce, n_node); atomicAdd(&counters->tri_tests_bounce, n_tri);
        }
    }
}

// per-sample reset of the queue / work counters
struct WfCounters {
    uint32_t n_bounce;        // size of the bounce queue
    uint32_t work_primary;    // work-fetch counters of the two persistent kernels
    uint32_t work_bounce;
    uint32_t pad;
};
__global__ void k_wf_tally(WfCounters* c) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { c->n_bounce = 0u; c->work_primary = 0u; c->work_bounce = 0u; }
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// Experiment knobs, read from the environment once (function-local static: thread-safe, the multi-GPU entry points call
// the launcher from one host thread per GPU).  Defaults are the measured optima quoted next to the constants above.
struct WfEnv {
    uint32_t descend_max, refill_min, refill_min_primary, pool_node_min;
    int smem_stack;          // > 0: force the bounded shared-memory stack with this many entries (+ overflow)
    int ctas, pool_ctas;     // > 0: cap on resident CTAs per SM (persistent grids)
    bool pool;               // RTB_WF_POOL=1: ray-pool bounce kernel for every frame
    static int geti(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
    static const WfEnv& get() {
        static const WfEnv env = [] {
            WfEnv e;
            e.descend_max = (uint32_t)std::max(1, geti("RTB_WF_DESCEND", (int)WF_DESCEND_MAX));
            e.refill_min = (uint32_t)std::min(32, std::max(1, geti("RTB_WF_REFILL", (int)WF_REFILL_MIN)));
            e.refill_min_primary = (uint32_t)std::min(32, std::max(1, geti("RTB_WF_REFILL_P", (int)WF_REFILL_MIN_PRIMARY)));
            e.pool_node_min = (uint32_t)std::max(1, geti("RTB_POOL_NODE_MIN", 1));
            e.smem_stack = geti("RTB_WF_STACK", 0);
            e.ctas = geti("RTB_WF_CTAS", 0);
            e.pool_ctas = geti("RTB_POOL_CTAS", 0);
            e.pool = geti("RTB_WF_POOL", 0) != 0;
            return e;
        }();
        return env;
    }
};

static uint32_t pool_ovf_depth(uint32_t stack4) { return stack4 > (uint32_t)POOL_STACK ? stack4 - POOL_STACK : 1u; }
static constexpr size_t POOL_MAX_WARPS = 148 * 8 * (WF_BLOCK / 32) * 2;   // resident warps of a B200, with slack

static bool pool_selected(uint32_t flags) {
    return (WfEnv::get().pool || (flags & RTB_FLAG_POOL)) && !(flags & RTB_FLAG_BRUTE);
}

size_t rtb_wf_workspace_bytes(uint32_t n_slots, uint32_t maxdepth, bool multisample, uint32_t stack4, uint32_t flags) {
    size_t b = 0;
    if (pool_selected(flags)) b += sizeof(uint32_t) * POOL_MAX_WARPS * POOL * pool_ovf_depth(stack4) + 256;   // deep stack entries
    b += 2 * (sizeof(float4) * (size_t)n_slots + 256);          // bounce ray queue (o, d)
    b += sizeof(float2) * (size_t)n_slots + 256;                // hit records of the primary rays
    b += sizeof(float4) * (size_t)n_slots * maxdepth + 256;     // mix stacks
    b += sizeof(uint64_t) * (size_t)n_slots + 256;              // RNG
    if (multisample) b += sizeof(float4) * (size_t)n_slots + 256;   // sample sums
    b += 512;
    return b;
}

int rtb_launch_wavefront(const SceneDev& sc, const ViewDev& vw, void* workspace, float4* d_rgba, uint32_t* d_prim,
                         float* d_t, TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches,
                         cudaEvent_t* stage_ev, float* stage_ms) {
    const uint32_t tiles_x8 = (vw.width + 7u) / 8u;
    const uint32_t n_slots = vw.my_tile_rows * 2u * tiles_x8 * 32u;
    if (n_slots == 0) return RTB_OK;
    const bool multi = (vw.s_end - vw.s_begin) > 1;
    // carve the workspace
    char* p = (char*)workspace;
    auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    float4* qo1 = (float4*)take(sizeof(float4) * n_slots); float4* qd1 = (float4*)take(sizeof(float4) * n_slots);
    float2* hit = (float2*)take(sizeof(float2) * n_slots);
    PathBuffers pb;
    pb.stack = (float4*)take(sizeof(float4) * (size_t)n_slots * vw.maxdepth);
    pb.rng_state = (uint64_t*)take(sizeof(uint64_t) * n_slots);
    pb.acc = multi ? (float4*)take(sizeof(float4) * n_slots) : nullptr;
    pb.rgba = d_rgba; pb.prim_out = d_prim; pb.t_out = d_t; pb.n_slots = n_slots;
    WfCounters* wc = (WfCounters*)take(sizeof(WfCounters));
    const bool pool = pool_selected(vw.flags);
    uint32_t* pool_ovf = pool ? (uint32_t*)take(sizeof(uint32_t) * POOL_MAX_WARPS * POOL * pool_ovf_depth(sc.stack4)) : nullptr;
    const WfEnv& env = WfEnv::get();

    // persistent grids: as many CTAs as fit, given the shared-memory traversal stacks (3 entries per BVH4 level, 4 B each, per thread)
    const bool stats = (vw.flags & RTB_FLAG_STATS) != 0;
    // the whole worst-case stack in shared memory when 8 CTAs of it fit an SM (<= 24 KB per CTA), else 8 entries + overflow
    // (whole stack = worst case + the sentinel entry + the one entry above the top the branch-free pushes may touch)
    const bool ovf = env.smem_stack > 0 || (size_t)(sc.stack4 + 2u) * WF_BLOCK * sizeof(uint32_t) > 24u * 1024u;
    const int smem_depth = ovf ? std::min<int>((int)sc.stack4, env.smem_stack > 0 ? env.smem_stack : WF_SMEM_STACK) : (int)sc.stack4 + 2;
    const size_t smem = (size_t)smem_depth * WF_BLOCK * sizeof(uint32_t);
    int dev = 0, sms = 0, per_sm_t = 0, per_sm_b = 0;
    RTB_CUDA(cudaGetDevice(&dev));
    RTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // kernel variants: <STATS, OVF>
    typedef void (*TraceFn)(const SceneDev, const ViewDev, uint32_t, uint32_t, float2*, uint32_t*, uint32_t, const WfTune, TraceCounters*);
    typedef void (*BounceFn)(const SceneDev, const ViewDev, const PathBuffers, const float4*, const float4*, const uint32_t*, uint32_t,
                             uint32_t*, uint32_t, const WfTune, TraceCounters*);
    const TraceFn trace_fn = stats ? (ovf ? k_wf_trace<true, true> : k_wf_trace<true, false>)
                                   : (ovf ? k_wf_trace<false, true> : k_wf_trace<false, false>);
    const BounceFn bounce_fn = stats ? (ovf ? k_wf_bounce<true, true> : k_wf_bounce<true, false>)
                                     : (ovf ? k_wf_bounce<false, true> : k_wf_bounce<false, false>);
    if (smem > 48 * 1024) {
        RTB_CUDA(cudaFuncSetAttribute(trace_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RTB_CUDA(cudaFuncSetAttribute(bounce_fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, trace_fn, WF_BLOCK, smem));
    RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, bounce_fn, WF_BLOCK, smem));
    if (env.ctas > 0) { per_sm_t = std::min(per_sm_t, env.ctas); per_sm_b = std::min(per_sm_b, env.ctas); }
    const int grid_t = sms * std::max(per_sm_t, 1), grid_b = sms * std::max(per_sm_b, 1);
    const uint32_t shade_blocks = std::min<uint32_t>((n_slots + 255u) / 256u, 148u * 16u);
    const uint32_t brute = (vw.flags & RTB_FLAG_BRUTE) ? 1u : 0u;
    const size_t smem_pool_bytes = sizeof(uint32_t) * (WF_BLOCK / 32) * POOL_WORDS;
    int grid_p = 0;
    if (pool) {
        int per_sm = 0;
        if (stats) RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_bounce_pool<true>, WF_BLOCK, smem_pool_bytes));
        else RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_bounce_pool<false>, WF_BLOCK, smem_pool_bytes));
        if (env.pool_ctas > 0) per_sm = std::min(per_sm, env.pool_ctas);
        grid_p = sms * std::max(per_sm, 1);
        if ((size_t)grid_p * (WF_BLOCK / 32) > POOL_MAX_WARPS) grid_p = (int)(POOL_MAX_WARPS / (WF_BLOCK / 32));
    }
    const WfTune tune = {env.descend_max, env.refill_min, smem_depth, env.pool_node_min};
    const WfTune tune_p = {env.descend_max, env.refill_min_primary, smem_depth, 1u};
    RTB_CUDA(cudaMemsetAsync(wc, 0, sizeof(WfCounters), stream));
    auto mark = [&](int k) { if (stage_ev) cudaEventRecord(stage_ev[k], stream); };
    for (uint32_t smp = vw.s_begin; smp < vw.s_end; ++smp) {
        mark(0);
        mark(1);
        trace_fn<<<grid_t, WF_BLOCK, smem, stream>>>(sc, vw, smp, n_slots, hit, &wc->work_primary, brute, tune_p, d_counters);
        mark(2);
        k_wf_shade<<<shade_blocks, 256, 0, stream>>>(sc, vw, pb, hit, n_slots, smp, qo1, qd1, &wc->n_bounce);
        mark(3);
        if (launches) *launches += 2;
        if (vw.maxdepth > 1 && pool) {
            if (stats)
                k_wf_bounce_pool<true><<<grid_p, WF_BLOCK, smem_pool_bytes, stream>>>(sc, vw, pb, qo1, qd1, &wc->n_bounce, smp, &wc->work_bounce, pool_ovf, pool_ovf_depth(sc.stack4), tune, d_counters);
            else
                k_wf_bounce_pool<false><<<grid_p, WF_BLOCK, smem_pool_bytes, stream>>>(sc, vw, pb, qo1, qd1, &wc->n_bounce, smp, &wc->work_bounce, pool_ovf, pool_ovf_depth(sc.stack4), tune, d_counters);
            if (launches) ++*launches;
        } else if (vw.maxdepth > 1) {
            bounce_fn<<<grid_b, WF_BLOCK, smem, stream>>>(sc, vw, pb, qo1, qd1, &wc->n_bounce, smp, &wc->work_bounce, brute, tune, d_counters);
            if (launches) ++*launches;
        }
        mark(4);
        if (stage_ev) {
            RTB_CUDA(cudaEventSynchronize(stage_ev[4]));
            for (int k = 0; k < RTB_N_STAGES; ++k) {
                float ms = 0.f;
                RTB_CUDA(cudaEventElapsedTime(&ms, stage_ev[k], stage_ev[k + 1]));
                stage_ms[k] += ms;
            }
        }
        if (smp + 1 < vw.s_end) {
            k_wf_tally<<<1, 32, 0, stream>>>(wc);
            if (launches) ++*launches;
        }
    }
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
