// rtb_wavefront.cu — the hot path as a wavefront pipeline (the default renderer).
//
// The one-kernel version (rtb_trace.cu) keeps only ~9 of 32 lanes busy per issued instruction
// (ncu, profiles/r1_v1_*): rays of one warp leave the scene at different times and the surviving
// Matte/mirror paths bounce up to maxdepth times.  Here the same arithmetic is split into stages that
// each run with full warps:
//
//   k_wf_raygen   one thread per pixel slot: Viewport::pixel_ray (raytrace.rs:1374-1394) -> ray queue 0
//   k_wf_trace    PERSISTENT kernel: warps pull rays from the level's queue, every lane that finishes its
//                 ray is refilled from the queue (ballot + one atomic per 128 rays per warp), so the
//                 traversal loop runs with (almost) all 32 lanes; closest hit -> hit record
//   k_wf_shade    one thread per ray: Triangle::intersects classification, color_ray (:1199-1254):
//                 terminal paths fold their mix_color stack innermost-first and add the sample to the
//                 pixel; bouncing paths push (colour, alpha) and append the next ray to the next queue
//                 (warp-aggregated atomic = stream compaction)
//   k_wf_tally    per-sample counter roll-up / reset
//
// Per-pixel state lives in HBM between stages (ray 32 B, hit 8 B, mix stack 16 B/level, RNG 8 B); at 4K
// that is ~0.7 GB of traffic per frame, i.e. ~0.1 ms at the measured 6.5 TB/s.
#include <algorithm>
#include <cstdlib>

#include "rtb_device.cuh"

using namespace rtbdev;

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t INVALID_SLOT = 0xffffffffu;
constexpr uint32_t LEAF_FLAG = 0x80000000u;
constexpr uint32_t WF_CHUNK = 128;      // rays reserved per warp per global atomic
constexpr int WF_BLOCK = 128;           // threads per CTA of the persistent trace kernel
// Tunables (env RTB_WF_DESCEND / RTB_WF_REFILL override for experiments).  Measured on B200, 4K teapot frame:
// (descend, refill) = (2,4) 3.53 ms, (4,8) 3.15 ms, (8,16) 3.06 ms, (unbounded,16) 3.30 ms.  Retiring surplus warps
// on small queues made things worse in proportion (the deep levels are latency-, not issue-bound).
constexpr uint32_t WF_DESCEND_MAX = 8;  // internal-node steps per lane per round before leaves are processed
constexpr uint32_t WF_REFILL_MIN = 16;   // refill idle lanes once at least this many are idle

// slot -> pixel.  Slots enumerate 8x4 warp tiles inside the 8-row bands this rank owns.
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
// stage 0: primary rays
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_wf_raygen(const ViewDev vw, uint32_t smp, uint32_t n_slots,
                                                   float4* __restrict__ qo, float4* __restrict__ qd,
                                                   uint64_t* __restrict__ rng_state) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_slots) return;
    const Pixel px = slot_to_pixel(vw, slot);
    if (!px.inside) {
        qo[slot] = make_float4(0.f, 0.f, 0.f, __uint_as_float(INVALID_SLOT));
        return;
    }
    Rng g;
    g.seed(vw.seed, (uint64_t)px.row * vw.width + px.col, smp);
    float u_off = 0.5f, v_off = 0.5f;
    if (vw.spp != 1) { u_off = g.next_f32(); v_off = g.next_f32(); }   // :1382-1386
    V3 o, d;
    gen_primary(vw, px.row, px.col, u_off, v_off, &o, &d);
    qo[slot] = make_float4(o.x, o.y, o.z, __uint_as_float(slot));
    qd[slot] = make_float4(d.x, d.y, d.z, 0.f);
    rng_state[slot] = g.state;
}

// ---------------------------------------------------------------------------
// stage 1: closest hit, persistent threads with per-lane refill
// ---------------------------------------------------------------------------
struct TravState {
    V3 o, d;
    float ix, iy, iz, ox, oy, oz;
    float tbest;
    Hit h;
    uint32_t cur;      // LEAF_FLAG | first<<3 | count, or the left node index of an internal sibling pair
    int sp;
};

template <bool STATS>
__global__ void __launch_bounds__(WF_BLOCK, 4)
k_wf_trace(const SceneDev sc, const float4* __restrict__ qo, const float4* __restrict__ qd,
           const uint32_t* __restrict__ n_ptr, uint32_t n_const, float2* __restrict__ hit_out,
           uint32_t* __restrict__ work_counter, uint32_t brute, uint32_t descend_max, uint32_t refill_min,
           TraceCounters* __restrict__ counters) {
    const uint32_t n = n_ptr ? *n_ptr : n_const;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    uint32_t chunk_next = 0, chunk_end = 0;       // warp-uniform
    bool exhausted = (n == 0u);

    bool active = false;
    uint32_t ray_id = 0;
    TravState s;
    s.sp = 0; s.cur = 0; s.tbest = FLT_MAX; s.h.t = FLT_MAX; s.h.slot = -1; s.h.orig = 0xffffffffu;
    // Traversal stack in SHARED memory, one 4-byte code per entry, laid out [depth][thread] so a warp never
    // bank-conflicts.  As thread-local arrays (first version) the stacks went through L1, took ~40% of it and
    // pushed the BVH nodes out: L1 hit rate 81%, 53-59% of stall samples on long_scoreboard for bounce rays.
    extern __shared__ uint32_t smem_stack[];
    uint32_t* const stack = smem_stack + threadIdx.x;     // entry k lives at stack[k * WF_BLOCK]
    unsigned long long n_node = 0, n_tri = 0;

    // root description (uniform)
    const float4 r0 = __ldg(sc.nodes + 0), r1 = __ldg(sc.nodes + 1);
    uint32_t root_code;
    if (sc.n_prims == 0u) root_code = LEAF_FLAG;                                     // leaf with 0 primitives
    else if (brute) root_code = 0xfffffffeu;                                        // handled separately
    else if (__float_as_uint(r1.w) != 0u) root_code = LEAF_FLAG | (__float_as_uint(r0.w) << 3) | __float_as_uint(r1.w);
    else root_code = __float_as_uint(r0.w);

    auto finish = [&]() {
        __stcs(hit_out + ray_id, make_float2(s.h.t, __int_as_float(s.h.slot)));
        active = false;
    };
    // next work item from the stack (skipping entries the current bound already rules out) or finish the ray
    auto pop = [&]() {
        if (s.sp == 0) { finish(); return; }
        --s.sp;
        s.cur = stack[s.sp * WF_BLOCK];
    };

    for (;;) {
        // ---- refill idle lanes from the queue ----
        const unsigned need = __ballot_sync(FULL, !active);
        if (!exhausted && ((uint32_t)__popc(need) >= refill_min)) {
            if (chunk_next >= chunk_end) {
                uint32_t base = 0;
                if (lane == 0) base = atomicAdd(work_counter, WF_CHUNK);
                base = __shfl_sync(FULL, base, 0);
                if (base >= n) { exhausted = true; }
                else { chunk_next = base; chunk_end = min(base + WF_CHUNK, n); }
            }
            if (!exhausted) {
                const uint32_t idx = chunk_next + __popc(need & lt_mask);
                if (!active && idx < chunk_end) {
                    const float4 ro = __ldcs(qo + idx);      // streamed once: keep it out of L1
                    if (__float_as_uint(ro.w) != INVALID_SLOT) {
                        const float4 rd = __ldcs(qd + idx);
                        ray_id = idx;
                        s.o = mk(ro.x, ro.y, ro.z);
                        s.d = mk(rd.x, rd.y, rd.z);
                        s.ix = fminf(fmaxf(1.0f / s.d.x, -1e30f), 1e30f);
                        s.iy = fminf(fmaxf(1.0f / s.d.y, -1e30f), 1e30f);
                        s.iz = fminf(fmaxf(1.0f / s.d.z, -1e30f), 1e30f);
                        s.ox = -s.o.x * s.ix; s.oy = -s.o.y * s.iy; s.oz = -s.o.z * s.iz;
                        s.tbest = FLT_MAX;
                        s.h.t = FLT_MAX; s.h.slot = -1; s.h.orig = 0xffffffffu;
                        s.sp = 0;
                        s.cur = root_code;
                        active = true;
                    }
                }
                chunk_next = min(chunk_next + (uint32_t)__popc(need), chunk_end);
            }
        }
        if (__ballot_sync(FULL, active) == 0u) {
            if (exhausted) break;
            continue;
        }

        if (brute) {   // validation mode: linear scan, no BVH
            if (active) {
                for (uint32_t k = 0; k < sc.n_prims; ++k) {
                    float t;
                    if (STATS) ++n_tri;
                    const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
                    if (tri_test(q, s.o, s.d, s.h.slot >= 0, s.h.t, &t)) {
                        const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
                        if (s.h.slot < 0 || t < s.h.t || (t == s.h.t && orig < s.h.orig)) { s.h.t = t; s.h.slot = (int)k; s.h.orig = orig; }
                    }
                }
                finish();
            }
            continue;
        }

        // ---- traversal rounds until enough lanes are idle again ----
        for (;;) {
            // phase 1: at most `descend_max` internal-node steps per round.  Unbounded descent (classic
            // while-while) left lanes that reached a leaf waiting for the slowest lane of the warp: 9-16 of 32
            // lanes active on bounce rays (ncu, profiles/r1_v2_*).
#pragma unroll 1
            for (uint32_t it = 0; it < descend_max; ++it) {
                const bool go = active && !(s.cur & LEAF_FLAG);
                if (!__any_sync(FULL, go)) break;
                if (!go) continue;
                const float4* np = sc.nodes + 2u * s.cur;
                const float4 a0 = __ldg(np + 0), a1 = __ldg(np + 1), b0 = __ldg(np + 2), b1 = __ldg(np + 3);
                if (STATS) n_node += 2;
                float ta, tb;
                bool hit_a, hit_b;
                {
                    const float x0 = fmaf(a0.x, s.ix, s.ox), x1 = fmaf(a1.x, s.ix, s.ox);
                    const float y0 = fmaf(a0.y, s.iy, s.oy), y1 = fmaf(a1.y, s.iy, s.oy);
                    const float z0 = fmaf(a0.z, s.iz, s.oz), z1 = fmaf(a1.z, s.iz, s.oz);
                    ta = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
                    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000005f;
                    hit_a = (ta <= tf) && (ta <= s.tbest);
                }
                {
                    const float x0 = fmaf(b0.x, s.ix, s.ox), x1 = fmaf(b1.x, s.ix, s.ox);
                    const float y0 = fmaf(b0.y, s.iy, s.oy), y1 = fmaf(b1.y, s.iy, s.oy);
                    const float z0 = fmaf(b0.z, s.iz, s.oz), z1 = fmaf(b1.z, s.iz, s.oz);
                    tb = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), 0.0f));
                    const float tf = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)) * 1.0000005f;
                    hit_b = (tb <= tf) && (tb <= s.tbest);
                }
                // child code: internal -> its left-child index; leaf -> LEAF_FLAG | first<<3 | count
                const uint32_t ca_cnt = __float_as_uint(a1.w), cb_cnt = __float_as_uint(b1.w);
                const uint32_t ca = ca_cnt ? (LEAF_FLAG | (__float_as_uint(a0.w) << 3) | ca_cnt) : __float_as_uint(a0.w);
                const uint32_t cb = cb_cnt ? (LEAF_FLAG | (__float_as_uint(b0.w) << 3) | cb_cnt) : __float_as_uint(b0.w);
                if (hit_a && hit_b) {
                    const bool a_first = ta <= tb;
                    stack[s.sp * WF_BLOCK] = a_first ? cb : ca;
                    ++s.sp;
                    s.cur = a_first ? ca : cb;
                } else if (hit_a) {
                    s.cur = ca;
                } else if (hit_b) {
                    s.cur = cb;
                } else {
                    pop();
                }
            }
            // phase 2: the leaf this lane holds (if it reached one)
            if (active && (s.cur & LEAF_FLAG)) {
                const uint32_t first = (s.cur & ~LEAF_FLAG) >> 3, cnt = s.cur & 7u;
                for (uint32_t k = first; k < first + cnt; ++k) {
                    float t;
                    if (STATS) ++n_tri;
                    const float4* q = sc.tri + (size_t)RTB_TRI_F4 * k;
                    if (tri_test(q, s.o, s.d, s.h.slot >= 0, s.h.t, &t)) {
                        const uint32_t orig = __float_as_uint(__ldg(q + 1).w);
                        if (s.h.slot < 0 || t < s.h.t || (t == s.h.t && orig < s.h.orig)) {
                            s.h.t = t; s.h.slot = (int)k; s.h.orig = orig;
                            if (t < s.tbest) s.tbest = t;      // a NaN t never tightens the bound
                        }
                    }
                }
                pop();
            }
            const unsigned act = __ballot_sync(FULL, active);
            if (act == 0u) break;
            if (!exhausted && (32u - __popc(act)) >= refill_min) break;
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
// stage 2: shade one queue level
// ---------------------------------------------------------------------------
struct ShadeArgs {
    const float4* qo_in; const float4* qd_in; const float2* hit;
    const uint32_t* n_in_ptr; uint32_t n_in_const;
    float4* qo_out; float4* qd_out; uint32_t* n_out;
    float4* stack;            // [maxdepth][n_slots]
    uint64_t* rng_state;      // [n_slots]
    float4* acc;              // [n_slots] (multi-sample only)
    float4* rgba; uint32_t* prim_out; float* t_out;
    uint32_t n_slots, level, smp;
};

__global__ void __launch_bounds__(256) k_wf_shade(const SceneDev sc, const ViewDev vw, const ShadeArgs a) {
    const uint32_t n = a.n_in_ptr ? *a.n_in_ptr : a.n_in_const;
    const unsigned lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n + 31u) & ~31u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        bool emit = false;
        V3 no = mk(0.f, 0.f, 0.f), nd = mk(0.f, 0.f, 0.f);
        uint32_t slot = INVALID_SLOT;
        if (i < n) {
            const float4 ro = a.qo_in[i];
            slot = __float_as_uint(ro.w);
        }
        if (slot != INVALID_SLOT) {
            const float4 ro = a.qo_in[i], rd = a.qd_in[i];
            const V3 o = mk(ro.x, ro.y, ro.z), d = mk(rd.x, rd.y, rd.z);
            const float2 hr = a.hit[i];
            const int prim_slot = __float_as_int(hr.y);
            const float t = hr.x;
            V3 term;
            uint32_t levels = a.level;     // entries on this path's mix stack
            if (a.level == 0u && a.smp == 0u && (a.prim_out || a.t_out)) {
                const Pixel px = slot_to_pixel(vw, slot);
                if (a.prim_out) a.prim_out[px.out_idx] = prim_slot >= 0 ? __float_as_uint(__ldg(sc.tri + (size_t)RTB_TRI_F4 * prim_slot + 1).w) : 0u;
                if (a.t_out) a.t_out[px.out_idx] = prim_slot >= 0 ? t : 0.0f;
            }
            if (prim_slot < 0) {
                term = sky_color();                                            // project_ray miss :1284
            } else {
                Rng g;
                g.state = a.rng_state[slot];
                V3 color;
                float alpha = 0.f;
                if (shade_hit(sc, prim_slot, t, o, d, g, &color, &alpha, &no, &nd) == 0) {
                    term = color;
                } else {
                    a.stack[(size_t)a.level * a.n_slots + slot] = make_float4(color.x, color.y, color.z, alpha);
                    levels = a.level + 1u;
                    if (a.level + 1u < vw.maxdepth) { emit = true; a.rng_state[slot] = g.state; }
                    else term = mk(0.f, 0.f, 0.f);                             // project_ray(depth 0): black :1261
                }
            }
            if (!emit) {
                // unwind the recursion innermost-first, then add the sample to the pixel (:1422-1426)
                V3 c = term;
                for (int k = (int)levels - 1; k >= 0; --k) {
                    const float4 e = a.stack[(size_t)k * a.n_slots + slot];
                    c = mix_color(mk(e.x, e.y, e.z), c, e.w);
                }
                V3 sum = mk(0.f, 0.f, 0.f);
                if (a.smp != vw.s_begin) { const float4 p = a.acc[slot]; sum = mk(p.x, p.y, p.z); }
                sum = vadd(sum, c);
                if (a.smp + 1u == vw.s_end) {
                    if (!(vw.flags & RTB_FLAG_SUM_ONLY)) sum = vmul(sum, __fdiv_rn(1.0f, (float)vw.spp));
                    const Pixel px = slot_to_pixel(vw, slot);
                    a.rgba[px.out_idx] = make_float4(sum.x, sum.y, sum.z, 0.f);
                } else {
                    a.acc[slot] = make_float4(sum.x, sum.y, sum.z, 0.f);
                }
            }
        }
        // stream compaction of the surviving paths into the next queue
        const unsigned m = __ballot_sync(FULL, emit);
        if (m) {
            uint32_t base = 0;
            const int leader = __ffs(m) - 1;
            if ((int)lane == leader) base = atomicAdd(a.n_out, (uint32_t)__popc(m));
            base = __shfl_sync(FULL, base, leader);
            if (emit) {
                const uint32_t j = base + __popc(m & ((1u << lane) - 1u));
                a.qo_out[j] = make_float4(no.x, no.y, no.z, __uint_as_float(slot));
                a.qd_out[j] = make_float4(nd.x, nd.y, nd.z, 0.f);
            }
        }
    }
}

// per-sample roll-up: total bounce rays += sum n[1..], then reset the per-level counters
struct WfCounters {
    uint32_t n[RTB_MAX_DEPTH + 1];
    uint32_t work[RTB_MAX_DEPTH + 1];
};
__global__ void k_wf_tally(WfCounters* c, TraceCounters* totals) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned long long s = 0;
        for (int k = 1; k <= RTB_MAX_DEPTH; ++k) s += c->n[k];
        atomicAdd(&totals->rays, s);   // several lanes (streams) of one GPU share the totals
        for (int k = 0; k <= RTB_MAX_DEPTH; ++k) { c->n[k] = 0u; c->work[k] = 0u; }
    }
}

}  // namespace

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
size_t rtb_wf_workspace_bytes(uint32_t n_slots, uint32_t maxdepth, bool multisample) {
    size_t b = 0;
    b += 4 * sizeof(float4) * (size_t)n_slots;                  // two ray queues (o, d)
    b += sizeof(float2) * (size_t)n_slots;                      // hit records
    b += sizeof(float4) * (size_t)n_slots * maxdepth;           // mix stacks
    b += sizeof(uint64_t) * (size_t)n_slots;                    // RNG
    if (multisample) b += sizeof(float4) * (size_t)n_slots;     // sample sums
    b += 256 + sizeof(WfCounters);
    return b;
}

int rtb_launch_wavefront(const SceneDev& sc, const ViewDev& vw, void* workspace, float4* d_rgba, uint32_t* d_prim,
                         float* d_t, TraceCounters* d_counters, cudaStream_t stream, uint32_t* launches) {
    const uint32_t tiles_x8 = (vw.width + 7u) / 8u;
    const uint32_t n_slots = vw.my_tile_rows * 2u * tiles_x8 * 32u;
    if (n_slots == 0) return RTB_OK;
    const bool multi = (vw.s_end - vw.s_begin) > 1;
    // carve the workspace
    char* p = (char*)workspace;
    auto take = [&](size_t bytes) { void* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    float4* qo[2]; float4* qd[2];
    qo[0] = (float4*)take(sizeof(float4) * n_slots); qd[0] = (float4*)take(sizeof(float4) * n_slots);
    qo[1] = (float4*)take(sizeof(float4) * n_slots); qd[1] = (float4*)take(sizeof(float4) * n_slots);
    float2* hit = (float2*)take(sizeof(float2) * n_slots);
    float4* stack = (float4*)take(sizeof(float4) * (size_t)n_slots * vw.maxdepth);
    uint64_t* rng = (uint64_t*)take(sizeof(uint64_t) * n_slots);
    float4* acc = multi ? (float4*)take(sizeof(float4) * n_slots) : nullptr;
    WfCounters* wc = (WfCounters*)take(sizeof(WfCounters));

    // persistent grid: as many CTAs as fit, given the shared-memory traversal stacks ((height+2) x 4 B per thread)
    const bool stats = (vw.flags & RTB_FLAG_STATS) != 0;
    const size_t smem = (size_t)(sc.height + 2u) * WF_BLOCK * sizeof(uint32_t);
    int dev = 0, sms = 0, per_sm = 0;
    RTB_CUDA(cudaGetDevice(&dev));
    RTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (smem > 48 * 1024) {
        RTB_CUDA(cudaFuncSetAttribute(k_wf_trace<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        RTB_CUDA(cudaFuncSetAttribute(k_wf_trace<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (stats) RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_trace<true>, WF_BLOCK, smem));
    else RTB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_wf_trace<false>, WF_BLOCK, smem));
    const int tb = sms * (per_sm > 0 ? per_sm : 1);
    const uint32_t shade_blocks = std::min<uint32_t>((n_slots + 255u) / 256u, 148u * 16u);
    const uint32_t brute = (vw.flags & RTB_FLAG_BRUTE) ? 1u : 0u;
    static uint32_t descend_max = 0, refill_min = 0;
    if (descend_max == 0) {
        const char* e1 = getenv("RTB_WF_DESCEND");
        const char* e2 = getenv("RTB_WF_REFILL");
        descend_max = e1 ? (uint32_t)std::max(1, atoi(e1)) : WF_DESCEND_MAX;
        refill_min = e2 ? (uint32_t)std::min(32, std::max(1, atoi(e2))) : WF_REFILL_MIN;
    }

    RTB_CUDA(cudaMemsetAsync(wc, 0, sizeof(WfCounters), stream));
    for (uint32_t smp = vw.s_begin; smp < vw.s_end; ++smp) {
        k_wf_raygen<<<(n_slots + 255u) / 256u, 256, 0, stream>>>(vw, smp, n_slots, qo[0], qd[0], rng);
        if (launches) ++*launches;
        for (uint32_t level = 0; level < vw.maxdepth; ++level) {
            const int in = level & 1u, out = in ^ 1;
            const uint32_t* n_ptr = level ? &wc->n[level] : nullptr;
            if (stats)
                k_wf_trace<true><<<tb, WF_BLOCK, smem, stream>>>(sc, qo[in], qd[in], n_ptr, n_slots, hit, &wc->work[level], brute, descend_max, refill_min, d_counters);
            else
                k_wf_trace<false><<<tb, WF_BLOCK, smem, stream>>>(sc, qo[in], qd[in], n_ptr, n_slots, hit, &wc->work[level], brute, descend_max, refill_min, d_counters);
            ShadeArgs a;
            a.qo_in = qo[in]; a.qd_in = qd[in]; a.hit = hit; a.n_in_ptr = n_ptr; a.n_in_const = n_slots;
            a.qo_out = qo[out]; a.qd_out = qd[out]; a.n_out = &wc->n[level + 1];
            a.stack = stack; a.rng_state = rng; a.acc = acc; a.rgba = d_rgba; a.prim_out = d_prim; a.t_out = d_t;
            a.n_slots = n_slots; a.level = level; a.smp = smp;
            k_wf_shade<<<shade_blocks, 256, 0, stream>>>(sc, vw, a);
            if (launches) *launches += 2;
        }
        k_wf_tally<<<1, 32, 0, stream>>>(wc, d_counters);
        if (launches) ++*launches;
    }
    RTB_CUDA(cudaGetLastError());
    return RTB_OK;
}
