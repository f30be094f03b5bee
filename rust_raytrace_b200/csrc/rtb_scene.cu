// rtb_scene.cu — scene assembly on the GPU: the step in front of the BVH build.
//
//   k_assemble     obj_parser::parse_obj's per-face work (obj_parser.rs:47-73): scale, change of basis, offset, then
//                  make_triangle (raytrace.rs:340-383) — one thread per (instance, face).  A 1 M-triangle scene of
//                  instanced meshes is uploaded as one mesh + one small record per instance instead of 138 MB of
//                  finished `Triangle`s, and the 35-field records are produced at HBM speed.
//   k_cull_flags   box_contains_polygon (raytrace.rs:753-779) against the octree root cube: triangles the reference's
//                  octree never holds (build_bounding_box :795-805) are invisible there and are culled here.
//   k_compact_keep ordered compaction of the survivors (the builder's `keep` array), after an exclusive scan.
//
// Arithmetic contract as in rtb_device.cuh: every operation the reference performs is an explicit round-to-nearest
// intrinsic in the reference's order, so the records are bit-identical to the ones raytrace_lib (and the host
// mirror csrc/host/raytrace_host.cpp, and the oracle) produce; tests/test_gpu_parity.py compares them bit for bit.
#include <cfloat>

#include "rtb_device.cuh"
#include "rtb_sort.cuh"

using namespace rtbdev;

namespace {

__device__ __forceinline__ float fm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float comp(V3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }
__device__ __forceinline__ V3 ld3(const float* p) { return mk(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
__device__ __forceinline__ float vlen(V3 a) { return __fsqrt_rn(vdot(a, a)); }
// Vec3::cross (raytrace.rs:80-90): (yzx * o.zxy) - (zxy * o.yzx)
__device__ __forceinline__ V3 vcross(V3 a, V3 b) {
    return mk(fs(fm(a.y, b.z), fm(a.z, b.y)), fs(fm(a.z, b.x), fm(a.x, b.z)), fs(fm(a.x, b.y), fm(a.y, b.x)));
}

// ---- the two-ray meeting point make_triangle uses for its "incenter" (Ray::intersect, raytrace.rs:212-267) ----
struct Ray2 { V3 orig, dir; };
__device__ __forceinline__ Ray2 mk_ray(V3 o, V3 d) { Ray2 r; r.orig = o; r.dir = vunit(d); return r; }
__device__ __forceinline__ V3 ray_at(const Ray2& r, float t) { return vadd(vmul(r.dir, t), r.orig); }

// ray parameters of the crossing in the plane of components (i, j); false when the projections are near-parallel
__device__ bool solve_2d(const Ray2& s, const Ray2& r, int i, int j, float* ts, float* tr) {
    const float rdi = comp(r.dir, i), rdj = comp(r.dir, j), sdi = comp(s.dir, i), sdj = comp(s.dir, j);
    const float det = fs(fm(rdi, sdj), fm(rdj, sdi));
    if (fabsf(det) < 0.0001f) return false;
    const float dx = fs(comp(r.orig, i), comp(s.orig, i));
    const float dy = fs(comp(r.orig, j), comp(s.orig, j));
    *ts = fd(fs(fm(dy, rdi), fm(dx, rdj)), det);
    *tr = fd(fs(fm(dy, sdi), fm(dx, sdj)), det);
    return true;
}
__device__ bool rays_meet(const Ray2& s, const Ray2& r, V3* where) {
    float ts = 0.f, tr = 0.f;
    if (!solve_2d(s, r, 0, 1, &ts, &tr) && !solve_2d(s, r, 0, 2, &ts, &tr) && !solve_2d(s, r, 1, 2, &ts, &tr))
        return false;
    const V3 ps = ray_at(s, ts), pr = ray_at(r, tr);
    const V3 d = vsub(pr, ps);
    if (vdot(d, d) < 0.01f) { *where = ps; return true; }
    return false;
}

struct SurfaceDev { uint32_t kind; float color[3]; float alpha, scattering; };

// make_triangle (raytrace.rs:340-383).  false where the reference panics (the medians do not meet, :357).
__device__ bool make_triangle_dev(const V3 pts[3], const SurfaceDev& sf, float edge_thickness, RtbTriangle* out) {
    const V3 ab = vsub(pts[1], pts[0]), ac = vsub(pts[2], pts[0]), bc = vsub(pts[2], pts[1]);
    const Ray2 from_a = mk_ray(pts[0], vadd(ac, ab));
    const Ray2 from_b = mk_ray(pts[1], vadd(bc, vmul(ab, -1.0f)));
    V3 centre;
    if (!rays_meet(from_a, from_b, &centre)) return false;
    RtbTriangle t;
    V3 side_dirs[3];
#pragma unroll
    for (int e = 0; e < 3; ++e) {
        const V3 edge = vsub(pts[(e + 1) % 3], pts[e]);
        const V3 to_centre = vsub(centre, pts[e]);
        const V3 foot = vmul(edge, fd(vdot(edge, to_centre), vdot(edge, edge)));
        const V3 out_vec = vsub(foot, to_centre);
        side_dirs[e] = vunit(out_vec);
        t.side_lens[e] = vlen(out_vec);
        st3(&t.sides[3 * e], side_dirs[e]);
        st3(&t.corners[3 * e], pts[e]);
    }
    st3(t.norm, vunit(vcross(side_dirs[0], side_dirs[1])));
    st3(t.incenter, centre);
    float r2 = 0.0f;
#pragma unroll
    for (int e = 0; e < 3; ++e) { const V3 d = vsub(pts[e], centre); r2 = fmaxf(r2, vdot(d, d)); }
    t.bounding_r2 = r2;
    t.edge_thickness = edge_thickness;
    t.kind = sf.kind;
    t.color[0] = sf.color[0]; t.color[1] = sf.color[1]; t.color[2] = sf.color[2];
    t.alpha = sf.alpha;
    t.scattering = sf.scattering;
    *out = t;
    return true;
}

// One thread per (instance, face): triangle 1 + inst*nfaces + f of the scene array (0 is the dummy).
__global__ void k_assemble(const float* __restrict__ verts, uint32_t nverts, const uint32_t* __restrict__ faces,
                           uint32_t nfaces, const RtbMeshInstance* __restrict__ inst, uint32_t n_inst,
                           RtbTriangle* __restrict__ out, uint32_t* __restrict__ first_bad) {
    const uint64_t gid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (uint64_t)nfaces * n_inst) return;
    const uint32_t in = (uint32_t)(gid / nfaces), f = (uint32_t)(gid - (uint64_t)in * nfaces);
    const RtbMeshInstance& I = inst[in];
    const V3 r0 = ld3(I.transform_rows), r1 = ld3(I.transform_rows + 3), r2 = ld3(I.transform_rows + 6);
    const V3 off = ld3(I.offset);
    V3 pts[3];
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const uint32_t vi = faces[3 * f + k];                       // 1-based, obj_parser.rs:64-66
        if (vi < 1u || vi > nverts) { ok = false; pts[k] = mk(0.f, 0.f, 0.f); continue; }
        const V3 raw = vmul(ld3(verts + 3 * (size_t)(vi - 1u)), I.scale);
        pts[k] = vadd(mk(vdot(r0, raw), vdot(r1, raw), vdot(r2, raw)), off);   // change_basis, raytrace.rs:117-121
    }
    SurfaceDev sf;
    sf.kind = I.kind; sf.color[0] = I.color[0]; sf.color[1] = I.color[1]; sf.color[2] = I.color[2];
    sf.alpha = I.alpha; sf.scattering = I.scattering;
    if (ok) ok = make_triangle_dev(pts, sf, I.edge_thickness, out + gid);
    if (!ok) atomicMin(first_bad, (uint32_t)gid);
}

// ---- octree-root membership (raytrace.rs:636-779) ---------------------------------------------------
__device__ __forceinline__ bool point_in_cube(V3 c, float h, V3 p) {
    const V3 d = vsub(p, c);
    return fabsf(d.x) < h && fabsf(d.y) < h && fabsf(d.z) < h;
}
struct Line { V3 orig, dir, inv; };
__device__ __forceinline__ Line mk_line(V3 o, V3 d) {
    Line l;
    l.orig = o; l.dir = vunit(d);
    l.inv = mk(fd(1.0f, l.dir.x), fd(1.0f, l.dir.y), fd(1.0f, l.dir.z));
    return l;
}
// face_contains_triangle (raytrace.rs:645-729): the line in which the plane of one cube face meets the plane of the
// triangle must cross both the face and the triangle.
__device__ bool face_cuts_triangle(V3 c, V3 n1, float h, const RtbTriangle& t) {
    const V3 n2 = ld3(t.norm), tri_c = ld3(t.incenter);
    const float h1 = vdot(n1, vadd(c, vmul(n1, h)));
    const float h2 = vdot(n2, tri_c);
    const float nn = vdot(n1, n2);
    const float den = fs(1.0f, fm(nn, nn));
    const float k1 = fd(fs(h1, fm(h2, nn)), den);
    const float k2 = fd(fs(h2, fm(h1, nn)), den);
    const Line first = mk_line(vadd(vmul(n1, k1), vmul(n2, k2)), vcross(n1, n2));
    float tmin = FLT_MAX;
    for (int ax = 0; ax < 3; ++ax) {
        if (comp(n1, ax) != 0.0f) continue;
        const float a = fm(fs(fs(comp(c, ax), h), comp(first.orig, ax)), comp(first.inv, ax));
        const float b = fm(fs(fa(comp(c, ax), h), comp(first.orig, ax)), comp(first.inv, ax));
        tmin = fminf(tmin, fminf(a, b));
    }
    const Line line = (tmin > 0.0f) ? first : mk_line(vadd(vmul(first.dir, fm(tmin, 2.0f)), first.orig), first.dir);
    tmin = -FLT_MAX;
    float tmax = FLT_MAX;
    for (int ax = 0; ax < 3; ++ax) {
        if (comp(n1, ax) != 0.0f) continue;
        const float a = fm(fs(fs(comp(c, ax), h), comp(line.orig, ax)), comp(line.inv, ax));
        const float b = fm(fs(fa(comp(c, ax), h), comp(line.orig, ax)), comp(line.inv, ax));
        tmin = fmaxf(tmin, fminf(a, b));
        tmax = fminf(tmax, fmaxf(a, b));
    }
    if (tmax < tmin) return false;
    V3 off[3];
    for (int k = 0; k < 3; ++k) {
        const V3 corner = ld3(&t.corners[3 * k]);
        const float s = fd(vdot(vsub(corner, line.orig), line.dir), vdot(line.dir, line.dir));
        off[k] = vsub(vadd(vmul(line.dir, s), line.orig), corner);
    }
    return vdot(off[0], off[1]) < 0.0f || vdot(off[0], off[2]) < 0.0f || vdot(off[1], off[2]) < 0.0f;
}
__device__ bool box_contains_polygon_dev(V3 orig, float len2, const RtbTriangle& t) {
    if (point_in_cube(orig, len2, ld3(t.incenter))) return true;
    for (int k = 0; k < 3; ++k)
        if (point_in_cube(orig, len2, ld3(&t.corners[3 * k]))) return true;
    for (int f = 0; f < 6; ++f) {
        const float sgn = (f & 1) ? -1.0f : 1.0f;
        const V3 n = mk((f >> 1) == 0 ? sgn : 0.0f, (f >> 1) == 1 ? sgn : 0.0f, (f >> 1) == 2 ? sgn : 0.0f);
        if (face_cuts_triangle(orig, n, len2, t)) return true;
    }
    return false;
}

// flags[i] = 1 when triangle i enters the tree: never triangle 0 (raytrace.rs:791), the others when the root cube
// holds them (or always when root_len2 <= 0)
__global__ void k_cull_flags(const RtbTriangle* __restrict__ tris, uint32_t n, float3 root, float root_len2,
                             uint32_t* __restrict__ flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t keep = 0u;
    if (i >= 1u && i < n) keep = (root_len2 > 0.f && !(tris[i].kind & RTB_PRIM_SPHERE)) ? (box_contains_polygon_dev(mk(root.x, root.y, root.z), root_len2, tris[i]) ? 1u : 0u) : 1u;
    flags[i] = keep;    // flags[n] = 0 closes the exclusive scan
}
__global__ void k_compact_keep(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ pos, uint32_t n,
                               uint32_t* __restrict__ keep) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) keep[pos[i]] = i;
}

inline uint32_t cdiv64(uint64_t a, uint32_t b) { return (uint32_t)((a + b - 1) / b); }

}  // namespace

// out[0 .. nfaces*n_inst) on the device; *h_first_bad = index of the first triangle the reference would panic on, or
// 0xffffffff.  Synchronises the stream.
int rtb_launch_assemble(const float* d_verts, uint32_t nverts, const uint32_t* d_faces, uint32_t nfaces,
                        const RtbMeshInstance* d_inst, uint32_t n_inst, RtbTriangle* d_out, cudaStream_t stream,
                        uint32_t* h_first_bad) {
    *h_first_bad = 0xffffffffu;
    const uint64_t total = (uint64_t)nfaces * n_inst;
    if (total == 0) return RTB_OK;
    uint32_t* d_bad = nullptr;
    RTB_CUDA(cudaMalloc(&d_bad, sizeof(uint32_t)));
    cudaError_t e = cudaMemsetAsync(d_bad, 0xff, sizeof(uint32_t), stream);
    if (e == cudaSuccess) {
        k_assemble<<<cdiv64(total, 128), 128, 0, stream>>>(d_verts, nverts, d_faces, nfaces, d_inst, n_inst, d_out, d_bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_first_bad, d_bad, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_bad);
    RTB_CUDA(e);
    return RTB_OK;
}

// Root-cube cull + ordered compaction: *d_keep_out (cudaMalloc'ed here, caller frees) holds the indices of the
// *n_keep triangles that enter the tree, ascending (stream-ordered allocation: free with cudaFreeAsync / cudaFree).
// Synchronises the stream.
int rtb_launch_cull(const RtbTriangle* d_tris, uint32_t n, const float root_orig[3], float root_len2,
                    cudaStream_t stream, uint32_t** d_keep_out, uint32_t* n_keep) {
    *d_keep_out = nullptr; *n_keep = 0;
    uint32_t *flags = nullptr, *pos = nullptr, *keep = nullptr;
    uint8_t* tmp = nullptr;
    auto cleanup = [&] { cudaFreeAsync(flags, stream); cudaFreeAsync(pos, stream); cudaFreeAsync(tmp, stream); };
    cudaError_t e = cudaMallocAsync(&flags, sizeof(uint32_t) * ((size_t)n + 1), stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&pos, sizeof(uint32_t) * ((size_t)n + 1), stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&tmp, rtbsort::scan_tmp_bytes<uint32_t>(n + 1), stream);
    if (e != cudaSuccess) { cleanup(); RTB_CUDA(e); }
    const float3 root = root_orig ? make_float3(root_orig[0], root_orig[1], root_orig[2]) : make_float3(0.f, 0.f, 0.f);
    k_cull_flags<<<(n + 1 + 127) / 128, 128, 0, stream>>>(d_tris, n, root, root_orig ? root_len2 : 0.f, flags);
    rtbsort::exclusive_sum<uint32_t>(flags, pos, n + 1, tmp, stream);
    uint32_t total = 0;
    e = cudaMemcpyAsync(&total, pos + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaMallocAsync(&keep, sizeof(uint32_t) * (total ? total : 1), stream);
    if (e == cudaSuccess) {
        k_compact_keep<<<(n + 127) / 128 + 1, 128, 0, stream>>>(flags, pos, n, keep);
        e = cudaStreamSynchronize(stream);
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) { cudaFreeAsync(keep, stream); RTB_CUDA(e); }
    *d_keep_out = keep;
    *n_keep = total;
    return RTB_OK;
}
