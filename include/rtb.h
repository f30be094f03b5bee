/*
 * rtb.h — C ABI of the B200-native renderer core ("rtb" = ray-trace B200).
 *
 * Drop-in boundary behind the reference's `RayCaster` trait
 * (raytrace_lib/src/raytrace.rs:1128-1165).  A Rust shim
 * (`B200RayCaster: RayCaster`, see INTEGRATION.md and
 * rust_raytrace_b200/rust/) converts `Scene.tris` into RtbTriangle[], the
 * `Viewport` into RtbView, and hands `data: &mut [Color]` (16-byte f32x4
 * pixels) to rtb_render() as `rgba_out`.
 *
 * Replaces, on the reference side:
 *   - DefaultRayCaster::walk_rays_internal   raytrace.rs:1175-1195
 *   - Viewport::walk_ray_set / pixel_ray     raytrace.rs:1374-1440
 *   - project_ray / color_ray                raytrace.rs:1199-1295
 *   - BoundingBox::get_object_intersection_for_ray (octree) raytrace.rs:910-1050
 *     -> replaced by a GPU-built LBVH with identical closest-hit semantics
 *   - the WIP cxx FFI `exec_cuda_raytrace`   cuda_raytrace_lib/src/cuda_rt.h:7-14
 *
 * Conventions: plain C, caller owns every host buffer, the library owns device
 * memory for the lifetime of the scene handle.  Every function returns
 * RTB_OK (0) or a negative RtbStatus and never aborts; rtb_last_error() gives
 * a thread-local message.  There is no CPU fallback: without a CUDA device
 * every entry point that needs one fails with RTB_ERR_NO_DEVICE.
 */
#ifndef RTB_H
#define RTB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RTB_VERSION 1
#define RTB_MAX_DEPTH 16   /* Viewport.maxdepth above this is rejected (reference uses 5, main.rs:172) */
#define RTB_MAX_GPUS 8

typedef enum RtbStatus {
    RTB_OK = 0,
    RTB_ERR_NO_DEVICE = -1,
    RTB_ERR_CUDA = -2,
    RTB_ERR_INVALID = -3,
    RTB_ERR_NOMEM = -4
} RtbStatus;

/* SurfaceKind discriminants, raytrace.rs:303-308. */
enum { RTB_SOLID = 0, RTB_MATTE = 1, RTB_REFLECTIVE = 2 };

/* One `Triangle` (raytrace.rs:326-337) in the reference's field order, flattened
 * to 35 4-byte fields (140 B).  `surface` is flattened to kind/color/alpha/scattering
 * (Solid uses color only; Matte color+alpha; Reflective all three).  `num` is the
 * array index (populate_triangle_numbers, raytrace.rs:393-397) and is implicit.
 * Index 0 is the reference's dummy triangle: never tested, prim id 0 = miss. */
typedef struct RtbTriangle {
    float incenter[3];
    float norm[3];
    float bounding_r2;
    float sides[9];
    float side_lens[3];
    float corners[9];
    float edge_thickness;
    uint32_t kind;
    float color[3];
    float alpha;
    float scattering;
} RtbTriangle;

/* `Viewport` (raytrace.rs:1305-1318) plus the render controls the reference
 * hard-codes.  orig/cam/vu/vv are the private fields computed by
 * create_viewport (raytrace.rs:1343-1370). */
typedef struct RtbView {
    uint32_t width, height;
    float orig[3];
    float cam[3];
    float vu[3];
    float vv[3];
    uint32_t maxdepth;        /* >= 1, <= RTB_MAX_DEPTH */
    uint32_t spp;             /* samples_per_pixel; 1 => pixel centre, no jitter (raytrace.rs:1382) */
    uint64_t seed;            /* counter-based RNG seed (the reference's ThreadRng is unseeded) */
    uint32_t sample_begin;    /* progressive / multi-GPU sample partition: samples        */
    uint32_t sample_end;      /*   [sample_begin, sample_end) of spp; 0,0 => all           */
    uint32_t flags;           /* RTB_FLAG_* */
    uint32_t reserved;
} RtbView;

enum {
    RTB_FLAG_SUM_ONLY = 1u,   /* write the un-normalised sample sum (for a later cross-GPU reduce) */
    RTB_FLAG_STATS    = 2u,   /* also count node/triangle tests (slower kernel variant)            */
    RTB_FLAG_BRUTE    = 4u,   /* validation: ignore the BVH, test every primitive (the GPU analogue of
                                 build_trivial_bounding_box, raytrace.rs:847-856)                  */
    RTB_FLAG_MEGAKERNEL = 8u, /* A/B: the one-kernel-per-frame renderer (rtb_trace.cu) instead of the
                                 default wavefront pipeline (rtb_wavefront.cu); same results        */
    RTB_FLAG_TIMING   = 16u,  /* rtb_render_device only: bracket every pipeline stage with CUDA events on the
                                 launching stream and report RtbStats.ms_stage (one piece, one lane)  */
    RTB_FLAG_RESERVED32 = 32u, /* was RTB_FLAG_POOL (round 1's shared-memory ray-pool bounce kernel, measured slower and
                                 removed); ignored                                   */
    RTB_FLAG_BVH8     = 128u, /* A/B: traverse the 8-wide compressed BVH (80-byte nodes, 8-bit child boxes, octant-ordered hit
                                 masks) instead of the 4-wide uncompressed one; same results, measured 35 % slower on
                                 cache-resident scenes.  The scene must have been created with
                                 RTB_BVH8=1 in the environment, else RTB_ERR_INVALID                                */
    RTB_FLAG_COPY_ONLY = 256u, /* measurement: rtb_render / rtb_render_rgb8 issue every copy and event of the frame but launch no
                                 kernel — the device-to-host floor of the call (bench.py `e2e.d2h_floor_ms`)          */
    RTB_FLAG_FUSED    = 64u   /* A/B: primary phase and bounce phase of the path kernel as ONE launch per sample (the bounce
                                 phase consumes the queue while the primary phase still fills it) instead of two; same
                                 results, measured 5 % slower on one GPU and equal on a 1/8 share      */
};

/* Stage timers of the wavefront renderer, indices into RtbStats.ms_stage: the primary phase of the path kernel (ray
 * generation, traversal, shading of the primary hits) is RTB_STAGE_TRACE, the bounce phase RTB_STAGE_BOUNCE (with
 * RTB_FLAG_FUSED the one launch is reported under RTB_STAGE_TRACE).  RAYGEN and SHADE are no separate kernels any more. */
enum { RTB_STAGE_RAYGEN = 0, RTB_STAGE_TRACE = 1, RTB_STAGE_SHADE = 2, RTB_STAGE_BOUNCE = 3, RTB_N_STAGES = 4 };

typedef struct RtbStats {
    uint64_t rays;            /* project_ray calls with depth>0 — the reference's "Rays" (raytrace.rs:1278) */
    uint64_t node_tests;      /* AABB slab tests   (RTB_FLAG_STATS only) */
    uint64_t tri_tests;       /* exact triangle tests (RTB_FLAG_STATS only) */
    double   ms_render;       /* device time of the frame's kernels, max over GPUs (CUDA events) */
    double   ms_total;        /* host wall time of the call incl. copies */
    uint32_t kernel_launches; /* kernels launched by this call */
    uint32_t n_gpus;
    uint64_t bounce_rays;        /* rays after the primary ones (rays - primary rays): the bounce phase     */
    uint64_t node_tests_bounce;  /* the share of node_tests / tri_tests spent on bounce rays               */
    uint64_t tri_tests_bounce;   /*   (RTB_FLAG_STATS only)                                               */
    double   ms_stage[4];        /* RTB_FLAG_TIMING: device ms per stage, summed over samples (CUDA events) */
    double   ms_reduce;          /* rtb_render_progressive: device ms of the cross-GPU reduce + copy home, max over GPUs */
} RtbStats;

typedef struct RtbSceneInfo {
    uint32_t n_tris;          /* triangles passed in (incl. the dummy at 0) */
    uint32_t n_prims;         /* triangles in the BVH after the root-cube cull */
    uint32_t n_nodes;         /* 32-byte BVH nodes */
    uint32_t n_leaves;
    uint32_t max_leaf;
    uint32_t tree_height;
    float    scene_lo[3], scene_hi[3];
    double   ms_upload;       /* H2D + SoA repack */
    double   ms_build;        /* LBVH build (morton, sort, hierarchy, refit, emit), device time */
    uint32_t build_launches;
    uint32_t n_gpus;
    uint32_t n_refs;          /* primitive references in the BVH (>= n_prims: a primitive whose box is long compared
                                 to the scene is entered as several references with clipped boxes) */
    uint32_t reserved;
} RtbSceneInfo;

typedef struct rtb_scene rtb_scene;

/* Select the devices this process drives.  n_gpus = 0 -> all visible devices;
 * device_ids NULL -> 0..n_gpus-1.  Idempotent; re-initialising with a different
 * set is allowed when no scene is alive. */
int rtb_init(int n_gpus, const int* device_ids);
int rtb_device_count(void);          /* devices selected by rtb_init, or <0 */
int rtb_visible_device_count(void);  /* CUDA devices this process can see (no selection is changed), or RTB_ERR_NO_DEVICE */
void rtb_shutdown(void);
const char* rtb_last_error(void);

/* Upload `n` triangles, repack to the device layout and build the LBVH on every
 * selected GPU.  root_orig/root_len2: the reference's octree root cube
 * (build_bounding_box args, main.rs:160-164); triangles for which
 * box_contains_polygon(root) (raytrace.rs:753-779) is false are invisible in
 * the reference and are culled here too.  root_len2 <= 0 disables the cull. */
int rtb_scene_create(const RtbTriangle* tris, uint32_t n, const float root_orig[3], float root_len2,
                     rtb_scene** out);

/* Scene assembly on the GPU (the step in front of the path: obj_parser::parse_obj's per-face work,
 * obj_parser.rs:47-73, and make_triangle, raytrace.rs:340-383, as one kernel).  One indexed mesh
 * (verts: 3 f32 per vertex; faces: three 1-based vertex indices per face, as in an OBJ file) is instantiated
 * n_inst times: point = change_basis(v * scale, transform) + offset (rows of `transform` as create_transform
 * returns them, raytrace.rs:1320-1341), every face becomes one `Triangle` with the instance's surface. */
typedef struct RtbMeshInstance {
    float transform_rows[9];
    float offset[3];
    float scale;
    float edge_thickness;
    uint32_t kind;            /* RTB_SOLID / RTB_MATTE / RTB_REFLECTIVE */
    float color[3];
    float alpha;
    float scattering;
} RtbMeshInstance;

/* Like rtb_scene_create for the triangle array [dummy, instance 0 faces, instance 1 faces, ..., extra[0..n_extra)]
 * (main.rs:116-152: dummy, parse_obj(teapot), make_disk, make_disk), without that array ever existing on the host:
 * the `Triangle` records are computed on every selected GPU, bit-identical to the reference's make_triangle.
 * RTB_ERR_INVALID where the reference would panic (degenerate face, raytrace.rs:357) or a face index is out of range. */
int rtb_scene_create_instanced(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                               const RtbMeshInstance* inst, uint32_t n_inst, const RtbTriangle* extra,
                               uint32_t n_extra, const float root_orig[3], float root_len2, rtb_scene** out);
/* The assembly kernel alone: writes the nfaces*n_inst records to the HOST array `out` (inspection, parity tests, or a
 * host that wants its Vec<Triangle> computed at GPU speed). */
int rtb_assemble_triangles(const float* verts, uint32_t nverts, const uint32_t* faces, uint32_t nfaces,
                           const RtbMeshInstance* inst, uint32_t n_inst, RtbTriangle* out);
/* The root-cube cull kernel alone (box_contains_polygon, raytrace.rs:753-779, per triangle; triangle 0 never
 * passes): ascending indices of the kept triangles to keep_out (nullable, capacity n), their number to *n_keep. */
int rtb_cull_triangles(const RtbTriangle* tris, uint32_t n, const float root_orig[3], float root_len2,
                       uint32_t* keep_out, uint32_t* n_keep);

/* ---- EXTENSION (BASELINE config 1 "circles scene, primary + shadow rays") -------------------
 * Neither analytic spheres nor shadow rays exist in the mounted reference any more: the sphere primitive survives as
 * the CollisionFace::{Side, Face} vestiges (raytrace.rs:311-318), the shadow test as the commented-out block of
 * color_ray (raytrace.rs:1203-1224) with LightSource::get_shadow_ray (:594-610).  Their semantics are therefore
 * defined by this build (rust_raytrace_b200/csrc/rtb_ext.cu, restated in the oracle) and checked oracle-vs-GPU only.
 * Scenes that use them are rendered by the one-kernel extension renderer, through the same rtb_render* entry points. */
typedef struct RtbSphere {
    float center[3];
    float radius;             /* > 0 */
    uint32_t kind;            /* RTB_SOLID / RTB_MATTE / RTB_REFLECTIVE */
    float color[3];
    float alpha;
    float scattering;
} RtbSphere;
/* rtb_scene_create plus n_spheres analytic spheres; primitive ids: triangles 1..n-1, sphere j = n + j. */
int rtb_scene_create_ext(const RtbTriangle* tris, uint32_t n, const RtbSphere* spheres, uint32_t n_spheres,
                         const float root_orig[3], float root_len2, rtb_scene** out);
/* Scene.lights = Some(LightSource{orig, len2}) (raytrace.rs:594-597): every hit casts one shadow ray towards a random
 * point of the cube [orig, orig+len2)^3 and is black where any other object intersects that ray.  orig NULL = none. */
int rtb_scene_set_light(rtb_scene* s, const float orig[3], float len2);

int rtb_scene_info(const rtb_scene* s, RtbSceneInfo* out);
void rtb_scene_destroy(rtb_scene* s);
/* Debug/inspection: copy the BVH of GPU 0 back (nodes: n_nodes*8 floats; prim_order: n_refs u32
 * = original triangle index of each leaf-order slot).  Either pointer may be NULL. */
int rtb_scene_download_bvh(const rtb_scene* s, float* nodes, uint32_t* prim_order);

/* Render one frame into HOST buffers (the RayCaster::walk_rays_internal replacement).
 *   rgba_out : width*height*4 f32, row-major `row*width+col`, lane 3 = 0 — may be the Rust `&mut [Color]`.
 *   prim_out : nullable, width*height u32 — primitive id of the primary ray of sample 0, 0 = miss.
 *   t_out    : nullable, width*height f32 — its hit time (0 on miss).
 * Work is split over the selected GPUs in interleaved 16x8-pixel tiles; every GPU copies its own
 * tiles back over its own PCIe link.  Pin the buffers with rtb_host_register for full D2H speed. */
int rtb_render(rtb_scene* s, const RtbView* view, float* rgba_out, uint32_t* prim_out, float* t_out,
               RtbStats* stats);

/* The same frame delivered the way main.rs consumes it (walk_rays then write_png, main.rs:191-227): quantised on the
 * GPU with write_png's `(c * 255.) as u8` (raytrace.rs:1468-1473) and copied home as width*height*3 bytes, RGB,
 * row-major — 3 instead of 16 bytes per pixel over PCIe.  rtbh_write_png_rgb8 (rtb_host.h) puts it in a PNG file. */
int rtb_render_rgb8(rtb_scene* s, const RtbView* view, uint8_t* rgb_out, RtbStats* stats);

/* Render the tile subset `tile_rank` of `tile_world` (tile i belongs to rank i % world) of one frame
 * into DEVICE buffers (full-frame indexing, pixels of other ranks untouched) on GPU slot `gpu`
 * (index into the rtb_init set) and CUDA stream `stream` (a cudaStream_t cast to void*, NULL = the
 * library's stream).  Asynchronous when stats == NULL.  Used by bench.py's resident-input timing and
 * by the one-process-per-GPU launch (torchrun), where the caller owns the buffers. */
int rtb_render_device(rtb_scene* s, const RtbView* view, int gpu, uint32_t tile_rank, uint32_t tile_world,
                      float* d_rgba, uint32_t* d_prim, float* d_t, void* stream, RtbStats* stats);

/* Multi-sample frame with the SAMPLES partitioned over the selected GPUs (sample s -> GPU s % n),
 * per-GPU f32 sum buffers combined over NVLink peer memory: every GPU reduces and normalises one
 * horizontal band reading its peers' buffers directly, then copies the band to rgba_out.  */
int rtb_render_progressive(rtb_scene* s, const RtbView* view, float* rgba_out, RtbStats* stats);

/* One-process-per-GPU sample partition (torchrun): every rank renders its samples [sample_begin, sample_end) with
 * RTB_FLAG_SUM_ONLY through rtb_render_device, the sum buffers are combined by an NCCL reduce (the caller's
 * torch.distributed / ncclReduce), and the root applies walk_ray_set's final `* (1/spp)` (raytrace.rs:1426) with this
 * call: d_rgba[i].xyz *= 1/spp, lane 3 = 0, on GPU slot `gpu` and stream `stream` (NULL = the library's stream). */
int rtb_scale_device(float* d_rgba, uint64_t npix, uint32_t spp, int gpu, void* stream);

/* (c*255.) as u8 quantiser of write_png (raytrace.rs:1468-1473) on the GPU: rgba f32 host -> rgb8 host. */
int rtb_quantize_rgb8(const float* rgba, uint64_t npix, uint8_t* rgb_out);

/* Which image rows does `rank` of `world` render?  (Band partition: 8-row band b belongs to rank
 * b % world.)  Writes up to `cap` row indices in ascending order, returns the number of rows owned.
 * Host-only; needs no GPU. */
int rtb_partition_rows(uint32_t height, uint32_t rank, uint32_t world, uint32_t* rows_out, uint32_t cap);

/* Device self-test of the BVH builder's hand-written radix sort (stable, 64-bit keys, 32-bit values) and exclusive
 * scans against the host on n pseudo-random pairs with key_bits significant key bits. */
int rtb_selftest_sort(uint32_t n, int key_bits, uint64_t seed);

/* Host self-test of the division by a per-frame constant that maps a pixel slot to its pixel (slot_to_pixel; Granlund &
 * Montgomery's round-up method): q = n / d through the precomputed multiplier must equal the machine's n / d for divisor d and
 * `samples` pseudo-random 32-bit numerators plus the boundary cases.  Returns the number of mismatches (0 = exact), or
 * RTB_ERR_INVALID for d == 0.  Host-only; needs no GPU. */
int rtb_selftest_udiv(uint32_t d, uint32_t samples, uint64_t seed);

/* ---- one process per GPU, one frame on one GPU (torchrun; reference: the row queue all workers write one `data` slice
 * from, raytrace.rs:1179-1194) ------------------------------------------------------------------------------------
 * The root rank allocates the frame with rtb_device_alloc and exports it (rtb_ipc_export, a 64-byte
 * cudaIpcMemHandle_t to be sent to the other ranks by any means); every other rank maps it with rtb_ipc_open and passes
 * the mapped pointer to rtb_render_device as d_rgba: its kernels then store their bands straight into the root GPU's
 * memory over NVLink peer access, and the frame is complete on the root GPU as soon as every rank's stream has drained —
 * no gather kernel, no copy, no collective. */
int rtb_device_alloc(int gpu, size_t bytes, void** d_ptr);           /* zero-filled device memory on GPU slot `gpu` */
int rtb_device_free(int gpu, void* d_ptr);
int rtb_ipc_export(void* d_ptr, unsigned char handle_out[64]);
int rtb_ipc_open(int gpu, const unsigned char handle[64], void** d_ptr_out);
int rtb_ipc_close(int gpu, void* d_ptr);

/* Pin / unpin a caller-owned host buffer (cudaHostRegister) so D2H runs at PCIe speed. */
int rtb_host_register(void* ptr, size_t bytes);
int rtb_host_unregister(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* RTB_H */
