//! `B200RayCaster` and the transcription of include/rtb.h (crate feature `gpu`, on by default).
use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};
use std::sync::mpsc::Sender;
use std::sync::Mutex;
use std::time::Duration;

use raytrace_lib::progress::ProgressStat;
use raytrace_lib::raytrace::{Color, RayCaster, Scene, SurfaceKind, Triangle, Vec3, Viewport};

// ---------------------------------------------------------------------------------------------
// include/rtb.h, transcribed
// ---------------------------------------------------------------------------------------------
pub const RTB_SOLID: u32 = 0;
pub const RTB_MATTE: u32 = 1;
pub const RTB_REFLECTIVE: u32 = 2;
pub const RTB_FLAG_STATS: u32 = 2;

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbTriangle {
    pub incenter: [f32; 3],
    pub norm: [f32; 3],
    pub bounding_r2: f32,
    pub sides: [f32; 9],
    pub side_lens: [f32; 3],
    pub corners: [f32; 9],
    pub edge_thickness: f32,
    pub kind: u32,
    pub color: [f32; 3],
    pub alpha: f32,
    pub scattering: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbView {
    pub width: u32,
    pub height: u32,
    pub orig: [f32; 3],
    pub cam: [f32; 3],
    pub vu: [f32; 3],
    pub vv: [f32; 3],
    pub maxdepth: u32,
    pub spp: u32,
    pub seed: u64,
    pub sample_begin: u32,
    pub sample_end: u32,
    pub flags: u32,
    pub reserved: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RtbStats {
    pub rays: u64,
    pub node_tests: u64,
    pub tri_tests: u64,
    pub ms_render: f64,
    pub ms_total: f64,
    pub kernel_launches: u32,
    pub n_gpus: u32,
    pub bounce_rays: u64,
    pub node_tests_bounce: u64,
    pub tri_tests_bounce: u64,
    pub ms_stage: [f64; 4],
    pub ms_reduce: f64,
}

#[repr(C)]
pub struct RtbSceneOpaque {
    _private: [u8; 0],
}

extern "C" {
    fn rtb_init(n_gpus: c_int, device_ids: *const c_int) -> c_int;
    fn rtb_visible_device_count() -> c_int;
    fn rtb_last_error() -> *const c_char;
    fn rtb_scene_create(tris: *const RtbTriangle, n: u32, root_orig: *const f32, root_len2: f32,
                        out: *mut *mut RtbSceneOpaque) -> c_int;
    fn rtb_scene_destroy(s: *mut RtbSceneOpaque);
    fn rtb_render(s: *mut RtbSceneOpaque, view: *const RtbView, rgba_out: *mut f32, prim_out: *mut u32,
                  t_out: *mut f32, stats: *mut RtbStats) -> c_int;
    fn rtb_render_progressive(s: *mut RtbSceneOpaque, view: *const RtbView, rgba_out: *mut f32,
                              stats: *mut RtbStats) -> c_int;
    fn rtb_render_rgb8(s: *mut RtbSceneOpaque, view: *const RtbView, rgb_out: *mut u8, stats: *mut RtbStats) -> c_int;
    fn rtb_scene_set_light(s: *mut RtbSceneOpaque, orig: *const f32, len2: f32) -> c_int;
    fn rtbh_write_png_rgb8(path: *const c_char, width: u32, height: u32, rgb: *const u8) -> c_int;
    fn rtb_host_register(ptr: *mut c_void, bytes: usize) -> c_int;
    fn rtb_host_unregister(ptr: *mut c_void) -> c_int;
}

fn last_error() -> String {
    unsafe { CStr::from_ptr(rtb_last_error()).to_string_lossy().into_owned() }
}

fn v3(v: &Vec3) -> [f32; 3] {
    let a = v.v.to_array();
    [a[0], a[1], a[2]]
}

/// `Triangle` (raytrace.rs:326-337) -> the flat ABI record.  Pure copies, no arithmetic.
pub fn flatten_triangle(t: &Triangle) -> RtbTriangle {
    let mut o = RtbTriangle::default();
    o.incenter = v3(&t.incenter);
    o.norm = v3(&t.norm);
    o.bounding_r2 = t.bounding_r2;
    for i in 0..3 {
        o.sides[3 * i..3 * i + 3].copy_from_slice(&v3(&t.sides[i]));
        o.corners[3 * i..3 * i + 3].copy_from_slice(&v3(&t.corners[i]));
        o.side_lens[i] = t.side_lens[i];
    }
    o.edge_thickness = t.edge_thickness;
    match t.surface {
        SurfaceKind::Solid { color } => {
            o.kind = RTB_SOLID;
            o.color = v3(&color);
        }
        SurfaceKind::Matte { color, alpha } => {
            o.kind = RTB_MATTE;
            o.color = v3(&color);
            o.alpha = alpha;
        }
        SurfaceKind::Reflective { scattering, color, alpha } => {
            o.kind = RTB_REFLECTIVE;
            o.color = v3(&color);
            o.alpha = alpha;
            o.scattering = scattering;
        }
    }
    o
}

/// The private `orig/cam/vu/vv` of `Viewport` (raytrace.rs:1310-1314) recovered bit-exactly from its derived
/// `Debug` output: Rust prints an f32 with the shortest digits that round-trip, so `parse::<f32>()` of the
/// printed text returns the original bits.  (A four-line upstream patch adding getters makes this
/// unnecessary; see INTEGRATION.md.)  Format: `Viewport { width: W, height: H, orig: Vec3 { v: [a, b, c, d] },
/// cam: Vec3 { v: [..] }, vu: Vec3 { v: [..] }, vv: Vec3 { v: [..] }, maxdepth: M, samples_per_pixel: S }`.
pub fn view_from_viewport(v: &Viewport, seed: u64) -> RtbView {
    let dbg = format!("{:?}", v);
    let field = |name: &str| -> [f32; 3] {
        let key = format!("{}: Vec3 {{ v: [", name);
        let start = dbg.find(&key).expect("Viewport Debug layout changed") + key.len();
        let end = start + dbg[start..].find(']').unwrap();
        let mut out = [0f32; 3];
        for (i, tok) in dbg[start..end].split(',').take(3).enumerate() {
            out[i] = tok.trim().parse::<f32>().expect("f32 in Viewport Debug");
        }
        out
    };
    RtbView {
        width: v.width as u32,
        height: v.height as u32,
        orig: field("orig"),
        cam: field("cam"),
        vu: field("vu"),
        vv: field("vv"),
        maxdepth: v.maxdepth as u32,
        spp: v.samples_per_pixel as u32,
        seed,
        ..Default::default()
    }
}

pub const RTB_MAX_GPUS: usize = 8;

struct Uploaded {
    handle: *mut RtbSceneOpaque,
    key: SceneKey,
}
unsafe impl Send for Uploaded {}

/// What the cached device scene was built from.  Address and length of `Scene.tris` alone would hand out a stale scene
/// after an in-place edit or a new `Vec` at the old address, so the key carries a hash of the flattened triangle bytes and
/// the octree root cube the visibility cull uses (raytrace.rs:795-805).  Hashing 6,721 triangles costs ~10 us per frame.
#[derive(Clone, Copy, PartialEq, Eq)]
struct SceneKey {
    n_tris: usize,
    content: u64,
    root: [u32; 4],
    n_gpus: usize,
}

fn fnv1a(bytes: &[u8]) -> u64 {
    let mut h: u64 = 0xcbf29ce484222325;
    for chunk in bytes.chunks(8) {
        let mut w = [0u8; 8];
        w[..chunk.len()].copy_from_slice(chunk);
        h = (h ^ u64::from_le_bytes(w)).wrapping_mul(0x100000001b3);
    }
    h
}

/// `threads` of `walk_rays` (main.rs:83 passes 16) reinterpreted as a GPU count: 0 = all visible devices, anything larger
/// than what exists is clamped — never an error.
fn gpu_count(threads: usize) -> Result<usize, String> {
    let visible = unsafe { rtb_visible_device_count() };
    if visible <= 0 {
        return Err(last_error());
    }
    let visible = visible as usize;
    let want = if threads == 0 { visible } else { threads };
    Ok(want.min(visible).min(RTB_MAX_GPUS))
}

/// The caster.  `threads` of `walk_rays` is reinterpreted as the number of GPUs (0 = all visible, clamped to what exists).
/// The uploaded scene (device SoA + LBVH) is cached across frames, keyed by the CONTENT of `Scene.tris` and the root cube;
/// `invalidate()` drops it explicitly.
pub struct B200RayCaster {
    pub seed: u64,
    /// samples of a multi-spp frame are partitioned over the GPUs and reduced over NVLink (rtb_render_progressive)
    pub progressive: bool,
    /// EXTENSION: `LightSource {orig, len2}` (raytrace.rs:594-597) — turns the commented-out shadow block of color_ray
    /// (raytrace.rs:1203-1224) on; the current `Scene` has no `lights` field to carry it
    pub light: Option<([f32; 3], f32)>,
    cache: Mutex<Option<Uploaded>>,
    /// Pin the caller's image buffer (cudaHostRegister) for the duration of each call: full D2H speed, but registering
    /// 133 MB costs milliseconds per frame.  The registration never outlives `walk_rays_internal` — the caller owns the
    /// `Vec` and may free or move it at any time afterwards.  A caller that renders many frames into one buffer should
    /// pin it itself once (`pin_buffer` / `unpin_buffer`) and leave this off.
    pub pin_per_call: bool,
}

unsafe impl Send for B200RayCaster {}
unsafe impl Sync for B200RayCaster {}

impl B200RayCaster {
    pub fn new() -> Self {
        B200RayCaster { seed: 0, progressive: false, light: None, cache: Mutex::new(None), pin_per_call: false }
    }

    /// Forget the cached device scene (the next frame uploads and builds again).
    pub fn invalidate(&self) {
        if let Some(u) = self.cache.lock().unwrap().take() {
            unsafe { rtb_scene_destroy(u.handle) };
        }
    }

    /// Caller-owned pinning of an image buffer that lives across many frames; undo with `unpin_buffer` BEFORE the buffer
    /// is freed or reallocated.
    pub fn pin_buffer(data: &mut [Color]) -> bool {
        unsafe { rtb_host_register(data.as_mut_ptr() as *mut c_void, data.len() * 16) == 0 }
    }
    pub fn unpin_buffer(data: &mut [Color]) {
        unsafe { rtb_host_unregister(data.as_mut_ptr() as *mut c_void) };
    }

    fn scene_handle(&self, s: &Scene, threads: usize) -> Result<*mut RtbSceneOpaque, String> {
        let n_gpus = gpu_count(threads)?;
        let flat: Vec<RtbTriangle> = s.tris.iter().map(flatten_triangle).collect();
        let bytes = unsafe { std::slice::from_raw_parts(flat.as_ptr() as *const u8, flat.len() * std::mem::size_of::<RtbTriangle>()) };
        // the octree root cube (pub fields, raytrace.rs:618-623) drives the same visibility cull as :795-805
        let root = v3(&s.boxes.orig);
        let key = SceneKey {
            n_tris: flat.len(),
            content: fnv1a(bytes),
            root: [root[0].to_bits(), root[1].to_bits(), root[2].to_bits(), s.boxes.len2.to_bits()],
            n_gpus,
        };
        let mut c = self.cache.lock().unwrap();
        if let Some(u) = c.as_ref() {
            if u.key == key {
                return Ok(u.handle);
            }
            unsafe { rtb_scene_destroy(u.handle) };
            *c = None;
        }
        if unsafe { rtb_init(n_gpus as c_int, std::ptr::null()) } != 0 {
            return Err(last_error());
        }
        let mut h: *mut RtbSceneOpaque = std::ptr::null_mut();
        let rc = unsafe { rtb_scene_create(flat.as_ptr(), flat.len() as u32, root.as_ptr(), s.boxes.len2, &mut h) };
        if rc != 0 {
            return Err(last_error());
        }
        *c = Some(Uploaded { handle: h, key });
        Ok(h)
    }
}

impl B200RayCaster {
    fn apply_light(&self, h: *mut RtbSceneOpaque) {
        match self.light {
            Some((o, len2)) => unsafe { rtb_scene_set_light(h, o.as_ptr(), len2) },
            None => unsafe { rtb_scene_set_light(h, std::ptr::null(), 0.0) },
        };
    }

    /// main.rs:191-227 in one call: the frame of `walk_rays` quantised on the GPU with write_png's `(c * 255.) as u8`
    /// (raytrace.rs:1468-1473), 3 bytes per pixel over PCIe instead of 16, written as an 8-bit RGB PNG.
    /// Returns the number of rays (the reference's "Rays" stat).
    pub fn render_png(&self, v: &Viewport, s: &Scene, n_gpus: usize, path: &str) -> Result<u64, String> {
        let h = self.scene_handle(s, n_gpus)?;
        self.apply_light(h);
        let view = view_from_viewport(v, self.seed);
        let mut rgb = vec![0u8; v.width * v.height * 3];
        let mut st = RtbStats::default();
        if unsafe { rtb_render_rgb8(h, &view, rgb.as_mut_ptr(), &mut st) } != 0 {
            return Err(last_error());
        }
        let cpath = std::ffi::CString::new(path).map_err(|e| e.to_string())?;
        if unsafe { rtbh_write_png_rgb8(cpath.as_ptr(), v.width as u32, v.height as u32, rgb.as_ptr()) } != 0 {
            return Err(format!("cannot write {}", path));
        }
        Ok(st.rays)
    }
}

impl Drop for B200RayCaster {
    fn drop(&mut self) {
        if let Some(u) = self.cache.lock().unwrap().take() {
            unsafe { rtb_scene_destroy(u.handle) };
        }
    }
}

impl RayCaster for B200RayCaster {
    fn walk_rays_internal(&self, v: &Viewport, s: &Scene, data: &mut [Color], threads: usize,
                          progress_tx: Sender<(usize, usize, usize, HashMap<String, ProgressStat>)>) {
        assert_eq!(data.len(), v.width * v.height);
        assert_eq!(std::mem::size_of::<Color>(), 16);
        let h = self.scene_handle(s, threads).unwrap_or_else(|e| panic!("b200: {}", e));
        self.apply_light(h);
        let view = view_from_viewport(v, self.seed);
        let mut st = RtbStats::default();
        let bytes = data.len() * 16;
        let p = data.as_mut_ptr() as *mut f32;
        // Scene.debug_en: also fetch the primary hits, to fill Scene.debug_ctx the way project_ray does (below)
        let mut prim: Vec<u32> = if s.debug_en { vec![0u32; data.len()] } else { Vec::new() };
        let mut hit_t: Vec<f32> = if s.debug_en { vec![0f32; data.len()] } else { Vec::new() };
        let (prim_p, t_p) = if s.debug_en { (prim.as_mut_ptr(), hit_t.as_mut_ptr()) } else { (std::ptr::null_mut(), std::ptr::null_mut()) };
        // registered for this call only (failure only costs D2H speed); see `pin_per_call`
        let pinned = self.pin_per_call && unsafe { rtb_host_register(p as *mut c_void, bytes) } == 0;
        unsafe {
            let rc = if self.progressive && view.spp > 1 {
                rtb_render_progressive(h, &view, p, &mut st)
            } else {
                rtb_render(h, &view, p, prim_p, t_p, &mut st)
            };
            if pinned {
                rtb_host_unregister(p as *mut c_void);
            }
            if rc != 0 {
                panic!("b200: rtb_render failed: {}", last_error());   // the reference's error convention (unwrap)
            }
        }
        // The reference's own parity mechanism is the two-caster differential run of main.rs:181-227: with
        // `Scene.debug_en` every caster records {ray, tri_hit, hit_t} per pixel in `Scene.debug_ctx` and
        // `DebugCtx::compare_to` (debug.rs:150-221) classifies the mismatches.  project_ray does that for the CPU caster
        // (raytrace.rs:1266-1289); this does the same for the GPU frame, so `compare_to` can be run between a
        // DefaultRayCaster scene and a B200RayCaster scene unchanged.  (1 spp only: with more samples pixel_ray draws
        // from ThreadRng.)  check_tris gets the hit itself — the GPU has no octree leaf lists, and an empty list makes
        // RayDebugCtx's Display panic (debug.rs:29, `reduce(..).unwrap()`).
        if s.debug_en && v.samples_per_pixel == 1 && !(self.progressive && view.spp > 1) {
            let mut ctx = s.debug_ctx.lock().unwrap();
            for row in 0..v.height {
                for col in 0..v.width {
                    let ray = v.pixel_ray((row, col));
                    ctx.register_ray(&ray, (row, col));
                    ctx.add_ray(&ray);
                    let i = row * v.width + col;
                    ctx.update_ray_triangles(&ray, &vec![prim[i] as usize]);
                    if prim[i] != 0 {
                        ctx.update_ray_hit(&ray, prim[i] as usize, hit_t[i]);
                    }
                }
            }
        }
        // keeps ProgressCtx::print_stats (progress.rs:157-185) meaningful: "Rays" feeds total_rays / Mrays/s
        let mut m: HashMap<String, ProgressStat> = HashMap::new();
        m.insert("Rays".to_string(), ProgressStat::Count(st.rays as usize));
        m.insert("GPU Render".to_string(), ProgressStat::Time(Duration::from_secs_f64(st.ms_render * 1e-3)));
        m.insert("GPU Total".to_string(), ProgressStat::Time(Duration::from_secs_f64(st.ms_total * 1e-3)));
        let _ = progress_tx.send((0, v.height - 1, v.width * v.height, m));
        // progress_tx dropped here -> RayCaster::walk_rays' wait loop ends (raytrace.rs:1150-1158)
    }
}
