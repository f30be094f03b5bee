//! b200_raytrace_lib — `B200RayCaster`, a third sibling of `DefaultRayCaster` (raytrace.rs:1167-1196) and
//! `CudaRayCaster` (cuda_raytrace.rs:544-573) behind the reference's `RayCaster` trait (raytrace.rs:1128-1165).
//!
//! The whole hot path (pixel_ray -> closest hit -> color_ray recursion -> per-pixel accumulation) runs on the
//! GPU(s) inside `rtb_render` (include/rtb.h).  This file only marshals: `Scene.tris` -> `RtbTriangle[]`,
//! `Viewport` -> `RtbView`, `data: &mut [Color]` -> `float* rgba_out` (a `Color` IS four f32, lane 3 = 0).
//!
//! NOT COMPILED IN THE BUILD IMAGE (no Rust toolchain there); the identical ABI is driven from C++ and
//! Python (rust_raytrace_b200/raytrace.py mirrors this file function by function).
#![feature(portable_simd)]

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
}

#[repr(C)]
pub struct RtbSceneOpaque {
    _private: [u8; 0],
}

extern "C" {
    fn rtb_init(n_gpus: c_int, device_ids: *const c_int) -> c_int;
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

struct Uploaded {
    handle: *mut RtbSceneOpaque,
    tris_ptr: usize,
    tris_len: usize,
    n_gpus: usize,
}
unsafe impl Send for Uploaded {}

/// The caster.  `threads` of `walk_rays` is reinterpreted as the number of GPUs (0 = all visible).
/// The uploaded scene (device SoA + LBVH) is cached across frames, keyed by the address and length of `Scene.tris`.
pub struct B200RayCaster {
    pub seed: u64,
    /// samples of a multi-spp frame are partitioned over the GPUs and reduced over NVLink (rtb_render_progressive)
    pub progressive: bool,
    /// EXTENSION: `LightSource {orig, len2}` (raytrace.rs:594-597) — turns the commented-out shadow block of color_ray
    /// (raytrace.rs:1203-1224) on; the current `Scene` has no `lights` field to carry it
    pub light: Option<([f32; 3], f32)>,
    cache: Mutex<Option<Uploaded>>,
    /// (address, bytes) of the caller's image buffer currently pinned with rtb_host_register: pinning 133 MB costs
    /// milliseconds, so it is done once per buffer, not once per frame
    pinned: Mutex<Option<(usize, usize)>>,
}

unsafe impl Send for B200RayCaster {}
unsafe impl Sync for B200RayCaster {}

impl B200RayCaster {
    pub fn new() -> Self {
        B200RayCaster { seed: 0, progressive: false, light: None, cache: Mutex::new(None), pinned: Mutex::new(None) }
    }

    fn scene_handle(&self, s: &Scene, n_gpus: usize) -> Result<*mut RtbSceneOpaque, String> {
        let mut c = self.cache.lock().unwrap();
        let key = (s.tris.as_ptr() as usize, s.tris.len());
        if let Some(u) = c.as_ref() {
            if (u.tris_ptr, u.tris_len, u.n_gpus) == (key.0, key.1, n_gpus) {
                return Ok(u.handle);
            }
            unsafe { rtb_scene_destroy(u.handle) };
            *c = None;
        }
        if unsafe { rtb_init(n_gpus as c_int, std::ptr::null()) } != 0 {
            return Err(last_error());
        }
        let flat: Vec<RtbTriangle> = s.tris.iter().map(flatten_triangle).collect();
        // the octree root cube (pub fields, raytrace.rs:618-623) drives the same visibility cull as :795-805
        let root = v3(&s.boxes.orig);
        let mut h: *mut RtbSceneOpaque = std::ptr::null_mut();
        let rc = unsafe { rtb_scene_create(flat.as_ptr(), flat.len() as u32, root.as_ptr(), s.boxes.len2, &mut h) };
        if rc != 0 {
            return Err(last_error());
        }
        *c = Some(Uploaded { handle: h, tris_ptr: key.0, tris_len: key.1, n_gpus });
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

    fn pin(&self, ptr: *mut c_void, bytes: usize) {
        let mut p = self.pinned.lock().unwrap();
        if *p == Some((ptr as usize, bytes)) {
            return;
        }
        if let Some((old, _)) = p.take() {
            unsafe { rtb_host_unregister(old as *mut c_void) };
        }
        if unsafe { rtb_host_register(ptr, bytes) } == 0 {      // failure only costs D2H speed
            *p = Some((ptr as usize, bytes));
        }
    }
}

impl Drop for B200RayCaster {
    fn drop(&mut self) {
        if let Some((old, _)) = self.pinned.lock().unwrap().take() {
            unsafe { rtb_host_unregister(old as *mut c_void) };
        }
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
        self.pin(p as *mut c_void, bytes);
        unsafe {
            let rc = if self.progressive && view.spp > 1 {
                rtb_render_progressive(h, &view, p, &mut st)
            } else {
                rtb_render(h, &view, p, std::ptr::null_mut(), std::ptr::null_mut(), &mut st)
            };
            if rc != 0 {
                panic!("b200: rtb_render failed: {}", last_error());   // the reference's error convention (unwrap)
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
