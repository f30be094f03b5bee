//! raytrace_b200 — raytrace/src/main.rs:89-227 with the third caster.
//!
//!     cargo +nightly run --release --bin raytrace_b200 -- [obj] [WxH] [gpus]
//!
//! Same scene, same camera, `DefaultRayCaster` first and `B200RayCaster` second (where main.rs runs `CudaRayCaster`),
//! `print_stats`, `write_png("test.png")`, the two debug CSVs and `DebugCtx::compare_to` -> debug_diffs.txt, whose last
//! line is the reference's own verdict, `Found N errors`.  The SDL window of main.rs:229-272 is left out.
//! Deterministic materials by default (`RTB_SHIPPED=1` for the shipped ones, where only primary hits are comparable).
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Rust toolchain); `rust_raytrace_b200/raytrace_b200` (C++) is the
//! same driver over the same ABI and is what the GPU tests run.
use std::fs;
use std::path::Path;

use b200_raytrace_lib::scene::{main_scene, main_tris, main_viewport};
use b200_raytrace_lib::B200RayCaster;
use raytrace_lib::raytrace::{self, make_vec, DefaultRayCaster, RayCaster};

fn main() -> std::io::Result<()> {
    let args: Vec<String> = std::env::args().collect();
    let obj = args.get(1).map(String::as_str).unwrap_or("../raytrace/teapot_tri.obj");
    let (width, height) = args.get(2)
        .and_then(|s| s.split_once('x'))
        .map(|(w, h)| (w.parse::<u32>().unwrap(), h.parse::<u32>().unwrap()))
        .unwrap_or((64, 64));                                        // main.rs:108-110
    let gpus: usize = args.get(3).map(|s| s.parse().unwrap()).unwrap_or(1);
    let det = std::env::var("RTB_SHIPPED").map(|v| v != "1").unwrap_or(true);

    let file = fs::File::create(Path::new("test.png"))?;
    let v = main_viewport(width, height, 5, 1);
    let s_default = main_scene(main_tris(obj, det), true, true);
    let mut data = vec![make_vec(&[0., 0., 0.]); (width * height) as usize];
    let progress_default = DefaultRayCaster {}.walk_rays(&v, &s_default, &mut data, 1, false);
    progress_default.print_stats();
    let reference_frame = data.clone();

    // main.rs:202-209 moves tris and boxes into the second scene; the B200 caster reads only tris and the root cube
    let s_b200 = main_scene(main_tris(obj, det), false, true);
    let caster_b200 = B200RayCaster::new();
    let progress_ctx = caster_b200.walk_rays(&v, &s_b200, &mut data, gpus, false);
    progress_ctx.print_stats();
    let _ = raytrace::write_png(file, (width, height), &data);

    let differing = data.iter().zip(reference_frame.iter())
        .filter(|(a, b)| (0..4).any(|k| a.v[k].to_bits() != b.v[k].to_bits()))
        .count();
    println!("pixels whose RGBA bits differ from DefaultRayCaster: {} of {}", differing, data.len());

    {
        let mut f = fs::File::create("debug_default.csv").unwrap();
        let ctx = s_default.debug_ctx.lock().unwrap();
        ctx.write_debug_header(&mut f);
        ctx.write_all_debug_context(&mut f);
    }
    {
        let mut f = fs::File::create("debug_b200.csv").unwrap();
        let ctx = s_b200.debug_ctx.lock().unwrap();
        ctx.write_debug_header(&mut f);
        ctx.write_all_debug_context(&mut f);
    }
    let mut f = fs::File::create("debug_diffs.txt").unwrap();
    let a = s_default.debug_ctx.lock().unwrap();
    let b = s_b200.debug_ctx.lock().unwrap();
    a.compare_to(&b, &mut f);                                        // ends with "Found N errors" (debug.rs:220)
    println!("{}", fs::read_to_string("debug_diffs.txt").unwrap().lines().last().unwrap_or(""));
    Ok(())
}
