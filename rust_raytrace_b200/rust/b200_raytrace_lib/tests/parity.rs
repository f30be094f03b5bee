//! cargo +nightly test --release -- --nocapture     (needs a B200 and nvcc; see build.rs)
//!
//! The reference's parity mechanism, as a test: main.rs's scene in the deterministic material set rendered by
//! `DefaultRayCaster` (octree, CPU) and by `B200RayCaster` (LBVH, GPU); every pixel's RGBA must have the same bits, and
//! `DebugCtx::compare_to` (debug.rs:150-221) must end with "Found 0 errors".  On a mismatch the classified report of
//! compare_to is printed together with a per-pixel list in the same style.
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Rust toolchain).
use b200_raytrace_lib::scene::{main_scene, main_tris, main_viewport};
use b200_raytrace_lib::B200RayCaster;
use raytrace_lib::raytrace::{make_vec, DefaultRayCaster, RayCaster};

fn obj_path() -> String {
    std::env::var("RTB_TEAPOT_OBJ").unwrap_or_else(|_| "../raytrace/teapot_tri.obj".to_string())
}

fn differential(width: u32, height: u32) {
    let v = main_viewport(width, height, 5, 1);
    let s_cpu = main_scene(main_tris(&obj_path(), true), true, true);
    let s_gpu = main_scene(main_tris(&obj_path(), true), false, true);
    let mut cpu = vec![make_vec(&[0., 0., 0.]); (width * height) as usize];
    let mut gpu = cpu.clone();
    let p_cpu = DefaultRayCaster {}.walk_rays(&v, &s_cpu, &mut cpu, 1, false);
    let p_gpu = B200RayCaster::new().walk_rays(&v, &s_gpu, &mut gpu, 1, false);
    p_cpu.print_stats();
    p_gpu.print_stats();

    let mut report: Vec<u8> = Vec::new();
    s_cpu.debug_ctx.lock().unwrap().compare_to(&s_gpu.debug_ctx.lock().unwrap(), &mut report);
    let report = String::from_utf8(report).unwrap();
    let mut bad = 0usize;
    for (i, (a, b)) in cpu.iter().zip(gpu.iter()).enumerate() {
        if (0..4).any(|k| a.v[k].to_bits() != b.v[k].to_bits()) {
            if bad < 20 {
                println!("({},{}): Colour Mismatch {:?} vs {:?}", i / width as usize, i % width as usize, a.v, b.v);
            }
            bad += 1;
        }
    }
    if bad != 0 || !report.trim_end().ends_with("Found 0 errors") {
        println!("{}", report);
    }
    assert!(report.trim_end().ends_with("Found 0 errors"), "DebugCtx::compare_to reports mismatches at {}x{}", width, height);
    assert_eq!(bad, 0, "{} pixels differ in RGBA bits at {}x{}", bad, width, height);
}

#[test]
fn main_rs_frame_64() {
    differential(64, 64);            // main.rs:108-110
}

#[test]
fn reference_size_640x480() {
    differential(640, 480);          // main.rs:104-106
}
