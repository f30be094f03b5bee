//! dump_golden — the REFERENCE's own renderer as a golden-vector generator.
//!
//!     cargo +nightly run --release --bin dump_golden -- ../raytrace/teapot_tri.obj out_dir
//!
//! Renders main.rs's scene (main.rs:116-173) with `DefaultRayCaster` (the octree path, 1 thread, `debug_en = true`) at
//! 64x64 — main.rs's own size — and 640x480, with the shipped and the deterministic materials, and writes per frame
//!   golden_{W}x{H}_{shipped|det}.csv    the reference's debug CSV (debug.rs:118-140): Pixel_x(row);Pixel_y(col);ray;tri_hit;hit_t
//!   golden_{W}x{H}_{shipped|det}.rgba   the frame, W*H*4 little-endian f32 (`Color` lanes 0..3)
//! `tools/compare_golden.py out_dir` then checks these against this repository's committed golden vectors
//! (tests/golden/main_scene_64.npz), its CPU oracle and — on a GPU box — the CUDA path: the step that turns
//! "bit-exact to our restatement" into "bit-exact to the reference".  Rust prints f32 with the shortest digits that
//! round-trip, so the CSV carries hit_t exactly.
//!
//! NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Rust toolchain); needs no GPU and no librtb to run — build it with
//! `--no-default-features` if librtb cannot be built on the machine at hand.
use std::fs;
use std::io::Write;

use b200_raytrace_lib::scene::{main_scene, main_tris, main_viewport};
use raytrace_lib::raytrace::{make_vec, DefaultRayCaster, RayCaster};

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let obj = args.get(1).map(String::as_str).unwrap_or("../raytrace/teapot_tri.obj");
    let out = args.get(2).map(String::as_str).unwrap_or("golden_out");
    fs::create_dir_all(out).unwrap();
    // a pixel whose ray never entered an octree leaf has an empty check_tris list, and RayDebugCtx's Display unwraps a
    // `reduce` over it (debug.rs:29): such rows are written as misses instead of taking the process down
    std::panic::set_hook(Box::new(|_| {}));
    for &(w, h) in &[(64u32, 64u32), (640u32, 480u32)] {
        for &det in &[false, true] {
            let tag = if det { "det" } else { "shipped" };
            let s = main_scene(main_tris(obj, det), true, true);
            let v = main_viewport(w, h, 5, 1);
            let mut data = vec![make_vec(&[0., 0., 0.]); (w * h) as usize];
            let ctx = DefaultRayCaster {}.walk_rays(&v, &s, &mut data, 1, false);
            ctx.print_stats();
            let mut f = fs::File::create(format!("{}/golden_{}x{}_{}.rgba", out, w, h, tag)).unwrap();
            for c in &data {
                for k in 0..4 {
                    f.write_all(&c.v[k].to_le_bytes()).unwrap();
                }
            }
            let mut csv = fs::File::create(format!("{}/golden_{}x{}_{}.csv", out, w, h, tag)).unwrap();
            let dbg = s.debug_ctx.lock().unwrap();
            dbg.write_debug_header(&mut csv);
            for row in 0..h as usize {
                for col in 0..w as usize {
                    let mut line: Vec<u8> = Vec::new();
                    let ok = std::panic::catch_unwind(std::panic::AssertUnwindSafe(|| {
                        let mut l: Vec<u8> = Vec::new();
                        dbg.write_px_debug_context((row, col), &mut l);
                        l
                    }));
                    match ok {
                        Ok(l) if !l.is_empty() => line = l,
                        _ => writeln!(&mut line, "{};{};0,0,0;0,0,0;0;0;", row, col).unwrap(),
                    }
                    csv.write_all(&line).unwrap();
                }
            }
            println!("wrote {}/golden_{}x{}_{}.{{csv,rgba}}", out, w, h, tag);
        }
    }
}
