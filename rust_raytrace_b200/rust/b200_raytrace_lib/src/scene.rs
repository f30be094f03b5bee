//! main.rs's benchmark scene and camera (raytrace/src/main.rs:116-173), shared by the binaries and the parity test of
//! this crate.  `deterministic` swaps in the materials of the deterministic parity mode: teapot
//! `Solid(252,119,0)` (the commented line main.rs:123), disks `Reflective` with scattering 0, disk sides `Solid` — no RNG
//! draw influences the image, so DefaultRayCaster (ThreadRng) and B200RayCaster (counter RNG) must agree bit for bit.
use raytrace_lib::debug::make_debug_ctx;
use raytrace_lib::obj_parser;
use raytrace_lib::raytrace::{self, make_color, make_disk, make_dummy_triangle, make_vec, populate_triangle_numbers, Scene,
                             SurfaceKind, Triangle, Viewport};

pub fn main_tris(obj_path: &str, deterministic: bool) -> Vec<Triangle> {
    let orange = make_color((252, 119, 0));
    let grey = make_color((230, 230, 230));
    let dark = make_color((40, 40, 40));
    let (teapot, d1, d2, side) = if deterministic {
        (SurfaceKind::Solid { color: orange },
         SurfaceKind::Reflective { scattering: 0.0, color: grey, alpha: 0.7 },
         SurfaceKind::Reflective { scattering: 0.0, color: grey, alpha: 0.7 },
         SurfaceKind::Solid { color: dark })
    } else {
        (SurfaceKind::Matte { color: orange, alpha: 0.2 },
         SurfaceKind::Reflective { scattering: 0.0002, color: grey, alpha: 0.7 },
         SurfaceKind::Reflective { scattering: 0.002, color: grey, alpha: 0.7 },
         SurfaceKind::Matte { color: dark, alpha: 0.2 })
    };
    let mut tris: Vec<Triangle> = Vec::new();
    tris.push(make_dummy_triangle());
    tris.extend(obj_parser::parse_obj(obj_path, &make_vec(&[0., 0.5, 5.]), 1.0,
                                      raytrace::create_transform(&make_vec(&[0., 0.3, 1.]).unit(), 270_f32.to_radians()),
                                      &teapot, 0.05));
    tris.extend(make_disk(&make_vec(&[4., 4., 7.]), &make_vec(&[-0.3, -0.55, -0.5]).unit(), 2., 0.1, 50, &d1, &side, -1.));
    tris.extend(make_disk(&make_vec(&[4., -3., 5.]), &make_vec(&[-0.5, 2.0, -0.5]).unit(), 1., 0.04, 50, &d2, &side, -1.));
    populate_triangle_numbers(&mut tris);
    tris
}

/// `Scene` with the octree of main.rs:160-164 (`with_octree`) or with the trivial single-box accelerator
/// (`build_trivial_bounding_box`, raytrace.rs:847-856): the B200 caster only reads the root cube.
pub fn main_scene(tris: Vec<Triangle>, with_octree: bool, debug_en: bool) -> Scene {
    let boxes = if with_octree {
        raytrace::build_bounding_box(&tris, &make_vec(&[0., 0., 20.1]), 20., 10, 19)
    } else {
        raytrace::build_trivial_bounding_box(&tris, &make_vec(&[0., 0., 20.1]), 20.)
    };
    Scene { tris, boxes, debug_ctx: make_debug_ctx().into(), debug_en }
}

/// main.rs:166-173 with `aspect = height / width` as in main.rs:96-110.
pub fn main_viewport(width: u32, height: u32, maxdepth: usize, samples: usize) -> Viewport {
    let aspect = height as f32 / width as f32;
    raytrace::create_viewport((width, height), (1., 1. * aspect), &make_vec(&[2., 0., 0.]), &make_vec(&[0., 0., 1.]).unit(),
                              90., 0_f32.to_radians(), maxdepth, samples)
}
