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

/// main.rs's scene and camera, shared by the binaries and the parity test (needs no GPU).
pub mod scene;

/// The caster and the FFI (links librtb, built by build.rs with nvcc).  `--no-default-features` leaves it out, so that
/// `dump_golden` — which only runs the reference's own CPU renderer — builds on a machine without CUDA.
#[cfg(feature = "gpu")]
mod caster;
#[cfg(feature = "gpu")]
pub use caster::*;
