// build.rs — builds librtb (the C-ABI CUDA library of include/rtb.h) with nvcc for sm_100a and links it.
//
// Replaces cuda_raytrace_lib/build.rs:3-23 (cxx bridge, sm_75, -G -O0).  No cxx, no C++ ABI: the kernels are
// compiled straight from the .cu files by nvcc and exported as plain `extern "C"`.
//
//   RTB_ROOT   path of the b200 repository (default: three levels above this crate, i.e. the checkout that
//              contains include/rtb.h and rust_raytrace_b200/csrc/)
//   NVCC       nvcc binary (default: nvcc on PATH)
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    if env::var("CARGO_FEATURE_GPU").is_err() {
        return;                       // --no-default-features: nothing to compile or link (dump_golden only)
    }
    let crate_dir = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let root = env::var("RTB_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| crate_dir.join("../../..").canonicalize().expect("RTB_ROOT"));
    let csrc = root.join("rust_raytrace_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());

    let units = ["rtb_api.cu", "rtb_lbvh.cu", "rtb_scene.cu", "rtb_ext.cu", "rtb_trace.cu", "rtb_wavefront.cu", "host/raytrace_host.cpp"];
    let mut objs = Vec::new();
    for u in units {
        let src = csrc.join(u);
        let obj = out.join(PathBuf::from(u).file_stem().unwrap()).with_extension("o");
        // -fmad is left at its default: every value the reference also computes goes through __f*_rn
        // intrinsics in the kernels (never contracted); only the conservative BVH slab test uses FMA.
        let st = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math"])
            .arg(format!("-I{}", root.join("include").display()))
            .arg(format!("-I{}", csrc.display()))
            .args(["-x", "cu", "-c"])
            .arg(&src)
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("failed to run nvcc");
        assert!(st.success(), "nvcc failed on {}", src.display());
        println!("cargo:rerun-if-changed={}", src.display());
        objs.push(obj);
    }
    let lib = out.join("librtb.a");
    let _ = std::fs::remove_file(&lib);
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar");
    assert!(st.success());

    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=rtb");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rustc-link-lib=dylib=pthread");
    println!("cargo:rerun-if-changed={}", root.join("include/rtb.h").display());
    println!("cargo:rerun-if-env-changed=RTB_ROOT");
}
