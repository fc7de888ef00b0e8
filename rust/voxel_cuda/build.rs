// build.rs -- compiles the CUDA sources of voxel_game_b200/csrc with nvcc for sm_100a and links the
// result.  NOT RUN IN THE BUILD IMAGE (no cargo there); it is the Rust twin of
// voxel_game_b200/csrc/Makefile, which is what the repo's own build() uses.
//
// -fmad=false is load-bearing: bit-exact block IDs need every f64 multiply and add to round
// separately, like the reference's libnoise (rustc never contracts a*b+c).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("voxel_game_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["vx_api.cu", "vx_gen.cu", "vx_mesh.cu", "vx_edit.cu", "vx_light.cu", "vx_tables.cpp"];
    let lib = out.join("libvoxel_cuda.so");
    let status = Command::new(&nvcc)
        .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
               "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-o"])
        .arg(&lib)
        .args(sources.iter().map(|s| csrc.join(s)))
        .status()
        .expect("nvcc not found: the chunk pipeline has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    for s in sources {
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/voxel_cuda.h").display());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=voxel_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
