// build.rs -- builds libvoxel_cuda.so for sm_100a by running voxel_game_b200/csrc/Makefile (the ONE
// place that lists the CUDA sources and the nvcc flags; the repo's own build() runs the same file)
// with the output redirected into OUT_DIR, and links the result.  NOT RUN IN THE BUILD IMAGE (no
// cargo there); tests/test_abi_and_host.py checks that this file, the Makefile and INTEGRATION.md
// name the same translation units.
//
// -fmad=false (in the Makefile) is load-bearing: bit-exact block IDs need every f64 multiply and
// add to round separately, like the reference's libnoise (rustc never contracts a*b+c).
use std::{env, fs, path::PathBuf, process::Command};

/// Words of a `NAME := a b \` ... Makefile assignment (continuation lines included).
fn make_list(makefile: &str, name: &str) -> Vec<String> {
    let mut out = Vec::new();
    let mut inside = false;
    for line in makefile.lines() {
        let body = if inside {
            line
        } else if let Some(rest) = line.strip_prefix(name) {
            match rest.trim_start().strip_prefix(":=") {
                Some(r) => r,
                None => continue,
            }
        } else {
            continue;
        };
        inside = body.trim_end().ends_with('\\');
        out.extend(body.trim_end().trim_end_matches('\\').split_whitespace().map(String::from));
        if !inside {
            break;
        }
    }
    out
}

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("voxel_game_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libvoxel_cuda.so");
    let mut make = Command::new("make");
    make.arg("-C").arg(&csrc).arg(format!("OUT={}", lib.display()));
    // undefined symbols (a translation unit missing from SRC) fail the build, not the first call
    // (-Xlinker --no-undefined is part of the Makefile's flags)
    if let Ok(nvcc) = env::var("NVCC") {
        make.arg(format!("NVCC={nvcc}"));
    }
    let status = make.status().expect("make/nvcc not found: the chunk pipeline has no CPU fallback");
    assert!(status.success(), "building libvoxel_cuda.so failed (see {}/build.log)", csrc.display());

    let makefile = fs::read_to_string(csrc.join("Makefile")).expect("csrc/Makefile");
    println!("cargo:rerun-if-changed={}", csrc.join("Makefile").display());
    for f in make_list(&makefile, "SRC").into_iter().chain(make_list(&makefile, "HDR")) {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=voxel_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
