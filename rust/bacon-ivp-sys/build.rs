// UNVERIFIED (no Rust toolchain in the build image).
// Compiles the hand-written .cu kernels with nvcc for sm_100a through the repo's Makefile and links the result.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("bacon_b200/csrc");
    // make -C bacon_b200/csrc  ==  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo ... -shared
    let status = Command::new("make").arg("-C").arg(&csrc).arg("-j8").status().expect("failed to run make/nvcc");
    assert!(status.success(), "nvcc build of libbacon_ivp.so failed");
    let libdir = root.join("bacon_b200");
    println!("cargo:rustc-link-search=native={}", libdir.display());
    println!("cargo:rustc-link-lib=dylib=bacon_ivp");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", libdir.display());
    for f in ["engine.cu", "rhs_builtin.cu", "rk_fast.cuh", "rk_strict.cuh", "bdf.cuh", "drive.cuh", "hist_stage.cuh",
              "rk_warp_linear.cuh", "tableaux.cuh", "launch.cuh", "ivp_common.cuh"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    println!("cargo:rerun-if-changed={}", root.join("include/bacon_ivp.h").display());
}
