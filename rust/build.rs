// build.rs -- compiles the CUDA sources for sm_100a and links them (plus cudart) into the crate.
// Mirrors veils_b200/build.py; set VEILS_CUDA_PREBUILT=/path/to/dir to link an existing
// libveils_cuda.so instead of invoking nvcc.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    if let Ok(dir) = env::var("VEILS_CUDA_PREBUILT") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=veils_cuda");
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("veils_b200").join("csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libveils_cuda.so");
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared", "-o"]).arg(&lib);
    for f in ["api.cu", "plan.cpp", "dft_generic.cu", "pow2_tile.cu", "pow2_pk.cu", "pad.cu", "detrend.cu", "wide_fwd.cu",
              "wide_inv.cu", "smooth.cu"] {
        cmd.arg(csrc.join(f));
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=veils_cuda");
}
