// build.rs — compiles the CUDA side of the crate for sm_100a and links it.
// NOT BUILT IN THIS REPOSITORY'S CI: the image has no Rust toolchain.
// Place next to Cargo.toml of box2d-rs with `box2d_rs_b200/csrc` and `include/` vendored under `b2gpu/`.
use std::process::Command;

fn main() {
    let out = std::env::var("OUT_DIR").unwrap();
    let sources = ["b2g_api.cu", "b2g_runtime.cu", "b2g_world.cu", "b2g_checkpoint.cu", "b2g_events.cu"];
    let mut cmd = Command::new(std::env::var("NVCC").unwrap_or_else(|_| "nvcc".into()));
    // --fmad=false is mandatory: rustc never contracts a*b+c, and pair sets are downstream of the solver.
    cmd.args([
        "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "--fmad=false", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-shared", "-o",
    ])
    .arg(format!("{out}/libb2gpu.so"));
    for s in sources {
        cmd.arg(format!("b2gpu/csrc/{s}"));
        println!("cargo:rerun-if-changed=b2gpu/csrc/{s}");
    }
    println!("cargo:rerun-if-changed=b2gpu/include/b2gpu.h");
    let status = cmd.status().expect("nvcc not found: the GPU step engine needs the CUDA toolkit");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={out}");
    println!("cargo:rustc-link-lib=dylib=b2gpu");
}
