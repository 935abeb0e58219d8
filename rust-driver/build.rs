// Builds the CUDA library in-tree and links it (replaces the reference's 4-line build.rs that only sets
// MACOSX_DEPLOYMENT_TARGET).  NOT compiled in this environment: no Rust toolchain is available.
use std::process::Command;

fn main() {
    let dir = std::path::Path::new(env!("CARGO_MANIFEST_DIR")).join("../mirror_maze_b200");
    let status = Command::new("make").arg("-C").arg(&dir).arg("libmirror_maze_cuda.so").status().expect("make not found");
    assert!(status.success(), "building libmirror_maze_cuda.so failed");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=mirror_maze_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-changed=../mirror_maze_b200/csrc");
    println!("cargo:rerun-if-changed=../include/mirror_maze_cuda.h");
}
