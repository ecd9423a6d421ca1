// build.rs -- replaces the OpenCL search-path / link lines of the reference's build script.
// SWB200_ROOT must point at a checkout of the swb200 repository (this one).
use std::process::Command;

fn main() {
    let root = std::env::var("SWB200_ROOT").expect("set SWB200_ROOT to the swb200 checkout");
    let status = Command::new("make")
        .args(["-C", &root, "mini_parallel_b200/libswb200.so"])
        .status()
        .expect("failed to run make (nvcc, CUDA 12.9+, sm_100a)");
    assert!(status.success(), "nvcc build of libswb200.so failed");
    println!("cargo:rustc-link-search=native={}/mini_parallel_b200", root);
    println!("cargo:rustc-link-lib=dylib=swb200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}/mini_parallel_b200", root);
    println!("cargo:rerun-if-changed={}/mini_parallel_b200/csrc", root);
    println!("cargo:rerun-if-changed={}/include/swb200.h", root);
    println!("cargo:rerun-if-env-changed=SWB200_ROOT");
}
