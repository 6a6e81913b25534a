// build.rs — tells cargo where libabfit.so lives.  ABFIT_LIB_DIR = directory holding libabfit.so
// (alphabeta-rs_b200/ of this repository after `make -C alphabeta-rs_b200`).  UNTESTED (no Rust toolchain here).
use std::env;
use std::path::PathBuf;

fn main() {
    println!("cargo:rerun-if-env-changed=ABFIT_LIB_DIR");
    let dir = env::var("ABFIT_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        // default: the in-tree build next to this crate
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../alphabeta-rs_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=abfit");
    // libabfit.so links the CUDA runtime statically and loads NVRTC with dlopen at first use; nothing else to link.
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
