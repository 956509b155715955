// Links libplonky2_b200.so (built by `python eth-lc-plonky2_b200/build.py` with nvcc for sm_100a).
// PLONKY2_B200_LIB_DIR = the directory holding the library (default: ../../eth-lc-plonky2_b200 relative to this crate).
use std::path::PathBuf;

fn main() {
    let dir = std::env::var("PLONKY2_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../eth-lc-plonky2_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=plonky2_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=PLONKY2_B200_LIB_DIR");
}
