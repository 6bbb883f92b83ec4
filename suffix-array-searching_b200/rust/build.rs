// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links the shared library.
// No other architecture, no CPU fallback.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = root.join("../csrc");
    let include = root.join("../../include");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let lib = out.join("libsst_b200.so");
    let mut cmd = Command::new(&nvcc);
    cmd.args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "--expt-relaxed-constexpr",
              "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(format!("-I{}", include.display()));
    // every csrc/*.cu, in a stable order (the same set csrc/Makefile builds with $(wildcard *.cu);
    // tests/test_abi.py::test_build_rs_compiles_every_source checks that no file list is hard-coded here)
    let mut sources: Vec<PathBuf> = std::fs::read_dir(&csrc)
        .expect("csrc/ not found")
        .filter_map(|e| e.ok().map(|e| e.path()))
        .filter(|p| p.extension().map_or(false, |x| x == "cu"))
        .collect();
    sources.sort();
    assert!(!sources.is_empty(), "no CUDA sources under {}", csrc.display());
    for f in &sources {
        cmd.arg(f);
        println!("cargo:rerun-if-changed={}", f.display());
    }
    println!("cargo:rerun-if-changed={}", csrc.join("common.cuh").display());
    println!("cargo:rerun-if-changed={}", include.join("sst_b200.h").display());
    cmd.arg("-lpthread");
    let status = cmd.status().expect("failed to run nvcc");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=sst_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", out.display());
}
