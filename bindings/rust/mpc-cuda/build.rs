// Builds libmpc_cuda.so from the CUDA sources with nvcc for sm_100a and links it.
// Not compiled in the build container of this repository (no Rust toolchain there); the same nvcc
// recipe is exercised by zk-mpc_b200/build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../..");
    let csrc = root.join("zk-mpc_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    for entry in std::fs::read_dir(&csrc).unwrap() {
        let p = entry.unwrap().path();
        if p.extension().map_or(false, |e| e == "cu") {
            let o = out.join(p.file_stem().unwrap()).with_extension("o");
            let st = Command::new(&nvcc)
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                       "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-c"])
                .arg(&p).arg("-o").arg(&o).status().expect("nvcc");
            assert!(st.success(), "nvcc failed on {:?}", p);
            objs.push(o);
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
    let lib = out.join("libmpc_cuda.so");
    let st = Command::new(&nvcc).args(["-shared", "-o"]).arg(&lib).args(&objs)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC"]).status().expect("nvcc link");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=mpc_cuda");
    println!("cargo:rerun-if-changed={}", root.join("include/mpc_cuda.h").display());
}
