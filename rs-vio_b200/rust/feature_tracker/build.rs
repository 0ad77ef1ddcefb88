//! build.rs -- links (or, with `--features build-cuda`, compiles) librsvio_b200.so.
//! north_star: "a new build.rs compiles hand-written .cu kernels for sm_100a".
use std::{env, path::PathBuf, process::Command};

fn main() {
    let crate_dir = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let pkg_root = crate_dir.join("../..").canonicalize().expect("rs-vio_b200/ not found");
    let lib_dir = env::var("RSVIO_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| pkg_root.join("lib"));

    if env::var("CARGO_FEATURE_BUILD_CUDA").is_ok() {
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
        let csrc = pkg_root.join("csrc");
        let out = PathBuf::from(env::var("OUT_DIR").unwrap());
        let mut objs = Vec::new();
        for f in ["pyramid.cu", "lk.cu", "detect.cu", "tracks.cu", "convert.cu", "trackb.cu", "detectb.cu", "microbench.cu", "api.cu", "apib.cu"] {
            let obj = out.join(f.replace(".cu", ".o"));
            // same flags as rs-vio_b200/Makefile: sm_100a only, no FMA contraction (bit-exact pyramid)
            let st = Command::new(&nvcc)
                .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
                       "-Xcompiler", "-fPIC", "-c"])
                .arg(csrc.join(f)).arg("-o").arg(&obj).status().expect("nvcc not found");
            assert!(st.success(), "nvcc failed on {f}");
            objs.push(obj);
            println!("cargo:rerun-if-changed={}", csrc.join(f).display());
        }
        let so = out.join("librsvio_b200.so");
        let st = Command::new(&nvcc).args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-o"])
            .arg(&so).args(&objs).status().expect("nvcc not found");
        assert!(st.success(), "nvcc link failed");
        println!("cargo:rustc-link-search=native={}", out.display());
    } else {
        println!("cargo:rustc-link-search=native={}", lib_dir.display());
    }
    println!("cargo:rustc-link-lib=dylib=rsvio_b200");
    println!("cargo:rerun-if-env-changed=RSVIO_B200_LIB_DIR");
}
