// UNBUILT here (no Rust toolchain in the image).  Compiles the CUDA sources of ../lp_b200/csrc with
// nvcc for sm_100a into liblpb200.so and links the crate against it: the "thin extern "C" FFI built
// by build.rs/nvcc" of the north star.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let csrc = root.join("lp_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("liblpb200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = ["vec_kernels.cu", "dmma_gemm.cu", "cholesky.cu", "batched.cu", "presolve.cu", "lpb_api.cu"];
    let mut cmd = Command::new(nvcc);
    cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-shared",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-o"]).arg(&lib);
    for s in sources {
        cmd.arg(csrc.join(s));
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    cmd.arg("-lnccl");
    assert!(cmd.status().expect("nvcc not found").success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=lpb200");
    println!("cargo:rerun-if-changed={}", root.join("include").join("lpb200.h").display());
}
