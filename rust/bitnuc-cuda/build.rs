// build.rs -- compiles the hand-written CUDA kernels with `nvcc -arch=sm_100a` and links them.
// Mirrors bitnuc_b200/build.py (the build that is exercised in this repository).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("bitnuc_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let sources = ["api.cu", "codec.cu", "kmer.cu", "hamming.cu", "counts.cu", "batch.cu", "split.cu", "gather.cu", "windows.cu", "fastq.cu", "synth.cu", "multi.cu"];
    let mut objects = Vec::new();
    for src in sources {
        let obj = out.join(format!("{src}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found: bitnuc-cuda has no CPU fallback and cannot build without the CUDA toolkit");
        assert!(status.success(), "nvcc failed on {src}");
        objects.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libbitnuc_cuda_kernels.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().expect("ar");
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=bitnuc_cuda_kernels");
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    println!("cargo:rustc-link-search=native={cuda}/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rustc-link-lib=dl"); // api.cu resolves cuCtxGetCurrent through dlopen/dlsym
    println!("cargo:rerun-if-changed={}", root.join("include/bitnuc_cuda.h").display());
}
