// utils/build.rs -- replacement for /root/reference/utils/build.rs:1-19.
// The reference compiles the spqlios FFT (3 .cpp + 2 .s) with the `cc` crate into libspqlios.a.  This version compiles
// the CUDA engine with nvcc for sm_100a into librustfhe_b200.a and links cudart.  NOT BUILT IN THIS REPOSITORY'S
// IMAGE (no cargo/rustc); kept thin and mechanical on purpose.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("TFHE_B200_CSRC").unwrap_or_else(|_| "../../rustfhe_b200/csrc".into()));
    let obj = out.join("engine.o");
    let keys = out.join("hostkeys.o");
    let wire = out.join("wire.o");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    for (src, dst) in [("engine.cu", &obj), ("hostkeys.cpp", &keys), ("wire.cpp", &wire)] {
        let ok = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(src))
            .arg("-o")
            .arg(dst)
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {src}");
    }
    let lib = out.join("librustfhe_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).arg(&obj).arg(&keys).arg(&wire).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=rustfhe_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
}
