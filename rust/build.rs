// build.rs — builds libptcore with nvcc for sm_100a and links it (text only: no Rust toolchain in this image).
// Appended to the reference's own build.rs (which copies res/ and preprocesses shaders, build.rs:4-61).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("ptcore");       // this repository, vendored
    let csrc = root.join("wgpu-pathtracer_b200/csrc");
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    let common = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                  "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off"];
    let mut objs = vec![];
    for src in ["pt_kernels.cu", "pt_step.cu", "pt_host.cpp", "bvh_build.cpp", "scene_prep.cpp"] {
        let obj = out.join(src).with_extension("o");
        let st = Command::new("nvcc").args(arch).args(common).arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj).status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {}", src);
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libptcore.so");
    let st = Command::new("nvcc").args(arch).arg("-shared").arg("-o").arg(&lib).args(&objs).args(["-cudart", "static", "-lpthread"]).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptcore");
}
