// build.rs — builds libptcore with nvcc for sm_100a and links it (text only: no Rust toolchain in this image).
// Appended to the reference's own build.rs (which copies res/ and preprocesses shaders, build.rs:4-61).
// The source list and the link line mirror wgpu-pathtracer_b200/csrc/Makefile (OBJS / the libptcore.so rule);
// tests/test_rust_boundary.py fails when the two drift apart.
use std::{env, path::PathBuf, process::Command};

const CUDA_SOURCES: [&str; 4] = ["pt_kernels.cu", "pt_step.cu", "bvh_gpu.cu", "img_decode.cu"];
const HOST_SOURCES: [&str; 3] = ["pt_host.cpp", "bvh_build.cpp", "scene_prep.cpp"];
const LINK_LIBS: [&str; 3] = ["-lpthread", "-ldl", "-lz"];      // -ldl: NCCL (ptc_reduce) and nvJPEG (JPEG textures) are bound with dlopen / dlsym; -lz: the PNG decoder inflates with zlib

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("ptcore");       // this repository, vendored
    let csrc = root.join("wgpu-pathtracer_b200/csrc");
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    // -fmad=false + -ffp-contract=off: device and host arithmetic must round like the CPU oracle
    let common = ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
                  "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off"];
    let mut objs = vec![];
    for src in CUDA_SOURCES.iter().chain(HOST_SOURCES.iter()) {
        let obj = out.join(src).with_extension("o");
        let mut cmd = Command::new("nvcc");
        cmd.args(arch).args(common);
        if *src == "scene_prep.cpp" { cmd.args(["-Xcompiler", "-mfma"]); }                    // explicit fmaf must be one instruction on the host too
        let st = cmd.arg("-c").arg(csrc.join(src)).arg("-o").arg(&obj).status().expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {}", src);
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    for hdr in ["img_decode.h", "pt_device.cuh", "pt_kernels.cuh", "bvh_build.h", "bvh_gpu.h", "lbvh.cuh", "scene_prep.h", "pt_error.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(hdr).display());
    }
    let lib = out.join("libptcore.so");
    let st = Command::new("nvcc").args(arch).arg("-shared").arg("-o").arg(&lib).args(&objs)
        .args(["-cudart", "static", "-Xlinker", "--no-undefined"]).args(LINK_LIBS).status().unwrap();
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=ptcore");
}
