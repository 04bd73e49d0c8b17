// build.rs -- compiles the sm_100a kernels with nvcc and links them into the dawnsearch binary.
// UNCOMPILED IN THIS REPO: the build image has no cargo/rustc.  It runs the same nvcc command
// as dawnsearch_b200/csrc/Makefile (which IS exercised by __graft_entry__.build()).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from("dawnsearch_b200/csrc");
    let srcs = ["dawn_index.cu", "scan_topk.cu", "finalize.cu", "ingest.cu"];
    let mut objs = Vec::new();
    for s in srcs {
        let obj = out.join(s.replace(".cu", ".o"));
        let ok = Command::new("nvcc")
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s)).arg("-o").arg(&obj)
            .status().expect("nvcc not found").success();
        assert!(ok, "nvcc failed on {s}");          // no CPU fallback: a failed CUDA build fails the crate
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
        objs.push(obj);
    }
    let lib = out.join("libdawn_b200.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=dawn_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
}
