// build.rs -- compiles the hand-written CUDA of ../../audio_pattern_discovery_b200/csrc for
// sm_100a into libapd_b200.so and links it.  Mirrors audio_pattern_discovery_b200/build.py
// (the script that is actually exercised in this repository).
use std::env;
use std::path::PathBuf;
use std::process::Command;

const DPADS: [u32; 8] = [4, 8, 12, 16, 20, 24, 28, 32];

fn run(cmd: &mut Command) {
    let status = cmd.status().expect("failed to spawn build tool");
    assert!(status.success(), "build step failed: {:?}", cmd);
}

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../../audio_pattern_discovery_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let arch = ["-gencode", "arch=compute_100a,code=sm_100a"];
    let flags = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off"];
    let mut objs = vec![];
    for d in DPADS.iter() {
        let o = out.join(format!("dtw_inst_{}.o", d));
        run(Command::new(&nvcc).args(&arch).args(&flags).arg(format!("-DAPD_DPAD={}", d))
            .arg("-c").arg("-o").arg(&o).arg(csrc.join("dtw_inst.cu")));
        objs.push(o);
    }
    for name in ["apd_api", "pair_path"].iter() {
        let o = out.join(format!("{}.o", name));
        run(Command::new(&nvcc).args(&arch).args(&flags).arg("-c").arg("-o").arg(&o)
            .arg(csrc.join(format!("{}.cu", name))));
        objs.push(o);
    }
    let o = out.join("host_plan.o");
    run(Command::new("g++").args(&["-O2", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-c", "-o"])
        .arg(&o).arg(csrc.join("host_plan.cpp")));
    objs.push(o);
    let lib = out.join("libapd_b200.so");
    run(Command::new(&nvcc).args(&arch).arg("-shared").arg("-o").arg(&lib).args(&objs)
        .args(&["-Xcompiler", "-fPIC", "-cudart", "static"]));
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=apd_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/apd.h").display());
}
