// build.rs -- compiles the hand-written CUDA of ../../audio_pattern_discovery_b200/csrc for
// sm_100a into libapd_b200.so and links it.  The source lists and flags below are the ones of
// audio_pattern_discovery_b200/build.py (the script that is exercised in this repository);
// tests/test_rust_sources.py parses both files and fails if they drift apart.
use std::env;
use std::path::PathBuf;
use std::process::Command;

const DPADS: [u32; 8] = [4, 8, 12, 16, 20, 24, 28, 32];
// translation units besides dtw_inst.cu (compiled once per padded frame width)
const CUDA_UNITS: [&str; 4] = ["apd_api", "pair_path", "percentile", "ae_encode"];
const CXX_UNITS: [&str; 3] = ["host_plan", "upgma", "matrix_io"];
const NVCC_FLAGS: [&str; 5] = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fno-fast-math,-ffp-contract=off"];
const CXX_FLAGS: [&str; 8] = ["-O2", "-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math", "-Wall",
                              "-Wno-unknown-pragmas"];

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
    let mut objs = vec![];
    for d in DPADS.iter() {
        let o = out.join(format!("dtw_inst_{}.o", d));
        run(Command::new(&nvcc).args(&arch).args(&NVCC_FLAGS).arg(format!("-DAPD_DPAD={}", d))
            .arg("-c").arg("-o").arg(&o).arg(csrc.join("dtw_inst.cu")));
        objs.push(o);
    }
    for name in CUDA_UNITS.iter() {
        let o = out.join(format!("{}.o", name));
        run(Command::new(&nvcc).args(&arch).args(&NVCC_FLAGS).arg("-c").arg("-o").arg(&o)
            .arg(csrc.join(format!("{}.cu", name))));
        objs.push(o);
    }
    for name in CXX_UNITS.iter() {
        let o = out.join(format!("{}.o", name));
        run(Command::new("g++").args(&CXX_FLAGS).arg("-c").arg("-o").arg(&o).arg(csrc.join(format!("{}.cpp", name))));
        objs.push(o);
    }
    let lib = out.join("libapd_b200.so");
    run(Command::new(&nvcc).args(&arch).arg("-shared").arg("-o").arg(&lib).args(&objs)
        .args(&["-Xcompiler", "-fPIC", "-cudart", "static"]));
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=apd_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", manifest.join("../../include/apd.h").display());
}
