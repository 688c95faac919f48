fn main() {
    // libaudioflow_gpu.so is built by `make -C audio-flow-rs_b200` (nvcc, sm_100a)
    let dir = std::env::var("AUDIOFLOW_GPU_LIB_DIR").unwrap_or_else(|_| "../lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=audioflow_gpu");
    println!("cargo:rerun-if-env-changed=AUDIOFLOW_GPU_LIB_DIR");
}
