// Links the prebuilt C-ABI library (python -m chapterhouseqe_b200.build produces libchdb_gpu.so).
fn main() {
    let dir = std::env::var("CHDB_GPU_LIB_DIR").unwrap_or_else(|_| "../../chapterhouseqe_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=chdb_gpu");
    println!("cargo:rerun-if-env-changed=CHDB_GPU_LIB_DIR");
}
