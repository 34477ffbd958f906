// Links against libspgpu.so (built by `make` at the root of the backend repository; CUDA 12.9, sm_100a).
// SPGPU_LIB_DIR names the directory that holds it; at run time the loader finds it through the rpath
// written here (or LD_LIBRARY_PATH).
fn main() {
    println!("cargo:rerun-if-env-changed=SPGPU_LIB_DIR");
    let dir = std::env::var("SPGPU_LIB_DIR")
        .expect("set SPGPU_LIB_DIR to the directory that holds libspgpu.so (spartan_parallel_b200/ in the backend repository)");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=spgpu");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
