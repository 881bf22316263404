// Link against the prebuilt libperceive_cuda.so.  PERCEIVE_CUDA_LIB_DIR points at the directory
// holding it (perceive_b200/ in this repository after `python -m perceive_b200._build`).
fn main() {
    let dir = std::env::var("PERCEIVE_CUDA_LIB_DIR").expect("set PERCEIVE_CUDA_LIB_DIR to the directory of libperceive_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=perceive_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PERCEIVE_CUDA_LIB_DIR");
}
