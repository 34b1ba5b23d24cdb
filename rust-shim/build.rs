// Links libezkvm.so (built by `python -m encrypt_zkvm_b200.build`: nvcc -gencode arch=compute_100a,code=sm_100a).
fn main() {
    let dir = std::env::var("EZKVM_LIB_DIR").expect("set EZKVM_LIB_DIR to the directory holding libezkvm.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=ezkvm");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=EZKVM_LIB_DIR");
}
