// hephaestus-jit/build.rs — links libhj_b200.so (built by `make -C hephaestus-jit_b200`, nvcc
// -gencode arch=compute_100a,code=sm_100a).  NVRTC comes from the CUDA toolkit, NCCL is dlopen'ed
// by the library at first use, cudart is linked statically into it.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no cargo / rustc): shipped as the source a
// maintainer drops into the reference tree; the same ABI is exercised by the ctypes mirror and tests/.
fn main() {
    let dir = std::env::var("HJ_B200_LIB_DIR").expect("set HJ_B200_LIB_DIR to the directory of libhj_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=hj_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=HJ_B200_LIB_DIR");
}
