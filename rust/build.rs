// rayrs-lib/build.rs — link librayrs_b200.so (built by `python -m rayrs_b200.build` in the rayrs-b200 repository).
// RAYRS_B200_LIB names the directory that holds the library; it is also put on the binary's rpath.
fn main() {
    let dir = std::env::var("RAYRS_B200_LIB").expect("set RAYRS_B200_LIB to the directory of librayrs_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rayrs_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=RAYRS_B200_LIB");
}
