// Links libcollision_b200.so (built by collision_rs_b200/csrc/Makefile).
// Point CB2_LIB_DIR at the directory that holds it.
fn main() {
    if let Ok(dir) = std::env::var("CB2_LIB_DIR") {
        println!("cargo:rustc-link-search=native={}", dir);
    }
    println!("cargo:rustc-link-lib=dylib=collision_b200");
    println!("cargo:rerun-if-env-changed=CB2_LIB_DIR");
}
