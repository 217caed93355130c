// Links libfeotensor_b200.so (built by `make -C feotensor_b200/csrc`); point FEOTENSOR_B200_LIB_DIR at feotensor_b200/lib.
fn main() {
    let dir = std::env::var("FEOTENSOR_B200_LIB_DIR").unwrap_or_else(|_| "../feotensor_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=feotensor_b200");
    println!("cargo:rerun-if-env-changed=FEOTENSOR_B200_LIB_DIR");
}
