// Links libfastlanes_b200.so (built by `make lib` at the repository root).
fn main() {
    let dir = std::env::var("FASTLANES_B200_LIB_DIR").unwrap_or_else(|_| "../../fastlanes_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fastlanes_b200");
    println!("cargo:rerun-if-env-changed=FASTLANES_B200_LIB_DIR");
}
