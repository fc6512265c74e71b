// Links libfsgpu.so (built by `make -C frankensearch_b200/csrc`); FSGPU_LIB_DIR points at its directory.
fn main() {
    if let Ok(dir) = std::env::var("FSGPU_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
    }
    println!("cargo:rustc-link-lib=dylib=fsgpu");
    println!("cargo:rerun-if-env-changed=FSGPU_LIB_DIR");
}
