fn main() {
    // point cargo at instant-distance_b200/lib (built by `make -C instant-distance_b200/csrc`)
    let dir = std::env::var("IDB_LIB_DIR").unwrap_or_else(|_| "../lib".into());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=instant_distance_b200");
}
