// Points rustc at the in-tree libwpt.so (wasm_pathtracer_b200/libwpt.so).
fn main() {
    let dir = std::env::var("WPT_LIB_DIR").unwrap_or_else(|_| "../../../wasm_pathtracer_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=wpt");
}
