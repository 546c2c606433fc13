"""wasm_pathtracer_b200 — B200-native path-tracing core behind the reference's
`wasm_interface.rs` surface (sourcedennis/wasm-pathtracer).

The product is `libwpt.so` (hand-written sm_100a CUDA + C++ host code, C ABI declared in
`include/wpt.h`). This package is the thin Python host layer over that ABI:

* `PathTracer`   — the reference's exported functions with the same names and arguments
                   (`src/wasm_interface.rs:65-384`), plus the additions of `wpt.h`.
* `build`        — in-tree nvcc build of `libwpt.so`.
* `dist`         — one-process-per-GPU row partitioning over `torch.distributed`.

There is no CPU rendering path: `PathTracer` raises if the library or a CUDA device is missing.
"""
from .api import (PathTracer, WptError, WptConfig, load_library, library_path, parse_obj,
                  NO_NEE, NORMAL_NEE, PNEE, SCENE_MUSEUM, SCENE_BUNNY, SCENE_EXT_WHITTED, CAM_MUSEUM, CAM_BUNNY, CAM_WHITTED, DEVICE_NONE)
from .build import build_library

__all__ = ["PathTracer", "WptError", "WptConfig", "load_library", "library_path", "parse_obj", "build_library",
           "NO_NEE", "NORMAL_NEE", "PNEE", "SCENE_MUSEUM", "SCENE_BUNNY", "SCENE_EXT_WHITTED", "CAM_MUSEUM", "CAM_BUNNY", "CAM_WHITTED", "DEVICE_NONE"]
