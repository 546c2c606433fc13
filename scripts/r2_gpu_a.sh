# round 2, GPU call: failing tests with tracebacks, bench line, photon warm-up trace
timeout -k 10 600 python -m pytest "tests/test_baseline_shapes.py::test_config3_bvh4_with_pnee_300k_photons" tests/test_baseline_shapes.py::test_config4_museum_adaptive_4k_two_rounds tests/test_reference_images.py tests/test_gpu_parity.py tests/test_gpu_features.py -m gpu -x -q 2>&1 | tail -60 > gpurun_out/r2_t3.log
timeout -k 10 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
WPT_TRACE_PHOTONS=1 timeout -k 10 100 python scripts/run_configs.py 3 > gpurun_out/r2_cfg3.log 2>&1
