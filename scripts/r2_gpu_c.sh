O=gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_features.py tests/test_baseline_shapes.py tests/test_render_tool.py tests/test_extension_scene.py -m gpu -q 2>&1 | tail -8 > $O/r2_t4.log
WPT_TRACE_ROUNDS=1 timeout -k 5 120 python scripts/target_trace.py > $O/r2_target_trace2.log 2>&1
timeout -k 10 300 python bench.py --steps 5 --warmup 3 --no-cpu > $O/r2_bench_c.json 2> $O/r2_bench_c.err
