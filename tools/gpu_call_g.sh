# Call G: third session.  New: prefix-cache engine loop, FP8 KV cache (kernel + runner), bench fp8 leg.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fp8_kv_gpu.py tests/test_runner_gpu.py::test_engine_loop_with_prefix_cache -q -x -s > gpurun_out/g_new_tests.log 2>&1
echo "new tests rc=$?"
tail -n 30 gpurun_out/g_new_tests.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/g_all_tests.log 2>&1
echo "all gpu tests rc=$?"
tail -n 15 gpurun_out/g_all_tests.log
B200_BENCH_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?"
tail -n 5 gpurun_out/g_bench.err
cat gpurun_out/g_bench.json
