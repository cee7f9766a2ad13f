# Call J: e4m3 decode attention with page pairs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fp8_kv_gpu.py tests/test_kernels_gpu.py -q -x -k "paged or fp8 or rope" > gpurun_out/j_tests.log 2>&1
echo "tests rc=$?"
tail -n 6 gpurun_out/j_tests.log
timeout 120 python tools/attn_fp8_bench.py 2>&1 | tail -n 1
timeout 120 python tools/attn_fp8_bench.py 64 4096 2>&1 | tail -n 1
timeout 120 python tools/attn_fp8_bench.py 8 4096 2>&1 | tail -n 1
timeout 120 python tools/attn_fp8_bench.py 32 1024 2>&1 | tail -n 1
