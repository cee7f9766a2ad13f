# Call P: decode GEMM with five dequant groups (side build g5) vs four
mkdir -p gpurun_out
B200_LIB_PATH=$PWD/build/g5/libb200_vllm.so timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_runner_gpu.py -q -x -k "wna16 or decode_steps" > gpurun_out/p_tests_g5.log 2>&1
echo "g5 tests rc=$?"; tail -n 4 gpurun_out/p_tests_g5.log
timeout 200 python tools/gemm_bench.py 32 2>&1 | tail -n 2 | sed "s/^/g4 /"
B200_LIB_PATH=$PWD/build/g5/libb200_vllm.so timeout 200 python tools/gemm_bench.py 32 2>&1 | tail -n 2 | sed "s/^/g5 /"
B200_LIB_PATH=$PWD/build/g5/libb200_vllm.so timeout 200 python tools/gemm_bench.py 16 2>&1 | tail -n 2 | sed "s/^/g5 M=16 /"
timeout 200 python tools/gemm_bench.py 16 2>&1 | tail -n 2 | sed "s/^/g4 M=16 /"
B200_LIB_PATH=$PWD/build/g5/libb200_vllm.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-fp8-kv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('g5 bench', d['ms_per_step'], d['value'], d['e2e']['value'])"
