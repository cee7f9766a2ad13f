# Call W: native engine step (b200_engine_step): engine-loop tests + e2e A/B
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_runner_gpu.py -q -x -k "engine_loop or prefix_cache or host_path" > gpurun_out/w_tests.log 2>&1; echo "tests rc=$?"; tail -n 5 gpurun_out/w_tests.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --prefill-tokens 0 --no-fp8-kv 2>gpurun_out/w_bench.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('native step', d['ms_per_step'], d['value'], d['e2e'])"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --prefill-tokens 0 --no-fp8-kv --py-step 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('python step', d['ms_per_step'], d['value'], d['e2e'])"
tail -n 3 gpurun_out/w_bench.err
