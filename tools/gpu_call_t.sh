# Call T: batches 1-4 on the tcgen05 decode GEMM instead of the flat GEMV
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_runner_gpu.py tests/test_kernels_gpu.py -q -x -k "small_batches or host_path or wna16 or decode_steps" > gpurun_out/t_tests.log 2>&1; echo "tests rc=$?"; tail -n 4 gpurun_out/t_tests.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -n 1
for m in 4 0; do
  B200_GEMV_MAX_M=$m B200_GEMM_CHAIN=0 timeout 100 python tools/gemm_bench.py 1 2>&1 | tail -n 1 | sed "s/^/gemv_max_m=$m /"
  B200_GEMV_MAX_M=$m B200_GEMM_CHAIN=0 timeout 100 python tools/gemm_bench.py 4 2>&1 | tail -n 1 | sed "s/^/gemv_max_m=$m /"
done
for b in 1 2 4; do
  timeout 200 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/s_bench_b$b.json 2> gpurun_out/s_bench_b$b.err; echo "b=$b rc=$?"
  python -c "import json; d=json.load(open('gpurun_out/s_bench_b$b.json')); print($b, d['ms_per_step'], d['value'], d['e2e']['value'], (d['fp8_kv'] or {}).get('ms_per_step'))"
done
