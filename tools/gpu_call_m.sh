# Call M: decode GEMM split-K through thread-block clusters (DSMEM) vs the global meeting point
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_runner_gpu.py -q -x -k "wna16 or decode_steps or host_path" > gpurun_out/m_tests.log 2>&1
echo "tests rc=$?"
tail -n 8 gpurun_out/m_tests.log
for c in 0 1; do
  B200_DEC_CLUSTER=$c B200_GEMM_CHAIN=1 timeout 200 python tools/gemm_bench.py 32 2>&1 | tail -n 2 | sed "s/^/cluster=$c /"
done
for c in 0 1; do
  B200_DEC_CLUSTER=$c timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-fp8-kv 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cluster=$c', d['ms_per_step'], d['value'], d['e2e']['value'])"
done
