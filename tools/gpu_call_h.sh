# Call H: fp8 attention with 3 CTAs/SM + split QK chains; fixed fp8 tests; bench with fp8 leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fp8_kv_gpu.py -q -x > gpurun_out/h_fp8_tests.log 2>&1
echo "fp8 tests rc=$?"
tail -n 12 gpurun_out/h_fp8_tests.log
B200_BENCH_TRACE=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/h_bench.json"))
print("fp16:", d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["frac"])
print("fp8 :", json.dumps(d["fp8_kv"]))
PY
