# Call D: first run of the tcgen05 prefill attention (own processes), penalties fix, suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_prefill_tc_gpu.py -m gpu -q -s --timeout 300 > gpurun_out/d_attn_tc.log 2>&1; ATC=$?; echo "attn tc rc=$ATC"
timeout 300 python tools/attn_bench.py 8192 > gpurun_out/d_attn_bench.log 2>&1; echo "attn bench rc=$?"
timeout 300 python tools/attn_bench.py 2048 >> gpurun_out/d_attn_bench.log 2>&1
if [ $ATC -ne 0 ]; then export B200_PREFILL_ATTN=mma; fi
echo "B200_PREFILL_ATTN=$B200_PREFILL_ATTN"
timeout 600 python -m pytest tests/test_sampling_gpu.py -m gpu -q --timeout 300 > gpurun_out/d_sampling.log 2>&1; echo "sampling rc=$?"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_sampling_gpu.py --deselect tests/test_attn_prefill_tc_gpu.py > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
B200_PROFILE=prefill timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/d_launches_prefill.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 4 > gpurun_out/d_ncu_prefill.log 2>&1; echo "prefill launch list rc=$?"
B200_PROFILE=prefill timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_prefill_tc -c 1 -o gpurun_out/d_attn_tc_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 1 > gpurun_out/d_ncu_attn.log 2>&1; echo "ncu attn rc=$?"
tail -n 25 gpurun_out/d_attn_tc.log; cat gpurun_out/d_attn_bench.log; tail -n 3 gpurun_out/d_sampling.log gpurun_out/d_tests.log; cat gpurun_out/d_bench.json; tail -n 3 gpurun_out/d_bench.err
