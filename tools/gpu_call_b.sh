# Call B: first run of the dense tcgen05 GEMM (own processes: a trap kills the CUDA context), then the suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dense_gemm_gpu.py -m gpu -q --timeout 300 -k "tc_matches and tc1" > gpurun_out/b_dense_tc1.log 2>&1; TC1=$?; echo "dense tc1 rc=$TC1"
timeout 600 python -m pytest tests/test_dense_gemm_gpu.py -m gpu -q --timeout 300 -k "tc_matches and tc2" > gpurun_out/b_dense_tc2.log 2>&1; TC2=$?; echo "dense tc2 rc=$TC2"
timeout 600 python -m pytest tests/test_dense_gemm_gpu.py -m gpu -q --timeout 300 -k "not tc_matches" > gpurun_out/b_dense_rest.log 2>&1; echo "dense rest rc=$?"
if [ $TC2 -eq 0 ]; then export B200_DENSE=tc2; elif [ $TC1 -eq 0 ]; then export B200_DENSE=tc1; else export B200_DENSE=cublas; fi
echo "B200_DENSE=$B200_DENSE"
timeout 600 python tools/dense_bench.py 8192 > gpurun_out/b_dense_bench.log 2>&1; echo "dense bench rc=$?"
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_dense_gemm_gpu.py > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
B200_PREFILL_DENSE_MIN_T=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench_oldprefill.json 2> gpurun_out/b_bench_oldprefill.err; echo "bench old prefill rc=$?"
B200_PROFILE=decode timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/b_ncu_decode.log 2>&1; echo "decode launch list rc=$?"
B200_PROFILE=prefill timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/b_launches_prefill.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 4 > gpurun_out/b_ncu_prefill.log 2>&1; echo "prefill launch list rc=$?"
B200_PROFILE=prefill timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:dense_gemm_tc -c 3 -o gpurun_out/b_dense_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --layers 2 > gpurun_out/b_ncu_dense.log 2>&1; echo "ncu dense rc=$?"
tail -n 3 gpurun_out/b_dense_tc1.log gpurun_out/b_dense_tc2.log gpurun_out/b_dense_rest.log gpurun_out/b_tests.log; cat gpurun_out/b_dense_bench.log; cat gpurun_out/b_bench.json; tail -n 3 gpurun_out/b_bench.err
