# Round-1 baseline call: GPU tests, smoke, bench (both arms), launch list, ncu --set full of the decode GEMM
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.log 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/a_bench_ref.json 2> gpurun_out/a_bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/gemm_bench.py 32 > gpurun_out/a_gemm_bench.log 2>&1; echo "gemm_bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/a_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wna16_tc_dec -s 40 -c 8 -o gpurun_out/a_gemm_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/a_ncu_gemm.log 2>&1; echo "ncu gemm rc=$?"
tail -3 gpurun_out/a_tests.log gpurun_out/a_smoke.log gpurun_out/a_gemm_bench.log; cat gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
