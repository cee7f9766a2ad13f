mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "not wna16_gemm" --timeout 120 > gpurun_out/t1.log 2>&1; echo "t1 rc=$?"
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "wna16_gemm" --timeout 120 > gpurun_out/t2.log 2>&1; echo "t2 rc=$?"
timeout 600 python -m pytest tests/test_runner_gpu.py -m gpu -q -s --timeout 200 > gpurun_out/t3.log 2>&1; echo "t3 rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"
tail -5 gpurun_out/t1.log gpurun_out/t2.log gpurun_out/t3.log gpurun_out/smoke.log; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
