# Call C: sampler tests (own process), dense GEMM rasterisation sweep, full suite
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sampling_gpu.py -m gpu -q --timeout 300 > gpurun_out/c_sampling.log 2>&1; echo "sampling rc=$?"
for g in 4 8 16 32; do B200_DENSE_GROUP_P=$g timeout 300 python tools/dense_bench.py 8192 2>&1 | grep -E "^M=" | sed "s/^/group_p=$g /" >> gpurun_out/c_dense_group.log; done; echo "group sweep done"
timeout 300 python tools/dense_bench.py 8192 > gpurun_out/c_dense_bench.log 2>&1; echo "dense bench rc=$?"
timeout 300 python tools/dense_bench.py 2048 > gpurun_out/c_dense_bench_2048.log 2>&1
timeout 900 python -m pytest tests -m gpu -q --timeout 300 --deselect tests/test_sampling_gpu.py > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?"
tail -n 15 gpurun_out/c_sampling.log; cat gpurun_out/c_dense_group.log gpurun_out/c_dense_bench.log gpurun_out/c_dense_bench_2048.log; tail -n 3 gpurun_out/c_tests.log
