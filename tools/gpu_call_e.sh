# Call E (2 GPUs): TP tests incl. the fused GEMM -> NVLink push -> reduce+norm plan, TP=2 bench fused vs classic
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tp_gpu.py -m gpu -q -s --timeout 300 > gpurun_out/e_tp_tests.log 2>&1; echo "tp tests rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/e_bench_tp2_fused.json 2> gpurun_out/e_bench_tp2_fused.err; echo "bench tp2 fused rc=$?"
B200_TP_FUSED=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/e_bench_tp2_classic.json 2> gpurun_out/e_bench_tp2_classic.err; echo "bench tp2 classic rc=$?"
tail -n 12 gpurun_out/e_tp_tests.log; cat gpurun_out/e_bench_tp2_fused.json; tail -n 5 gpurun_out/e_bench_tp2_fused.err; cat gpurun_out/e_bench_tp2_classic.json; tail -n 3 gpurun_out/e_bench_tp2_classic.err
