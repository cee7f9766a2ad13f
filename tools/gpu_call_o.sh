# Call O (2 GPUs): TP=2 tests + both TP bench arms after the third session's changes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_tp_gpu.py -m gpu -q > gpurun_out/o_tp_tests.log 2>&1; echo "tp tests rc=$?"; tail -n 3 gpurun_out/o_tp_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/o_bench_tp2.json 2> gpurun_out/o_bench_tp2.err; echo "bench tp2 rc=$?"
cat gpurun_out/o_bench_tp2.json | cut -c1-900
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29712 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/o_bench_ref_tp2.json 2> gpurun_out/o_bench_ref_tp2.err; echo "bench ref tp2 rc=$?"
cut -c1-300 gpurun_out/o_bench_ref_tp2.json
