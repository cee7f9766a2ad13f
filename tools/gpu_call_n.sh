# Call N: round-1 v7 validation: full GPU suite, smoke, both bench arms, launch list, ncu --set full of the e4m3 attention
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/n_tests.log 2>&1; echo "tests rc=$?"
tail -n 6 gpurun_out/n_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/n_smoke.log
timeout 900 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err; echo "bench ref rc=$?"
B200_PROFILE=decode timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/n_launches_decode.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --prefill-tokens 0 --no-fp8-kv > gpurun_out/n_ncu_decode.log 2>&1; echo "decode launch list rc=$?"
B200_ATTN_ONLY=fp8 timeout 600 ncu --set full --clock-control none --import-source on -k regex:paged_attn_decode_kernel --launch-skip 8 -c 1 -o gpurun_out/n_attn_fp8_full -f python tools/attn_fp8_bench.py > gpurun_out/n_ncu_attn_fp8.log 2>&1; echo "ncu fp8 attn rc=$?"
cat gpurun_out/n_bench.json; tail -n 2 gpurun_out/n_bench.err; cat gpurun_out/n_bench_ref.json
