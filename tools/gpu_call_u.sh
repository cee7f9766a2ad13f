# Call U: pages-per-split sweep of the decode attention at small batches (fp16 cache)
mkdir -p gpurun_out
for b in 1 4 8; do
  for pps in 0 2 3 4 6 9 13; do
    if [ $pps -eq 0 ]; then B200_ATTN_ONLY=f16 timeout 60 python tools/attn_fp8_bench.py $b 4096 2>&1 | tail -n 1;
    else B200_ATTN_ONLY=f16 B200_ATTN_PPS=$pps timeout 60 python tools/attn_fp8_bench.py $b 4096 2>&1 | tail -n 1; fi
  done
done > gpurun_out/u_attn_pps_small_batch.txt 2>&1
cat gpurun_out/u_attn_pps_small_batch.txt
