# Call I: e4m3 decode-attention variants (occupancy 2 / 3) x pages per split
mkdir -p gpurun_out
for v in 2 3 4; do
  B200_ATTN_DEBUG=1 B200_ATTN_F8=$v timeout 120 python tools/attn_fp8_bench.py 2>&1 | grep -v "^$" | tail -n 4
  for pps in 7 9 11 13 17 22; do
    B200_ATTN_F8=$v B200_ATTN_PPS=$pps timeout 120 python tools/attn_fp8_bench.py 2>&1 | tail -n 1
  done
done > gpurun_out/i_attn_fp8_variants.txt 2>&1
cat gpurun_out/i_attn_fp8_variants.txt
