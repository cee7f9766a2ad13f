#!/bin/bash
# A/B of the decode GEMM kernels per Llama-3-8B projection: round-1 tcgen05 (2), stream-K tcgen05 (5), stream-K mma.sync (6)
for M in 1 16 32; do
  for shape in "4096 6144 0" "4096 4096 0" "4096 28672 1" "14336 4096 0"; do
    set -- $shape
    for v in 2 5 6; do timeout 120 python tools/gemm_shape_bench.py $1 $2 $M $3 $v 2>&1 | tail -1; done
  done
done
