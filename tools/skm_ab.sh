#!/bin/bash
# A/B of the stream-K decode GEMM flavours per Llama-3-8B projection: tcgen05 (variant 5) vs mma.sync registers (6);
# SKM_VARIANTS / SKM_BATCHES override the lists (the round-1 kernel is variant 2)
for M in ${SKM_BATCHES:-1 8 16 32}; do
  for shape in "4096 28672 1" "14336 4096 0"; do
    set -- $shape
    for v in ${SKM_VARIANTS:-5 6}; do timeout 120 python tools/gemm_shape_bench.py $1 $2 $M $3 $v 2>&1 | tail -1; done
  done
done
