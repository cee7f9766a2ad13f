# Call K: pipeline trace of the decode GEMM (instrumented side build), all four Llama-3-8B shapes
mkdir -p gpurun_out
for s in qkv o gate_up down; do
  B200_LIB_PATH=$PWD/build/trace/libb200_vllm.so timeout 120 python tools/dec_trace.py $s 32 > gpurun_out/k_trace_$s.txt 2>&1
  echo "$s rc=$?"
done
head -n 40 gpurun_out/k_trace_o.txt
