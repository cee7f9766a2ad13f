# Call S: batch sweep of the decode step at ctx 4096 (north star: batch 1-64), fp16 and e4m3 KV cache
mkdir -p gpurun_out
for b in 1 8 16 64; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 5 --no-cpu-baseline --prefill-tokens 0 > gpurun_out/s_bench_b$b.json 2> gpurun_out/s_bench_b$b.err
  echo "b=$b rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/s_bench_b$b.json"))
f8 = d.get("fp8_kv") or {}
print("b=$b", "ms", round(d["ms_per_step"], 3), "tok/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "step_frac", round(d["step_roofline"]["frac"], 3),
      "attn_frac", round(d["roofline"]["frac"], 3), "| fp8 ms", round(f8.get("ms_per_step", 0), 3), "tok/s", round(f8.get("value", 0)),
      "attn_frac", round((f8.get("attention") or {}).get("frac", 0), 3), f8.get("error", ""))
PY
done
