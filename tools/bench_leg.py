"""Run ONE leg of bench.py in isolation (debugging aid):  python tools/bench_leg.py fp8|prefill|prefill_awq [--layers L]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

ap = argparse.ArgumentParser()
ap.add_argument("leg", choices=["fp8", "prefill", "prefill_awq"])
ap.add_argument("--layers", type=int, default=0)
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--ctx", type=int, default=4096)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--seqs", type=int, default=8)
ap.add_argument("--tokens", type=int, default=8192)
a = ap.parse_args()
cx = bench.Ctx()
if a.leg == "fp8":
    peaks, _ = bench.measured_peaks()
    res = bench.run_fp8_leg(a, cx, a.steps, a.warmup, float(peaks["hbm_gbs"]))
elif a.leg == "prefill":
    res = bench.run_prefill_leg(a, cx, "fp16", a.seqs, a.tokens)
else:
    res = bench.run_prefill_leg(a, cx, "awq", 1, a.tokens)
print(json.dumps(res))
