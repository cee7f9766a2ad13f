"""Where does a TP rank's decode step go?  Builds ONE rank's shard of Llama-3-8B AWQ for tp = N on a single GPU, with no
communicator attached (the exchange steps are absent, the GEMM / attention shapes are the rank's own), and times
the b=32 ctx=4096 step as graph replays: the compute floor of a rank, to hold against the measured N-GPU step
(profiles/r2_{k,l,m}_bench_tp{2,4,8}.json).  `B200_PROFILE=decode ncu --profile-from-start off ...` lists the launches.
    python tools/tp_rank_step.py [N ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from vllm_rs_b200 import synthetic

cx = bench.Ctx()
B, ctx, K, W = 32, 4096, 20, 5
for tp in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]:
    bps = (ctx + K + W + 64) // 64 + 1
    nb = B * bps + 2
    r = synthetic.build_awq_runner("llama3-8b", num_blocks=nb, max_blocks_per_seq=bps, max_batch=B, tp_size=tp,
                                   tp_rank=0, device=cx.dev)
    synthetic.fill_kv_random(r)
    si = bench.decode_inputs(cx, r, B, ctx, bps, nb, 128256)
    ms, _ = bench.graph_time(cx, lambda: r.decode_device(*si, want_logits=False), K, W, prof="decode")
    a = synthetic.decode_algorithmic_bytes(r.shape, B, ctx)
    print(json.dumps({"tp": tp, "rank_step_ms_without_exchange": round(ms, 4), "launches": int(r.launches_per_step),
                      "rank_algorithmic_GB": round(a["total"] / 1e9, 3),
                      "rank_hbm_floor_ms": round(a["total"] / 6530.3e6, 4)}))
    del r
    torch.cuda.empty_cache()
