"""CPU oracle for the vllm.rs forward hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import anything from this package.  The product (vllm_rs_b200/) never does.

What it restates (every function cites the reference file:line it follows):
  * sched.py   -- Sequence / BlockManager / Scheduler / prepare_{decode,prefill}: integer state,
                  compared BIT-EXACT with the C++ host mirror (vllm_rs_b200/csrc/host/).
  * quant.py   -- AWQ / GPTQ checkpoint packing and dequantisation.
  * model.py   -- LLaMA / Qwen2 forward with a paged KV cache in torch fp32 on CPU, fp16
                  rounding at the points where the reference materialises fp16 tensors.
  * c/         -- the same per-step math in plain C + OpenMP (the CPU baseline that is timed).

PARITY UNPINNED for the floating-point path: the reference has no test, fixture or golden
vector that pins attention, linear or logits numerics, its kernels live
in the un-vendored `attention-rs@0bf5727` / `candle@68b6d74` git dependencies, and no Rust
toolchain exists here, so the reference cannot be run.  Mitigation: model.py is
cross-checked against HuggingFace `transformers` LlamaForCausalLM / Qwen2ForCausalLM with
identical random weights (tests/test_oracle_model.py, fixtures under tests/golden/), and the
AWQ/GPTQ formats are the public AutoAWQ / AutoGPTQ ones the loader's tensor shapes imply.
The integer path (block ids, slot mapping, block tables) is pinned by construction against
hand-derived traces from the reference source (tests/test_sched_parity.py).
"""
