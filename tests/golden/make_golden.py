"""Generate the HuggingFace cross-check fixtures for tests/test_oracle_cpu.py.

    python tests/golden/make_golden.py

Builds tiny random LlamaForCausalLM / Qwen2ForCausalLM models with transformers (fp32 math on
fp16-valued weights, eager attention), runs full-sequence forwards, and stores weights +
last-token logits.  The reference itself cannot be run here (Rust, un-vendored kernels), so
HF is the independent implementation the oracle is pinned against.
"""
import json
import sys
from pathlib import Path

import numpy as np
import torch
from transformers import LlamaConfig, LlamaForCausalLM, Qwen2Config, Qwen2ForCausalLM

OUT = Path(__file__).parent
sys.path.insert(0, str(OUT.parent.parent))


def make(name, kind, cfg, rope_scaling=None, seed=0, prompts=(37, 64, 65), tol=6e-3):
    torch.manual_seed(seed)
    common = dict(hidden_size=cfg["hidden"], num_hidden_layers=cfg["num_layers"],
                  num_attention_heads=cfg["num_heads"], num_key_value_heads=cfg["num_kv_heads"],
                  intermediate_size=cfg["inter"], vocab_size=cfg["vocab"], rms_norm_eps=cfg["rms_eps"],
                  rope_theta=cfg["rope_theta"], max_position_embeddings=cfg["max_pos"],
                  tie_word_embeddings=False, attn_implementation="eager")
    if kind == "llama":
        hc = LlamaConfig(**common, head_dim=cfg["head_dim"], attention_bias=False, mlp_bias=False,
                         rope_scaling=rope_scaling)
        m = LlamaForCausalLM(hc)
    else:
        hc = Qwen2Config(**common, use_sliding_window=False)
        m = Qwen2ForCausalLM(hc)
    m = m.to(torch.float32).eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1 and "norm" in "".join(n for n, q in m.named_parameters() if q is p):
                p.copy_(1 + 0.05 * torch.randn_like(p))
            elif p.dim() == 1:
                p.copy_(0.1 * torch.randn_like(p))
            else:
                p.copy_(0.05 * torch.randn_like(p))
        m.model.embed_tokens.weight.copy_(torch.randn_like(m.model.embed_tokens.weight))
        for p in m.parameters():                       # fp16-valued weights, fp32 math
            p.copy_(p.to(torch.float16).to(torch.float32))
    g = torch.Generator().manual_seed(seed + 1)
    toks = [torch.randint(0, cfg["vocab"], (n,), generator=g).tolist() for n in prompts]
    nxt = [int(torch.randint(0, cfg["vocab"], (1,), generator=g)) for _ in prompts]
    logits, logits_next = [], []
    with torch.no_grad():
        for t, n in zip(toks, nxt):
            logits.append(m(torch.tensor([t])).logits[0, -1])
            logits_next.append(m(torch.tensor([t + [n]])).logits[0, -1])
    arrs = {"embed": m.model.embed_tokens.weight, "final_norm": m.model.norm.weight,
            "lm_head": m.lm_head.weight, "hf_logits": torch.stack(logits),
            "hf_logits_next": torch.stack(logits_next)}
    for l, layer in enumerate(m.model.layers):
        a, f = layer.self_attn, layer.mlp
        arrs.update({f"l{l}.input_norm": layer.input_layernorm.weight,
                     f"l{l}.post_norm": layer.post_attention_layernorm.weight,
                     f"l{l}.wq": a.q_proj.weight.t(), f"l{l}.wk": a.k_proj.weight.t(),
                     f"l{l}.wv": a.v_proj.weight.t(), f"l{l}.wo": a.o_proj.weight.t(),
                     f"l{l}.wgate": f.gate_proj.weight.t(), f"l{l}.wup": f.up_proj.weight.t(),
                     f"l{l}.wdown": f.down_proj.weight.t()})
        if cfg["qkv_bias"]:
            arrs.update({f"l{l}.bq": a.q_proj.bias, f"l{l}.bk": a.k_proj.bias, f"l{l}.bv": a.v_proj.bias})
    out = {}
    for k, v in arrs.items():
        v = v.detach().contiguous()
        out[k] = v.numpy().astype(np.float32 if k.startswith("hf_") else np.float16)
    np.savez_compressed(OUT / f"{name}.npz", **out)
    mcfg = dict(cfg)
    mcfg["rope_scaling"] = rope_scaling
    (OUT / f"{name}.json").write_text(json.dumps({"cfg": mcfg, "prompts": toks, "next_tokens": nxt,
                                                  "tol": tol, "kind": kind}))
    print(name, {k: v.shape for k, v in out.items() if k.startswith("hf_")})


if __name__ == "__main__":
    base = dict(hidden=128, num_layers=2, num_heads=8, num_kv_heads=2, head_dim=16, inter=256,
                vocab=96, rope_theta=10000.0, rms_eps=1e-5, max_pos=256, qkv_bias=False, block_size=64)
    make("tinyllama_2l", "llama", base)
    make("qwen2_bias_2l", "qwen2", dict(base, qkv_bias=True, rope_theta=1000000.0), seed=3)
    make("llama3_rope_2l", "llama", dict(base, head_dim=32, num_heads=4, num_kv_heads=1,
                                         rope_theta=500000.0),
         rope_scaling={"rope_type": "llama3", "factor": 8.0, "low_freq_factor": 1.0,
                       "high_freq_factor": 4.0, "original_max_position_embeddings": 64}, seed=5)
    # head_dim 64 / 128: the head sizes the CUDA kernels take, so the GPU path is compared with an implementation
    # that is independent of oracle/ (tests/test_hf_fixture_gpu.py); dims are multiples of 64 (tcgen05 GEMM K tiles)
    make("llama_d64_2l", "llama", dict(base, hidden=128, head_dim=64, num_heads=2, num_kv_heads=1, inter=256,
                                       vocab=128), seed=7, prompts=(37, 64, 65, 130))
    make("qwen2_d128_bias_2l", "qwen2", dict(base, hidden=256, head_dim=128, num_heads=2, num_kv_heads=1,
                                             inter=256, vocab=128, qkv_bias=True, rope_theta=1000000.0),
         seed=9, prompts=(37, 64, 65, 130))
