"""The CUDA path against an implementation that is INDEPENDENT of oracle/: HuggingFace transformers golden
vectors at the head sizes the kernels take (head_dim 64 and 128), and the fp16-weights (`bits = 16`) runner path
(BASELINE.json configs[0] / configs[2]: packed qkv / gate_up dense linears, linear.rs:75-123, attention.rs:207,
224-228) against the oracle at TinyLlama layer shapes.  Fixtures: tests/golden/*.npz + make_golden.py."""
import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import sched as osched
from parity_util import Checkpoint, _dev_meta, load_hf_fixture, tp_parity, ulp16_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["llama_d64_2l", "qwen2_d128_bias_2l"])
def test_cuda_runner_matches_huggingface_fixture(lib, name):
    meta, cfg, w, z = load_hf_fixture(name)
    NB, MAXB, MAXBLK = 24, 8, 4
    runner = Checkpoint.from_dense(cfg, w).build_rank_runner(0, 1, "cuda", num_blocks=NB, max_batch=MAXB,
                                                             max_blocks=MAXBLK, max_prefill_tokens=512)
    oracle = om.LlamaOracle(cfg, w, NB)
    seqs, nxt = [], 0
    for toks in meta["prompts"]:
        s = osched.Sequence(toks, 64)
        n = s.num_blocks()
        s.block_table = list(range(nxt, nxt + n))
        nxt += n + 1
        seqs.append(s)
    pmeta = osched.prepare_prefill(seqs, 64)
    gmeta = dict(pmeta)
    bt = np.zeros((len(seqs), MAXBLK), dtype=np.uint32)
    for r, s in enumerate(seqs):
        bt[r, :len(s.block_table)] = s.block_table
    gmeta["block_tables"] = bt
    got, _ = runner.prefill_device(gmeta)
    got = got.float().cpu()
    hf = torch.from_numpy(z["hf_logits"])
    ref = oracle.prefill(pmeta)
    noise = (ref - hf).abs().max().item()               # the reference-faithful fp16 pipeline vs HF's fp32
    err = (got - hf).abs().max().item()
    print(f"{name} prefill: |cuda - hf| {err:.2e} ({ulp16_err(got, hf):.2f} ulp16), |oracle - hf| {noise:.2e} "
          f"({ulp16_err(ref, hf):.2f} ulp16), fixture tol {meta['tol']}")
    assert err < meta["tol"], err                        # the bar the oracle itself is held to
    assert err <= 3 * noise + 2e-3
    assert got.argmax(-1).tolist() == hf.argmax(-1).tolist()
    # one decode step over the cache the CUDA prefill wrote
    for s, t in zip(seqs, meta["next_tokens"]):
        s.append_token(t)
        if len(s) % 64 == 1:
            s.block_table.append(nxt)
            nxt += 1
    mo = osched.prepare_decode(seqs, 64)
    dec, _ = runner.decode_device(*_dev_meta(mo, MAXBLK, "cuda"))
    dec = dec.float().cpu()
    hf2 = torch.from_numpy(z["hf_logits_next"])
    err2 = (dec - hf2).abs().max().item()
    print(f"{name} decode: |cuda - hf| {err2:.2e} ({ulp16_err(dec, hf2):.2f} ulp16)")
    assert err2 < meta["tol"], err2
    assert dec.argmax(-1).tolist() == hf2.argmax(-1).tolist()


def test_fp16_weights_runner_tinyllama_shapes(lib):
    """bits = 16 runner (dense tcgen05 GEMMs, packed qkv / gate_up, classic launch plan) at TinyLlama-1.1B layer
    shapes (2048 hidden, 32 heads, 4 kv heads, head_dim 64, inter 5632), 2 layers: prefill + 4 decode steps."""
    cfg = om.ModelConfig(hidden=2048, num_layers=2, num_heads=32, num_kv_heads=4, head_dim=64, inter=5632,
                         vocab=4096, rope_theta=10000.0, rms_eps=1e-5, max_pos=2048)
    r = tp_parity(cfg, "fp16", rank=0, world=1, dev="cuda", steps=4, seed=5)
    print(r)
    assert r["ids_equal"]
    assert r["truth_err_ulp16"] <= max(3 * r["noise_ulp16"], 2.0), r
    assert r["max_err_ulp16"] <= max(3 * r["noise_ulp16"], 2.0) + r["noise_ulp16"], r
