"""PD KV hand-off on the GPU (transfer/mod.rs:260-311,370-470): RemoteTcp round trip through the wire frame and the
LocalIpc peer copy, byte-exact, with the achieved copy rates printed."""
import time

import numpy as np
import pytest
import torch

from oracle import transfer as ot

pytestmark = pytest.mark.gpu


def _caches(L, nb, Hkv, D, dtype, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    mk = lambda: (torch.randint(0, 255, (nb, 64, Hkv, D), generator=g, device=dev, dtype=torch.uint8)  # noqa: E731
                  if dtype == torch.uint8 else torch.randn(nb, 64, Hkv, D, generator=g, device=dev, dtype=dtype))
    return [mk() for _ in range(L)], [mk() for _ in range(L)]


@pytest.mark.parametrize("dtype", [torch.float16, torch.uint8])
def test_remote_tcp_round_trip_through_the_frame(dtype):
    from vllm_rs_b200 import transfer
    L, nb, Hkv, D = 4, 40, 8, 128
    sk, sv = _caches(L, nb, Hkv, D, dtype, "cuda", 1)
    ck, cv = _caches(L, nb, Hkv, D, dtype, "cuda", 2)
    before_k = [t.clone() for t in ck]
    rng = np.random.default_rng(3)
    table = rng.permutation(nb)[:17].tolist()                     # the server's blocks of the sequence
    local = rng.permutation(nb)[:20].tolist()                     # the client allocated a few more
    frame = transfer.transfer_kv_cache(sk, sv, table, seq_id=41, first_token=777, num_cached_tokens=128,
                                       sending_time=1760000000000)
    torch.cuda.synchronize()
    # the frame IS what the reference would put on the socket: oracle bincode of the gathered blocks
    layers = [(sk[l][table].cpu().numpy().tobytes(), sv[l][table].cpu().numpy().tobytes()) for l in range(L)]
    assert frame.numpy().tobytes() == ot.kv_remote_frame(41, 777, layers, len(table), 1760000000000, 128)
    got = transfer.receive_kv_cache(frame, ck, cv, local)
    torch.cuda.synchronize()
    assert (got["seq_id"], got["first_token"], got["num_blocks"], got["num_cached_tokens"]) == (41, 777, 17, 128)
    for l in range(L):
        for i, (s, d) in enumerate(zip(table, local)):
            assert torch.equal(ck[l][d], sk[l][s]) and torch.equal(cv[l][d], sv[l][s]), (l, i)
        untouched = sorted(set(range(nb)) - set(local[:len(table)]))
        assert torch.equal(ck[l][untouched], before_k[l][untouched])
    with pytest.raises(ValueError):
        transfer.receive_kv_cache(frame, ck, cv, local[:5])       # mod.rs:296-298


def test_local_ipc_copy_same_device_and_rate():
    """The LocalIpc leg on one device (server and client caches in the same memory) at Llama-3-8B block size:
    byte-exact, and the rate of the one-launch-per-tensor copy."""
    from vllm_rs_b200 import transfer
    L, nb, Hkv, D = 32, 160, 8, 128
    sk, sv = _caches(L, nb, Hkv, D, torch.float16, "cuda", 5)
    ck = [torch.zeros_like(t) for t in sk]
    cv = [torch.zeros_like(t) for t in sv]
    table = list(range(0, 128))
    local = list(range(16, 144))
    transfer.receive_kv_cache_local_ipc(sk, sv, table, ck, cv, local)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    keep = transfer.receive_kv_cache_local_ipc(sk, sv, table, ck, cv, local)
    e1.record()
    torch.cuda.synchronize()
    del keep
    nbytes = 2 * L * len(table) * 64 * Hkv * D * 2
    print(f"LocalIpc copy of a 8192-token sequence ({nbytes / 1e6:.0f} MB): {e0.elapsed_time(e1):.3f} ms = "
          f"{2 * nbytes / e0.elapsed_time(e1) / 1e6:.0f} GB/s read+write")
    for l in (0, 7, 31):
        assert torch.equal(ck[l][16:144], sk[l][0:128]) and torch.equal(cv[l][16:144], sv[l][0:128])
        assert not ck[l][:16].any() and not ck[l][144:].any()


def test_local_ipc_copy_between_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from vllm_rs_b200 import transfer
    L, nb, Hkv, D = 32, 160, 8, 128
    sk, sv = _caches(L, nb, Hkv, D, torch.float16, "cuda:1", 7)           # the prefill server's GPU
    ck = [torch.zeros(nb, 64, Hkv, D, dtype=torch.float16, device="cuda:0") for _ in range(L)]
    cv = [torch.zeros(nb, 64, Hkv, D, dtype=torch.float16, device="cuda:0") for _ in range(L)]
    rng = np.random.default_rng(9)
    table, local = rng.permutation(nb)[:128].tolist(), rng.permutation(nb)[:128].tolist()
    torch.cuda.synchronize("cuda:1")
    with torch.cuda.device(0):
        transfer.receive_kv_cache_local_ipc(sk, sv, table, ck, cv, local)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keep = transfer.receive_kv_cache_local_ipc(sk, sv, table, ck, cv, local)
        e1.record()
        torch.cuda.synchronize()
        del keep, t0
    nbytes = 2 * L * len(table) * 64 * Hkv * D * 2
    print(f"LocalIpc GPU1 -> GPU0 over NVLink, 8192-token sequence ({nbytes / 1e6:.0f} MB): "
          f"{e0.elapsed_time(e1):.3f} ms = {nbytes / e0.elapsed_time(e1) / 1e6:.0f} GB/s (NVLink 5: 900 GB/s per direction)")
    for l in (0, 15, 31):
        for s, d in list(zip(table, local))[:20]:
            assert torch.equal(ck[l][d].cpu(), sk[l][s].cpu()) and torch.equal(cv[l][d].cpu(), sv[l][s].cpu())
