"""hexb_mem_alloc / hexb_mem_free (compressible device memory) and a HexBatch whose state and outputs live in it: same results
as the oracle, bit for bit (compression changes the bytes that cross HBM, never the values)."""
import ctypes

import numpy as np
import pytest
import torch

from hex_gym_env_b200 import AGENT_RANDOM, VARIANT_B, HexBatch, _native

pytestmark = pytest.mark.gpu


def test_mem_alloc_roundtrip():
    n = (16 << 20) + 12345
    buf = _native.DeviceBuffer(n, 0, True)
    assert buf.ptr % (2 << 20) == 0 and isinstance(buf.compressed, bool)
    t = torch.as_tensor(buf, device="cuda:0")
    assert t.dtype == torch.uint8 and t.numel() == n and t.data_ptr() == buf.ptr
    pat = (torch.arange(n, device="cuda:0", dtype=torch.int64) * 2654435761 >> 7).to(torch.uint8)
    t.copy_(pat)
    t[1 << 20: 9 << 20] = 0          # a long run of zeros: the compressible case
    pat[1 << 20: 9 << 20] = 0
    torch.cuda.synchronize()
    assert torch.equal(t, pat)
    v = t[: 4 << 20].view(torch.int32)
    v += 1
    assert torch.equal(t[: 4 << 20].view(torch.int32), pat[: 4 << 20].view(torch.int32) + 1)
    del t, v
    L = _native.lib()
    assert L.hexb_mem_free(ctypes.c_void_p(12345)) != 0          # not one of ours
    p, got = ctypes.c_void_p(), ctypes.c_int32(-1)
    assert L.hexb_mem_alloc(0, 1, 0, ctypes.byref(p), ctypes.byref(got)) == 0 and p.value and got.value == 0   # ordinary on request
    assert L.hexb_mem_free(p) == 0
    assert L.hexb_mem_alloc(0, 0, 1, ctypes.byref(p), None) != 0  # zero bytes


def test_compressible_batch_matches_oracle():
    from oracle import hexref
    N, G, T = 11, 80000 + 17, 60
    env = HexBatch(N, G, variant=VARIANT_B, device=0, seed=5, agent_mode=AGENT_RANDOM, compressible=True)
    assert env.compressible, env.memory_kind          # state (11.6 MB) and obs / mask (9.7 MB each) come from hexb_mem_alloc
    ref = hexref.RefBatch(hexref.KIND_SELFPLAY_B, N, G, seed=5, agent_mode=2)
    obs, mask = env.reset()
    robs, rmask = ref.reset()
    assert np.array_equal(obs.cpu().numpy(), robs) and np.array_equal(mask.cpu().numpy(), rmask)
    for t in range(T):
        o = env.step(want_actions=True)
        r = ref.step()
        if t % 7 == 0 or t == T - 1:
            for k in ("obs", "mask", "reward", "done", "actions"):
                assert np.array_equal(o[k].cpu().numpy(), r[k]), (k, t)
    e, re_ = env.export_state(), ref.export()
    for k in ("board", "regions", "region_counter", "cur", "done", "winner", "agent"):
        assert np.array_equal(e[k].cpu().numpy(), re_[k]), k
    assert np.array_equal(env.stats().cpu().numpy(), ref.stats())
    sd = env.state_dict()                              # the checkpoint path reads and writes the same memory
    env2 = HexBatch(N, G, variant=VARIANT_B, device=0, seed=5, agent_mode=AGENT_RANDOM, compressible=False)
    env2.reset()
    env2.load_state_dict(sd)
    a, b = env.step(want_actions=True), env2.step(want_actions=True)
    for k in ("obs", "mask", "reward", "done", "actions"):
        assert torch.equal(a[k], b[k]), k
