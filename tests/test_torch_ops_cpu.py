"""CPU: the PyTorch operator layer (csrc/hexb_torch.cpp -> libhexb_torch.so) builds, registers torch.ops.hexb.* with mutable
output annotations for every buffer the C ABI writes, and refuses to run without a valid handle / CUDA tensors (there is no
CPU path behind the operators). Compute is covered by tests/test_gpu_torch_ops.py."""
import pytest
import torch

from hex_gym_env_b200 import torch_ops

OPS = ("version", "reset", "step", "rollout", "half_step", "ply", "encode", "sample_actions", "stats", "masked_sample", "gae", "set_launch_form", "set_eval")


@pytest.fixture(scope="module")
def ops():
    return torch_ops.load()


def test_operators_registered(ops):
    for name in OPS:
        assert hasattr(ops, name), name
    from hex_gym_env_b200 import _native
    assert ops.version() == _native.lib().hexb_version()


def test_schemas_mark_outputs_mutable(ops):
    s = str(ops.step.default._schema)
    for out in ("obs", "mask", "reward", "done", "term_obs", "actions_out"):
        assert "!)? %s" % out in s, (out, s)
    for inp in ("Tensor? actions", "Tensor? opp_u"):
        assert inp in s
    assert "Tensor(a!) out8" in str(ops.stats.default._schema)
    assert "Tensor(a!)? obs" in str(ops.rollout.default._schema)


def test_no_cpu_path(ops):
    with pytest.raises(RuntimeError, match="null environment handle"):
        ops.step(0, None, None, None, None, None, None, None, None)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.masked_sample(torch.zeros(4, 9), torch.ones(4, 9, dtype=torch.uint8), torch.zeros(4, dtype=torch.float64), None, None, None)
