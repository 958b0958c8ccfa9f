"""CPU: host-side pieces that need no GPU - the random start-board generator of HexVecEnv(sample_board=True), the bench
byte accounting, the scripted opponent stand-in."""
import numpy as np
import torch

import bench
from hex_gym_env_b200.vec_env import random_start_boards
from oracle.scripted import scripted_choice


def test_random_start_boards_invariants():
    gen = torch.Generator(device="cpu")
    gen.manual_seed(7)
    for N in (5, 8, 11):
        b = random_start_boards(3000, N, gen, torch.device("cpu"))
        assert b.shape == (3000, N, N) and b.dtype == torch.int8 and set(b.unique().tolist()) <= {0, 1, 2}
        nb, nw = (b == 0).flatten(1).sum(1), (b == 1).flatten(1).sum(1)
        assert bool((nb == nw).all())                                   # even stone count, BLACK to move (HexSingleGame.py:312)
        occ = b != 2
        h = occ.any(2).sum(1)
        w = occ.any(1).sum(1)
        assert int(h.max()) <= max(N - 2, N // 4) and int(w.max()) <= max(N - 2, N // 4)   # rectangle sides in [N//4, N-2]
        assert int((nb + nw).max()) >= (N // 4) ** 2 // 2 and float((nb + nw).float().mean()) > 1.0


def test_contract_and_moved_bytes():
    # SURVEY.md section 8(d) table
    assert [bench.contract_bytes(N) for N in (5, 6, 7, 11, 19)] == [207, 289, 367, 831, 2399]
    # this implementation: 2 * (C + 4 * (W + 2)) + 2C + 5
    assert bench.moved_bytes(11) == 2 * (121 + 24) + 242 + 5 == 537
    assert bench.moved_bytes(11, sampled=False) == 541


def test_scripted_choice_is_legal_and_deterministic():
    rs = np.random.RandomState(0)
    for _ in range(50):
        board = rs.choice([-1, 0, 1], size=(6, 6))
        mask = (board == 0).reshape(-1)
        if not mask.any():
            continue
        a = scripted_choice(board, mask)
        assert mask[a] and a == scripted_choice(board.astype(np.float64), mask.astype(np.uint8))


def test_product_never_touches_the_oracle_or_the_emulator():
    """The package (the product path) must not import, load or mention the test infrastructure: oracle/ and tests/emu."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "hex_gym_env_b200")
    pat = re.compile(r"^\s*(from|import)\s+(oracle|tests|emu)\b|libhexref|libhexb_emu|HEXB_HOST_EMU\s*1", re.M)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu")):
                src = open(os.path.join(dirpath, f)).read()
                assert not pat.search(src), os.path.join(dirpath, f)
    # and importing the package does not pull them in
    import subprocess
    import sys
    code = "import sys, hex_gym_env_b200; bad = [m for m in sys.modules if m.split('.')[0] in ('oracle', 'emu')]; assert not bad, bad"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(root))


def test_vec_env_rejects_option_mixes_before_touching_a_gpu():
    """HexVecEnv's option checks (which reference env each option belongs to) fire before any device work."""
    import pytest
    from hex_gym_env_b200.vec_env import HexVecEnv
    for kw in (dict(variant="hex-v1"), dict(output="jax"), dict(variant="hex-v0", sample_board=True),
               dict(variant="hex-v0", base_model=lambda o, m: None), dict(variant="selfplay", sample_board=True, base_model=lambda o, m: None),
               dict(variant="selfplay", opponent_model=lambda o, m: None)):
        with pytest.raises(ValueError):
            HexVecEnv(board_size=5, num_envs=4, **kw)


def test_lazy_infos_follow_the_sb3_shape():
    """output='torch': infos behaves like SB3's list of dicts - terminal_observation only for the finished games."""
    from hex_gym_env_b200.vec_env import _LazyInfos
    done = np.array([0, 1, 0], np.uint8)
    term = np.arange(3 * 4, dtype=np.int8).reshape(3, 2, 2)
    infos = _LazyInfos(done, term, 3)
    assert len(infos) == 3 and infos[0] == {} and infos[2] == {}
    assert set(infos[1]) == {"terminal_observation", "TimeLimit.truncated"} and infos[1]["TimeLimit.truncated"] is False
    assert np.array_equal(infos[1]["terminal_observation"], term[1])
    assert [bool(i) for i in infos] == [False, True, False]
    t = _LazyInfos(torch.tensor([1, 0]), torch.ones(2, 2, 2), 2)[0]["terminal_observation"]
    assert isinstance(t, torch.Tensor) and t.shape == (2, 2)


def test_rollout_helpers_have_no_cpu_path():
    """masked_sample / gae are kernels of libhexb.so: CPU tensors are refused, nothing is computed in torch instead."""
    import pytest
    from hex_gym_env_b200.rollout import gae, masked_sample
    with pytest.raises(RuntimeError):
        masked_sample(torch.zeros(4, 9), torch.ones(4, 9, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        gae(torch.zeros(3, 4), torch.zeros(4, 4), torch.zeros(3, 4, dtype=torch.uint8))
