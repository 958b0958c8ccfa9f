"""CPU: the kernel's phase code (hex_gym_env_b200/csrc/hexb_phases.cuh), replayed serially by the host emulator
(tests/emu), against the golden vectors of the unmodified reference and against the oracle. This is a check of the
device LOGIC in a container without a GPU; the -m gpu tests repeat the same drivers through libhexb.so on the B200."""
import os

import numpy as np
import pytest

import parity
from conftest import GOLDEN, golden_files
from oracle import hexref
from emu.emu import EmuBatch


def make(kind, N, G, seed=0, game_offset=0, agent_mode=0, opponent_first=False, auto_reset=True, eval_state=False,
         manual_opponent=False, pool_size=0):
    if kind == hexref.KIND_GAME_A:
        return EmuBatch(0, N, G, raw=True)
    variant = 0 if kind == hexref.KIND_ENV_A else 1
    return EmuBatch(variant, N, G, seed=seed, game_offset=game_offset, agent_mode=agent_mode, opponent_first=opponent_first,
                    auto_reset=auto_reset, eval_state=eval_state, manual_opponent=manual_opponent, pool_size=pool_size)


@pytest.mark.parametrize("name", golden_files("oppmodel_") + golden_files("evalpool_"))
def test_golden_scripted_opponent(name):
    parity.golden_oppmodel(make, name)


@pytest.mark.parametrize("name", golden_files("game_A"))
def test_golden_raw_games(name):
    parity.golden_raw_game(make, name)


@pytest.mark.parametrize("name", golden_files("selfplay_") + golden_files("envA_"))
def test_golden_rollouts(name):
    parity.golden_rollout(make, name)


@pytest.mark.parametrize("N", [3, 4, 5, 6, 7, 8, 9, 11, 13, 19])
@pytest.mark.parametrize("agent_mode", [0, 1, 2])
def test_selfplay_vs_oracle(N, agent_mode):
    G = 300 if N <= 11 else 140   # not a multiple of the 128-game tile: exercises the ragged last tile
    T = 2 * N * N // 3 + 10
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, N, G, T, seed=N * 10 + agent_mode, game_offset=12345678901 * agent_mode,
                         fused=(agent_mode != 1), agent_mode=agent_mode)


@pytest.mark.parametrize("N", [3, 5, 7, 10])
@pytest.mark.parametrize("opponent_first", [False, True])
def test_envA_vs_oracle(N, opponent_first):
    parity.versus_oracle(make, hexref.KIND_ENV_A, N, 200, N * N, seed=5 + N, fused=not opponent_first, opponent_first=opponent_first)


@pytest.mark.parametrize("kind", [hexref.KIND_SELFPLAY_B, hexref.KIND_ENV_A])
def test_no_auto_reset(kind):
    """Finished games stay finished: terminal observation (incl. the opponent's view after an agent win), stale rewards."""
    kw = dict(agent_mode=2) if kind == hexref.KIND_SELFPLAY_B else {}
    parity.versus_oracle(make, kind, 4, 200, 30, seed=3, fused=False, auto_reset=False, illegal_rate=0.1, **kw)


def test_eval_state_draws():
    parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, 5, 130, 40, seed=9, agent_mode=2, eval_state=True)


@pytest.mark.parametrize("N", [7, 19])
def test_snake_chain_worst_case(N):
    parity.snake_chain(make, N)


@pytest.mark.parametrize("N", [4, 11])
def test_preset_board_rebuild(N):
    """hexb_import_boards (HexGame.__init__ with a preset board, raster-order flood_fill) vs the oracle."""
    rs = np.random.RandomState(N)
    G = 50
    boards = rs.choice([0, 1, 2], size=(G, N, N), p=[0.3, 0.3, 0.4]).astype(np.int8)
    env = EmuBatch(0, N, G, raw=True)
    env.reset()
    env.import_boards(boards, np.zeros(G, np.int8))
    ref = hexref.RefBatch(hexref.KIND_GAME_A, N, G)
    ref.set_board(boards, cur=0)
    e, r = env.export(), ref.export()
    for k in ("board", "regions", "region_counter", "cur"):
        parity.eq(e[k], r[k], k)


@pytest.mark.parametrize("N,kind", [(5, hexref.KIND_SELFPLAY_B), (6, hexref.KIND_ENV_A)])
def test_half_step_vs_oracle(N, kind):
    from test_gpu_parity import test_half_step_vs_oracle as driver
    driver(make, N, kind)


@pytest.mark.parametrize("N,kind,kw", [(5, hexref.KIND_SELFPLAY_B, dict(agent_mode=2)), (11, hexref.KIND_SELFPLAY_B, dict(agent_mode=1)),
                                       (6, hexref.KIND_ENV_A, dict(opponent_first=True))])
def test_rollout_equals_steps(N, kind, kw):
    parity.rollout_equals_steps(make, kind, N, 100, N * N // 2 + 3, seed=N, **kw)


@pytest.mark.parametrize("N", [5, 8])
def test_sample_board_flow(N):
    parity.sample_board_flow(make, N, 150, 40, seed=N)


@pytest.mark.parametrize("name", golden_files("preset_"))
def test_golden_preset_boards(name):
    def make_raw(kind, N, G):
        z = np.load(os.path.join(GOLDEN, name))
        env = EmuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True)
        env.reset()
        env.import_boards(z["board_true"], np.zeros(G, np.int8))
        return env
    parity.golden_preset(make_raw, name)


@pytest.mark.parametrize("form", [0, 1])
@pytest.mark.parametrize("N", [3, 4, 7, 8, 11, 13, 16, 19])
def test_both_relabel_sweep_forms(N, form):
    """The device runs the one-row-per-pass sweep on deep launches and the several-rows-per-pass sweep on small ones; the CPU
    batches here are all small, so force each form in turn (odd and even boards: rows with and without shared edge words)."""
    from emu import emu
    emu.force_sweep(form)
    try:
        parity.versus_oracle(make, hexref.KIND_SELFPLAY_B, N, 97, N * N // 2 + 8, seed=100 + N, fused=True, agent_mode=2, check_state_every=2)
        parity.snake_chain(make, N) if N in (7, 11) else None
    finally:
        emu.force_sweep(-1)


@pytest.mark.parametrize("seed", range(80))
def test_api_fuzz(seed):
    parity.api_fuzz(make, seed)


@pytest.mark.parametrize("N,variant_a", [(3, False), (4, True), (5, True), (6, False), (7, True), (8, False), (9, True), (10, False), (11, False),
                                         (12, True), (13, False), (14, False), (15, True), (16, False), (17, False), (18, True), (19, False)])
def test_sampler_and_views(N, variant_a):
    parity.sampler_and_views(make, N, variant_a, seed=N)


@pytest.mark.parametrize("N", [3, 5, 8, 11])
def test_raw_random_games(N):
    parity.raw_random_games(make, N, 200, seed=N)


@pytest.mark.parametrize("name", golden_files("saturation_"))
def test_golden_label_saturation(name):
    """~110 distinct region labels per colour on 18x18 / 19x19 (the packed state holds 7-bit labels), then everything merges."""
    parity.golden_saturation(make, name)


@pytest.mark.parametrize("name", golden_files("presetreset_"))
def test_golden_preset_resets(name):
    parity.golden_preset_resets(lambda kind, N, G: EmuBatch(0 if kind == hexref.KIND_GAME_A else 1, N, G, raw=True), name)


def test_word_encode_matches_byte_definition():
    """encode_word_k (the step kernel's 4-cells-per-word observation / mask encode: obs = mask*ka + C*kb + kc, no carries between
    bytes) against encode_byte for EVERY valid label byte in every byte position, both variants, random neighbours."""
    import ctypes
    import random
    from emu import emu as emu_mod
    L = emu_mod.lib()
    rnd = random.Random(7)
    valid = [0] + list(range(1, 128)) + list(range(0x81, 0x100))      # empty, R labels, C labels (0x80 = label 0 never occurs)

    def byte_def(b, variant):
        o, m = ctypes.c_uint32(), ctypes.c_uint32()
        L.emu_encode_byte(b, variant, ctypes.byref(o), ctypes.byref(m))
        return o.value, m.value

    for variant in (0, 1):
        table = {b: byte_def(b, variant) for b in valid}
        for b in valid:
            for pos in range(4):
                for _ in range(6):
                    bytes_ = [rnd.choice(valid) for _ in range(4)]
                    bytes_[pos] = b
                    x = sum(v << (8 * k) for k, v in enumerate(bytes_))
                    o, m = ctypes.c_uint32(), ctypes.c_uint32()
                    L.emu_encode_word(x, variant, ctypes.byref(o), ctypes.byref(m))
                    for k, v in enumerate(bytes_):
                        assert ((o.value >> (8 * k)) & 0xff, (m.value >> (8 * k)) & 0xff) == table[v], (variant, hex(x), k)


@pytest.mark.parametrize("name", golden_files("oppredict_"))
def test_opponent_predict_batched(name):
    """hexb_set_opponent_eps + hexb_half_step (variant A, HexEnv.opponent_predict for a batch) in the device code."""
    parity.golden_oppredict_batched(make, name)


@pytest.mark.parametrize("kind,N,seed", [(hexref.KIND_SELFPLAY_B, 4, 1), (hexref.KIND_SELFPLAY_B, 5, 2), (hexref.KIND_SELFPLAY_B, 7, 3),
                                         (hexref.KIND_ENV_A, 4, 4), (hexref.KIND_ENV_A, 5, 5), (hexref.KIND_ENV_A, 6, 6)])
def test_opponent_modes_fuzz(kind, N, seed):
    parity.opponent_modes_fuzz(make, kind, N, 200, 60, seed)


def test_device_logic_under_address_and_ub_sanitizers():
    """A cross-section of this file (golden replays with learned opponents, the split step, the opponent modes, the API fuzz)
    once more against the emulator built with -fsanitize=address,undefined (python tools/emu_sanitize.py runs all of it: 242 green)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("emu_sanitize", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                               "tools", "emu_sanitize.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rc, text = mod.run("golden_scripted or half_step or opponent_modes or api_fuzz or opponent_predict")
    if rc is None:
        pytest.skip("no sanitizer runtime in this toolchain: " + text[-200:])
    assert rc == 0 and "passed" in text and "runtime error" not in text and "AddressSanitizer" not in text, text[-3000:]
