"""GPU (-m gpu): the stable-baselines3 VecEnv contract HexVecEnv has to satisfy, written out as assertions.

SB3 / sb3-contrib are not installed in the build image, so this is the recorded contract (stable-baselines3 2.2.1
`common/vec_env/base_vec_env.py`, `dummy_vec_env.py`; sb3-contrib 2.2.1 `common/maskable/utils.py`), the versions the reference's
saved models were trained with and what its scripts drive (scripts/experiments/*.py:34-47:
ActionMasker + MaskablePPO + MaskableEvalCallback). Every assertion names the SB3 call site that relies on it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def venv():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("-m gpu tests need a CUDA device")
    from hex_gym_env_b200.vec_env import HexVecEnv
    return HexVecEnv(board_size=5, num_envs=96, seed=3)


ABSTRACT = ["reset", "step_async", "step_wait", "close", "get_attr", "set_attr", "env_method", "env_is_wrapped"]   # VecEnv's abstract methods
CONCRETE = ["step", "seed", "set_options", "render", "get_images", "getattr_depth_check"]


def test_surface(venv):
    for m in ABSTRACT + CONCRETE:
        assert callable(getattr(venv, m)), m
    # VecEnv.__init__(num_envs, observation_space, action_space)
    assert venv.num_envs == 96
    assert tuple(venv.observation_space.shape) == (5, 5)
    assert venv.action_space.n == 25
    assert venv.render_mode is None
    assert venv.unwrapped is venv
    # SB3 >= 2.0: reset_infos / _seeds / _options, one entry per env
    assert isinstance(venv.reset_infos, list) and len(venv.reset_infos) == 96 and all(isinstance(d, dict) for d in venv.reset_infos)
    assert len(venv.seed(7)) == 96
    venv.set_options({"x": 1})
    assert len(venv._options) == 96


def test_reset_and_step_shapes_dtypes(venv):
    obs = venv.reset()
    # DummyVecEnv.reset -> np.ndarray [n_envs, *obs_shape]; on_policy_algorithm.collect_rollouts feeds it to obs_as_tensor
    assert isinstance(obs, np.ndarray) and obs.shape == (96, 5, 5) and obs.dtype == np.float32
    assert set(np.unique(obs).tolist()) <= {-1.0, 0.0, 1.0}
    # sb3_contrib.common.maskable.utils.get_action_masks(env) = np.stack(env.env_method("action_masks"))
    masks = venv.env_method("action_masks")
    assert isinstance(masks, list) and len(masks) == 96
    m = np.stack(masks)
    assert m.shape == (96, 25) and m.dtype == bool
    assert np.array_equal(m, obs.reshape(96, 25) == 0)             # legal == empty (HexSingleGame.py:205-206)
    # is_masking_supported(env): env.get_attr("action_masks") must not raise for a VecEnv
    assert len(venv.get_attr("action_masks")) == 96
    assert venv.env_method("action_masks", indices=[3, 5])[0].shape == (25,)
    assert venv.env_is_wrapped(object) == [False] * 96
    # step_async(actions: np.ndarray [n_envs]) ; step_wait() -> (obs, rews float [n_envs], dones bool [n_envs], infos list[dict])
    acts = np.array([int(np.flatnonzero(r)[0]) for r in m], dtype=np.int64)
    venv.step_async(acts)
    obs2, rew, done, infos = venv.step_wait()
    assert obs2.shape == (96, 5, 5) and obs2.dtype == np.float32
    assert isinstance(rew, np.ndarray) and rew.shape == (96,) and rew.dtype == np.float32
    assert isinstance(done, np.ndarray) and done.shape == (96,) and done.dtype == bool
    assert isinstance(infos, list) and len(infos) == 96 and all(isinstance(i, dict) for i in infos)


def test_auto_reset_and_terminal_observation(venv):
    """DummyVecEnv.step_wait: when an env is done, infos[i]["terminal_observation"] holds its last observation and the returned
    observation is already the first one of the next episode; MaskablePPO's rollout loop bootstraps from it on TimeLimit.truncated."""
    obs = venv.reset()
    seen_done = 0
    for t in range(40):
        m = np.stack(venv.env_method("action_masks"))
        acts = np.array([int(np.flatnonzero(r)[-1]) for r in m])
        obs, rew, done, infos = venv.step(acts)
        for i in np.flatnonzero(done):
            seen_done += 1
            info = infos[i]
            assert "terminal_observation" in info and info["terminal_observation"].shape == (5, 5)
            assert info["terminal_observation"].dtype == np.float32
            assert info.get("TimeLimit.truncated") is False
            assert (info["terminal_observation"] != 0).sum() > (obs[i] != 0).sum()     # the new episode's board is (almost) empty
            assert rew[i] in (-1.0, 0.0, 1.0)
        for i in np.flatnonzero(~done):
            assert infos[i] == {} and rew[i] == 0.0
        # the mask handed out next belongs to the observation just returned (new episode for finished envs)
        assert np.array_equal(np.stack(venv.env_method("action_masks")), obs.reshape(96, 25) == 0)
    assert seen_done > 20


def test_unknown_env_method_raises(venv):
    with pytest.raises(AttributeError):
        venv.env_method("no_such_method")


def test_torch_output_mode_keeps_everything_on_the_device():
    import torch
    from hex_gym_env_b200.vec_env import HexVecEnv
    v = HexVecEnv(board_size=6, num_envs=64, seed=1, output="torch")
    obs = v.reset()
    assert obs.is_cuda and obs.dtype == torch.float32 and tuple(obs.shape) == (64, 6, 6)
    m = v.action_masks()
    assert m.is_cuda and m.dtype == torch.bool
    a = torch.argmax(m.int(), dim=1).int()
    obs, rew, done, infos = v.step(a)
    assert obs.is_cuda and rew.is_cuda and done.is_cuda and done.dtype == torch.bool and len(infos) == 64
    v.close()
