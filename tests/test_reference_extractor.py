"""The flat observation row IS the output of the reference's own `FlattenedDictExtractor`
(gym_comm/extractors/CustomExtractor.py:77-128, imported from /root/reference with SB3's base class stubbed):
its `forward` over the per-key views of our rows returns the rows unchanged, and the feature width it computes
from the observation space is F.  Build container only."""
import argparse
import sys
import types

import pytest
import torch

from oracle import ref_harness

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="needs /root/reference")
D = dict(CAN_MOVE=True, ALLERGIC=False, BLIND=False)


def _reference_extractor():
    ref_harness._install_stubs()
    if "stable_baselines3.common.torch_layers" not in sys.modules:
        class BaseFeaturesExtractor(torch.nn.Module):           # stable_baselines3/common/torch_layers.py, the two lines used
            def __init__(self, observation_space, features_dim=0):
                super().__init__()
                self._observation_space, self._features_dim = observation_space, features_dim
        m = types.ModuleType("stable_baselines3.common.torch_layers")
        m.BaseFeaturesExtractor = BaseFeaturesExtractor
        sys.modules["stable_baselines3.common.torch_layers"] = m
    from gym_comm.extractors.CustomExtractor import FlattenedDictExtractor
    return FlattenedDictExtractor


@pytest.mark.parametrize("level,C", [("open-divider_tomato", 10), ("random-salad-superwide", 100)])
def test_flat_rows_are_the_reference_extractor_output(level, C):
    from gym_comm_b200.vec_env import OvercookedVecEnv
    from tests.parity_util import EmuVecEnv, emu_library
    ns = argparse.Namespace(level=level, num_agents=2, max_num_timesteps=30, communication_on=True, num_communication=C,
                            ego_led=False, fow_radius=2, ego_config=D, partner_config=D)
    E = 19
    env = EmuVecEnv(ns, num_envs=E, device="cpu", seed=2, auto_reset=True, lib=emu_library())
    ext = _reference_extractor()(env.observation_space)
    assert ext._features_dim == env.obs_width
    gen = torch.Generator().manual_seed(0)
    env.reset()
    for _ in range(12):
        a = torch.stack([torch.randint(0, 4, (E, 2), generator=gen), torch.randint(0, C, (E, 2), generator=gen)], -1).to(torch.int32)
        obs = env.step(a)[0]
        for k in range(2):                                       # each observer's row = the extractor over its dict
            views = env.obs_dict(obs[:, k])
            assert torch.equal(ext.forward(views), obs[:, k])
    env.close()
