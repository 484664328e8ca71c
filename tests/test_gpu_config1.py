"""GPU: BASELINE.json configs[0] -- the reference's own main_sac.py loop with its own ENetEnv (batch 64, prioritized=False,
--episodes 16 --steps 5) -- as a transcript.  The reference and its environment cannot travel to the GPU box and this container
has no GPU, so tools/record_main_sac_transcript.py ran the UNMODIFIED loop here (reference Agent, CPU) and recorded every call it
makes on the Agent API with the objects the environment really produces; this test replays those calls against the facade:
  * the constructor receives exactly main_sac.py's keyword arguments;
  * with the torch generator restored to the recorded state, construction draws the reference's initial weights, so until the
    replay holds a batch (63 transitions: learn() is a no-op, enet_sac.py:556) choose_action must return the reference's actions;
  * store_transition takes the env's dicts of CPU tensors / 0-dim tensor rewards, learn() runs once per step: 80 stores, 17 updates.
"""
import ast
import os

import numpy as np
import pytest
import torch

import smartcal_b200  # noqa: F401

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_reference_main_sac_loop_transcript_runs_against_the_facade():
    from smartcal_b200.elasticnet import enet_sac
    z = np.load(os.path.join(G, "main_sac_transcript.npz"))
    kw = {k: ast.literal_eval(v) for k, v in zip(z["ctor_keys"].tolist(), z["ctor_vals"].tolist())}
    assert kw["batch_size"] == 64 and kw["prioritized"] is False and kw["input_dims"] == [420]       # main_sac.py:34-36
    torch.set_rng_state(torch.from_numpy(z["ctor_rng"]))
    agent = enet_sac.Agent(**kw)                      # default rng='reference': host draws in the reference's order
    n = z["st_reward"].shape[0]
    worst = 0.0
    for i in range(n):
        obs = {"eig": torch.from_numpy(z["obs_eig"][i].copy()), "A": torch.from_numpy(z["obs_A"][i].copy())}
        if i < z["rng"].shape[0]:
            torch.set_rng_state(torch.from_numpy(z["rng"][i]))
        a = agent.choose_action(obs)
        assert a.shape == (2,) and a.dtype == np.float32 and np.all(np.abs(a) <= 1.0)
        if i < z["rng"].shape[0]:                     # same weights, same observation, same N(0,1) draw -> the reference's action
            worst = max(worst, float(np.max(np.abs(a - z["action"][i]))))
        st = {"eig": torch.from_numpy(z["st_eig"][i].copy()), "A": torch.from_numpy(z["st_A"][i].copy())}
        st2 = {"eig": torch.from_numpy(z["st_eig2"][i].copy()), "A": torch.from_numpy(z["st_A2"][i].copy())}
        reward = torch.tensor(z["st_reward"][i], dtype=torch.float32) if z["reward_is_tensor"][i] else float(z["st_reward"][i])
        agent.store_transition(st, z["st_action"][i], reward, st2, bool(z["st_done"][i]), z["st_hint"][i])
        assert agent.learn() is None
    assert worst < 5e-5, f"choose_action deviates from the reference's recorded actions by {worst}"
    mem = agent.replaymem
    assert mem.mem_cntr == n == 80 and agent.learn_counter == n - 63
    np.testing.assert_array_equal(mem.state_memory[:n, :20], z["st_eig"])
    np.testing.assert_array_equal(mem.state_memory[:n, 20:], z["st_A"])
    np.testing.assert_array_equal(mem.action_memory[:n], z["st_action"])
    np.testing.assert_allclose(mem.reward_memory[:n], z["st_reward"].astype(np.float32), rtol=0, atol=0)
    for net in (agent.actor, agent.critic_1, agent.target_critic_2):
        for k, v in net.state_dict().items():
            assert torch.isfinite(v).all(), k
