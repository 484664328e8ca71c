"""GPU: edge cases of the hot path -- the smallest and the largest shapes the ABI accepts, a replay that is still filling or has
wrapped, and minibatches with repeated PER leaves (collisions) -- each against the oracle or against the reference's documented
behaviour."""
import numpy as np
import pytest
import torch

import parity_utils as PU
import smartcal_b200
from smartcal_b200 import _lib as L
from smartcal_b200.synth import make_transitions, pack

pytestmark = pytest.mark.gpu


def _assert(rep):
    assert rep["ok"], f"{len(rep['failures'])} mismatches, first: {rep['failures'][:8]}"
    assert rep["launches"] > 0, "no CUDA kernels were launched by libsmartcal_b200"


def _agent(B=32, C=64, S=24, A=2, prioritized=False, **kw):
    from smartcal_b200.elasticnet import enet_sac
    torch.manual_seed(3)
    return enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
                          reward_scale=1, alpha=0.03, rng="device", prioritized=prioritized, **kw)


def _params(agent):
    return torch.cat([v.reshape(-1).cpu() for n in (agent.actor, agent.critic_1, agent.critic_2, agent.target_critic_1, agent.target_critic_2)
                      for v in n.state_dict().values()])


@pytest.mark.parametrize("prioritized", [False, True])
def test_learn_is_a_no_op_until_a_batch_is_stored(prioritized):
    """enet_sac.py:556-557 `if self.replaymem.mem_cntr < self.batch_size: return`: nothing moves, nothing is launched on the nets"""
    agent = _agent(B=16, C=64, prioritized=prioritized)
    tr = make_transitions(5, 64, 24, 2, "normal", 0.1)
    before = _params(agent)
    for i in range(15):
        agent.store_transition(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]), tr["hint"][i])
        assert agent.learn() is None
    agent._dev.sync()
    assert torch.equal(before, _params(agent))
    assert agent._dev.counters()[0] == 15 and agent._dev.counters()[1] == 0
    agent.store_transition(tr["state"][15], tr["action"][15], float(tr["reward"][15]), tr["new_state"][15], bool(tr["done"][15]), tr["hint"][15])
    agent.learn()
    agent._dev.sync()
    assert not torch.equal(before, _params(agent)) and agent._dev.counters()[1] == 1


def test_replay_ring_wraps_like_the_reference():
    """ReplayBuffer.store_transition (enet_sac.py:35-44): index = mem_cntr % mem_size, mem_cntr keeps counting"""
    C, S, A = 16, 24, 2
    agent = _agent(B=8, C=C)
    tr = make_transitions(6, 40, S, A, "normal", 0.2)
    for i in range(40):
        agent.store_transition(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]), tr["hint"][i])
    mem = agent.replaymem
    assert mem.mem_cntr == 40
    owner = {i % C: i for i in range(40)}                      # last writer of every slot
    idx = np.array([owner[s] for s in range(C)])
    assert np.array_equal(np.asarray(mem.state_memory), tr["state"][idx])
    assert np.array_equal(np.asarray(mem.new_state_memory), tr["new_state"][idx])
    assert np.array_equal(np.asarray(mem.action_memory), tr["action"][idx])
    assert np.array_equal(np.asarray(mem.reward_memory).reshape(-1), tr["reward"][idx].astype(np.float32))
    assert np.array_equal(np.asarray(mem.terminal_memory).reshape(-1).astype(bool), tr["done"][idx].astype(bool))


def test_smallest_batches_against_the_oracle():
    # one and two rows per minibatch: a single valid lane in a 128-row tensor-core tile, LayerNorm over features only
    _assert(PU.sac_parity_report(S=24, A=2, B=1, C=8, steps=2, seed=41))
    _assert(PU.sac_parity_report(S=24, A=2, B=2, C=8, steps=2, seed=42, prioritized=True))
    _assert(PU.td3_parity_report(S=24, A=2, B=2, C=8, steps=2, seed=43, prioritized=True))


def test_widest_action_and_odd_state_dims_against_the_oracle():
    # n_actions = 32 is the ABI's maximum (AMAX); state_dim = 97 is neither a multiple of 4 nor of 32 (row pitch / k-block padding)
    _assert(PU.sac_parity_report(S=97, A=32, B=40, C=64, steps=2, seed=44, prioritized=True))
    _assert(PU.sac_parity_report(S=5, A=1, B=33, C=64, steps=2, seed=45))


def test_largest_batch_against_the_oracle():
    # batch_size = 8192 is the ABI's maximum: 64 row blocks per pass, PER capacity 8192 fully sampled
    _assert(PU.sac_parity_report(S=420, A=2, B=8192, C=8192, steps=1, seed=46, prioritized=True, trace=False))


def test_per_minibatch_with_repeated_leaves():
    """Few stored transitions, one dominant priority: the stratified draw returns the same leaf many times.  PER.batch_update
    (enet_sac.py:314-323) then writes that leaf once per occurrence, in order -- the last occurrence's priority stays and the
    ancestors see every intermediate change.  Against the sequential oracle: same leaves written, tree equal to the float32-pow
    agreement of the two platforms (the priorities are float32 powers; 2 ulp there is 1e-7 of the root)."""
    from oracle.per_ref import PerRef
    C, B = 64, 32
    agent = _agent(B=B, C=C, prioritized=True)
    dev = agent._dev
    orc = PerRef(C, 100.0)
    tr = make_transitions(7, C, 24, 2, "normal", 0.0)
    rs = np.random.RandomState(2)
    for i in range(6):
        e = 50.0 if i == 3 else float(rs.uniform(0.01, 0.1))
        agent.replaymem.store_transition_from_buffer(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], False, tr["hint"][i], e)
        orc.store(e)
    u = rs.random_sample(B)
    idx, data, w = orc.sample(u)
    ti, wd = dev.per_sample(u, want_outputs=True)
    assert np.array_equal(ti, idx)
    assert len(np.unique(idx)) < B // 2, "the fixture is meant to repeat leaves"
    err = rs.uniform(0.0, 2.0, B).astype(np.float32)
    ps_ref = orc.update(idx, err)
    ps = dev.per_update(ti, err)
    assert np.abs(ps.view(np.int32).astype(np.int64) - np.asarray(ps_ref, np.float32).view(np.int32).astype(np.int64)).max() <= 2
    tree = np.asarray(dev.buffer("tree").cpu()).reshape(-1)
    assert np.array_equal(tree == 0.0, orc.tree.tree == 0.0)
    np.testing.assert_allclose(tree, orc.tree.tree, rtol=1e-6, atol=0.0)
    # the leaf that was drawn repeatedly holds the priority of its LAST occurrence in the minibatch
    last = {int(n): float(p) for n, p in zip(ti, ps)}
    for node, p in last.items():
        assert tree[node] == np.float64(np.float32(p))
