"""GPU: the CUDA path against the golden vectors recorded from the unmodified reference (no oracle in the loop),
and PER sample/update against the reference's recorded indices, weights and trees."""
import os

import numpy as np
import pytest
import torch

import smartcal_b200
from smartcal_b200 import _lib as L
from smartcal_b200.synth import make_transitions, pack

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


@pytest.mark.parametrize("name", ["per_c64_b16", "per_c4096_b256"])
def test_per_kernels_reproduce_reference_trees_and_indices(name):
    from smartcal_b200.core import DeviceAgent
    z = np.load(os.path.join(G, name + ".npz"))
    C, B, rounds, seed = [int(x) for x in z["cfg"]]
    dev = DeviceAgent(L.SC_KIND_SAC, 4, 2, B, C, 0.99, 0.005, 1e-3, 1e-3, prioritized=True)
    zero = np.zeros(dev.T, np.float32)
    # stores: error=None -> max leaf (or clip), error given -> min((|e|+eps)^alpha, clip) evaluated like the reference (python floats)
    for e, none in zip(z["store_err"], z["store_none"]):
        if none:
            dev.store(zero, 1, 0.0)
        else:
            dev.store(zero, 2, min((abs(float(e)) + 0.01) ** 0.6, 100.0))
    tree = dev.buffer("tree")[:, 0].cpu().numpy()
    assert np.array_equal(tree, z["tree_after_store"]), "sequential store path is not bit-exact"
    for r in range(rounds):
        ti, w = dev.per_sample(z[f"r{r}/u01"], want_outputs=True)
        assert np.array_equal(ti, z[f"r{r}/idx"]), f"round {r}: tree indices differ"
        assert np.abs(w.view(np.int32).astype(np.int64) - z[f"r{r}/w"].view(np.int32).astype(np.int64)).max() <= 2
        assert abs(dev.counters()[3] - float(z[f"r{r}/beta"][0])) < 1e-15
        ps = dev.per_update(ti, z[f"r{r}/abs_err"])
        ulp = np.abs(ps.view(np.int32).astype(np.int64) - z[f"r{r}/ps"].view(np.int32).astype(np.int64)).max()
        assert ulp <= 2, f"priorities differ by {ulp} ulp"
        tree = dev.buffer("tree")[:, 0].cpu().numpy()
        if ulp == 0:
            assert np.array_equal(tree, z[f"r{r}/tree"]), f"round {r}: tree not bit-exact after batch_update"
        else:   # float32 pow differs by an ulp on this host: re-pin so later rounds stay comparable
            np.testing.assert_allclose(tree, z[f"r{r}/tree"], rtol=1e-6)
            dev.buffer("tree").copy_(torch.as_tensor(z[f"r{r}/tree"]).reshape(-1, 1))
    # arbitrary (unsorted, duplicated) index list takes the exact sequential path
    ps = dev.per_update(z["unsorted/idx"], z["unsorted/abs_err"])
    np.testing.assert_allclose(dev.buffer("tree")[:, 0].cpu().numpy(), z["unsorted/tree"], rtol=1e-6)


@pytest.mark.parametrize("name", ["sac_s24_b32", "sac_enet_b256"])
def test_sac_learn_against_reference_golden(name):
    from smartcal_b200.elasticnet import enet_sac
    z = np.load(os.path.join(G, name + ".npz"))
    S, A, B, C, steps, seed, use_hint = [int(x) for x in z["cfg"]]
    tr = make_transitions(seed + 100, C, S, A, "normal", float(z["done_frac"][0]))
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C,
                           tau=0.005, reward_scale=20, alpha=0.03, rng="reference")
    for nm, view in (("actor", agent.actor), ("critic_1", agent.critic_1), ("critic_2", agent.critic_2)):
        for k, v in view.state_dict().items():
            assert np.array_equal(digest(v), z[f"init/{nm}/{k}/p"]), f"init {nm}.{k} differs from the reference"
    agent._dev.fill(pack(tr))
    # same seeds as the generator -> the facade must draw the very same indices and noise from numpy/torch
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    st = agent.learn(want_stats=True)
    idx_dev = agent._dev.buffer("batch_idx")[:, 0].cpu().numpy()
    assert np.array_equal(idx_dev, z["step0/idx"]), "facade did not reproduce np.random.choice draw of the reference"
    assert np.array_equal(agent._dev.buffer("eps_next").cpu().numpy(), z["step0/eps1"])
    assert np.array_equal(agent._dev.buffer("eps_cur").cpu().numpy(), z["step0/eps2"])
    np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"]], z["step0/losses"], rtol=2e-4)
    # first-step Adam moment m = 0.1*g pins every gradient; digests: sums relative to abs-sum, samples relative to max
    for nm, view in (("actor", agent.actor), ("critic_1", agent.critic_1), ("critic_2", agent.critic_2)):
        adam = view.adam_state()
        for k in adam:
            ref = z[f"step0/{nm}/{k}/m"]
            got = digest(adam[k][0])
            assert abs(got[0] - ref[0]) <= 2e-4 * ref[1] + 1e-12, f"{nm}.{k} sum(m)"
            assert abs(got[1] - ref[1]) <= 2e-4 * ref[1] + 1e-12, f"{nm}.{k} sum|m|"
            assert np.max(np.abs(got[2:] - ref[2:])) <= 2e-4 * max(np.max(np.abs(ref[2:])), ref[1] / adam[k][0].numel()) + 1e-12, f"{nm}.{k} samples"


def test_ddpg_learn_against_reference_golden():
    """enet_ddpg.Agent on the TD3 program (mirrored twin critic, sum-of-squares critic loss) vs the unmodified reference."""
    from smartcal_b200.elasticnet import enet_ddpg
    z = np.load(os.path.join(G, "ddpg_s24_b32.npz"))
    S, A, B, C, steps, seed = [int(x) for x in z["cfg"]]
    tr = make_transitions(seed + 100, C, S, A, "normal", float(z["done_frac"][0]))
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = enet_ddpg.Agent(gamma=0.99, lr_a=1e-4, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.001,
                            rng="reference")
    for nm, view in (("actor", agent.actor), ("critic", agent.critic)):
        for k, v in view.state_dict().items():
            assert np.array_equal(digest(v), z[f"init/{nm}/{k}/p"]), f"init {nm}.{k} differs from the reference"
    tr = dict(tr)
    tr["hint"] = np.zeros_like(tr["hint"])
    agent._dev.fill(pack(tr))
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    st = agent.learn(want_stats=True)
    assert np.array_equal(agent._dev.buffer("batch_idx")[:, 0].cpu().numpy(), z["step0/idx"])
    np.testing.assert_allclose([st["critic_loss"], st["actor_loss"]], z["step0/losses"], rtol=2e-4)
    for nm, view in (("actor", agent.actor), ("critic", agent.critic)):
        adam = view.adam_state()
        for k in adam:
            ref = z[f"step0/{nm}/{k}/m"]
            got = digest(adam[k][0])
            assert abs(got[0] - ref[0]) <= 2e-4 * ref[1] + 1e-12, f"{nm}.{k} sum(m)"
            assert abs(got[1] - ref[1]) <= 2e-4 * ref[1] + 1e-12, f"{nm}.{k} sum|m|"
            assert np.max(np.abs(got[2:] - ref[2:])) <= 2e-4 * max(np.max(np.abs(ref[2:])), ref[1] / adam[k][0].numel()) + 1e-12, f"{nm}.{k} samples"
    # the mirrored critic slot stays bit-identical to the critic, the targets moved with tau
    for k, v in agent.critic.state_dict().items():
        assert torch.equal(v, agent._mirror.state_dict()[k]), f"mirror critic diverged at {k}"
    for nm, view in (("target_actor", agent.target_actor), ("target_critic", agent.target_critic)):
        for k, v in view.state_dict().items():
            ref = z[f"step0/{nm}/{k}/p"]
            got = digest(v)
            assert np.max(np.abs(got[2:] - ref[2:])) <= 1e-6 * max(1.0, np.max(np.abs(ref[2:]))), f"{nm}.{k}"
    for s in range(1, steps):
        agent.learn()
    assert agent._dev.counters()[1] == steps
