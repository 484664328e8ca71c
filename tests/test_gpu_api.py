"""GPU: the reference-facing API surface (choose_action, store_transition, checkpoints, attribute names, PER class,
device-RNG mode, the main_sac.py-shaped loop) on the CUDA path."""
import os
import pickle

import numpy as np
import pytest
import torch

import parity_utils as PU
import smartcal_b200
from smartcal_b200 import _lib as L
from oracle import nets_ref as N

pytestmark = pytest.mark.gpu


def _obs(rs, n=4, m=5):
    return {"eig": torch.from_numpy(rs.standard_normal(n).astype(np.float32)), "A": torch.from_numpy(rs.standard_normal(n * m).astype(np.float32))}


def test_choose_action_matches_reference_sampling():
    """same torch seed -> same N(0,1) draw as Normal.sample() in the reference -> same action (enet_sac.py:547-553)"""
    orc, agent, tr, init_ok = PU.build_pair("sac", 420, 2, 32, 64, seed=21)
    assert init_ok
    rs = np.random.RandomState(0)
    for i in range(5):
        ob = {"eig": torch.from_numpy(rs.standard_normal(20).astype(np.float32)), "A": torch.from_numpy(rs.standard_normal(400).astype(np.float32))}
        torch.manual_seed(100 + i)
        a_dev = agent.choose_action(ob)
        torch.manual_seed(100 + i)
        eps = torch.randn(2)
        a_ref = orc.choose_action(torch.cat((ob["eig"], ob["A"])).numpy(), eps.numpy())
        assert a_dev.shape == (2,) and a_dev.dtype == np.float32
        np.testing.assert_allclose(a_dev, a_ref, rtol=5e-5, atol=5e-6)


def test_td3_choose_action_after_warmup():
    orc, agent, tr, _ = PU.build_pair("td3", 24, 2, 32, 64, seed=22, prioritized=False, reward="uniform")
    agent.warmup, agent.time_step = 0, 0
    ob = {"eig": torch.arange(4.0), "A": torch.linspace(-1, 1, 20)}
    np.random.seed(3)
    a_dev = agent.choose_action(ob)
    np.random.seed(3)
    nz = np.random.normal(scale=0.1, size=(2,)).astype(np.float32)
    with torch.no_grad():
        mu = N.td3_actor_forward(orc.actor, torch.cat((ob["eig"], ob["A"])).unsqueeze(0))[0].numpy()
    np.testing.assert_allclose(a_dev, np.clip(mu + nz, -1, 1), rtol=5e-5, atol=5e-6)


def test_store_transition_and_replay_attributes(tmp_path):
    from smartcal_b200.elasticnet import enet_sac
    torch.manual_seed(0)
    agent = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=8, n_actions=2, max_mem_size=16, tau=0.005)
    rs = np.random.RandomState(1)
    rows = []
    for i in range(20):     # wraps the ring of 16
        o, o2 = _obs(rs), _obs(rs)
        a, h = rs.uniform(-1, 1, 2).astype(np.float32), rs.uniform(-1, 1, 2).astype(np.float32)
        r = torch.tensor(float(i))
        agent.store_transition(o, a, r, o2, i % 3 == 0, h)
        rows.append((torch.cat((o["eig"], o["A"])).numpy(), a, float(i), torch.cat((o2["eig"], o2["A"])).numpy(), i % 3 == 0, h))
    mem = agent.replaymem
    assert mem.mem_cntr == 20 and mem.mem_size == 16
    for i in range(4, 20):
        slot = i % 16
        np.testing.assert_array_equal(mem.state_memory[slot], rows[i][0])
        np.testing.assert_array_equal(mem.action_memory[slot], rows[i][1])
        assert mem.reward_memory[slot] == rows[i][2] and bool(mem.terminal_memory[slot]) == rows[i][4]
        np.testing.assert_array_equal(mem.new_state_memory[slot], rows[i][3])
        np.testing.assert_array_equal(mem.hint_memory[slot], rows[i][5])
    assert agent.learn() is None and agent.learn_counter == 1       # 20 >= batch 8: one step ran
    # replay pickle is a plain host object with the reference attribute names
    os.chdir(tmp_path)
    mem.save_checkpoint()
    with open("replaymem_sac.model", "rb") as f:
        blob = pickle.load(f)
    assert blob.mem_cntr == 20 and blob.state_memory.shape == (16, 24) and blob.terminal_memory.dtype == bool


def test_checkpoint_roundtrip_and_reference_key_names(tmp_path):
    from smartcal_b200.elasticnet import enet_sac
    os.chdir(tmp_path)
    torch.manual_seed(1)
    a1 = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=8, n_actions=2, max_mem_size=16, name_prefix="x_")
    sd = a1.actor.state_dict()
    assert list(sd.keys()) == list(N.actor_shapes(24, 2, "sac").keys())
    assert list(a1.critic_1.state_dict().keys()) == list(N.critic_shapes(24, 2).keys())
    for k, shp in N.critic_shapes(24, 2).items():
        assert tuple(a1.critic_2.state_dict()[k].shape) == tuple(shp)
    a1.save_models()
    for f in ("x_a_eval_sac_actor.model", "x_q_eval_1_sac_critic.model", "x_q_eval_2_sac_critic.model", "replaymem_sac.model"):
        assert os.path.exists(f), f
    # the files are ordinary torch state_dicts: a plain torch load sees reference keys and CPU tensors
    raw = torch.load("x_a_eval_sac_actor.model")
    assert set(raw.keys()) == set(sd.keys()) and raw["fc1.weight"].device.type == "cpu"
    torch.manual_seed(2)
    a2 = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=8, n_actions=2, max_mem_size=16, name_prefix="x_")
    assert not torch.equal(a2.actor.state_dict()["fc1.weight"], sd["fc1.weight"])
    a2.load_models()
    for k in sd:
        assert torch.equal(a2.actor.state_dict()[k], sd[k])
    # load_models hard-copies the targets (enet_sac.py:645)
    for k, v in a2.critic_1.state_dict().items():
        assert torch.equal(a2.target_critic_1.state_dict()[k], v)
    with pytest.raises(RuntimeError):
        a2.actor.load_state_dict({"fc1.weight": sd["fc1.weight"]})          # strict
    a2.actor.load_state_dict({"fc1.weight": sd["fc1.weight"]}, strict=False)  # distributed_per_sac.py:128


def test_per_class_matches_reference_semantics():
    from smartcal_b200.elasticnet.enet_sac import PER
    from oracle.per_ref import PerRef
    with pytest.raises(AssertionError, match="power of 2"):
        PER(100, [4], 2)
    per = PER(64, [4], 2)
    orc = PerRef(64, 100.0)
    rs = np.random.RandomState(5)
    for i in range(70):
        e = None if i % 4 == 0 and i > 0 else float(rs.uniform(0, 3))
        per.store_transition_from_buffer(np.full(4, i, np.float32), np.zeros(2, np.float32), float(i), np.zeros(4, np.float32), False, np.zeros(2, np.float32), e)
        orc.store(e)
    np.random.seed(9)
    out = per.sample_buffer(16)
    assert np.array_equal(per.tree.tree, orc.tree.tree), "store path (max-leaf and given-error priorities) not bit-exact"
    np.random.seed(9)
    idx, data, w = orc.sample(np.random.random_sample(16))
    assert np.array_equal(out[6], idx) and np.abs(out[7].view(np.int32).astype(np.int64) - w.view(np.int32).astype(np.int64)).max() <= 2
    assert np.array_equal(out[0][:, 0], (data + 64 * (data < 6)).astype(np.float32))      # slots 0..5 were overwritten by stores 64..69
    assert len(per) == 64 and per.is_full() and per.mem_cntr == 70
    t = torch.arange(6.0).reshape(3, 2)
    assert float(per.mse(t[:, :1], t[:, 1:], torch.ones(3))) == 1.0


def test_device_rng_mode_runs_and_learns():
    """rng='device': no host draws; indices are a valid sample without replacement, losses finite and decreasing-ish"""
    from smartcal_b200.elasticnet import enet_sac
    from smartcal_b200.synth import make_transitions, pack
    torch.manual_seed(3)
    agent = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=32, n_actions=2, max_mem_size=128,
                           tau=0.005, reward_scale=1, alpha=0.03, rng="device")
    agent._dev.fill(pack(make_transitions(0, 128, 24, 2)))
    np_state = np.random.get_state()[1].copy()
    losses = []
    for _ in range(30):
        st = agent.learn(want_stats=True)
        losses.append(st["critic_1_loss"])
        idx = agent._dev.buffer("batch_idx")[:, 0].cpu().numpy()
        assert len(set(idx.tolist())) == 32 and idx.min() >= 0 and idx.max() < 128
    assert np.array_equal(np.random.get_state()[1], np_state), "device mode must not consume the host numpy stream"
    assert np.isfinite(losses).all() and np.mean(losses[-5:]) < np.mean(losses[:5])
    eps = agent._dev.buffer("eps_next").cpu().numpy()
    assert abs(eps.mean()) < 0.5 and 0.6 < eps.std() < 1.4


def test_main_sac_shaped_loop_with_stub_env():
    """the training loop of elasticnet/main_sac.py:47-76 against the facade, env replaced by a shape-compatible stub"""
    from smartcal_b200.elasticnet import enet_sac
    from smartcal_b200.elasticnet.distributed_per_sac import SyntheticENetEnv
    np.random.seed(0)
    torch.manual_seed(0)
    N_ = M_ = 20
    env = SyntheticENetEnv(M_, N_, provide_hint=False, seed=1)
    agent = enet_sac.Agent(gamma=0.99, batch_size=64, n_actions=2, tau=0.005, max_mem_size=1024, input_dims=[N_ + N_ * M_],
                           lr_a=1e-3, lr_c=1e-3, reward_scale=N_, alpha=0.03, prioritized=False, use_hint=False)
    scores = []
    for i in range(16):
        score, done, observation, loop = 0, False, env.reset(), 0
        while (not done) and loop < 5:
            action = agent.choose_action(observation)
            observation_, reward, done, info = env.step(action)
            agent.store_transition(observation, action, reward, observation_, done, np.zeros_like(action))
            score += reward
            agent.learn()
            observation = observation_
            loop += 1
        scores.append(score.cpu().data.item() / loop)
    assert agent.replaymem.mem_cntr == 80 and agent.learn_counter == 80 - 63
    assert np.isfinite(scores).all()


def test_per_presample_is_transparent(monkeypatch):
    """rng='device', prioritized: back-to-back learn() calls draw the next minibatch on the side stream during the current
    step.  That must be invisible: same parameters bit for bit as with SMARTCAL_NO_PRESAMPLE=1, including across a store
    between two steps (which discards the pending draw) and a raw tree access."""
    from smartcal_b200.elasticnet import enet_sac
    from smartcal_b200.synth import make_transitions, pack

    def run(no_presample):
        if no_presample:
            monkeypatch.setenv("SMARTCAL_NO_PRESAMPLE", "1")
        else:
            monkeypatch.delenv("SMARTCAL_NO_PRESAMPLE", raising=False)
        torch.manual_seed(5)
        agent = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=32, n_actions=2, max_mem_size=128,
                               tau=0.005, reward_scale=1, alpha=0.03, rng="device", prioritized=True)
        dev = agent._dev
        tr = pack(make_transitions(1, 129, 24, 2))
        dev.fill(tr[:128])
        batches = []
        for _ in range(5):
            dev.learn()
        batches.append(dev.buffer("batch_idx")[:, 0].cpu().numpy().copy())        # (raw buffer access: drops the pending draw)
        for _ in range(3):
            dev.learn()
        dev.store(tr[128])                                                          # a store between two steps
        for _ in range(4):
            dev.learn()
        batches.append(dev.buffer("batch_idx")[:, 0].cpu().numpy().copy())
        beta = dev.counters()[3]
        params = {k: v.detach().cpu().clone() for net in (agent.actor, agent.critic_1, agent.critic_2, agent.target_critic_1)
                  for k, v in net.state_dict().items()}
        tree = dev.buffer("tree")[:, 0].cpu().numpy().copy()
        return batches, beta, params, tree, dev.launch_count()

    b0, beta0, p0, t0, n0 = run(False)
    b1, beta1, p1, t1, n1 = run(True)
    assert n0 > n1, "the presampling path did not run"
    assert beta0 == beta1 == pytest.approx(0.4 + 12e-4)
    for x, y in zip(b0, b1):
        assert np.array_equal(x, y)
    assert np.array_equal(t0, t1)
    for (k, v), (k2, v2) in zip(p0.items(), p1.items()):
        assert k == k2 and torch.equal(v, v2), k
