"""GPU: round-2 hardening -- ensembles (E agents per launch), CUDA-graph replay, the fused store+learn call, free-running
(un-forced) trajectories against the oracle, statistical tests of the device RNG paths, PER checkpoints."""
import os
import pickle
import sys

import numpy as np
import pytest
import torch

import parity_utils as PU
import smartcal_b200
from smartcal_b200 import _lib as L
from smartcal_b200.core import DeviceEnsemble
from smartcal_b200.synth import make_transitions, pack

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HP = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=32, n_actions=2, max_mem_size=128, tau=0.005,
          reward_scale=1, alpha=0.03)


def _state(agent):
    nets = [agent.actor, agent.critic_1, agent.critic_2, agent.target_critic_1, agent.target_critic_2]
    out = {}
    for i, n in enumerate(nets):
        for k, v in n.state_dict().items():
            out[f"{i}.{k}"] = v.detach().cpu().clone()
        if i < 3:
            for k, (m, v) in n.adam_state().items():
                out[f"{i}.{k}.m"], out[f"{i}.{k}.v"] = m.detach().cpu().clone(), v.detach().cpu().clone()
    return out


def _make(seed, prioritized, kind="sac", **kw):
    from smartcal_b200.elasticnet import enet_sac, enet_td3
    torch.manual_seed(seed)
    hp = dict(HP)
    hp.update(kw)
    if kind == "sac":
        agent = enet_sac.Agent(**hp, rng="device", prioritized=prioritized)
    else:
        hp.pop("reward_scale"), hp.pop("alpha")
        agent = enet_td3.Agent(**hp, rng="device", prioritized=prioritized, warmup=0)
    C, S, A = hp["max_mem_size"], hp["input_dims"][0], hp["n_actions"]
    agent._dev.fill(pack(make_transitions(seed + 50, C, S, A, "uniform" if kind == "td3" else "normal", 0.1)))
    return agent


@pytest.mark.parametrize("kind,prioritized", [("sac", False), ("sac", True), ("td3", True)])
def test_ensemble_agents_bit_identical_to_standalone(kind, prioritized):
    """E = 3 agents with different seeds / weights / replay contents, stepped by ONE launch sequence (blockIdx.y = agent),
    end bit-identical (parameters, Adam moments, targets, PER tree, minibatch indices) to the same agents stepped alone."""
    steps, seeds = 7, (11, 12, 13)
    td3_noise = [[0.05 * (i + 1) * (-1) ** s for s in range(steps)] for i in range(3)]
    alone = []
    for i, sd in enumerate(seeds):
        a = _make(sd, prioritized, kind)
        for s in range(steps):
            a._dev.learn(td3_noise=td3_noise[i][s] if kind == "td3" else None)
        tree = a._dev.buffer("tree")[:, 0].cpu().numpy().copy() if prioritized else None
        alone.append((_state(a), tree, a._dev.buffer("batch_idx")[:, 0].cpu().numpy().copy()))
        a._dev.close()
    agents = [_make(sd, prioritized, kind) for sd in seeds]
    ens = DeviceEnsemble([a._dev for a in agents])
    for s in range(steps):
        ens.learn([dict(td3_noise=td3_noise[i][s]) for i in range(3)] if kind == "td3" else None)
    assert ens.launch_count() > 0
    for i, a in enumerate(agents):
        st = _state(a)
        for k, v in alone[i][0].items():
            assert torch.equal(st[k], v), (i, k)
        if prioritized:
            assert np.array_equal(a._dev.buffer("tree")[:, 0].cpu().numpy(), alone[i][1]), i
        assert np.array_equal(a._dev.buffer("batch_idx")[:, 0].cpu().numpy(), alone[i][2]), i
        assert a._dev.counters()[1] == steps
    # the agents really are different problems
    assert not torch.equal(_state(agents[0])["0.fc1.weight"], _state(agents[1])["0.fc1.weight"])
    ens.close()


def test_ensemble_rejects_mismatched_agents_and_divergence():
    a, b = _make(1, False), _make(2, False, batch_size=16)
    with pytest.raises(L.SmartcalError, match="differs from agent 0"):
        DeviceEnsemble([a._dev, b._dev])
    c = _make(3, False)
    ens = DeviceEnsemble([a._dev, c._dev])
    c._dev.set_batch(np.arange(32))            # only one agent has an externally supplied batch: the plans diverge
    with pytest.raises(L.SmartcalError, match="lock-step"):
        ens.learn()


def test_graph_replay_equals_eager_launches(monkeypatch):
    """identical plan shapes are replayed from a CUDA graph from their 3rd occurrence on; SMARTCAL_GRAPH=0 keeps every step on
    eager launches.  Same bits either way (PER presampling, the side streams and PDL edges are part of the graph)."""
    def run(graph):
        monkeypatch.setenv("SMARTCAL_GRAPH", "1" if graph else "0")
        a = _make(21, True)
        for _ in range(9):
            a._dev.learn()
        out = _state(a), a._dev.buffer("tree")[:, 0].cpu().numpy().copy()
        a._dev.close()
        return out

    (s1, t1), (s0, t0) = run(True), run(False)
    assert np.array_equal(t0, t1)
    for k in s0:
        assert torch.equal(s0[k], s1[k]), k


@pytest.mark.parametrize("prioritized", [False, True])
def test_fused_store_learn_equals_store_then_learn(prioritized):
    """facade store_transition() is deferred into the next learn() (sc_agent_store_learn: one pinned block, one H2D, one graph
    launch); the result is what an immediate store followed by learn() gives, in reference-RNG mode with host draws"""
    from smartcal_b200.elasticnet import enet_sac
    tr = make_transitions(7, 40, 24, 2, "normal", 0.1)

    def run(fused):
        torch.manual_seed(31)
        np.random.seed(31)
        agent = enet_sac.Agent(**HP, rng="reference", prioritized=prioritized)
        agent._dev.fill(pack(make_transitions(8, 128, 24, 2, "normal", 0.1)))
        stats = []
        for i in range(12):
            ob = {"eig": torch.from_numpy(tr["state"][i, :4].copy()), "A": torch.from_numpy(tr["state"][i, 4:].copy())}
            ob2 = {"eig": torch.from_numpy(tr["new_state"][i, :4].copy()), "A": torch.from_numpy(tr["new_state"][i, 4:].copy())}
            agent.store_transition(ob, tr["action"][i], float(tr["reward"][i]), ob2, bool(tr["done"][i]), tr["hint"][i])
            if not fused:
                agent._dev.flush()
            stats.append(agent.learn(want_stats=True))
        return _state(agent), stats, agent.replaymem.state_memory.copy(), agent._dev.buffer("batch_idx")[:, 0].cpu().numpy().copy()

    (sa, sta, ma, ia), (sb, stb, mb, ib) = run(True), run(False)
    assert np.array_equal(ma, mb) and np.array_equal(ia, ib)
    for x, y in zip(sta, stb):
        assert x == y
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k


# ---- free-running trajectories ------------------------------------------------------------------------------------------------
# Stated band: over 50 un-forced learn() calls at S=420, B=256 from identical initial state and identical minibatches / noise, the
# CUDA path's critic losses and the batch mean of Q1 stay within 5e-4 relative (+1e-5 / +2e-4 abs) of the CPU oracle's; with PER
# each side samples from ITS OWN tree with the same uniforms (float32-ulp priority differences may flip an occasional index: at
# most 0.2 % of the draws) and the band is 2e-3.  Measured on B200 (round 2): worst relative deviation 4e-6 (uniform) / 7e-5 (PER),
# no index flips in 12 800 draws -- Adam's amplification of rounding-level gradient noise does not compound visibly.
@pytest.mark.parametrize("prioritized", [False, True])
def test_free_running_50_steps_stay_within_band_of_oracle(prioritized):
    S, A, B, C, steps = 420, 2, 256, 512, 50
    orc, agent, tr, init_ok = PU.build_pair("sac", S, A, B, C, seed=41, prioritized=prioritized)
    assert init_ok
    dev = agent._dev
    np.random.seed(42)
    torch.manual_seed(42)
    worst = dict(c1=0.0, c2=0.0, q1=0.0)
    flips = 0
    band_l, band_q = (2e-3, 2e-3) if prioritized else (5e-4, 5e-4)
    for s in range(steps):
        e1, e2 = torch.randn(B, A), torch.randn(B, A)
        if prioritized:
            u = np.random.random_sample(B)
            idxs, data, w = orc.per.sample(u)
            res = orc.learn_on_batch(orc.replay.gather(data), e1, e2, None, is_weights=w, tree_idx=idxs)
            st = dev.learn(e1.numpy(), e2.numpy(), want_stats=True, per_u01=u)
            flips += int((dev.buffer("tree_idx")[:, 0].cpu().numpy() != idxs).sum())
        else:
            idx = np.random.choice(C, B, replace=False)
            res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2, None)
            st = dev.learn(e1.numpy(), e2.numpy(), want_stats=True, batch_idx=idx)
            assert np.array_equal(dev.buffer("batch_idx")[:, 0].cpu().numpy(), idx)
        for key, ref, got, band, ab in (("c1", res["critic_1_loss"], st["critic_1_loss"], band_l, 1e-5),
                                        ("c2", res["critic_2_loss"], st["critic_2_loss"], band_l, 1e-5),
                                        ("q1", float(res["q1"].mean()) if "q1" in res else None, st["q1_mean"], band_q, 2e-4)):
            if ref is None:
                continue
            dev_rel = abs(got - float(ref)) / (band * abs(float(ref)) + ab)
            worst[key] = max(worst[key], dev_rel)
    print(f"free-running {steps} steps prioritized={prioritized}: worst deviation / band = {worst}, PER index flips = {flips} of {steps * B}")
    assert all(np.isfinite(v) and v <= 1.0 for v in worst.values()), worst
    assert flips <= 0.002 * steps * B


# ---- device RNG: distribution tests at n >= 1e5 -----------------------------------------------------------------------------------
def test_device_normal_draws_are_standard_normal():
    """Philox4x32-10 + Box-Muller as the kernels draw it (rng='device'): 262 144 values, KS test against N(0,1) and moments"""
    from scipy import stats
    from smartcal_b200.elasticnet import enet_sac
    torch.manual_seed(5)
    B, A = 8192, 32
    agent = enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[8], batch_size=B, n_actions=A, max_mem_size=B, tau=0.005,
                           reward_scale=1, alpha=0.03, rng="device")
    agent._dev.fill(pack(make_transitions(0, B, 8, A)))
    agent._dev.learn()
    x = np.concatenate([agent._dev.buffer("eps_next").cpu().numpy().ravel(), agent._dev.buffer("eps_cur").cpu().numpy().ravel()]).astype(np.float64)
    assert x.size >= 5e5
    ks = stats.kstest(x, "norm")
    assert ks.pvalue > 1e-3, ks
    assert abs(x.mean()) < 4 / np.sqrt(x.size) and abs(x.std() - 1) < 4 / np.sqrt(2 * x.size)
    assert abs(stats.skew(x)) < 0.02 and abs(stats.kurtosis(x)) < 0.04
    # the two streams (and consecutive steps) are distinct draws
    n = x.size // 2
    assert abs(np.corrcoef(x[:n], x[n:])[0, 1]) < 0.01


def test_device_uniform_sampling_without_replacement_is_uniform():
    """key-sort sampling (k_sample_uniform): every batch is duplicate-free and slot frequencies over 4000 draws pass a chi-square test"""
    from scipy import stats
    a = _make(61, False)
    dev = a._dev
    C, B, n = 128, 32, 4000
    counts = np.zeros(C)
    for _ in range(n):
        L.check(dev.lib.sc_replay_sample_uniform_dev(dev.h))
        idx = dev.buffer("batch_idx")[:, 0].cpu().numpy()
        assert len(set(idx.tolist())) == B
        counts[idx] += 1
    chi2 = ((counts - n * B / C) ** 2 / (n * B / C * (1 - B / C))).sum()      # hypergeometric variance per slot
    assert stats.chi2.sf(chi2, C - 1) > 1e-3, chi2


def test_device_per_sampling_follows_the_stratified_law():
    """k_per_sample with device Philox uniforms: sample i always lands in stratum i of the cumulative priority (PER.sample_buffer,
    enet_sac.py:289-297), and inside a stratum leaves are hit in proportion to their priority (chi-square over 3000 draws)"""
    from scipy import stats
    a = _make(62, True)
    dev = a._dev
    C, B, n = 128, 32, 3000
    rs = np.random.RandomState(3)
    pr = rs.uniform(0.05, 3.0, C)
    dev.fill(pack(make_transitions(1, C, 24, 2)), pr)
    tree = dev.buffer("tree")[:, 0].cpu().numpy()
    leaves = tree[C - 1:]
    cum = np.concatenate(([0.0], np.cumsum(leaves)))
    seg = tree[0] / B
    hits = np.zeros((B, C))
    for _ in range(n):
        ti, w = dev.per_sample(None, want_outputs=True)
        d = ti - (C - 1)
        lo, hi = cum[d], cum[d + 1]
        i = np.arange(B)
        assert np.all(hi > seg * i - 1e-9) and np.all(lo < seg * (i + 1) + 1e-9), "a draw left its stratum"
        hits[i, d] += 1
        assert np.all(w > 0) and w.max() == 1.0
    worst_p = 1.0
    for i in (0, 7, 19, 31):
        ov = np.clip(np.minimum(cum[1:], seg * (i + 1)) - np.maximum(cum[:-1], seg * i), 0, None)     # leaf overlap with stratum i
        m = ov > 1e-12
        exp = n * ov[m] / ov[m].sum()
        chi2 = ((hits[i, m] - exp) ** 2 / exp).sum()
        worst_p = min(worst_p, stats.chi2.sf(chi2, m.sum() - 1))
        assert hits[i, ~m].sum() == 0
    assert worst_p > 1e-4, worst_p


# ---- PER checkpoints (PER.load_checkpoint crashed) ---------------------------------------------------------------------
def test_per_agent_save_then_load_roundtrip(tmp_path):
    os.chdir(tmp_path)
    a = _make(71, True)
    for _ in range(3):
        a._dev.learn()
    a.save_models()
    b = _make(72, True)
    b.load_models()
    da, db = a._dev, b._dev
    assert np.array_equal(da.buffer("tree")[:, 0].cpu().numpy(), db.buffer("tree")[:, 0].cpu().numpy())
    for n in ("state_memory", "new_state_memory", "action_memory", "reward_memory", "terminal_memory", "hint_memory"):
        assert torch.equal(da.buffer(n).cpu(), db.buffer(n).cpu()), n
    ca, cb = da.counters(), db.counters()
    assert ca[0] == cb[0] and ca[2] == cb[2] and ca[3] == cb[3]
    for k, v in a.actor.state_dict().items():
        assert torch.equal(v.cpu(), b.actor.state_dict()[k].cpu())
    # the pickle is a plain container: no device object inside, reference attribute names
    with open("prioritized_replaymem_sac.model", "rb") as f:
        blob = pickle.load(f)
    assert blob.tree.tree.shape == (255,) and blob.tree.capacity == 128 and blob.state_memory.shape == (128, 24)
    b._dev.learn()          # and the loaded agent steps


def test_loads_a_reference_written_per_pickle(tmp_path):
    """tests/golden/ref_prioritized_replaymem_sac.model was written by the UNMODIFIED reference's PER.save_checkpoint
    (tests/golden/make_ref_per_pickle.py).  With the facade directory on sys.path as `enet_sac` (the drop-in usage) it loads."""
    src = os.path.join(ROOT, "tests", "golden", "ref_prioritized_replaymem_sac.model")
    facade = os.path.join(ROOT, "smart-calibration_b200", "elasticnet")
    sys.path.insert(0, facade)
    try:
        import enet_sac as top          # the facade imported the way callers of the reference import it
        torch.manual_seed(1)
        agent = top.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=8, n_actions=2, max_mem_size=16, tau=0.005,
                          prioritized=True)
        os.chdir(tmp_path)
        with open(src, "rb") as f, open("prioritized_replaymem_sac.model", "wb") as g:
            g.write(f.read())
        agent.replaymem.load_checkpoint()
        z = np.load(os.path.join(ROOT, "tests", "golden", "ref_prioritized_replaymem_sac.npz"))
        assert np.array_equal(agent.replaymem.tree.tree, z["tree"])
        assert np.array_equal(agent.replaymem.state_memory, z["state_memory"])
        assert np.array_equal(agent.replaymem.reward_memory, z["reward_memory"])
        assert agent.replaymem.mem_cntr == int(z["mem_cntr"]) and agent.replaymem.tree.data_pointer == int(z["data_pointer"])
    finally:
        sys.path.remove(facade)
        for m in ("enet_sac", "_host"):
            sys.modules.pop(m, None)


def test_agent_on_non_current_device_and_scalar_properties():
    """handles select their own device at every ABI entry; alpha / admm_rho assignments reach the device; rho is read
    from the device"""
    from smartcal_b200.elasticnet import enet_sac
    torch.manual_seed(3)
    agent = enet_sac.Agent(**HP, rng="device", use_hint=True)
    agent._dev.fill(pack(make_transitions(0, 128, 24, 2)))
    agent.alpha = 0.07
    agent.admm_rho = 0.5
    assert float(agent.alpha) == pytest.approx(0.07) and agent.admm_rho == 0.5
    for _ in range(11):
        agent.learn()
    st = agent.learn(want_stats=True)
    assert float(agent.rho) == pytest.approx(st["rho"]) and np.isfinite(st["rho"])
    if torch.cuda.device_count() > 1:
        with torch.cuda.device(1):
            agent.learn()               # current device 1, agent lives on device 0
        assert torch.cuda.current_device() == 0


def test_deferred_stats_are_the_previous_steps_losses():
    """want_stats='deferred': no synchronisation inside learn(); the call for step n+1 returns exactly what a synchronous read of step n gives"""
    def run(mode):
        a = _make(81, True)
        out = []
        for _ in range(8):
            out.append(a.learn(want_stats=mode))
        a._dev.sync()
        return out

    sync, deferred = run(True), run("deferred")
    assert deferred[0]["did_learn"] == 0
    for n in range(7):
        for k in ("critic_1_loss", "critic_2_loss", "actor_loss", "q1_mean", "learn_counter"):
            assert deferred[n + 1][k] == sync[n][k], (n, k)


@pytest.mark.parametrize("S", [24, 96])
def test_operand_copies_follow_parameter_writes(S):
    """The tcgen05 GEMMs read the wide weights as tf32 big / small copies kept next to the fp32 parameters.  Writing parameters
    from outside -- load_state_dict through fresh views, or an in-place write through a view taken earlier followed by
    params_changed() -- must give the same trajectory as an agent that was built with those weights (S = 24: layer 1 is
    updated by the FP32 kernel, S = 96: by the tensor-core kernel)."""
    def run(agent, n):
        for _ in range(n):
            agent.learn()
        agent._dev.sync()
    a, b, c = (_make(31, True, input_dims=[S]) for _ in range(3))
    cached = b._dev.tensor(L.NET_CRITIC_1, "fc12.weight")         # a view b's owner holds on to
    cached_a = b._dev.tensor(L.NET_ACTOR, "fc1.weight")
    run(a, 3), run(b, 3), run(c, 3)
    assert all(torch.equal(v, _state(b)[k]) for k, v in _state(a).items())
    g = torch.Generator().manual_seed(5)
    d12 = 0.01 * torch.randn(cached.shape, generator=g)
    d1 = 0.01 * torch.randn(cached_a.shape, generator=g)
    # a: load_state_dict (fresh views)
    sd = a.critic_1.state_dict(); sd["fc12.weight"] = sd["fc12.weight"] + d12.to(sd["fc12.weight"].device); a.critic_1.load_state_dict(sd)
    sd = a.actor.state_dict(); sd["fc1.weight"] = sd["fc1.weight"] + d1.to(sd["fc1.weight"].device); a.actor.load_state_dict(sd)
    # b: in-place through the cached views + explicit notification
    b._dev.flush()
    cached.add_(d12.to(cached.device)); cached_a.add_(d1.to(cached_a.device))
    b._dev.params_changed()
    run(a, 3), run(b, 3)
    sa, sb = _state(a), _state(b)
    assert all(torch.equal(v, sb[k]) for k, v in sa.items())
    # and the write did change the trajectory (c never saw it)
    run(c, 3)
    sc = _state(c)
    assert any(not torch.equal(v, sc[k]) for k, v in sa.items())
