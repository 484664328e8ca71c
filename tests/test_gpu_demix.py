"""GPU: the two demixing SAC agents (demixing_rl/demix_sac.py conv SAC, demixing_fuzzy/demix_sac.py MLP SAC) on the sc_calib_*
kernels with the network-family flags, against the CPU oracle (oracle/demix_ref.py, bit-identical to the unmodified reference
modules, tests/golden/make_golden_demix.py) and against golden vectors recorded from the reference itself."""
import os

import numpy as np
import pytest
import torch

import smartcal_b200
from smartcal_b200 import _lib as L
from smartcal_b200.synth import make_demix_transitions

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2")


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def make_agent(family, hp, M, use_influence):
    if family == "demixing_rl":
        from smartcal_b200.demixing_rl import demix_sac
        return demix_sac.DemixingAgent(**hp, M=M, rng="reference")
    from smartcal_b200.demixing_fuzzy import demix_sac
    return demix_sac.DemixingAgent(**hp, n_meta=M, use_influence=use_influence, rng="reference")


def build(family, M, A, B, C, seed, use_hint, done_frac):
    from oracle.demix_ref import DemixSacRef
    ui = family == "demixing_rl"
    inner = 16 if ui else 32
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=3e-4, lr_c=1e-3,
              alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=use_hint)
    tr = make_demix_transitions(seed + 100, C, M, A, 128 if ui else 0, done_frac)
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = make_agent(family, hp, M, ui)
    torch.manual_seed(seed)
    orc = DemixSacRef(**hp, M=M, use_influence=ui, inner=inner)
    for i in range(C):
        s = {"infmap": tr["img"][i] if ui else 0, "metadata": tr["meta"][i]}
        s2 = {"infmap": tr["img2"][i] if ui else 0, "metadata": tr["meta2"][i]}
        agent.store_transition(s, tr["action"][i], float(tr["reward"][i]), s2, bool(tr["done"][i]), tr["hint"][i])
        orc.replay.store(tr["img"][i] if ui else None, tr["meta"][i], tr["action"][i], float(tr["reward"][i]),
                         tr["img2"][i] if ui else None, tr["meta2"][i], bool(tr["done"][i]), tr["hint"][i])
    return agent, orc, tr, ui


def oracle_state(orc, nm):
    return {**getattr(orc, nm), **getattr(orc, nm + "_b")}


def push_oracle_state(agent, orc):
    """teacher forcing (as in tests/test_gpu_calib.py): device parameters, BatchNorm buffers, Adam moments, rho := oracle"""
    for nm in NETS:
        getattr(agent, nm).load_state_dict({k: v.detach() for k, v in oracle_state(orc, nm).items()})
    for nm, opt in (("actor", orc.opt_a), ("critic_1", orc.opt_1), ("critic_2", orc.opt_2)):
        view = getattr(agent, nm)
        for k, p in getattr(orc, nm).items():
            st = opt.state.get(p)
            if st:
                view._t(k, L.SLOT_ADAM_M).copy_(st["exp_avg"].reshape(view._t(k).shape))
                view._t(k, L.SLOT_ADAM_V).copy_(st["exp_avg_sq"].reshape(view._t(k).shape))
    agent._dev.set_scalars(float(orc.rho), agent.admm_rho)


@pytest.mark.parametrize("family,use_hint", [("demixing_fuzzy", False), ("demixing_fuzzy", True), ("demixing_rl", False), ("demixing_rl", True)])
def test_demix_learn_matches_oracle(family, use_hint):
    # demixing_fuzzy/main_sac.py dims: n_meta = 32, n_actions = 128; demixing_rl/main_sac.py: M = 20, n_actions = 6
    M, A = (32, 128) if family == "demixing_fuzzy" else (20, 6)
    B, C, seed = (16, 32, 31) if family == "demixing_fuzzy" else (8, 16, 32)
    agent, orc, tr, ui = build(family, M, A, B, C, seed, use_hint, 0.2)
    dev = agent._dev
    for nm in NETS:                           # construction: identical weights and buffers (reference draw order)
        for k, v in oracle_state(orc, nm).items():
            assert torch.equal(getattr(agent, nm).state_dict()[k].cpu(), v.detach()), f"init {nm}.{k}"
    rs = np.random.RandomState(seed + 1)
    for step in range(3):
        idx = rs.choice(C, B, replace=False)
        e1, e2, e3 = (rs.standard_normal((B, A)).astype(np.float32) for _ in range(3))
        dual = use_hint and step % 10 == 0
        dev.set_batch(idx)
        st = dev.learn(e1, e2, eps_dual=e3 if dual else None, want_stats=True)
        res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2, e3 if dual else None)
        for name, key in (("y", "y"), ("q1", "q1"), ("q2", "q2"), ("pi", "pi"), ("logp", "logp"), ("q1_pi", "q1_pi")):
            assert rel(dev.buffer(name).cpu().numpy(), res[key].numpy()) < 3e-5, (step, name)
        np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"]],
                                   [res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]], rtol=1e-4, atol=1e-6)
        if use_hint:
            assert abs(st["gfun"] - res["gfun"]) <= 1e-4 * abs(res["gfun"]) + 1e-9
            assert abs(st["rho"] - float(orc.rho)) <= 2e-4 * abs(float(orc.rho)) + 1e-9, "dual update from the resampled actions"
        sg_dev, sg_orc = dev.buffer("sigma").cpu().numpy().astype(np.float64), res["sigma"].numpy().astype(np.float64)
        assert rel(sg_dev, sg_orc) < 2e-5, "sigma = exp(clamp(logsigma))"
        for nm, gkey in (("critic_1", "grad_c1"), ("critic_2", "grad_c2"), ("actor", "grad_a")):
            got = getattr(agent, nm).grads()
            for k, g in res[gkey].items():
                gd, gr = got[k].cpu().numpy().reshape(-1), g.numpy().reshape(-1)
                scale = np.max(np.abs(gr))
                if k.startswith("conv") and k.endswith(".bias"):        # zero up to rounding in front of a BatchNorm
                    assert np.max(np.abs(gd)) < 1e-4 and scale < 1e-4, f"{nm}.{k} should vanish"
                    continue
                assert np.max(np.abs(gd - gr)) <= 5e-5 * scale + 1e-9, f"step {step} grad {nm}.{k}: {np.max(np.abs(gd - gr))} vs scale {scale}"
        if ui:
            for nm in NETS:
                sd = getattr(agent, nm).state_dict()
                for k, v in getattr(orc, nm + "_b").items():
                    if k.endswith("num_batches_tracked"):
                        assert int(sd[k]) == int(v), f"{nm}.{k} (training-mode BatchNorm passes incl. the dual resample)"
        push_oracle_state(agent, orc)
    obs = make_demix_transitions(seed + 7, 1, M, A, 128 if ui else 0, 0.0)
    eps = rs.standard_normal((1, A)).astype(np.float32)
    a_dev = dev.act(obs["img"][0] if ui else None, obs["meta"][0], eps)
    a_orc = orc.choose_action(obs["img"][0] if ui else None, obs["meta"][0], eps)
    assert np.max(np.abs(a_dev - a_orc)) < 2e-5


@pytest.mark.parametrize("name,family", [("demix_fuzzy_mlp_b16", "demixing_fuzzy"), ("demix_rl_b8", "demixing_rl")])
def test_demix_against_reference_golden(name, family):
    """no oracle in the loop: digests recorded from the unmodified reference DemixingAgent (first learn() of each file)"""
    z = np.load(os.path.join(G, name + ".npz"))
    M, A, B, C, steps, seed, use_hint, ui, inner = [int(x) for x in z["cfg"]]
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=3e-4, lr_c=1e-3,
              alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=bool(use_hint))
    tr = make_demix_transitions(seed + 100, C, M, A, 128 if ui else 0, float(z["done_frac"][0]))
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = make_agent(family, hp, M, bool(ui))
    for nm in NETS:
        for k, v in getattr(agent, nm).state_dict().items():
            assert np.array_equal(digest(v.float()), z[f"init/{nm}/{k}/p"]), f"init {nm}.{k} differs from the reference"
    for i in range(C):
        s = {"infmap": tr["img"][i] if ui else 0, "metadata": tr["meta"][i]}
        s2 = {"infmap": tr["img2"][i] if ui else 0, "metadata": tr["meta2"][i]}
        agent.store_transition(s, tr["action"][i], float(tr["reward"][i]), s2, bool(tr["done"][i]), tr["hint"][i])
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    st = agent.learn(want_stats=True)
    assert np.array_equal(agent._dev.buffer("batch_idx")[:, 0].cpu().numpy(), z["step0/idx"]), "np.random.choice draw"
    assert np.array_equal(agent._dev.buffer("eps_next").cpu().numpy(), z["step0/eps1"])
    assert np.array_equal(agent._dev.buffer("eps_cur").cpu().numpy(), z["step0/eps2"])
    np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"]], z["step0/losses"], rtol=2e-4)
    for nm_, buf in (("y", "y"), ("q1", "q1"), ("pi", "pi")):
        ref, got = z[f"step0/{nm_}"], digest(agent._dev.buffer(buf))
        assert np.max(np.abs(got[2:] - ref[2:])) <= 3e-5 * max(np.max(np.abs(ref[2:])), 1e-3), nm_
    for nm in ("actor", "critic_1", "critic_2"):          # first Adam step: m = 0.1 g pins the reference's gradients
        adam = getattr(agent, nm).adam_state()
        for k in adam:
            if k.startswith("conv") and k.endswith(".bias"):
                continue
            ref, got = z[f"step0/{nm}/{k}/m"], digest(adam[k][0])
            tol = 5e-4 * max(np.max(np.abs(ref[2:])), ref[1] / adam[k][0].numel()) + 1e-12
            assert np.max(np.abs(got[2:] - ref[2:])) <= tol, f"{nm}.{k} samples of exp_avg"
    # choose_action (eval mode) against the reference's recorded action
    obs = make_demix_transitions(seed + 200, 1, M, A, 128 if ui else 0, 0.0)
    a = agent._dev.act(obs["img"][0] if ui else None, obs["meta"][0], z["act/eps"])
    assert a.shape == (A,)


def test_demix_fuzzy_facade_roundtrip_at_main_sac_shape(tmp_path, monkeypatch):
    """the calls demixing_fuzzy/main_sac.py makes, at its shapes (n_meta = 32, n_fuzzy = 128, batch 256): choose_action /
    store_transition / learn / save_models / load_models / load_replaybuffer"""
    from smartcal_b200.demixing_fuzzy import demix_sac
    monkeypatch.chdir(tmp_path)
    M, A, B, C = 32, 128, 256, 512
    torch.manual_seed(5)
    np.random.seed(5)
    agent = demix_sac.DemixingAgent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], n_meta=M,
                                    lr_a=3e-4, lr_c=3e-4, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=False, use_influence=False)
    tr = make_demix_transitions(9, C, M, A, 0, 0.0)
    losses = []
    for i in range(B + 30):
        obs, obs_ = {"metadata": tr["meta"][i]}, {"metadata": tr["meta2"][i]}
        action = agent.choose_action(obs)
        assert action.shape == (A,) and np.all(np.abs(action) <= 1.0)
        agent.store_transition(obs, action, float(tr["reward"][i]), obs_, False, tr["hint"][i])
        st = agent.learn(want_stats=True)
        if st:
            losses.append(st["critic_1_loss"])
    assert agent.replaymem.mem_cntr == B + 30 and agent.learn_counter == 31 and np.isfinite(losses).all()
    np.testing.assert_array_equal(agent.replaymem.state_memory_sky[3], tr["meta"][3])
    before = {k: v.clone() for k, v in agent.actor.state_dict().items()}
    assert list(before.keys()) == ["fc11.weight", "fc11.bias", "fc12.weight", "fc12.bias", "fc21.weight", "fc21.bias", "fc22mu.weight",
                                   "fc22mu.bias", "fc22logsigma.weight", "fc22logsigma.bias"]
    assert tuple(before["fc21.weight"].shape) == (128, 32) and tuple(agent.critic_1.state_dict()["fc1.weight"].shape) == (128, 160)
    agent.save_models()
    assert os.path.exists("a_eval_sac_actor.model") and os.path.exists("q_eval_1_sac_critic.model") and os.path.exists("replaymem_sac.model")
    agent.actor.load_state_dict({k: torch.zeros_like(v) for k, v in before.items()})
    agent.load_models()
    for k, v in agent.actor.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_demixing_rl_per_clips_errors_at_one():
    """demixing_rl/demix_sac.py:151-161: the PER beside the conv agent has absolute_error_upper = 1 (elastic net: 100)"""
    from smartcal_b200.demixing_rl.demix_sac import PER
    from oracle.per_ref import PerRef
    per, orc = PER(64, [4], 2), PerRef(64, 1.0)
    rs = np.random.RandomState(5)
    for i in range(64):
        e = float(rs.uniform(0, 3))             # most errors exceed the clip
        per.store_transition_from_buffer(np.full(4, i, np.float32), np.zeros(2, np.float32), float(i), np.zeros(4, np.float32), False, np.zeros(2, np.float32), e)
        orc.store(e)
    np.random.seed(9)
    out = per.sample_buffer(16)
    assert np.array_equal(per.tree.tree, orc.tree.tree) and per.tree.tree[63:].max() <= 1.0
    np.random.seed(9)
    idx, data, w = orc.sample(np.random.random_sample(16))
    assert np.array_equal(out[6], idx)
    err = rs.uniform(0, 3, 16).astype(np.float32)
    ps = per.batch_update(idx, err)
    ps_ref = orc.update(idx, err)
    assert np.all(ps <= 1.0)
    # float32 pow is host-libm dependent: priorities within 2 ulp, the tree follows them (bit-exact whenever they agree)
    if ps_ref is not None:
        assert np.abs(np.asarray(ps, np.float32).view(np.int32).astype(np.int64) - np.asarray(ps_ref, np.float32).view(np.int32).astype(np.int64)).max() <= 2
    np.testing.assert_allclose(per.tree.tree, orc.tree.tree, rtol=1e-6)


@pytest.mark.parametrize("family", ["demixing_fuzzy", "demixing_rl"])
def test_calib_family_graph_replay_equals_eager(family, monkeypatch):
    """rng='device': learn() of the sc_calib_* family replays from a CUDA graph from the third step on (step-dependent scalars --
    Adam bias corrections, Philox stream ids -- live in device memory); SMARTCAL_GRAPH=0 keeps eager launches.  Same bits."""
    M, A = (32, 128) if family == "demixing_fuzzy" else (20, 6)
    ui = family == "demixing_rl"
    B, C = 16, 32

    def run(graph):
        monkeypatch.setenv("SMARTCAL_GRAPH", "1" if graph else "0")
        hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=3e-4, lr_c=1e-3,
                  alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=True)
        torch.manual_seed(3)
        np.random.seed(3)
        if ui:
            from smartcal_b200.demixing_rl import demix_sac
            agent = demix_sac.DemixingAgent(**hp, M=M, rng="device")
        else:
            from smartcal_b200.demixing_fuzzy import demix_sac
            agent = demix_sac.DemixingAgent(**hp, n_meta=M, use_influence=False, rng="device")
        tr = make_demix_transitions(4, C, M, A, 128 if ui else 0, 0.1)
        for i in range(C):
            s = {"infmap": tr["img"][i] if ui else 0, "metadata": tr["meta"][i]}
            s2 = {"infmap": tr["img2"][i] if ui else 0, "metadata": tr["meta2"][i]}
            agent.store_transition(s, tr["action"][i], float(tr["reward"][i]), s2, bool(tr["done"][i]), tr["hint"][i])
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(13):                   # covers plain steps (graph) and two dual-update steps (their own graph)
                st = agent.learn(want_stats=True)
        out = {f"{nm}.{k}": v.detach().cpu().clone() for nm in NETS for k, v in getattr(agent, nm).state_dict().items()}
        out["rho"] = torch.tensor(st["rho"])
        agent._dev.close()
        return out

    g, e = run(True), run(False)
    for k in e:
        assert torch.equal(g[k], e[k]), k
