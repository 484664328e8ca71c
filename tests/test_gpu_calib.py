"""GPU: the calibration conv-SAC path (sc_calib_*) against the CPU oracle (oracle/calib_ref.py, bit-identical to the
reference) and against the golden vectors recorded from the unmodified reference."""
import os

import numpy as np
import pytest
import torch

import smartcal_b200
from smartcal_b200 import _lib as L
from smartcal_b200.synth import make_calib_transitions

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NETS = (("actor", L.NET_ACTOR), ("critic_1", L.NET_CRITIC_1), ("critic_2", L.NET_CRITIC_2),
        ("target_critic_1", L.NET_TARGET_CRITIC_1), ("target_critic_2", L.NET_TARGET_CRITIC_2))


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


def build(M, A, B, C, seed, use_hint, done_frac):
    from smartcal_b200.calibration import calib_sac
    from oracle.calib_ref import CalibSacRef
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3, lr_c=1e-3,
              reward_scale=M, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=use_hint)
    tr = make_calib_transitions(seed + 100, C, M, A, 128, done_frac)
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = calib_sac.Agent(**hp, rng="reference")
    torch.manual_seed(seed)
    orc = CalibSacRef(**hp)
    for i in range(C):
        agent.store_transition({"img": tr["img"][i], "sky": tr["sky"][i]}, tr["action"][i], float(tr["reward"][i]),
                               {"img": tr["img2"][i], "sky": tr["sky2"][i]}, bool(tr["done"][i]), tr["hint"][i])
        orc.replay.store(tr["img"][i], tr["sky"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i], tr["sky2"][i],
                         bool(tr["done"][i]), tr["hint"][i])
    return agent, orc, tr


def oracle_state(orc, nm):
    return {**getattr(orc, nm), **getattr(orc, nm + "_b")}


def push_oracle_state(agent, orc):
    """teacher forcing: device parameters, BatchNorm buffers and Adam moments := oracle (first-step Adam amplifies
    noise-level gradients such as the conv biases in front of BatchNorm to +-lr, so trajectories are re-pinned per step)"""
    for nm, net in NETS:
        getattr(agent, nm).load_state_dict({k: v.detach() for k, v in oracle_state(orc, nm).items()})
    for nm, opt in (("actor", orc.opt_a), ("critic_1", orc.opt_1), ("critic_2", orc.opt_2)):
        view = getattr(agent, nm)
        for k, p in getattr(orc, nm).items():
            st = opt.state.get(p)
            if st:
                view._t(k, L.SLOT_ADAM_M).copy_(st["exp_avg"].reshape(view._t(k).shape))
                view._t(k, L.SLOT_ADAM_V).copy_(st["exp_avg_sq"].reshape(view._t(k).shape))
    agent._dev.set_scalars(float(orc.rho), agent.admm_rho)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


@pytest.mark.parametrize("use_hint", [False, True])
def test_calib_learn_matches_oracle(use_hint):
    M, A, B, C, seed = 3, 6, 8, 16, 21
    agent, orc, tr = build(M, A, B, C, seed, use_hint, 0.2)
    dev = agent._dev
    for nm, net in NETS:                      # construction: identical weights and buffers
        for k, v in oracle_state(orc, nm).items():
            assert torch.equal(getattr(agent, nm).state_dict()[k].cpu(), v.detach()), f"init {nm}.{k}"
    rs = np.random.RandomState(seed + 1)
    for step in range(3):
        idx = rs.choice(C, B, replace=False)
        e1, e2 = rs.standard_normal((B, A)).astype(np.float32), rs.standard_normal((B, A)).astype(np.float32)
        dev.set_batch(idx)
        st = dev.learn(e1, e2, want_stats=True)
        res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2)
        assert rel(dev.buffer("y").cpu().numpy(), res["y"].numpy()) < 2e-5, "TD target"
        assert rel(dev.buffer("q1").cpu().numpy(), res["q1"].numpy()) < 2e-5 and rel(dev.buffer("q2").cpu().numpy(), res["q2"].numpy()) < 2e-5
        assert rel(dev.buffer("pi").cpu().numpy(), res["pi"].numpy()) < 2e-5, "policy sample"
        assert rel(dev.buffer("logp").cpu().numpy(), res["logp"].numpy()) < 2e-5, "log-prob"
        np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"]],
                                   [res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]], rtol=1e-4, atol=1e-6)
        if use_hint:
            assert abs(st["gfun"] - res["gfun"]) <= 1e-4 * abs(res["gfun"]) + 1e-9
            assert abs(st["rho"] - float(orc.rho)) <= 1e-4 * abs(float(orc.rho)) + 1e-9
        # gradients of every parameter (slot 3), relative to the largest entry of the tensor; conv biases in front of a
        # BatchNorm have a mathematically zero gradient (pure rounding noise on both sides), so they get an absolute bound.
        # Actor: d logp / d sigma = -1/sigma and sigma = clamp(fc22sigma(x), 1e-6, 1) is a cancelling sum that can land at
        # 1e-5..1e-4, where a 1e-8 summation-order difference is a 1e-4..1e-3 relative difference of the term that dominates
        # every upstream actor gradient.  The bound therefore scales with the measured relative sigma difference.
        sg_dev, sg_orc = dev.buffer("sigma").cpu().numpy().astype(np.float64), res["sigma"].numpy().astype(np.float64)
        assert np.max(np.abs(sg_dev - sg_orc)) < 5e-7, "sigma differs by more than summation-order noise"
        sig_amp = float(np.max(np.abs(sg_dev - sg_orc) / sg_orc))
        for nm, gkey in (("critic_1", "grad_c1"), ("critic_2", "grad_c2"), ("actor", "grad_a")):
            rtol = 2e-5 if nm != "actor" else max(2e-5, 3.0 * sig_amp)
            got = getattr(agent, nm).grads()
            for k, g in res[gkey].items():
                gd = got[k].cpu().numpy().reshape(-1)
                gr = g.numpy().reshape(-1)
                scale = np.max(np.abs(gr))
                if k.startswith("conv") and k.endswith(".bias"):
                    assert np.max(np.abs(gd)) < 1e-4 and scale < 1e-4, f"{nm}.{k} should vanish"
                    continue
                assert np.max(np.abs(gd - gr)) <= rtol * scale + 1e-9, f"step {step} grad {nm}.{k}: {np.max(np.abs(gd - gr))} vs scale {scale} (rtol {rtol})"
        # BatchNorm running statistics and counters follow the reference's eight passes
        for nm, net in NETS:
            sd = getattr(agent, nm).state_dict()
            for k, v in getattr(orc, nm + "_b").items():
                if k.endswith("num_batches_tracked"):
                    assert int(sd[k]) == int(v), f"{nm}.{k}"
                elif k.endswith("running_mean"):
                    # the conv bias in front of a BatchNorm has a zero gradient up to rounding, so its first Adam steps are
                    # +-lr with a sign that is noise on both sides; the batch mean of the next pass moves with it (x momentum)
                    assert np.max(np.abs(sd[k].cpu().numpy() - v.numpy())) <= 0.1 * 2.2e-3 + 1e-5, f"{nm}.{k}"
                else:
                    assert rel(sd[k].cpu().numpy(), v.numpy()) < 1e-4, f"{nm}.{k}"
        push_oracle_state(agent, orc)
    # eval-mode action
    obs = make_calib_transitions(seed + 7, 1, M, A, 128, 0.0)
    eps = rs.standard_normal((1, A)).astype(np.float32)
    a_dev = dev.act(obs["img"][0], obs["sky"][0], eps)
    a_orc = orc.choose_action(obs["img"][0], obs["sky"][0], eps)
    assert np.max(np.abs(a_dev - a_orc)) < 2e-5


def test_calib_against_reference_golden():
    """no oracle in the loop: digests recorded from the unmodified calib_sac.Agent"""
    z = np.load(os.path.join(G, "calib_m3_b8.npz"))
    M, A, B, C, steps, seed, use_hint = [int(x) for x in z["cfg"]]
    from smartcal_b200.calibration import calib_sac
    tr = make_calib_transitions(seed + 100, C, M, A, 128, float(z["done_frac"][0]))
    torch.manual_seed(seed)
    np.random.seed(seed)
    agent = calib_sac.Agent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3,
                            lr_c=1e-3, reward_scale=M, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=bool(use_hint), rng="reference")
    for nm, net in NETS:
        for k, v in getattr(agent, nm).state_dict().items():
            assert np.array_equal(digest(v.float()), z[f"init/{nm}/{k}/p"]), f"init {nm}.{k} differs from the reference"
    for i in range(C):
        agent.store_transition({"img": tr["img"][i], "sky": tr["sky"][i]}, tr["action"][i], float(tr["reward"][i]),
                               {"img": tr["img2"][i], "sky": tr["sky2"][i]}, bool(tr["done"][i]), tr["hint"][i])
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    st = agent.learn(want_stats=True)
    assert np.array_equal(agent._dev.buffer("batch_idx")[:, 0].cpu().numpy(), z["step0/idx"]), "np.random.choice draw"
    assert np.array_equal(agent._dev.buffer("eps_next").cpu().numpy(), z["step0/eps1"])
    assert np.array_equal(agent._dev.buffer("eps_cur").cpu().numpy(), z["step0/eps2"])
    np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"]], z["step0/losses"], rtol=2e-4)
    for name, buf in (("y", "y"), ("q1", "q1"), ("pi", "pi")):
        ref, got = z[f"step0/{name}"], digest(agent._dev.buffer(buf))
        assert np.max(np.abs(got[2:] - ref[2:])) <= 3e-5 * max(np.max(np.abs(ref[2:])), 1e-3), name
    # first Adam step: m = 0.1 g pins the gradients of the reference
    for nm in ("actor", "critic_1", "critic_2"):
        adam = getattr(agent, nm).adam_state()
        for k in adam:
            if k.startswith("conv") and k.endswith(".bias"):
                continue
            ref, got = z[f"step0/{nm}/{k}/m"], digest(adam[k][0])
            tol = 5e-4 * max(np.max(np.abs(ref[2:])), ref[1] / adam[k][0].numel()) + 1e-12
            assert np.max(np.abs(got[2:] - ref[2:])) <= tol, f"{nm}.{k} samples of exp_avg"
    for nm in ("actor", "critic_1", "target_critic_1"):
        sd = getattr(agent, nm).state_dict()
        for l in (1, 2, 3):
            assert int(sd[f"bn{l}.num_batches_tracked"]) == int(z[f"step0/{nm}/bn{l}.num_batches_tracked/p"][0])
            ref, got = z[f"step0/{nm}/bn{l}.running_var/p"], digest(sd[f"bn{l}.running_var"])
            assert np.max(np.abs(got - ref)) <= 1e-4 * np.max(np.abs(ref)), f"{nm}.bn{l}.running_var"


def test_calib_config3_dimensions_one_step():
    """BASELINE.json configs[2] dimensions (M=10: sky 11x7, 20 actions, fc1 in 97 wide) at a batch the CPU oracle finishes quickly"""
    M, A, B, C, seed = 10, 20, 32, 64, 33
    agent, orc, tr = build(M, A, B, C, seed, True, 0.1)
    dev = agent._dev
    rs = np.random.RandomState(seed + 1)
    idx = rs.choice(C, B, replace=False)
    e1, e2 = rs.standard_normal((B, A)).astype(np.float32), rs.standard_normal((B, A)).astype(np.float32)
    dev.set_batch(idx)
    st = dev.learn(e1, e2, want_stats=True)
    res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2)
    for name, key in (("y", "y"), ("q1", "q1"), ("q2", "q2"), ("pi", "pi"), ("logp", "logp"), ("q1_pi", "q1_pi"), ("q2_pi", "q2_pi")):
        assert rel(dev.buffer(name).cpu().numpy(), res[key].numpy()) < 3e-5, name
    np.testing.assert_allclose([st["critic_1_loss"], st["critic_2_loss"], st["actor_loss"], st["gfun"]],
                               [res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"], res["gfun"]], rtol=2e-4, atol=1e-7)
    sg_dev, sg_orc = dev.buffer("sigma").cpu().numpy().astype(np.float64), res["sigma"].numpy().astype(np.float64)
    sig_amp = float(np.max(np.abs(sg_dev - sg_orc) / sg_orc))
    for nm, gkey in (("critic_1", "grad_c1"), ("critic_2", "grad_c2"), ("actor", "grad_a")):
        got = getattr(agent, nm).grads()
        rtol = 3e-5 if nm != "actor" else max(3e-5, 3.0 * sig_amp)
        for k, g in res[gkey].items():
            if k.startswith("conv") and k.endswith(".bias"):
                continue
            gd, gr = got[k].cpu().numpy().reshape(-1), g.numpy().reshape(-1)
            assert np.max(np.abs(gd - gr)) <= rtol * np.max(np.abs(gr)) + 1e-9, f"grad {nm}.{k}"


def test_calib_facade_api_roundtrip(tmp_path, monkeypatch):
    """the calls calibration/main_sac.py makes: choose_action / store_transition / learn / save_models / load_models"""
    from smartcal_b200.calibration import calib_sac
    monkeypatch.chdir(tmp_path)
    M, A, B, C = 3, 6, 4, 8
    torch.manual_seed(5)
    np.random.seed(5)
    agent = calib_sac.Agent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3,
                            lr_c=1e-3, reward_scale=M, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=True)
    tr = make_calib_transitions(9, C + 2, M, A, 128, 0.0)
    for i in range(C + 2):                                # wraps the ring
        obs, obs_ = {"img": tr["img"][i], "sky": tr["sky"][i]}, {"img": tr["img2"][i], "sky": tr["sky2"][i]}
        action = agent.choose_action(obs)
        assert action.shape == (A,) and np.all(np.abs(action) <= 1.0)
        agent.store_transition(obs, action, float(tr["reward"][i]), obs_, False, tr["hint"][i])
        agent.learn()
    assert agent.replaymem.mem_cntr == C + 2 and agent.learn_counter == C + 2 - (B - 1)
    np.testing.assert_array_equal(agent.replaymem.state_memory_img[1], tr["img"][C + 1])          # slot 1 was overwritten on wrap
    assert float(agent.rho) >= 0.0
    before = {k: v.clone() for k, v in agent.actor.state_dict().items()}
    agent.save_models()
    assert os.path.exists("a_eval_sac_actor.model") and os.path.exists("q_eval_2_sac_critic.model") and os.path.exists("replaymem_sac.model")
    sd = torch.load("a_eval_sac_actor.model")
    assert list(sd.keys()) == list(before.keys()) and tuple(sd["conv2.weight"].shape) == (32, 16, 5, 5)
    agent.actor.load_state_dict({k: torch.zeros_like(v) for k, v in before.items()})
    agent.load_models()
    for k, v in agent.actor.state_dict().items():
        assert torch.equal(v, before[k]), k
    for k, v in agent.target_critic_1.state_dict().items():
        if "running" not in k and "num_batches" not in k:
            assert torch.equal(v, agent.critic_1.state_dict()[k]), k        # load_models hard-copies the targets (calib_sac.py:398-399)
