"""CPU: the oracle restatement (oracle/) against the golden vectors recorded from the UNMODIFIED reference
(tests/golden/*.npz, made by tests/golden/make_golden.py).  On the generating image the match is bit-exact; other
hosts may block their sgemm differently, hence the small tolerances below (relative to 1+|x|)."""
import os
import numpy as np
import pytest
import torch

from oracle.per_ref import PerRef
from oracle.sac_ref import SacRef
from oracle.td3_ref import Td3Ref
from smartcal_b200.synth import make_transitions

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NSAMP = 32
TOL = 5e-5       # digests after an Adam step: a handful of near-zero-gradient elements may move by ~lr


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(NSAMP, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


def close(got, ref, tol, what):
    # sums (entries 0,1) scale with the tensor; compare relative to the abs-sum, samples relative to 1+|x|
    scale = np.concatenate([[1 + abs(ref[1])] * 2, 1 + np.abs(ref[2:])])
    d = np.max(np.abs(got - ref) / scale)
    assert d <= tol, f"{what}: {d}"


def check_nets(z, prefix, nets, tol):
    for nm, prm, opt in nets:
        for k, p in prm.items():
            close(digest(p), z[f"{prefix}/{nm}/{k}/p"], tol, f"{prefix}/{nm}/{k}/p")
            st = opt.state.get(p) if opt is not None else None
            if st:
                close(digest(st["exp_avg"]), z[f"{prefix}/{nm}/{k}/m"], tol, f"{prefix}/{nm}/{k}/m")
                close(digest(st["exp_avg_sq"]), z[f"{prefix}/{nm}/{k}/v"], tol, f"{prefix}/{nm}/{k}/v")


@pytest.mark.parametrize("name", ["per_c64_b16", "per_c4096_b256"])
def test_per_oracle_bit_exact(name):
    z = np.load(os.path.join(G, name + ".npz"))
    C, B, rounds, seed = [int(x) for x in z["cfg"]]
    orc = PerRef(C, 100.0)
    for e, none in zip(z["store_err"], z["store_none"]):
        orc.store(None if none else float(e))
    assert np.array_equal(orc.tree.tree, z["tree_after_store"])
    for r in range(rounds):
        idx, data, w = orc.sample(z[f"r{r}/u01"])
        assert np.array_equal(idx, z[f"r{r}/idx"])
        assert np.array_equal(data, z[f"r{r}/idx"] - (C - 1))
        assert np.array_equal(w, z[f"r{r}/w"])
        assert float(orc.beta) == float(z[f"r{r}/beta"][0])
        ps = orc.update(idx, z[f"r{r}/abs_err"])
        # float32 power is libm/SVML dependent across hosts: allow 2 ulp, then pin the tree through the recorded priorities
        d = np.abs(ps.view(np.int32).astype(np.int64) - z[f"r{r}/ps"].view(np.int32).astype(np.int64))
        assert d.max() <= 2
        if d.max() == 0:
            assert np.array_equal(orc.tree.tree, z[f"r{r}/tree"])
        else:
            orc.tree.tree[:] = z[f"r{r}/tree"]
    # idxs non-decreasing for stratified draws (what the CUDA per-level fold relies on; it has an exact fallback)
    for r in range(rounds):
        assert np.all(np.diff(z[f"r{r}/idx"]) >= 0)
    orc.tree.tree[:] = z[f"r{rounds - 1}/tree"]
    ps = orc.update(z["unsorted/idx"], z["unsorted/abs_err"])
    np.testing.assert_allclose(orc.tree.tree, z["unsorted/tree"], rtol=1e-6)


@pytest.mark.parametrize("name", ["sac_s24_b32", "sac_s22_hint_b32", "sac_enet_b256"])
def test_sac_oracle_matches_reference(name):
    z = np.load(os.path.join(G, name + ".npz"))
    S, A, B, C, steps, seed, use_hint = [int(x) for x in z["cfg"]]
    hp = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
              reward_scale=20, alpha=0.03)
    tr = make_transitions(seed + 100, C, S, A, "normal", float(z["done_frac"][0]))
    torch.manual_seed(seed)
    orc = SacRef(**hp, use_hint=bool(use_hint))
    check_nets(z, "init", [("actor", orc.actor, None), ("critic_1", orc.critic_1, None), ("critic_2", orc.critic_2, None)], 0.0)
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], tr["reward"][i], tr["new_state"][i], tr["done"][i], tr["hint"][i])
    for s in range(steps):
        e3 = z[f"step{s}/eps3"] if f"step{s}/eps3" in z else None
        res = orc.learn_on_batch(orc.replay.gather(z[f"step{s}/idx"]), z[f"step{s}/eps1"], z[f"step{s}/eps2"], e3)
        tol = TOL * (s + 1)
        check_nets(z, f"step{s}", [("actor", orc.actor, orc.opt_a), ("critic_1", orc.critic_1, orc.opt_1), ("critic_2", orc.critic_2, orc.opt_2),
                                   ("target_critic_1", orc.target_critic_1, None), ("target_critic_2", orc.target_critic_2, None)], tol)
        np.testing.assert_allclose([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]], z[f"step{s}/losses"], rtol=1e-4)
        if use_hint:
            np.testing.assert_allclose(float(orc.rho), z[f"step{s}/rho"][0], rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("name", ["td3_per_s24_b32", "td3_uniform_s24_b32", "td3_per_enet_b64", "td3_per_hint_s24_b32"])
def test_td3_oracle_matches_reference(name):
    z = np.load(os.path.join(G, name + ".npz"))
    S, A, B, C, steps, seed, prioritized = [int(x) for x in z["cfg"]]
    hp = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
              update_actor_interval=2, warmup=100, noise=0.1)
    tr = make_transitions(seed + 100, C, S, A, "uniform", float(z["done_frac"][0]))
    torch.manual_seed(seed)
    use_hint, rho = (int(z["hint"][0]), float(z["hint"][1])) if "hint" in z else (0, 0.0)
    orc = Td3Ref(**hp, prioritized=bool(prioritized), use_hint=bool(use_hint))
    if rho:
        orc.admm_rho = rho
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]), tr["hint"][i])
    if prioritized:
        assert np.array_equal(orc.per.tree.tree, z["tree_after_fill"])
    for s in range(steps):
        nz = float(z[f"step{s}/noise"][0])
        if prioritized:
            idxs, data, w = orc.per.sample(z[f"step{s}/u01"])
            assert np.array_equal(idxs, z[f"step{s}/tree_idx"])
            res = orc.learn_on_batch(orc.replay.gather(data), nz, is_weights=w, tree_idx=idxs)
            np.testing.assert_allclose(orc.per.tree.tree, z[f"step{s}/tree"], rtol=1e-5)
            orc.per.tree.tree[:] = z[f"step{s}/tree"]
        else:
            res = orc.learn_on_batch(orc.replay.gather(z[f"step{s}/idx"]), nz)
        check_nets(z, f"step{s}", [("actor", orc.actor, orc.opt_a), ("critic_1", orc.critic_1, orc.opt_1), ("critic_2", orc.critic_2, orc.opt_2),
                                   ("target_actor", orc.target_actor, None), ("target_critic_1", orc.target_critic_1, None),
                                   ("target_critic_2", orc.target_critic_2, None)], TOL * (s + 1))


def test_ddpg_oracle_matches_reference():
    from oracle.ddpg_ref import DdpgRef
    z = np.load(os.path.join(G, "ddpg_s24_b32.npz"))
    S, A, B, C, steps, seed = [int(x) for x in z["cfg"]]
    hp = dict(gamma=0.99, lr_a=1e-4, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.001)
    tr = make_transitions(seed + 100, C, S, A, "normal", float(z["done_frac"][0]))
    torch.manual_seed(seed)
    orc = DdpgRef(**hp)
    check_nets(z, "init", [("actor", orc.actor, None), ("critic", orc.critic, None)], 0.0)
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]))
    for s in range(steps):
        res = orc.learn_on_batch(orc.replay.gather(z[f"step{s}/idx"]))
        check_nets(z, f"step{s}", [("actor", orc.actor, orc.opt_a), ("critic", orc.critic, orc.opt_c),
                                   ("target_actor", orc.target_actor, None), ("target_critic", orc.target_critic, None)], TOL * (s + 1))
        np.testing.assert_allclose([res["critic_loss"], res["actor_loss"]], z[f"step{s}/losses"], rtol=1e-4)


@pytest.mark.parametrize("name", ["calib_m3_b8", "calib_m3_hint_b8"])
def test_calib_oracle_matches_reference(name):
    """conv SAC of calibration/calib_sac.py: the oracle restatement replays the reference's recorded draws"""
    from oracle.calib_ref import CalibSacRef
    from smartcal_b200.synth import make_calib_transitions
    z = np.load(os.path.join(G, name + ".npz"))
    M, A, B, C, steps, seed, use_hint = [int(x) for x in z["cfg"]]
    steps = min(steps, 2)                       # each CPU step costs ~1 s
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3, lr_c=1e-3,
              reward_scale=M, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=bool(use_hint))
    tr = make_calib_transitions(seed + 100, C, M, A, 128, float(z["done_frac"][0]))
    torch.manual_seed(seed)
    orc = CalibSacRef(**hp)
    for i in range(C):
        orc.replay.store(tr["img"][i], tr["sky"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i], tr["sky2"][i],
                         bool(tr["done"][i]), tr["hint"][i])

    def check(prefix, tol):
        for nm in ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2"):
            sd = {**getattr(orc, nm), **getattr(orc, nm + "_b")}
            for k, v in sd.items():
                ref = z[f"{prefix}/{nm}/{k}/p"]
                a = v.detach().float().numpy().astype(np.float64).reshape(-1)
                idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
                got = np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])
                assert np.max(np.abs(got - ref) / (1.0 + np.abs(ref))) <= tol, f"{prefix}/{nm}/{k}"

    check("init", 0.0)
    for s in range(steps):
        res = orc.learn_on_batch(orc.replay.gather(z[f"step{s}/idx"]), z[f"step{s}/eps1"], z[f"step{s}/eps2"])
        check(f"step{s}", TOL * (s + 1))
        np.testing.assert_allclose([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]], z[f"step{s}/losses"], rtol=1e-4)


@pytest.mark.parametrize("name", ["demix_fuzzy_mlp_b16", "demix_fuzzy_mlp_hint_b16", "demix_rl_b8", "demix_rl_hint_b8"])
def test_demix_oracle_matches_reference(name):
    """the two demixing SAC agents (demixing_rl/demix_sac.py, demixing_fuzzy/demix_sac.py): the oracle restatement replays the
    draws recorded from the unmodified reference modules (tests/golden/make_golden_demix.py)"""
    from oracle.demix_ref import DemixSacRef
    from smartcal_b200.synth import make_demix_transitions
    z = np.load(os.path.join(G, name + ".npz"))
    M, A, B, C, steps, seed, use_hint, ui, inner = [int(x) for x in z["cfg"]]
    steps = min(steps, 2 if ui else 11)         # each conv CPU step costs ~1 s
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=3e-4, lr_c=1e-3,
              alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=bool(use_hint))
    tr = make_demix_transitions(seed + 100, C, M, A, 128 if ui else 0, float(z["done_frac"][0]))
    torch.manual_seed(seed)
    orc = DemixSacRef(**hp, M=M, use_influence=bool(ui), inner=inner)
    for i in range(C):
        orc.replay.store(tr["img"][i] if ui else None, tr["meta"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i] if ui else None,
                         tr["meta2"][i], bool(tr["done"][i]), tr["hint"][i])

    def check(prefix, tol):
        for nm in ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2"):
            sd = {**getattr(orc, nm), **getattr(orc, nm + "_b")}
            for k, v in sd.items():
                ref = z[f"{prefix}/{nm}/{k}/p"]
                a = v.detach().float().numpy().astype(np.float64).reshape(-1)
                idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
                got = np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])
                assert np.max(np.abs(got - ref) / (1.0 + np.abs(ref))) <= tol, f"{prefix}/{nm}/{k}"

    check("init", 0.0)
    for s in range(steps):
        e3 = z[f"step{s}/eps3"] if f"step{s}/eps3" in z else None
        res = orc.learn_on_batch(orc.replay.gather(z[f"step{s}/idx"]), z[f"step{s}/eps1"], z[f"step{s}/eps2"], e3)
        check(f"step{s}", TOL * (s + 1))
        np.testing.assert_allclose([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]], z[f"step{s}/losses"], rtol=1e-4)
        if use_hint:
            assert abs(float(orc.rho) - float(z[f"step{s}/rho"][0])) <= 1e-6 * (1 + abs(float(orc.rho)))


@pytest.mark.parametrize("name", ["demix_td3_b8", "demix_td3_hint_b8"])
def test_demix_td3_oracle_matches_reference(name):
    """the demixing conv TD3 agent (demixing_rl/demix_td3.py): the oracle restatement replays the draws recorded from the unmodified
    reference module (tests/golden/make_golden_demix_td3.py) -- PER tree / indices / IS weights, the eval-mode target passes, the
    training-mode priority pre-pass, the delayed actor update (with hints: the five-step adaptive-ADMM loop) and the state_dict soft
    update incl. BatchNorm buffers"""
    from oracle.demix_td3_ref import DemixTd3Ref
    from smartcal_b200.synth import make_demix_transitions
    z = np.load(os.path.join(G, name + ".npz"))
    M, A, B, C, steps, seed, use_hint = [int(x) for x in z["cfg"]]
    steps = min(steps, 2)                       # one critic-only step and one actor update (each conv CPU step costs ~1 s)
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=1e-3, lr_c=1e-3,
              update_actor_interval=2, warmup=4, noise=0.1, use_hint=bool(use_hint))
    tr = make_demix_transitions(seed + 100, C, M, A, 128, float(z["done_frac"][0]))
    torch.manual_seed(seed)
    orc = DemixTd3Ref(**hp, M=M)
    for i in range(C):
        orc.store(tr["img"][i], tr["meta"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i], tr["meta2"][i], bool(tr["done"][i]), tr["hint"][i])

    def dg(v):
        a = (v.detach().double() if v.dtype == torch.int64 else v.detach().float()).numpy().astype(np.float64).reshape(-1)
        idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
        return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])

    def check(prefix, tol):
        for nm in DemixTd3Ref.NETS:
            for k, v in {**getattr(orc, nm), **getattr(orc, nm + "_b")}.items():
                ref = z[f"{prefix}/{nm}/{k}/p"]
                assert np.max(np.abs(dg(v) - ref) / (1.0 + np.abs(ref))) <= tol, f"{prefix}/{nm}/{k}"
        ref = z[f"{prefix}/tree"]
        assert np.max(np.abs(dg(torch.as_tensor(orc.per.tree.tree)) - ref) / (1.0 + np.abs(ref))) <= tol, f"{prefix}/tree"

    check("init", 0.0)
    for s in range(steps):
        idxs, data, w = orc.per.sample(z[f"step{s}/u01"])
        assert np.array_equal(idxs, z[f"step{s}/tree_idx"])
        assert np.abs(w.view(np.int32).astype(np.int64) - z[f"step{s}/is_w"].view(np.int32).astype(np.int64)).max() <= 2
        res = orc.learn_on_batch(orc.replay.gather(data), float(z[f"step{s}/noise"][0]), w, idxs)
        check(f"step{s}", TOL * (s + 1))
        np.testing.assert_allclose([res["critic_1_loss"], res["critic_2_loss"]], z[f"step{s}/losses"][:2], rtol=1e-4)
        assert res["actor_step"] == (s % 2 == 1)
        if res["actor_step"]:
            np.testing.assert_allclose(res["actor_loss"], z[f"step{s}/losses"][2], rtol=1e-4)


def test_demix_td3_parity_cases_are_well_conditioned():
    """tools/screen_demix_td3_seeds.py: the committed conv-TD3 parity cases (golden seeds 41 / 43, oracle-test seed 51) are steps where the
    reference's float32 gradients agree with their float64 evaluation to rounding; seed 42 -- the case that was dropped -- has a ReLU knife
    edge in its first step (percent-level disagreement), which is what the screen is for."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("screen_demix_td3_seeds", os.path.join(root, "tools", "screen_demix_td3_seeds.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.screen(41, False, "golden", steps=2) < 1e-4
    assert m.screen(43, True, "golden", steps=2) < 1e-4
    assert m.screen(51, True, "oracle-test", steps=2) < 1e-4
    assert m.screen(42, False, "golden", steps=1) > 1e-3
