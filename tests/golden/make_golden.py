"""Golden-vector generator: runs the UNMODIFIED reference (imported from /root/reference, CPU) on seeded
synthetic transitions and records (a) every random draw it consumed, (b) digests of every parameter, Adam moment
and target after each learn(), (c) PER indices / weights / trees.  While generating it also runs the oracle
restatement (oracle/) on the same draws and asserts agreement, which is what pins the oracle.

Run here (the reference is not available on the GPU box):   python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (committed).  Deterministic: reruns reproduce the files bit for bit on this image.
"""
import os
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/elasticnet"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
os.environ["CUDA_VISIBLE_DEVICES"] = ""
warnings.filterwarnings("ignore")

import smartcal_b200  # noqa: E402
from smartcal_b200.synth import make_transitions, make_calib_transitions  # noqa: E402
from oracle.per_ref import PerRef  # noqa: E402
from oracle.sac_ref import SacRef  # noqa: E402
from oracle.td3_ref import Td3Ref  # noqa: E402
from oracle.ddpg_ref import DdpgRef  # noqa: E402
from oracle.calib_ref import CalibSacRef  # noqa: E402

import enet_sac as ref_sac  # noqa: E402  (the reference)
import enet_td3 as ref_td3  # noqa: E402
import enet_ddpg as ref_ddpg  # noqa: E402

NSAMP = 32


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(NSAMP, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


def net_digests(out, prefix, module, with_adam):
    for name, p in module.named_parameters():
        out[f"{prefix}/{name}/p"] = digest(p)
        if with_adam:
            st = module.optimizer.state.get(p, None)
            if st:
                out[f"{prefix}/{name}/m"] = digest(st["exp_avg"])
                out[f"{prefix}/{name}/v"] = digest(st["exp_avg_sq"])


def oracle_digests(out, prefix, params, opt):
    for name, p in params.items():
        out[f"{prefix}/{name}/p"] = digest(p)
        if opt is not None:
            st = opt.state.get(p, None)
            if st:
                out[f"{prefix}/{name}/m"] = digest(st["exp_avg"])
                out[f"{prefix}/{name}/v"] = digest(st["exp_avg_sq"])


def fill_reference_replay(mem, tr):
    n = tr["state"].shape[0]
    mem.state_memory[:n] = tr["state"]
    mem.new_state_memory[:n] = tr["new_state"]
    mem.action_memory[:n] = tr["action"]
    mem.reward_memory[:n] = tr["reward"]
    mem.terminal_memory[:n] = tr["done"]
    mem.hint_memory[:n] = tr["hint"]
    mem.mem_cntr = n


def compare(a, b, what, tol=0.0):
    worst = 0.0
    for k in a:
        d = np.max(np.abs(a[k] - b[k]) / (1.0 + np.abs(a[k])))
        worst = max(worst, d)
        assert d <= tol, f"{what}: oracle differs from the reference at {k}: {d}"
    return worst


def sac_case(name, S, A, B, C, steps, seed, use_hint, done_frac):
    hp = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
              reward_scale=20, alpha=0.03)
    tr = make_transitions(seed + 100, C, S, A, "normal", done_frac)
    np.random.seed(seed)
    torch.manual_seed(seed)
    agent = ref_sac.Agent(**hp, prioritized=False, use_hint=use_hint)
    fill_reference_replay(agent.replaymem, tr)
    torch.manual_seed(seed)
    orc = SacRef(**hp, use_hint=use_hint)
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], tr["reward"][i], tr["new_state"][i], tr["done"][i], tr["hint"][i])
    out = dict(cfg=np.array([S, A, B, C, steps, seed, int(use_hint)], np.int64), done_frac=np.array([done_frac]))
    net_digests(out, "init/actor", agent.actor, False)
    net_digests(out, "init/critic_1", agent.critic_1, False)
    net_digests(out, "init/critic_2", agent.critic_2, False)
    o0 = {}
    oracle_digests(o0, "init/actor", orc.actor, None)
    oracle_digests(o0, "init/critic_1", orc.critic_1, None)
    oracle_digests(o0, "init/critic_2", orc.critic_2, None)
    compare({k: out[k] for k in o0}, o0, name + " init")
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    worst = 0.0
    for s in range(steps):
        nst, tst = np.random.get_state(), torch.get_rng_state()
        cnt = agent.learn_counter
        agent.learn()
        nend, tend = np.random.get_state(), torch.get_rng_state()
        # replay the generators to capture what learn() drew, and check it accounts for ALL the draws
        np.random.set_state(nst)
        torch.set_rng_state(tst)
        idx = np.random.choice(C, B, replace=False)
        e1, e2 = torch.randn(B, A), torch.randn(B, A)
        e3 = torch.randn(B, A) if (use_hint and cnt % 10 == 0) else None
        assert np.array_equal(np.random.get_state()[1], nend[1]) and torch.equal(torch.get_rng_state(), tend), "draw order mismatch"
        out[f"step{s}/idx"] = idx.astype(np.int32)
        out[f"step{s}/eps1"], out[f"step{s}/eps2"] = e1.numpy(), e2.numpy()
        if e3 is not None:
            out[f"step{s}/eps3"] = e3.numpy()
        ref_d = {}
        for nm, mod, adam in (("actor", agent.actor, True), ("critic_1", agent.critic_1, True), ("critic_2", agent.critic_2, True),
                              ("target_critic_1", agent.target_critic_1, False), ("target_critic_2", agent.target_critic_2, False)):
            net_digests(ref_d, f"step{s}/{nm}", mod, adam)
        if use_hint:
            ref_d[f"step{s}/rho"] = np.array([float(agent.rho)])
        res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2, e3)
        orc_d = {}
        for nm, prm, opt in (("actor", orc.actor, orc.opt_a), ("critic_1", orc.critic_1, orc.opt_1), ("critic_2", orc.critic_2, orc.opt_2),
                             ("target_critic_1", orc.target_critic_1, None), ("target_critic_2", orc.target_critic_2, None)):
            oracle_digests(orc_d, f"step{s}/{nm}", prm, opt)
        if use_hint:
            orc_d[f"step{s}/rho"] = np.array([float(orc.rho)])
        worst = max(worst, compare(ref_d, orc_d, f"{name} step {s}", tol=1e-12))
        out.update(ref_d)
        out[f"step{s}/losses"] = np.array([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {steps} steps, oracle-vs-reference worst rel diff {worst:.3g}")


def td3_case(name, S, A, B, C, steps, seed, prioritized, done_frac, use_hint=False, admm_rho=None):
    hp = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
              update_actor_interval=2, warmup=100, noise=0.1)
    tr = make_transitions(seed + 100, C, S, A, "uniform", done_frac)
    np.random.seed(seed)
    torch.manual_seed(seed)
    agent = ref_td3.Agent(**hp, prioritized=prioritized, use_hint=use_hint)
    torch.manual_seed(seed)
    orc = Td3Ref(**hp, prioritized=prioritized, use_hint=use_hint)
    if admm_rho is not None:          # main_td3.py:22 sets agent.admm_rho after construction
        agent.admm_rho = admm_rho
        orc.admm_rho = admm_rho
    for i in range(C):
        args = (tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]), tr["hint"][i])
        if prioritized:
            agent.replaymem.store_transition_from_buffer(*args, float(tr["reward"][i]))   # enet_td3.py:205 error=reward
        else:
            i0 = agent.replaymem.mem_cntr % C
            for arr, v in zip(("state_memory", "action_memory", "reward_memory", "new_state_memory", "terminal_memory", "hint_memory"), args):
                getattr(agent.replaymem, arr)[i0] = v
            agent.replaymem.mem_cntr += 1
        orc.store_flat(*args)
    out = dict(cfg=np.array([S, A, B, C, steps, seed, int(prioritized)], np.int64), done_frac=np.array([done_frac]),
               hint=np.array([int(use_hint), 0.0 if admm_rho is None else admm_rho]))
    if prioritized:
        out["tree_after_fill"] = agent.replaymem.tree.tree.copy()
        assert np.array_equal(orc.per.tree.tree, agent.replaymem.tree.tree), "PER fill differs"
    net_digests(out, "init/actor", agent.actor, False)
    net_digests(out, "init/critic_1", agent.critic_1, False)
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    worst = 0.0
    for s in range(steps):
        nst = np.random.get_state()
        agent.learn()
        nend = np.random.get_state()
        np.random.set_state(nst)
        if prioritized:
            u = np.random.random_sample(B)
            out[f"step{s}/u01"] = u
        else:
            idx = np.random.choice(C, B, replace=False)
            out[f"step{s}/idx"] = idx.astype(np.int32)
        nz = np.random.normal(scale=0.2)
        assert np.array_equal(np.random.get_state()[1], nend[1]), "draw order mismatch"
        out[f"step{s}/noise"] = np.array([nz])
        ref_d = {}
        for nm, mod, adam in (("actor", agent.actor, True), ("critic_1", agent.critic_1, True), ("critic_2", agent.critic_2, True),
                              ("target_actor", agent.target_actor, False), ("target_critic_1", agent.target_critic_1, False),
                              ("target_critic_2", agent.target_critic_2, False)):
            net_digests(ref_d, f"step{s}/{nm}", mod, adam)
        if prioritized:
            ref_d[f"step{s}/tree"] = agent.replaymem.tree.tree.copy()
            idxs, data, w = orc.per.sample(u)
            res = orc.learn_on_batch(orc.replay.gather(data), nz, is_weights=w, tree_idx=idxs)
            out[f"step{s}/tree_idx"], out[f"step{s}/is_weights"] = idxs, w
            out[f"step{s}/per_errors"] = res["per_errors"].reshape(-1)
        else:
            res = orc.learn_on_batch(orc.replay.gather(idx), nz)
        orc_d = {}
        for nm, prm, opt in (("actor", orc.actor, orc.opt_a), ("critic_1", orc.critic_1, orc.opt_1), ("critic_2", orc.critic_2, orc.opt_2),
                             ("target_actor", orc.target_actor, None), ("target_critic_1", orc.target_critic_1, None),
                             ("target_critic_2", orc.target_critic_2, None)):
            oracle_digests(orc_d, f"step{s}/{nm}", prm, opt)
        if prioritized:
            orc_d[f"step{s}/tree"] = orc.per.tree.tree.copy()
        worst = max(worst, compare(ref_d, orc_d, f"{name} step {s}", tol=1e-12))
        out.update(ref_d)
        out[f"step{s}/losses"] = np.array([res["critic_1_loss"], res["critic_2_loss"], res.get("actor_loss", np.nan)])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {steps} steps, oracle-vs-reference worst rel diff {worst:.3g}")


def ddpg_case(name, S, A, B, C, steps, seed, done_frac):
    hp = dict(gamma=0.99, lr_a=1e-4, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.001)   # main_ddpg.py:18-19
    tr = make_transitions(seed + 100, C, S, A, "normal", done_frac)
    np.random.seed(seed)
    torch.manual_seed(seed)
    agent = ref_ddpg.Agent(**hp)
    torch.manual_seed(seed)
    orc = DdpgRef(**hp)
    mem = agent.replaymem
    mem.state_memory[:C], mem.new_state_memory[:C], mem.action_memory[:C] = tr["state"], tr["new_state"], tr["action"]
    mem.reward_memory[:C], mem.terminal_memory[:C] = tr["reward"], tr["done"]
    mem.mem_cntr = C
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]))
    out = dict(cfg=np.array([S, A, B, C, steps, seed], np.int64), done_frac=np.array([done_frac]))
    net_digests(out, "init/actor", agent.actor, False)
    net_digests(out, "init/critic", agent.critic, False)
    o0 = {}
    oracle_digests(o0, "init/actor", orc.actor, None)
    oracle_digests(o0, "init/critic", orc.critic, None)
    compare({k: out[k] for k in o0}, o0, name + " init")
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    worst = 0.0
    for s in range(steps):
        nst = np.random.get_state()
        agent.learn()
        nend = np.random.get_state()
        np.random.set_state(nst)
        idx = np.random.choice(C, B, replace=False)
        assert np.array_equal(np.random.get_state()[1], nend[1]), "draw order mismatch"
        out[f"step{s}/idx"] = idx.astype(np.int32)
        ref_d = {}
        for nm, mod, adam in (("actor", agent.actor, True), ("critic", agent.critic, True), ("target_actor", agent.target_actor, False),
                              ("target_critic", agent.target_critic, False)):
            net_digests(ref_d, f"step{s}/{nm}", mod, adam)
        res = orc.learn_on_batch(orc.replay.gather(idx))
        orc_d = {}
        for nm, prm, opt in (("actor", orc.actor, orc.opt_a), ("critic", orc.critic, orc.opt_c), ("target_actor", orc.target_actor, None),
                             ("target_critic", orc.target_critic, None)):
            oracle_digests(orc_d, f"step{s}/{nm}", prm, opt)
        worst = max(worst, compare(ref_d, orc_d, f"{name} step {s}", tol=1e-12))
        out.update(ref_d)
        out[f"step{s}/losses"] = np.array([res["critic_loss"], res["actor_loss"]])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {steps} steps, oracle-vs-reference worst rel diff {worst:.3g}")


def calib_case(name, M, B, C, steps, seed, use_hint, done_frac):
    """calibration/calib_sac.py (conv SAC).  The module is imported under its own directory (it has its own lbfgsnew.py)."""
    import contextlib
    import importlib.util
    import io
    cdir = "/root/reference/calibration"
    sys.path.insert(0, cdir)
    spec = importlib.util.spec_from_file_location("ref_calib_sac", os.path.join(cdir, "calib_sac.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    sys.path.remove(cdir)
    A = 2 * M
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3, lr_c=1e-3,
              reward_scale=M, alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=use_hint)       # calibration/main_sac.py:16-19
    tr = make_calib_transitions(seed + 100, C, M, A, 128, done_frac)
    np.random.seed(seed)
    torch.manual_seed(seed)
    agent = ref.Agent(**hp)
    torch.manual_seed(seed)
    orc = CalibSacRef(**hp)
    for i in range(C):
        agent.replaymem.store_transition({"img": tr["img"][i], "sky": tr["sky"][i]}, tr["action"][i], float(tr["reward"][i]),
                                         {"img": tr["img2"][i], "sky": tr["sky2"][i]}, bool(tr["done"][i]), tr["hint"][i])
        orc.replay.store(tr["img"][i], tr["sky"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i], tr["sky2"][i],
                         bool(tr["done"][i]), tr["hint"][i])
    nets = (("actor", "actor"), ("critic_1", "critic_1"), ("critic_2", "critic_2"), ("target_critic_1", "target_critic_1"),
            ("target_critic_2", "target_critic_2"))

    def ref_digests(prefix):
        d = {}
        for nm, attr in nets:
            mod = getattr(agent, attr)
            for k, v in mod.state_dict().items():
                d[f"{prefix}/{nm}/{k}/p"] = digest(v.float())
            if hasattr(mod, "optimizer") and not nm.startswith("target"):
                for k, p in mod.named_parameters():
                    st = mod.optimizer.state.get(p, None)
                    if st:
                        d[f"{prefix}/{nm}/{k}/m"], d[f"{prefix}/{nm}/{k}/v"] = digest(st["exp_avg"]), digest(st["exp_avg_sq"])
        return d

    def orc_digests(prefix):
        d = {}
        for nm, attr in nets:
            p, b = getattr(orc, attr), getattr(orc, attr + "_b")
            for k, v in {**p, **b}.items():
                d[f"{prefix}/{nm}/{k}/p"] = digest(v.float())
            opt = {"actor": orc.opt_a, "critic_1": orc.opt_1, "critic_2": orc.opt_2}.get(nm)
            if opt is not None:
                for k, v in p.items():
                    st = opt.state.get(v, None)
                    if st:
                        d[f"{prefix}/{nm}/{k}/m"], d[f"{prefix}/{nm}/{k}/v"] = digest(st["exp_avg"]), digest(st["exp_avg_sq"])
        return d

    out = dict(cfg=np.array([M, A, B, C, steps, seed, int(use_hint)], np.int64), done_frac=np.array([done_frac]))
    out.update(ref_digests("init"))
    compare(ref_digests("init"), orc_digests("init"), name + " init")
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    worst = 0.0
    for s in range(steps):
        nst, tst = np.random.get_state(), torch.get_rng_state()
        with contextlib.redirect_stdout(io.StringIO()):
            agent.learn()
        nend, tend = np.random.get_state(), torch.get_rng_state()
        np.random.set_state(nst)
        torch.set_rng_state(tst)
        idx = np.random.choice(C, B, replace=False)
        e1, e2 = torch.randn(B, A), torch.randn(B, A)
        assert np.array_equal(np.random.get_state()[1], nend[1]) and torch.equal(torch.get_rng_state(), tend), "draw order mismatch"
        out[f"step{s}/idx"] = idx.astype(np.int32)
        out[f"step{s}/eps1"], out[f"step{s}/eps2"] = e1.numpy(), e2.numpy()
        res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2)
        rd, od = ref_digests(f"step{s}"), orc_digests(f"step{s}")
        if use_hint:
            rd[f"step{s}/rho"], od[f"step{s}/rho"] = np.array([float(agent.rho)]), np.array([float(orc.rho)])
        worst = max(worst, compare(rd, od, f"{name} step {s}", tol=1e-12))
        out.update(rd)
        out[f"step{s}/losses"] = np.array([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]])
        out[f"step{s}/y"], out[f"step{s}/q1"], out[f"step{s}/pi"] = digest(res["y"]), digest(res["q1"]), digest(res["pi"])
    # choose_action in eval mode (running statistics)
    obs = make_calib_transitions(seed + 200, 1, M, A, 128, 0.0)
    tst = torch.get_rng_state()
    a_ref = agent.choose_action({"img": obs["img"][0], "sky": obs["sky"][0]})
    torch.set_rng_state(tst)
    eps = torch.randn(1, A)
    a_orc = orc.choose_action(obs["img"][0], obs["sky"][0], eps)
    assert np.array_equal(a_ref, a_orc), "choose_action differs"
    out["act/eps"], out["act/action"] = eps.numpy(), a_ref
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {steps} steps, oracle-vs-reference worst rel diff {worst:.3g}")


def per_case(name, C, B, rounds, seed):
    """Drive the reference PER class directly: stores (with and without errors), stratified samples, updates."""
    rs = np.random.RandomState(seed)
    S, A = 4, 2
    per = ref_sac.PER(C, [S], A)
    orc = PerRef(C, 100.0)
    z = np.zeros(S, np.float32)
    out = dict(cfg=np.array([C, B, rounds, seed], np.int64))
    nstore = C + C // 4            # wraps the ring
    errs = rs.uniform(0, 3, nstore)
    use_none = rs.uniform(0, 1, nstore) < 0.25
    use_none[0] = False
    out["store_err"], out["store_none"] = errs, use_none
    for i in range(nstore):
        e = None if use_none[i] else float(errs[i])
        per.store_transition_from_buffer(z, np.zeros(A, np.float32), 0.0, z, False, np.zeros(A, np.float32), e)
        orc.store(e)
    assert np.array_equal(per.tree.tree, orc.tree.tree)
    out["tree_after_store"] = per.tree.tree.copy()
    np.random.seed(seed)
    for r in range(rounds):
        st = np.random.get_state()
        res = per.sample_buffer(B)
        idxs, w = res[6], res[7]
        np.random.set_state(st)
        u = np.random.random_sample(B)
        oi, od, ow = orc.sample(u)
        assert np.array_equal(oi, idxs) and np.array_equal(ow, w), "oracle PER sample differs from the reference"
        abs_err = rs.uniform(0, 150, (B, 1)).astype(np.float32)      # some above the clip of 100
        per.batch_update(idxs, abs_err)
        ps = orc.update(idxs, abs_err)
        assert np.array_equal(per.tree.tree, orc.tree.tree), "oracle PER update differs from the reference"
        out[f"r{r}/u01"], out[f"r{r}/idx"], out[f"r{r}/w"] = u, idxs, w
        out[f"r{r}/abs_err"], out[f"r{r}/ps"], out[f"r{r}/tree"] = abs_err.reshape(-1), ps, per.tree.tree.copy()
        out[f"r{r}/beta"] = np.array([float(per.beta)])
    # an UNSORTED, duplicate-laden update (arbitrary caller-supplied index list)
    idxs = rs.randint(C - 1, 2 * C - 1, B).astype(np.int32)
    abs_err = rs.uniform(0, 5, (B, 1)).astype(np.float32)
    per.batch_update(idxs, abs_err)
    orc.update(idxs, abs_err)
    assert np.array_equal(per.tree.tree, orc.tree.tree)
    out["unsorted/idx"], out["unsorted/abs_err"], out["unsorted/tree"] = idxs, abs_err.reshape(-1), per.tree.tree.copy()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: oracle PER bit-identical to the reference over {rounds} rounds")


if __name__ == "__main__":
    torch.set_num_threads(8)
    per_case("per_c64_b16", 64, 16, 6, 11)
    per_case("per_c4096_b256", 4096, 256, 3, 12)
    sac_case("sac_s24_b32", 24, 2, 32, 64, 3, 1, False, 0.1)
    sac_case("sac_s22_hint_b32", 22, 2, 32, 64, 12, 2, True, 0.1)
    sac_case("sac_enet_b256", 420, 2, 256, 1024, 2, 3, False, 0.05)
    td3_case("td3_per_s24_b32", 24, 2, 32, 64, 4, 4, True, 0.1)
    td3_case("td3_uniform_s24_b32", 24, 2, 32, 64, 2, 5, False, 0.1)
    td3_case("td3_per_enet_b64", 420, 2, 64, 128, 2, 6, True, 0.05)
    td3_case("td3_per_hint_s24_b32", 24, 2, 32, 64, 4, 7, True, 0.1, use_hint=True, admm_rho=1.0)
    ddpg_case("ddpg_s24_b32", 24, 2, 32, 64, 3, 8, 0.1)
    calib_case("calib_m3_b8", 3, 8, 16, 3, 9, False, 0.2)
    calib_case("calib_m3_hint_b8", 3, 8, 16, 11, 10, True, 0.2)
