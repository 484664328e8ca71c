"""Golden vectors for the two demixing SAC agents: runs the UNMODIFIED reference modules (/root/reference/demixing_rl/demix_sac.py
and /root/reference/demixing_fuzzy/demix_sac.py, CPU) on seeded synthetic transitions, records every random draw they consumed
and digests of every parameter / Adam moment / BatchNorm buffer / target after each learn(), and asserts that the oracle
restatement (oracle/demix_ref.py) agrees on the same draws -- which is what pins the oracle.

Run in the build container (the reference does not travel to the GPU box):   python tests/golden/make_golden_demix.py
Outputs: tests/golden/demix_*.npz (committed)."""
import contextlib
import importlib.util
import io
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ["CUDA_VISIBLE_DEVICES"] = ""
warnings.filterwarnings("ignore")

import smartcal_b200  # noqa: E402,F401
from smartcal_b200.synth import make_demix_transitions  # noqa: E402
from oracle.demix_ref import DemixSacRef  # noqa: E402

NSAMP = 32


def digest(t):
    a = t.detach().cpu().numpy().astype(np.float64).reshape(-1)
    idx = np.linspace(0, a.size - 1, min(NSAMP, a.size)).astype(np.int64)
    return np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])


def load_ref(family):
    d = f"/root/reference/{family}"
    spec = importlib.util.spec_from_file_location(f"ref_{family}_sac", os.path.join(d, "demix_sac.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def case(name, family, M, A, B, C, steps, seed, use_hint, use_influence, done_frac):
    ref = load_ref(family)
    inner = ref.INNER_DIM if hasattr(ref, "INNER_DIM") else 16
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=3e-4, lr_c=1e-3,
              alpha=0.03, hint_threshold=0.01, admm_rho=1.0, use_hint=use_hint)
    tr = make_demix_transitions(seed + 100, C, M, A, 128 if use_influence else 0, done_frac)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if family == "demixing_rl":
        agent = ref.DemixingAgent(**hp, M=M)
    else:
        agent = ref.DemixingAgent(**hp, n_meta=M, use_influence=use_influence)
    torch.manual_seed(seed)
    orc = DemixSacRef(**hp, M=M, use_influence=use_influence, inner=inner)
    for i in range(C):
        s = {"infmap": tr["img"][i] if use_influence else 0, "metadata": tr["meta"][i]}
        s2 = {"infmap": tr["img2"][i] if use_influence else 0, "metadata": tr["meta2"][i]}
        agent.replaymem.store_transition(s, tr["action"][i], float(tr["reward"][i]), s2, bool(tr["done"][i]), tr["hint"][i])
        orc.replay.store(tr["img"][i] if use_influence else None, tr["meta"][i], tr["action"][i], float(tr["reward"][i]),
                         tr["img2"][i] if use_influence else None, tr["meta2"][i], bool(tr["done"][i]), tr["hint"][i])
    nets = ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2")

    def ref_digests(prefix):
        d = {}
        for nm in nets:
            mod = getattr(agent, nm)
            for k, v in mod.state_dict().items():
                d[f"{prefix}/{nm}/{k}/p"] = digest(v.float())
            if not nm.startswith("target"):
                for k, p in mod.named_parameters():
                    st = mod.optimizer.state.get(p, None)
                    if st:
                        d[f"{prefix}/{nm}/{k}/m"], d[f"{prefix}/{nm}/{k}/v"] = digest(st["exp_avg"]), digest(st["exp_avg_sq"])
        return d

    def orc_digests(prefix):
        d = {}
        for nm in nets:
            p, b = getattr(orc, nm), getattr(orc, nm + "_b")
            for k, v in {**p, **b}.items():
                d[f"{prefix}/{nm}/{k}/p"] = digest(v.float())
            opt = {"actor": orc.opt_a, "critic_1": orc.opt_1, "critic_2": orc.opt_2}.get(nm)
            if opt is not None:
                for k, v in p.items():
                    st = opt.state.get(v, None)
                    if st:
                        d[f"{prefix}/{nm}/{k}/m"], d[f"{prefix}/{nm}/{k}/v"] = digest(st["exp_avg"]), digest(st["exp_avg_sq"])
        return d

    def compare(a, b, what, tol):
        worst = 0.0
        assert set(a) == set(b), (what, sorted(set(a) ^ set(b))[:6])
        for k in a:
            d = np.max(np.abs(a[k] - b[k]) / (1.0 + np.abs(a[k])))
            worst = max(worst, d)
            assert d <= tol, f"{what}: {k} differs by {d}"
        return worst

    out = dict(cfg=np.array([M, A, B, C, steps, seed, int(use_hint), int(use_influence), inner], np.int64), done_frac=np.array([done_frac]))
    out.update(ref_digests("init"))
    compare(ref_digests("init"), orc_digests("init"), name + " init", 0.0)
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
        e3 = torch.randn(B, A) if (use_hint and s % 10 == 0) else None
        assert np.array_equal(np.random.get_state()[1], nend[1]) and torch.equal(torch.get_rng_state(), tend), "draw order mismatch"
        out[f"step{s}/idx"] = idx.astype(np.int32)
        out[f"step{s}/eps1"], out[f"step{s}/eps2"] = e1.numpy(), e2.numpy()
        if e3 is not None:
            out[f"step{s}/eps3"] = e3.numpy()
        res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2, e3)
        rd, od = ref_digests(f"step{s}"), orc_digests(f"step{s}")
        if use_hint:
            rd[f"step{s}/rho"], od[f"step{s}/rho"] = np.array([float(agent.rho)]), np.array([float(orc.rho)])
        worst = max(worst, compare(rd, od, f"{name} step {s}", 1e-12))
        out.update(rd)
        out[f"step{s}/losses"] = np.array([res["critic_1_loss"], res["critic_2_loss"], res["actor_loss"]])
        out[f"step{s}/y"], out[f"step{s}/q1"], out[f"step{s}/pi"] = digest(res["y"]), digest(res["q1"]), digest(res["pi"])
    obs = make_demix_transitions(seed + 200, 1, M, A, 128 if use_influence else 0, 0.0)
    tst = torch.get_rng_state()
    a_ref = agent.choose_action({"infmap": obs["img"][0] if use_influence else 0, "metadata": obs["meta"][0]})
    torch.set_rng_state(tst)
    eps = torch.randn(1, A)
    a_orc = orc.choose_action(obs["img"][0] if use_influence else None, obs["meta"][0], eps)
    assert np.array_equal(a_ref, a_orc), "choose_action differs"
    out["act/eps"], out["act/action"] = eps.numpy(), a_ref
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(f"{name}: {steps} steps, oracle-vs-reference worst rel diff {worst:.3g}")


if __name__ == "__main__":
    # demixing_fuzzy MLP agent at the shape of demixing_fuzzy/main_sac.py (n_meta = 32, n_actions = 128), small batch; with hints
    case("demix_fuzzy_mlp_b16", "demixing_fuzzy", 32, 128, 16, 32, 3, 21, False, False, 0.2)
    case("demix_fuzzy_mlp_hint_b16", "demixing_fuzzy", 32, 128, 16, 32, 11, 22, True, False, 0.2)
    # demixing_rl conv agent (M = 3K + 2 = 20, n_actions = K = 6), small batch; with hints
    case("demix_rl_b8", "demixing_rl", 20, 6, 8, 16, 3, 23, False, True, 0.2)
    case("demix_rl_hint_b8", "demixing_rl", 20, 6, 8, 16, 11, 24, True, True, 0.2)
