"""Shared parity machinery for the -m gpu tests, __graft_entry__.smoke() and the debug report: drives the CUDA
agent (through the Python facade -> C ABI) and the CPU oracle with IDENTICAL minibatch indices / PER uniforms /
N(0,1) draws and compares activations, gradients, Adam moments, parameters, targets, PER indices and trees.

Tolerances (fp32 path; stated once here and cited by the tests):
  ACT_RTOL  = 2e-5  forward activations / Q / y / logp, relative to the tensor's max magnitude
  GRAD_RTOL = 2e-4  gradients (and Adam moments), relative to the tensor's max magnitude
  parameters: |dp| <= PARAM_ATOL = 4e-6 where the gradient is well above noise; elements whose |g| is below
              1e-4 of the tensor max may legitimately move by up to lr per step in either direction (Adam divides by
              ~|g|) and are only checked against that bound
  PER: tree indices bit-exact; IS weights and float32 priorities within 2 ulp (float32 pow is host-libm dependent)
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import smartcal_b200  # noqa: E402
from smartcal_b200 import _lib as L  # noqa: E402
from smartcal_b200.synth import make_transitions, pack  # noqa: E402
from oracle import nets_ref as N  # noqa: E402
from oracle.sac_ref import SacRef  # noqa: E402
from oracle.td3_ref import Td3Ref  # noqa: E402

ACT_RTOL, GRAD_RTOL, PARAM_ATOL = 2e-5, 2e-4, 4e-6


class Report:
    def __init__(self, verbose=False):
        self.rows, self.failures, self.verbose = [], [], verbose
        self.max_grad_rel = 0.0
        self.allowed = {}

    def add(self, name, got, ref, rtol, kind="act"):
        got = np.asarray(got, np.float64).reshape(-1)
        ref = np.asarray(ref, np.float64).reshape(-1)
        if got.shape != ref.shape:
            self.failures.append((name, "shape", got.shape, ref.shape))
            return
        scale = max(np.max(np.abs(ref)), 1e-30)
        err = float(np.max(np.abs(got - ref))) if got.size else 0.0
        rel = err / scale
        bad = not np.isfinite(got).all() or rel > rtol
        self.rows.append((name, err, scale, rel, bad))
        if kind == "grad":
            self.max_grad_rel = max(self.max_grad_rel, rel)
        if bad:
            self.failures.append((name, rel, rtol))
        if self.verbose:
            print(f"  {'FAIL' if bad else 'ok  '} {name:34s} max|d|={err:.3e}  max|ref|={scale:.3e}  rel={rel:.3e}")

    def add_exact(self, name, got, ref):
        ok = np.array_equal(np.asarray(got), np.asarray(ref))
        self.rows.append((name, 0.0 if ok else 1.0, 1.0, 0.0 if ok else 1.0, not ok))
        if not ok:
            self.failures.append((name, "not bit-exact"))
        if self.verbose:
            print(f"  {'ok  ' if ok else 'FAIL'} {name:34s} bit-exact={ok}")

    def add_ulp(self, name, got, ref, max_ulp):
        g = np.asarray(got, np.float32).reshape(-1)
        r = np.asarray(ref, np.float32).reshape(-1)
        d = np.abs(g.view(np.int32).astype(np.int64) - r.view(np.int32).astype(np.int64))
        u = int(d.max()) if d.size else 0
        bad = u > max_ulp
        self.rows.append((name, float(u), 1.0, float(u), bad))
        if bad:
            self.failures.append((name, f"{u} ulp > {max_ulp}"))
        if self.verbose:
            print(f"  {'FAIL' if bad else 'ok  '} {name:34s} max ulp diff={u}")

    def add_params(self, name, got, ref, grad_ref, vhat_sqrt, lr):
        """Adam-aware parameter check for ONE step from identical state: the update is lr_t*mhat/(sqrt(vhat)+eps), which
        lies in [-lr, lr], so an error dg in this step's gradient moves it by at most lr*min(2, |dg|/sqrt(vhat)) (the cap 2 is
        a sign flip of a near-zero gradient on the first step, where the update is lr*sign(g)).  With |dg| <= GRAD_RTOL*max|g|
        (checked separately) the admissible deviation of element i is
            PARAM_ATOL*(1+max|p|) + lr*min(2, 2*GRAD_RTOL*max|g| / sqrt(vhat_i))."""
        got = np.asarray(got, np.float64).reshape(-1)
        ref = np.asarray(ref, np.float64).reshape(-1)
        g = np.abs(np.asarray(grad_ref, np.float64).reshape(-1))
        sv = np.maximum(np.asarray(vhat_sqrt, np.float64).reshape(-1), 1e-30)
        gmax = max(g.max(), 1e-30)
        allowed = PARAM_ATOL * (1 + np.abs(ref).max()) + lr * np.minimum(2.0, 2 * GRAD_RTOL * gmax / sv)
        self.allowed[name] = allowed
        d = np.abs(got - ref)
        ratio = float((d / allowed).max())
        strong = sv > 1e-2 * gmax
        e_strong = float(d[strong].max()) if strong.any() else 0.0
        bad = (not np.isfinite(got).all()) or ratio > 1.0
        self.rows.append((name, e_strong, 1.0, ratio, bad))
        if bad:
            self.failures.append((name, "param", ratio))
        if self.verbose:
            print(f"  {'FAIL' if bad else 'ok  '} {name:34s} max|dp|={d.max():.3e} (sqrt(vhat)>1%max|g|: {e_strong:.3e})  worst dev/allowed={ratio:.3f}")

    def add_target(self, name, got, ref, online_name, tau):
        """target = tau*p + (1-tau)*t inherits tau x the admissible deviation of p"""
        got = np.asarray(got, np.float64).reshape(-1)
        ref = np.asarray(ref, np.float64).reshape(-1)
        allowed = tau * self.allowed[online_name] + 1e-7 * (1 + np.abs(ref).max())
        d = np.abs(got - ref)
        ratio = float((d / allowed).max())
        bad = (not np.isfinite(got).all()) or ratio > 1.0
        self.rows.append((name, float(d.max()), 1.0, ratio, bad))
        if bad:
            self.failures.append((name, "target", ratio))
        if self.verbose:
            print(f"  {'FAIL' if bad else 'ok  '} {name:34s} max|d|={d.max():.3e}  worst dev/allowed={ratio:.3f}")

    def add_abs(self, name, got, ref, rtol, atol):
        got = np.asarray(got, np.float64).reshape(-1)
        ref = np.asarray(ref, np.float64).reshape(-1)
        err = float(np.max(np.abs(got - ref)))
        lim = rtol * max(np.max(np.abs(ref)), 1e-30) + atol
        bad = not np.isfinite(got).all() or err > lim
        self.rows.append((name, err, lim, err / lim, bad))
        if bad:
            self.failures.append((name, err, lim))
        if self.verbose:
            print(f"  {'FAIL' if bad else 'ok  '} {name:34s} max|d|={err:.3e}  limit={lim:.3e}")


def traced_block(p, fc, bn, x, tr, tag):
    z = F.linear(x, p[fc + ".weight"], p[fc + ".bias"])
    zh = F.layer_norm(z, (z.shape[-1],), None, None, 1e-5)
    u = zh * p[bn + ".weight"] + p[bn + ".bias"]
    h = F.elu(u)
    for t in (z, u):
        if t.requires_grad:
            t.retain_grad()
    tr["z" + tag], tr["zh" + tag], tr["u" + tag], tr["h" + tag] = z, zh, u, h
    tr["rs" + tag] = 1.0 / torch.sqrt(z.var(dim=-1, unbiased=False) + 1e-5)
    return h


def traced_critic(p, s, a):
    tr = {}
    x = traced_block(p, "fc12", "bn12", traced_block(p, "fc11", "bn11", s, tr, "11"), tr, "12")
    y = traced_block(p, "fc22", "bn22", traced_block(p, "fc21", "bn21", a, tr, "21"), tr, "22")
    tr["q"] = F.linear(torch.cat((x, y), dim=1), p["fc3.weight"], p["fc3.bias"])
    return tr


def upload_params(net_view, params):
    net_view.load_state_dict(OrderedDict((k, v.detach()) for k, v in params.items()))


def build_pair(kind, S, A, B, C, seed, prioritized=False, use_hint=False, reward="normal", done_frac=0.1, hp=None):
    """oracle + CUDA agent with identical weights and identical replay contents"""
    from smartcal_b200.elasticnet import enet_sac, enet_td3
    base = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005)
    if hp:
        base.update(hp)
    torch.manual_seed(seed)
    np.random.seed(seed)
    if kind == "sac":
        orc = SacRef(**base, reward_scale=20, alpha=0.03, use_hint=use_hint, prioritized=prioritized)
        torch.manual_seed(seed)
        agent = enet_sac.Agent(**base, reward_scale=20, alpha=0.03, use_hint=use_hint, prioritized=prioritized, rng="reference")
        nets = [("actor", orc.actor, agent.actor), ("critic_1", orc.critic_1, agent.critic_1), ("critic_2", orc.critic_2, agent.critic_2)]
    else:
        orc = Td3Ref(**base, update_actor_interval=2, warmup=0, noise=0.1, prioritized=prioritized, use_hint=use_hint)
        torch.manual_seed(seed)
        agent = enet_td3.Agent(**base, update_actor_interval=2, warmup=0, noise=0.1, prioritized=prioritized, use_hint=use_hint, rng="reference")
        nets = [("actor", orc.actor, agent.actor), ("critic_1", orc.critic_1, agent.critic_1), ("critic_2", orc.critic_2, agent.critic_2)]
    init_ok = all(torch.equal(pv.cpu(), orc_p[k].detach()) for _, orc_p, view in nets for k, pv in view.state_dict().items())
    tr = make_transitions(seed + 100, C, S, A, reward, done_frac)
    for i in range(C):
        orc.store_flat(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], bool(tr["done"][i]), tr["hint"][i])
    prios = None
    if prioritized:
        prios = orc.per.tree.leaves().copy()
    agent._dev.fill(pack(tr), prios)
    if prioritized:
        agent.replaymem.tree.tree = orc.per.tree.tree     # reference-order rounding of the inner nodes
    return orc, agent, tr, init_ok


def _np(t):
    return t.detach().cpu().numpy()


def compare_nets(rep, tag, agent, orc, grads, kind, lr=1e-3, tau=0.005, actor_stepped=True, targets_moved=True, actor_inner_steps=1):
    trip = [("actor", agent.actor, orc.actor, orc.opt_a, grads.get("grad_a")),
            ("critic_1", agent.critic_1, orc.critic_1, orc.opt_1, grads.get("grad_c1")),
            ("critic_2", agent.critic_2, orc.critic_2, orc.opt_2, grads.get("grad_c2"))]
    for nm, view, prm, opt, g in trip:
        sd = view.state_dict()
        adam = view.adam_state()
        for k, p in prm.items():
            st = opt.state.get(p, None)
            if st:
                rep.add(f"{tag}{nm}.{k}.exp_avg", _np(adam[k][0]), _np(st["exp_avg"]), GRAD_RTOL, "grad")
                rep.add(f"{tag}{nm}.{k}.exp_avg_sq", _np(adam[k][1]), _np(st["exp_avg_sq"]), 2 * GRAD_RTOL, "grad")
                t = float(st["step"])
                sv = np.sqrt(_np(st["exp_avg_sq"]) / (1 - 0.999 ** t))
                gref = _np(g[k]) if g is not None else _np(st["exp_avg"])
            else:
                sv = gref = np.ones_like(_np(p))
            rep.add_params(f"{tag}{nm}.{k}", _np(sd[k]), _np(p), gref, sv, lr * (actor_inner_steps if nm == "actor" else 1))
    tl = [("target_critic_1", agent.target_critic_1, orc.target_critic_1, "critic_1"),
          ("target_critic_2", agent.target_critic_2, orc.target_critic_2, "critic_2")]
    if kind == "td3":
        tl.append(("target_actor", agent.target_actor, orc.target_actor, "actor"))
    for nm, view, prm, on in tl:
        sd = view.state_dict()
        for k, p in prm.items():
            rep.add_target(f"{tag}{nm}.{k}", _np(sd[k]), _np(p), f"{tag}{on}.{k}", tau if targets_moved else 0.0)


def sync_device_to_oracle(agent, orc, kind):
    """teacher forcing: copy the oracle's parameters, Adam moments and targets into the HBM arena so that every step
    is compared from bit-identical state (Adam amplifies 1e-7 gradient noise into 1e-3 parameter differences on
    near-zero-gradient elements, SURVEY H3, which would otherwise compound across steps)."""
    trip = [(agent.actor, orc.actor, orc.opt_a), (agent.critic_1, orc.critic_1, orc.opt_1), (agent.critic_2, orc.critic_2, orc.opt_2)]
    for view, prm, opt in trip:
        upload_params(view, prm)
        moments = {}
        for k, p in prm.items():
            st = opt.state.get(p, None)
            if st:
                moments[k] = (st["exp_avg"], st["exp_avg_sq"])
        view.load_adam_state(moments)
    upload_params(agent.target_critic_1, orc.target_critic_1)
    upload_params(agent.target_critic_2, orc.target_critic_2)
    if kind == "td3":
        upload_params(agent.target_actor, orc.target_actor)


def sac_parity_report(S=24, A=2, B=32, C=64, steps=2, seed=7, prioritized=False, use_hint=False, verbose=False, trace=True):
    rep = Report(verbose)
    orc, agent, tr, init_ok = build_pair("sac", S, A, B, C, seed, prioritized, use_hint)
    rep.add_exact("init.state_dict (reference draw order)", init_ok, True)
    dev = agent._dev
    np.random.seed(seed + 1)
    torch.manual_seed(seed + 1)
    for s in range(steps):
        tag = f"s{s}."
        if prioritized:
            u = np.random.random_sample(B)
            idxs, data, w = orc.per.sample(u)
            ti, wd = dev.per_sample(u, want_outputs=True)
            rep.add_exact(tag + "per.tree_idx", ti, idxs)
            rep.add_ulp(tag + "per.is_weights", wd, w, 2)
            kw = dict(is_weights=w, tree_idx=idxs)
            batch_idx = data
        else:
            batch_idx = np.random.choice(C, B, replace=False)
            dev.set_batch(batch_idx)
            kw = {}
        e1, e2 = torch.randn(B, A), torch.randn(B, A)
        e3 = torch.randn(B, A) if (use_hint and orc.learn_counter % 10 == 0) else None
        batch = orc.replay.gather(batch_idx)
        if trace and s == 0:
            # forward/backward of the critic phase with the PRE-step oracle weights (before Adam mutates them)
            sT, aT = torch.as_tensor(batch[0]), torch.as_tensor(batch[1])
            with torch.no_grad():
                s2T = torch.as_tensor(batch[3])
                a2, lp2 = N.sac_sample(orc.actor, s2T, e1, 1)
                m = torch.min(N.critic_forward(orc.target_critic_1, s2T, a2), N.critic_forward(orc.target_critic_2, s2T, a2)) - orc.alpha * lp2
                m[torch.as_tensor(batch[4]).unsqueeze(1)] = 0.0
                y_pre = orc.scale * torch.as_tensor(batch[2]).unsqueeze(1) + orc.gamma * m
            wsum = float(np.sum(kw["is_weights"])) if prioritized else float(B)
            traces = []
            for prm in (orc.critic_1, orc.critic_2):
                trc = traced_critic(prm, sT, aT)
                ((wsum / B) * torch.mean((trc["q"] - y_pre) ** 2)).backward()
                traces.append({k: (v.detach().clone(), None if v.grad is None else v.grad.clone()) for k, v in trc.items()})
                for v in prm.values():
                    v.grad = None
        res = orc.learn_on_batch(batch, e1, e2, e3, **kw)
        st = dev.learn(e1.numpy(), e2.numpy(), None if e3 is None else e3.numpy(), want_stats=True)
        if trace and s == 0:
            for k, trc in ((1, traces[0]), (2, traces[1])):
                c = f"c{k}."
                for lay in ("11", "12", "21", "22"):
                    rep.add(tag + c + "h" + lay, _np(dev.buffer(c + "h" + lay)), _np(trc["h" + lay][0]), ACT_RTOL)
                    rep.add(tag + c + "zh" + lay, _np(dev.buffer(c + "zh" + lay)), _np(trc["zh" + lay][0]), 5 * ACT_RTOL)
                    rep.add(tag + c + "rs" + lay, _np(dev.buffer(c + "rs" + lay)), _np(trc["rs" + lay][0]), ACT_RTOL)
                    rep.add(tag + c + "du" + lay, _np(dev.buffer(c + "du" + lay)), _np(trc["u" + lay][1]), GRAD_RTOL, "grad")
                    rep.add(tag + c + "dz" + lay, _np(dev.buffer(c + "dz" + lay)), _np(trc["z" + lay][1]), GRAD_RTOL, "grad")
                rep.add(tag + c + "q", _np(dev.buffer(c + "q")), _np(trc["q"][0]), ACT_RTOL)
            rep.add(tag + "y", _np(dev.buffer("y")), _np(res["y"]), ACT_RTOL)
        rep.add(tag + "critic_1_loss", st["critic_1_loss"], res["critic_1_loss"], 1e-4)
        rep.add(tag + "critic_2_loss", st["critic_2_loss"], res["critic_2_loss"], 1e-4)
        rep.add(tag + "actor_loss", st["actor_loss"], res["actor_loss"], 1e-4)
        if e3 is None:      # on dual-update steps the buffers hold the extra no-grad resample (enet_sac.py:611)
            rep.add(tag + "pi", _np(dev.buffer("a.pi")), _np(res["pi"]), 5 * ACT_RTOL)
            rep.add(tag + "logp", _np(dev.buffer("a.logp")), _np(res["logp"]), 5 * ACT_RTOL)
        if prioritized:
            rep.add(tag + "per.errors", _np(dev.buffer("err")), res["per_errors"], 1e-4)
            from oracle.per_ref import priorities_from_errors_f32
            rep.add_ulp(tag + "per.priorities", _np(dev.buffer("per_ps")), priorities_from_errors_f32(_np(dev.buffer("err")), 100.0), 2)
            tree_dev = _np(dev.buffer("tree")).reshape(-1)
            rep.add(tag + "per.tree", tree_dev, orc.per.tree.tree, 1e-6)
            agent.replaymem.tree.tree = orc.per.tree.tree      # re-sync bits so that later indices stay comparable
        if use_hint:
            rep.add(tag + "rho", st["rho"], float(orc.rho), 1e-4)
        compare_nets(rep, tag, agent, orc, res, "sac")
        sync_device_to_oracle(agent, orc, "sac")
    rep_d = dict(ok=not rep.failures, failures=rep.failures, rows=rep.rows, max_grad_rel=rep.max_grad_rel,
                 launches=dev.launch_count())
    return rep_d


def td3_parity_report(S=24, A=2, B=32, C=64, steps=4, seed=9, prioritized=True, verbose=False, use_hint=False, admm_rho=None):
    rep = Report(verbose)
    orc, agent, tr, init_ok = build_pair("td3", S, A, B, C, seed, prioritized, use_hint, reward="uniform")
    if admm_rho is not None:       # main_td3.py:22
        orc.admm_rho = admm_rho
        agent.admm_rho = admm_rho
        agent._dev.set_scalars(0.0, 0.0, admm_rho)       # what Agent.learn() does when the attribute changed
        agent._admm_rho_sent = admm_rho
    rep.add_exact("init.state_dict (reference draw order)", init_ok, True)
    dev = agent._dev
    np.random.seed(seed + 1)
    for s in range(steps):
        tag = f"s{s}."
        if prioritized:
            u = np.random.random_sample(B)
            idxs, data, w = orc.per.sample(u)
            ti, wd = dev.per_sample(u, want_outputs=True)
            rep.add_exact(tag + "per.tree_idx", ti, idxs)
            rep.add_ulp(tag + "per.is_weights", wd, w, 2)
            kw = dict(is_weights=w, tree_idx=idxs)
            batch_idx = data
        else:
            batch_idx = np.random.choice(C, B, replace=False)
            dev.set_batch(batch_idx)
            kw = {}
        nz = np.random.normal(scale=0.2)
        res = orc.learn_on_batch(orc.replay.gather(batch_idx), nz, **kw)
        st = dev.learn(td3_noise=nz, want_stats=True)
        rep.add(tag + "y", _np(dev.buffer("y")), _np(res["y"]), ACT_RTOL)
        rep.add(tag + "q1", _np(dev.buffer("c1.q")), _np(res["q1"]), ACT_RTOL)
        rep.add(tag + "critic_1_loss", st["critic_1_loss"], res["critic_1_loss"], 1e-4)
        rep.add(tag + "critic_2_loss", st["critic_2_loss"], res["critic_2_loss"], 1e-4)
        rep.add_exact(tag + "actor_step", bool(st["actor_step"]), bool(res["actor_step"]))
        if res["actor_step"] and not use_hint:
            rep.add(tag + "actor_loss", st["actor_loss"], res["actor_loss"], 1e-4)
        if prioritized:
            rep.add(tag + "per.errors", _np(dev.buffer("err")), res["per_errors"], 1e-4)
            from oracle.per_ref import priorities_from_errors_f32
            rep.add_ulp(tag + "per.priorities", _np(dev.buffer("per_ps")), priorities_from_errors_f32(_np(dev.buffer("err")), 100.0), 2)
            rep.add(tag + "per.tree", _np(dev.buffer("tree")).reshape(-1), orc.per.tree.tree, 1e-6)
            agent.replaymem.tree.tree = orc.per.tree.tree
        compare_nets(rep, tag, agent, orc, res, "td3", targets_moved=bool(res["actor_step"]), actor_inner_steps=5 if use_hint else 1)
        sync_device_to_oracle(agent, orc, "td3")
    return dict(ok=not rep.failures, failures=rep.failures, rows=rep.rows, max_grad_rel=rep.max_grad_rel, launches=dev.launch_count())
