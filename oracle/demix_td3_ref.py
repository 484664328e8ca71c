"""CPU oracle for the demixing conv TD3 agent's update.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this file.  The product path never does.

Restates demixing_rl/demix_td3.py (line numbers below) on CPU torch with every random draw injected:
  PER :23-167 (the elastic-net PER with image + metadata arrays; absolute_error_upper = 100, :34)
  CriticNetwork :231-288  conv trunk + BatchNorm2d + ReLU | fc1/fc2 on cat((ACTION, metadata)) (:279; demix_sac.py has the
                other order) | head on cat((trunk, side))
  ActorNetwork  :291-352  conv trunk + BatchNorm2d + ELU | fc11/fc12 on metadata (ReLU) | fc21 (ELU) | tanh(fc22)
  DemixingAgent :355-407, update_network_parameters :408-432, choose_action :439-456, learn :459-603
RNG order inside one reference learn(): B x np.random.uniform (PER.sample_buffer :100) -> ONE scalar
np.random.normal(scale=0.2) (:493).
Behaviour kept on purpose:
  * `prioritized` is hard-wired True (:374); PER.mse and the actor loss broadcast (B,) x (B,1) -> (B,B), i.e. mean(w) * mean(.)
    (:137-140, :538, :553);
  * the target networks run in eval() mode (running statistics, :484-486), the online critics NEVER leave train() mode: the
    priority pre-pass (:506-511) is a training-mode forward that moves their BatchNorm running statistics, the training pass
    (:516-517) moves them again with the same batch statistics, and every critic_1 forward inside the actor update (:536 / :552)
    moves critic_1's a further time;
  * update_network_parameters (:408-432) blends EVERY state_dict entry -- BatchNorm running_mean / running_var included, and
    num_batches_tracked through float32 arithmetic truncated back to int64 by load_state_dict -- and runs only on actor steps (:603);
  * with use_hint (demixing_rl/main_td3.py:15,20) the actor takes Nadmm = 5 Adam steps per update inside the adaptive-ADMM loop
    (:543-600): y0 is initialised from the ACTIONS (:567), the penalty is re-estimated once, at iteration 3 (:569), against the
    constructor's admm_rho (:591), and is local to one learn() call;
  * the scalar target noise enters as a float64 0-dim tensor added to a float32 batch (:492-493).
Parity status: pinned against the unmodified reference module by tests/golden/demix_td3_*.npz (tests/golden/make_golden_demix_td3.py).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .calib_ref import _init_layer, _state, conv_out, split, trunk
from .demix_ref import DemixReplay, init_critic as _init_critic_sac
from .per_ref import PerRef
from .sac_ref import per_mse


def init_critic(input_dims, n_actions, M):
    """:232-271 -- same modules, shapes and init_layer calls as demix_sac.py's critic (fc1 is Linear(n_actions + M, 128))"""
    return _init_critic_sac(input_dims, n_actions, M, 16, True)


def init_actor(input_dims, n_actions, M):
    """:292-331"""
    w, h = input_dims[1], input_dims[2]
    mods = OrderedDict()
    mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
    mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["fc11"] = nn.Linear(M, 128)
    mods["fc12"] = nn.Linear(128, 16)
    mods["fc21"] = nn.Linear(conv_out(conv_out(conv_out(w))) * conv_out(conv_out(conv_out(h))) * 32 + 16, 128)
    mods["fc22"] = nn.Linear(128, n_actions)
    for k in ("conv1", "conv2", "conv3", "fc11", "fc12", "fc21"):
        _init_layer(mods[k])
    _init_layer(mods["fc22"], 0.003)
    return _state(mods)


def critic_forward(p, b, img, action, meta, train):
    """:273-283"""
    x = trunk(p, b, img, F.relu, train)
    y = F.relu(F.linear(torch.cat((action, torch.flatten(meta, start_dim=1)), 1), p["fc1.weight"], p["fc1.bias"]))
    y = F.relu(F.linear(y, p["fc2.weight"], p["fc2.bias"]))
    return F.linear(torch.cat((x, y), 1), p["head.weight"], p["head.bias"])


def actor_forward(p, b, img, meta, train):
    """:333-343"""
    x = trunk(p, b, img, F.elu, train)
    z = F.relu(F.linear(torch.flatten(meta, start_dim=1), p["fc11.weight"], p["fc11.bias"]))
    z = F.relu(F.linear(z, p["fc12.weight"], p["fc12.bias"]))
    x = F.elu(F.linear(torch.cat((x, z), 1), p["fc21.weight"], p["fc21.bias"]))
    return torch.tanh(F.linear(x, p["fc22.weight"], p["fc22.bias"]))


def blend_state_dict(onl_p, onl_b, tgt_p, tgt_b, tau):
    """:413-418: tau * online + (1 - tau) * target over every state_dict entry, written back through load_state_dict (copy_)"""
    with torch.no_grad():
        for src, dst in ((onl_p, tgt_p), (onl_b, tgt_b)):
            for k in dst:
                dst[k].copy_(tau * src[k].detach().clone() + (1 - tau) * dst[k].detach().clone())


class DemixTd3Ref:
    NETS = ("actor", "critic_1", "critic_2", "target_actor", "target_critic_1", "target_critic_2")

    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001, M=30,
                 update_actor_interval=2, warmup=1000, noise=0.1, use_hint=False):
        self.gamma, self.tau, self.B, self.A, self.M = gamma, tau, int(batch_size), int(n_actions), int(M)
        self.update_actor_interval, self.warmup, self.noise, self.use_hint = update_actor_interval, warmup, noise, use_hint
        self.admm_rho, self.Nadmm, self.adaptive_admm, self.corr_min = 0.1, 5, True, 0.5
        self.learn_step_cntr, self.time_step = 0, 0
        # construction order of the reference (:380-389): actor, critic_1, critic_2, target actor, target critics 1, 2
        self.actor, self.actor_b = split(init_actor(input_dims, n_actions, M))
        self.critic_1, self.critic_1_b = split(init_critic(input_dims, n_actions, M))
        self.critic_2, self.critic_2_b = split(init_critic(input_dims, n_actions, M))
        self.target_actor, self.target_actor_b = split(init_actor(input_dims, n_actions, M))
        self.target_critic_1, self.target_critic_1_b = split(init_critic(input_dims, n_actions, M))
        self.target_critic_2, self.target_critic_2_b = split(init_critic(input_dims, n_actions, M))
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_1 = torch.optim.Adam(list(self.critic_1.values()), lr=lr_c)
        self.opt_2 = torch.optim.Adam(list(self.critic_2.values()), lr=lr_c)
        self.replay = DemixReplay(max_mem_size, input_dims, M, n_actions, True)
        self.per = PerRef(max_mem_size, 100.0)
        self._sync_targets(1.0)
        self.last = {}
        self.dtype = torch.float32

    def double_shadow(self):
        """A float64 copy of the current state (parameters, buffers; fresh optimizers), used by tests/ and tools/ to measure how well
        conditioned a step is: where the float32 gradients of the reference differ from their float64 evaluation by more than rounding
        (a ReLU pre-activation within an ulp of zero flips a whole gradient path), a second float32 implementation cannot be held to
        rounding-level agreement either."""
        import copy
        sh = copy.copy(self)
        sh.dtype = torch.float64
        for nm in self.NETS:
            setattr(sh, nm, OrderedDict((k, v.detach().double().requires_grad_(True)) for k, v in getattr(self, nm).items()))
            setattr(sh, nm + "_b", OrderedDict((k, v.clone() if v.dtype == torch.int64 else v.double()) for k, v in getattr(self, nm + "_b").items()))
        for on, nm in (("opt_a", "actor"), ("opt_1", "critic_1"), ("opt_2", "critic_2")):
            old = getattr(self, on)
            new = torch.optim.Adam(list(getattr(sh, nm).values()), lr=old.param_groups[0]["lr"])
            for p_old, p_new in zip(getattr(self, nm).values(), getattr(sh, nm).values()):       # carry the moments and step counts over
                st = old.state.get(p_old)
                if st:
                    new.state[p_new] = {"step": st["step"].clone(), "exp_avg": st["exp_avg"].double(), "exp_avg_sq": st["exp_avg_sq"].double()}
            setattr(sh, on, new)
        sh.per = copy.deepcopy(self.per)
        sh.last = {}
        return sh

    def _sync_targets(self, tau):
        for nm in ("actor", "critic_1", "critic_2"):
            blend_state_dict(getattr(self, nm), getattr(self, nm + "_b"), getattr(self, "target_" + nm), getattr(self, "target_" + nm + "_b"), tau)

    def store(self, img, meta, action, reward, img2, meta2, done, hint):
        """PER.store_transition :60-78 with error=None: priority = max leaf (100 while the tree is empty); slot = tree.add"""
        slot = self.per.store(None)
        r = self.replay
        r.img[slot], r.img2[slot], r.meta[slot], r.meta2[slot] = img, img2, meta, meta2
        r.action[slot], r.hint[slot], r.reward[slot], r.done[slot] = action, hint, reward, done
        r.mem_cntr += 1

    def choose_action(self, img, meta, noise):
        """:439-456 after the warm-up: eval-mode actor, additive N(0, noise^2) draw (injected, float64 -> float32 tensor), clamp"""
        with torch.no_grad():
            mu = actor_forward(self.actor, self.actor_b, torch.as_tensor(img)[None, ], torch.as_tensor(meta)[None, ], False)
            mu = torch.clamp(mu + torch.tensor(np.asarray(noise), dtype=torch.float), -1, 1)
        return mu.numpy()

    def learn_on_batch(self, batch, target_noise, is_weights, tree_idx):
        img, meta, a, r, img2, meta2, done, hint = (torch.as_tensor(x) for x in batch)
        img, meta, a, r, img2, meta2, hint = (x.to(self.dtype) for x in (img, meta, a, r, img2, meta2, hint))
        w = torch.as_tensor(is_weights).to(self.dtype)
        out = {}
        with torch.no_grad():       # (the reference leaves this graph attached; the online networks' gradients are unaffected)
            ta = actor_forward(self.target_actor, self.target_actor_b, img2, meta2, False)
            ta = ta + torch.clamp(torch.tensor(np.float64(target_noise)), -0.5, 0.5)
            ta = torch.clamp(ta, -1, 1)
            q1_ = critic_forward(self.target_critic_1, self.target_critic_1_b, img2, ta, meta2, False)
            q2_ = critic_forward(self.target_critic_2, self.target_critic_2_b, img2, ta, meta2, False)
            q1_[done] = 0.0
            q2_[done] = 0.0
            y = (r + self.gamma * torch.min(q1_.view(-1), q2_.view(-1))).view(self.B, 1)
            # priority pre-pass :506-511 (training-mode BatchNorm: running statistics move)
            e1 = torch.abs(critic_forward(self.critic_1, self.critic_1_b, img, a, meta, True) - y).numpy()
            e2 = torch.abs(critic_forward(self.critic_2, self.critic_2_b, img, a, meta, True) - y).numpy()
        err = 0.5 * (e1 + e2)
        out["per_errors"] = err
        out["per_ps"] = self.per.update(tree_idx, err)

        q1 = critic_forward(self.critic_1, self.critic_1_b, img, a, meta, True)
        q2 = critic_forward(self.critic_2, self.critic_2_b, img, a, meta, True)
        l1, l2 = per_mse(y, q1, w), per_mse(y, q2, w)
        self.opt_1.zero_grad()
        self.opt_2.zero_grad()
        (l1 + l2).backward()
        out.update(y=y, q1=q1.detach().clone(), q2=q2.detach().clone(), critic_1_loss=l1.item(), critic_2_loss=l2.item(), target_action=ta)
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_1.items())
        out["grad_c2"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_2.items())
        self.opt_1.step()
        self.opt_2.step()

        self.learn_step_cntr += 1
        out["actor_step"] = self.learn_step_cntr % self.update_actor_interval == 0
        if out["actor_step"]:
            if not self.use_hint:
                self.opt_a.zero_grad()
                act = actor_forward(self.actor, self.actor_b, img, meta, True)
                q = critic_forward(self.critic_1, self.critic_1_b, img, act, meta, True)
                loss = -torch.mean(q * w)
                loss.backward()
                out.update(actor_loss=loss.item(), pi=act.detach().clone(), q1_pi=q.detach().clone())
                out["grad_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.actor.items())
                self.opt_a.step()
            else:
                self._admm_actor(img, meta, hint, w, out)
            self._sync_targets(self.tau)
        self.last = out
        return out

    def _admm_actor(self, img, meta, hint, w, out):
        """:543-600"""
        lag_y = torch.zeros(hint.numel(), dtype=self.dtype)
        lag_y0 = act0 = None
        rho = self.admm_rho
        out["admm"] = []
        for it in range(self.Nadmm):
            self.opt_a.zero_grad()
            act = actor_forward(self.actor, self.actor_b, img, meta, True)
            q = critic_forward(self.critic_1, self.critic_1_b, img, act, meta, True)
            loss = -torch.mean(q * w)
            diff = (act - hint).view(-1)
            pen = ((torch.dot(lag_y, diff) + rho / 2 * F.mse_loss(act, hint)) * w).mean() / act.numel()
            loss = loss + pen
            loss.backward()
            rec = dict(pi=act.detach().clone(), q1_pi=q.detach().clone(), actor_loss=loss.item(), penalty=pen.item(), rho=float(rho),
                       grad_a=OrderedDict((k, v.grad.clone()) for k, v in self.actor.items()))
            out["admm"].append(rec)
            self.opt_a.step()
            with torch.no_grad():
                lag_y = lag_y + rho * (act - hint).view(-1)
                if self.adaptive_admm:
                    if it == 0:
                        lag_y0 = act.view(-1).detach().clone()
                        act0 = act.view(-1).detach().clone()
                    elif it % 3 == 0 and it < self.Nadmm - 1:
                        lag_y1 = lag_y + rho * (act - hint).view(-1)
                        dy, du = lag_y1 - lag_y0, act.view(-1).detach() - act0
                        d11, d12, d22 = torch.dot(dy, dy), torch.dot(dy, du), torch.dot(du, du)
                        lag_y0, act0 = lag_y1, act.view(-1).detach().clone()
                        if d11 > 0 and d12 > 0 and d22 > 0:
                            corr = d12 / torch.sqrt(d11 * d22)
                            a_sd, a_mg = d11 / d12, d12 / d22
                            a_hat = a_mg if 2 * a_mg > a_sd else a_sd - 0.5 * a_mg
                            if corr > self.corr_min and a_hat < 10 * self.admm_rho and a_hat > 0.1 * self.admm_rho:
                                rho = a_hat
            rec["lagrange"] = lag_y.clone()
        out.update(actor_loss=loss.item(), pi=act.detach().clone(), admm_rho_final=float(rho))

    def learn(self):
        """reference-shaped learn() on the global numpy stream (CPU baseline)"""
        if self.replay.mem_cntr < self.B:
            return None
        idxs, data, w = self.per.sample(np.random.random_sample(self.B))
        return self.learn_on_batch(self.replay.gather(data), np.random.normal(scale=0.2), w, idxs)
