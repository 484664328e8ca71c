"""CPU oracle for the TD3 agent update.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Restates elasticnet/enet_td3.py:124-364 on CPU torch with injected randomness:
  RNG order inside one reference learn(): B x np.random.uniform (PER, enet_sac.py:294) or
  np.random.choice (:48)  ->  ONE scalar np.random.normal(scale=0.2) (enet_td3.py:249).
Quirks kept on purpose: target-smoothing noise is a single scalar per step (:248-249);
priorities are refreshed BEFORE the critic step (:264-269); PER.mse and the actor loss broadcast
(B,)x(B,1)->(B,B), i.e. mean(w)*mean(.) (:281-282,:307); the critic closure runs twice but each
critic receives exactly one Adam step on gradients taken at its pre-step weights (:274-293);
actor AND all three targets move only every `update_actor_interval`-th step (:298,:364).
The ADMM hint loop (:310-361) is restated in learn_on_batch(use_hint=True).

Parity status: pinned against the reference by tests/golden/td3_*.npz.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import nets_ref as N
from .per_ref import PerRef
from .sac_ref import HostReplay, _leafs, per_mse, soft_update_like_reference


class Td3Ref:
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001,
                 update_actor_interval=2, warmup=1000, noise=0.1, prioritized=False, use_hint=False,
                 per_clip=100.0):
        S, A = int(input_dims[0]), int(n_actions)
        self.S, self.A, self.B = S, A, int(batch_size)
        self.gamma, self.tau = gamma, tau
        self.update_actor_interval, self.warmup, self.noise = update_actor_interval, warmup, noise
        self.use_hint = use_hint
        self.admm_rho, self.Nadmm, self.adaptive_admm, self.corr_min = 0.1, 5, True, 0.5
        self.learn_step_cntr, self.time_step = 0, 0
        # construction order of the reference (:151-159): actor, critic_1, critic_2, target actor, targets 1, 2
        self.actor = _leafs(N.init_actor(S, A, "td3"))
        self.critic_1 = _leafs(N.init_critic(S, A))
        self.critic_2 = _leafs(N.init_critic(S, A))
        self.target_actor = _leafs(N.init_actor(S, A, "td3"))
        self.target_critic_1 = _leafs(N.init_critic(S, A))
        self.target_critic_2 = _leafs(N.init_critic(S, A))
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_1 = torch.optim.Adam(list(self.critic_1.values()), lr=lr_c)
        self.opt_2 = torch.optim.Adam(list(self.critic_2.values()), lr=lr_c)
        self.replay = HostReplay(max_mem_size, S, A)
        self.per = PerRef(max_mem_size, per_clip) if prioritized else None
        self._sync_targets(1.0)
        self.last = {}

    def _sync_targets(self, tau):
        soft_update_like_reference(self.actor, self.target_actor, tau)
        soft_update_like_reference(self.critic_1, self.target_critic_1, tau)
        soft_update_like_reference(self.critic_2, self.target_critic_2, tau)

    def store_flat(self, s, a, r, s2, d, h):
        if self.per is not None:
            slot = self.per.store(float(r))         # ref :205 passes the reward as the error
        else:
            slot = self.replay.mem_cntr % self.replay.mem_size
        self.replay.store_flat(slot, s, a, r, s2, d, h)
        self.replay.mem_cntr += 1

    def learn_on_batch(self, batch, target_noise, is_weights=None, tree_idx=None):
        s, a, r, s2, done, hint = batch
        s, s2, a, hint, r = (torch.as_tensor(x) for x in (s, s2, a, hint, r))
        done = torch.as_tensor(done)
        out = {}
        w = None if is_weights is None else torch.as_tensor(is_weights)

        with torch.no_grad():   # the reference leaves this graph attached; online grads are unaffected
            ta = N.td3_actor_forward(self.target_actor, s2)
            ta = ta + torch.clamp(torch.tensor(np.float64(target_noise)), -0.5, 0.5)
            ta = torch.clamp(ta, -1, 1)
            q1_ = N.critic_forward(self.target_critic_1, s2, ta)
            q2_ = N.critic_forward(self.target_critic_2, s2, ta)
            q1_[done] = 0.0
            q2_[done] = 0.0
            y = (r + self.gamma * torch.min(q1_.view(-1), q2_.view(-1))).view(self.B, 1)

        if w is not None:
            with torch.no_grad():
                e1 = torch.abs(N.critic_forward(self.critic_1, s, a) - y).numpy()
                e2 = torch.abs(N.critic_forward(self.critic_2, s, a) - y).numpy()
            err = 0.5 * (e1 + e2)
            out["per_errors"] = err
            if self.per is not None:
                out["per_ps"] = self.per.update(tree_idx, err)

        q1 = N.critic_forward(self.critic_1, s, a)
        q2 = N.critic_forward(self.critic_2, s, a)
        if w is not None:
            l1, l2 = per_mse(y, q1, w), per_mse(y, q2, w)
        else:
            l1, l2 = F.mse_loss(y, q1), F.mse_loss(y, q2)
        self.opt_1.zero_grad()
        self.opt_2.zero_grad()
        (l1 + l2).backward()
        out.update(y=y, q1=q1.detach().clone(), q2=q2.detach().clone(), critic_1_loss=l1.item(), critic_2_loss=l2.item())
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_1.items())
        out["grad_c2"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_2.items())
        self.opt_1.step()
        self.opt_2.step()

        self.learn_step_cntr += 1
        out["actor_step"] = self.learn_step_cntr % self.update_actor_interval == 0
        if out["actor_step"]:
            if not self.use_hint:
                self.opt_a.zero_grad()
                q = N.critic_forward(self.critic_1, s, N.td3_actor_forward(self.actor, s))
                loss = -torch.mean(q) if w is None else -torch.mean(q * w)
                loss.backward()
                out["actor_loss"] = loss.item()
                out["grad_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.actor.items())
                self.opt_a.step()
            else:
                self._admm_actor(s, hint, w, out)
            self._sync_targets(self.tau)
        self.last = out
        return out

    def _admm_actor(self, s, hint, w, out):
        """ref enet_td3.py:311-361 (five actor steps, adaptive penalty)."""
        lag_y = torch.zeros(hint.numel())
        lag_y0 = act0 = None
        rho = self.admm_rho
        for it in range(self.Nadmm):
            self.opt_a.zero_grad()
            act = N.td3_actor_forward(self.actor, s)
            q = N.critic_forward(self.critic_1, s, act)
            loss = -torch.mean(q) if w is None else -torch.mean(q * w)
            diff = (act - hint).view(-1)
            pen = torch.dot(lag_y, diff) + rho / 2 * F.mse_loss(act, hint)
            pen = (pen if w is None else pen * w).mean() / act.numel()
            loss = loss + pen
            loss.backward()
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
        out["actor_loss"] = loss.item()
        out["admm_rho_final"] = float(rho)

    def learn(self):
        """Reference-shaped learn() on the global numpy stream (CPU baseline)."""
        if self.replay.mem_cntr < self.B:
            return None
        if self.per is not None:
            idxs, data, w = self.per.sample(np.random.random_sample(self.B))
            batch, kw = self.replay.gather(data), dict(is_weights=w, tree_idx=idxs)
        else:
            idx = np.random.choice(min(self.replay.mem_cntr, self.replay.mem_size), self.B, replace=False)
            batch, kw = self.replay.gather(idx), {}
        return self.learn_on_batch(batch, np.random.normal(scale=0.2), **kw)
