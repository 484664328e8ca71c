"""CPU oracle for the SAC agent update (learn / choose_action).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Restates elasticnet/enet_sac.py:478-628 (Agent.__init__, update_network_parameters,
choose_action, learn) on CPU torch, with every random draw INJECTED so that the CUDA path
and this oracle consume identical minibatch indices and N(0,1) noise:

  RNG order inside one reference learn() (enet_sac.py): np.random.choice(max_mem,B,replace=False) :48
  -> Normal.sample (B,A) :570 -> Normal.rsample (B,A) :589 -> [every 10th call with hints]
  Normal.sample (B,A) :611.

"SAC + PER" has no reference implementation (the `prioritized` flag is dead in enet_sac.Agent,
:488-490).  The `per=` mode below is the DERIVED composition: reference SAC math + reference PER
class wired the way enet_td3.py wires it (:231-232 sample, :264-269 priority update before the
critic step, :281-282 PER.mse with its (B,)x(B,1)->(B,B) broadcast, :307 actor loss x is_weight
with the same broadcast).  Tests that use it are labelled "derived".

Parity status: pinned against the reference by tests/golden/sac_*.npz.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

from . import nets_ref as N
from .per_ref import PerRef


def _leafs(params, device="cpu"):
    return OrderedDict((k, v.clone().to(device).requires_grad_(True)) for k, v in params.items())


def soft_update_like_reference(online, target, tau):
    """ref update_network_parameters enet_sac.py:523-542: tau*p.clone() + (1-tau)*t.clone(),
    then a load_state_dict style copy.  Kept op-for-op because this is ~40 % of the CPU step."""
    fresh = {}
    for k in target:
        fresh[k] = tau * online[k].clone() + (1 - tau) * target[k].clone()
    with torch.no_grad():
        for k in target:
            target[k].copy_(fresh[k])


def per_mse(expected, targets, is_weights):
    """ref PER.mse enet_sac.py:326-329 -- note (B,)*(B,1) broadcasts to (B,B)."""
    td = expected - targets
    wse = is_weights * td * td
    return torch.sum(wse) / torch.numel(wse)


class HostReplay:
    """Host ring of the six reference arrays (enet_sac.py:23-44)."""

    def __init__(self, cap, S, A):
        self.mem_size, self.mem_cntr = cap, 0
        self.state_memory = np.zeros((cap, S), np.float32)
        self.new_state_memory = np.zeros((cap, S), np.float32)
        self.action_memory = np.zeros((cap, A), np.float32)
        self.reward_memory = np.zeros(cap, np.float32)
        self.terminal_memory = np.zeros(cap, bool)
        self.hint_memory = np.zeros((cap, A), np.float32)

    def store_flat(self, slot, s, a, r, s2, d, h):
        self.state_memory[slot] = s
        self.new_state_memory[slot] = s2
        self.action_memory[slot] = a
        self.reward_memory[slot] = r
        self.terminal_memory[slot] = d
        self.hint_memory[slot] = h

    def gather(self, idx):
        return (self.state_memory[idx], self.action_memory[idx], self.reward_memory[idx],
                self.new_state_memory[idx], self.terminal_memory[idx], self.hint_memory[idx])


class SacRef:
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100,
                 tau=0.001, reward_scale=2, alpha=0.1, use_hint=False, prioritized=False, per_clip=100.0, device="cpu"):
        # device="cuda" runs the same op sequence as torch-eager CUDA kernels (the reference's own use_cuda=True path,
        # enet_sac.py:12-16): bench.py's `gpu_eager_baseline`, never a parity oracle
        S, A = int(input_dims[0]), int(n_actions)
        self.dev = torch.device(device)
        self.S, self.A, self.B = S, A, int(batch_size)
        self.gamma, self.tau, self.scale = gamma, tau, reward_scale
        self.max_action = 1
        # construction order of the reference Agent: actor, critic_1, critic_2, targets 1, 2 (:493-499)
        self.actor = _leafs(N.init_actor(S, A, "sac"), self.dev)
        self.critic_1 = _leafs(N.init_critic(S, A), self.dev)
        self.critic_2 = _leafs(N.init_critic(S, A), self.dev)
        self.target_critic_1 = _leafs(N.init_critic(S, A), self.dev)
        self.target_critic_2 = _leafs(N.init_critic(S, A), self.dev)
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_1 = torch.optim.Adam(list(self.critic_1.values()), lr=lr_c)
        self.opt_2 = torch.optim.Adam(list(self.critic_2.values()), lr=lr_c)
        self.alpha = torch.tensor(alpha, device=self.dev)
        self.zero = torch.tensor(0., device=self.dev)
        self.use_hint = use_hint
        self.hint_threshold, self.admm_rho = 0.1, 0.01
        self.rho = torch.tensor(0.0, device=self.dev)
        self.replay = HostReplay(max_mem_size, S, A)
        self.per = PerRef(max_mem_size, per_clip) if prioritized else None
        soft_update_like_reference(self.critic_1, self.target_critic_1, 1.0)
        soft_update_like_reference(self.critic_2, self.target_critic_2, 1.0)
        self.learn_counter = 0
        self.last = {}

    # ---- replay -------------------------------------------------------------------------
    def store_flat(self, s, a, r, s2, d, h, error=None):
        if self.per is not None:
            slot = self.per.store(error)
        else:
            slot = self.replay.mem_cntr % self.replay.mem_size
        self.replay.store_flat(slot, s, a, r, s2, d, h)
        self.replay.mem_cntr += 1

    # ---- inference ----------------------------------------------------------------------
    def choose_action(self, state_vec, eps):
        """ref choose_action :547-553 on an already concatenated (S,) state, eps (A,)."""
        with torch.no_grad():
            a, _ = N.sac_sample(self.actor, torch.as_tensor(state_vec).to(self.dev), torch.as_tensor(eps).to(self.dev), self.max_action)
        return a.cpu().numpy()[0]

    # ---- update -------------------------------------------------------------------------
    def _gfun(self, actions, hint):
        return torch.max(self.zero, (F.mse_loss(actions, hint) - self.hint_threshold).mean()).pow(2)

    def learn_on_batch(self, batch, eps_next, eps_cur, eps_dual=None, is_weights=None, tree_idx=None):
        """One reference learn() (:555-628) on an explicit minibatch and explicit noise."""
        s, a, r, s2, done, hint = batch
        T = lambda x: torch.as_tensor(x).to(self.dev)
        s, s2, a, hint = (T(x) for x in (s, s2, a, hint))
        r = self.scale * T(r).unsqueeze(1)
        done = T(done).unsqueeze(1)
        eps_next, eps_cur = T(eps_next), T(eps_cur)
        out = {}

        with torch.no_grad():
            a2, logp2 = N.sac_sample(self.actor, s2, eps_next, self.max_action)
            m = torch.min(N.critic_forward(self.target_critic_1, s2, a2),
                          N.critic_forward(self.target_critic_2, s2, a2)) - self.alpha * logp2
            m[done] = 0.0
            y = r + self.gamma * m
        q1 = N.critic_forward(self.critic_1, s, a)
        q2 = N.critic_forward(self.critic_2, s, a)
        if is_weights is not None:                       # derived SAC+PER (TD3 wiring)
            w = T(is_weights)
            err = 0.5 * (torch.abs(q1 - y).detach().cpu().numpy() + torch.abs(q2 - y).detach().cpu().numpy())
            out["per_errors"] = err
            if self.per is not None:
                out["per_ps"] = self.per.update(tree_idx, err)
            l1, l2 = per_mse(y, q1, w), per_mse(y, q2, w)
        else:
            l1, l2 = F.mse_loss(q1, y), F.mse_loss(q2, y)
        self.opt_1.zero_grad()
        self.opt_2.zero_grad()
        (l1 + l2).backward()
        out.update(y=y, q1=q1.detach().clone(), q2=q2.detach().clone(), critic_1_loss=l1.item(), critic_2_loss=l2.item())
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_1.items())
        out["grad_c2"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_2.items())
        self.opt_1.step()
        self.opt_2.step()

        act, logp = N.sac_sample(self.actor, s, eps_cur, self.max_action)
        cv = torch.min(N.critic_forward(self.critic_1, s, act), N.critic_forward(self.critic_2, s, act))
        if is_weights is not None:
            actor_loss = torch.mean((self.alpha * logp - cv) * w)
        else:
            actor_loss = (self.alpha * logp - cv).mean()
        if self.use_hint:
            g = self._gfun(act, hint)
            actor_loss = actor_loss + 0.5 * self.admm_rho * g * g + self.rho * g
            out["gfun"] = g.item()
        self.opt_a.zero_grad()
        actor_loss.backward()
        out.update(actor_loss=actor_loss.item(), pi=act.detach().clone(), logp=logp.detach().clone())
        out["grad_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.actor.items())
        self.opt_a.step()

        if self.learn_counter % 10 == 0 and self.use_hint:
            with torch.no_grad():
                act, _ = N.sac_sample(self.actor, s, T(eps_dual), self.max_action)
                self.rho = self.rho + self.admm_rho * self._gfun(act, hint)
        self.learn_counter += 1
        soft_update_like_reference(self.critic_1, self.target_critic_1, self.tau)
        soft_update_like_reference(self.critic_2, self.target_critic_2, self.tau)
        self.last = out
        return out

    def learn(self):
        """Reference-shaped learn(): host sampling, global numpy/torch RNG streams (CPU baseline)."""
        if self.replay.mem_cntr < self.B:
            return None
        BA = (self.B, self.A)
        if self.per is not None:
            idxs, data, w = self.per.sample(np.random.random_sample(self.B))
            batch = self.replay.gather(data)
            kw = dict(is_weights=w, tree_idx=idxs)
        else:
            idx = np.random.choice(min(self.replay.mem_cntr, self.replay.mem_size), self.B, replace=False)
            batch, kw = self.replay.gather(idx), {}
        e1, e2 = torch.randn(BA), torch.randn(BA)
        e3 = torch.randn(BA) if (self.use_hint and self.learn_counter % 10 == 0) else None
        return self.learn_on_batch(batch, e1, e2, e3, **kw)
