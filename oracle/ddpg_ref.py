"""CPU oracle for the DDPG agent update.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Restates elasticnet/enet_ddpg.py:192-303 on CPU torch with the minibatch injected:
  RNG inside one reference learn(): np.random.choice(max_mem, B, replace=False) (:68) only.
  (choose_action adds Ornstein-Uhlenbeck noise drawn from np.random.normal on the host, :243-249 / :32-36.)
Reference behaviour kept on purpose: the target is r + gamma*Q'(s', mu'(s'))*(1-done) with `done` a float32 array
(:267-271); the critic loss is the SUM of squared Bellman errors, ||q - y||_2^2 (:281), not the mean; the actor loss is
mean(-Q(s, mu(s))) (:293-294); actor, critic and both targets move on every call (:288,:300,:303); the critic has the
same fc22-keeps-default-init quirk as enet_sac (:111-115).

Parity status: pinned against the reference by tests/golden/ddpg_*.npz.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import nets_ref as N
from .sac_ref import HostReplay, _leafs, soft_update_like_reference


class DdpgRef:
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001):
        S, A = int(input_dims[0]), int(n_actions)
        self.S, self.A, self.B = S, A, int(batch_size)
        self.gamma, self.tau = gamma, tau
        # construction order of the reference (:202-209): actor, critic, target actor, target critic
        self.actor = _leafs(N.init_actor(S, A, "td3"))
        self.critic = _leafs(N.init_critic(S, A))
        self.target_actor = _leafs(N.init_actor(S, A, "td3"))
        self.target_critic = _leafs(N.init_critic(S, A))
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_c = torch.optim.Adam(list(self.critic.values()), lr=lr_c)
        self.replay = HostReplay(max_mem_size, S, A)
        self._sync_targets(1.0)
        self.last = {}

    def _sync_targets(self, tau):
        soft_update_like_reference(self.critic, self.target_critic, tau)
        soft_update_like_reference(self.actor, self.target_actor, tau)

    def store_flat(self, s, a, r, s2, d):
        slot = self.replay.mem_cntr % self.replay.mem_size
        self.replay.store_flat(slot, s, a, r, s2, d, np.zeros(self.A, np.float32))
        self.replay.mem_cntr += 1

    def choose_action(self, state_vec, ou_noise):
        with torch.no_grad():
            mu = N.td3_actor_forward(self.actor, torch.as_tensor(state_vec))
        return (mu + torch.tensor(ou_noise, dtype=torch.float)).numpy()

    def learn_on_batch(self, batch):
        s, a, r, s2, done = batch[:5]
        s, s2, a, r = (torch.as_tensor(x) for x in (s, s2, a, r))
        done = torch.as_tensor(np.asarray(done, np.float32))
        out = {}
        with torch.no_grad():
            q_ = N.critic_forward(self.target_critic, s2, N.td3_actor_forward(self.target_actor, s2))
            y = (r + self.gamma * q_.view(-1) * (1 - done)).view(self.B, 1)
        self.opt_c.zero_grad()
        q = N.critic_forward(self.critic, s, a)
        critic_loss = torch.norm(q - y, 2) ** 2
        critic_loss.backward()
        out.update(y=y, q1=q.detach().clone(), critic_loss=critic_loss.item())
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic.items())
        self.opt_c.step()
        self.opt_a.zero_grad()
        pi = N.td3_actor_forward(self.actor, s)
        actor_loss = torch.mean(-N.critic_forward(self.critic, s, pi))
        actor_loss.backward()
        out.update(actor_loss=actor_loss.item(), pi=pi.detach().clone())
        out["grad_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.actor.items())
        self.opt_a.step()
        self._sync_targets(self.tau)
        self.last = out
        return out

    def learn(self):
        if self.replay.mem_cntr < self.B:
            return None
        idx = np.random.choice(min(self.replay.mem_cntr, self.replay.mem_size), self.B, replace=False)
        return self.learn_on_batch(self.replay.gather(idx))
