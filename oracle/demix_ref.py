"""CPU oracle for the two demixing SAC agents' update.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this file.  The product path never does.

Restates demixing_rl/demix_sac.py (conv SAC, INNER_DIM 16) and demixing_fuzzy/demix_sac.py (same agent with
`use_influence` and INNER_DIM 32; `diff` of the two files shows nothing else) on CPU torch with every random draw injected.
Line numbers: demixing_rl/demix_sac.py.
  init_layer :18-21; CriticNetwork :372-431  [conv trunk + BatchNorm2d + ReLU] | fc1/fc2 on cat((metadata, action)) | head
  ActorNetwork  :435-527  [conv trunk + BatchNorm2d + ELU] | fc11/fc12 on metadata (ReLU) | fc21 (ELU) | fc22mu, fc22logsigma
                sigma = exp(clamp(logsigma, -20, 2)); tanh-squashed sample; logp -= log(max_action (1 - a^2) + 1e-6)
  DemixingAgent.learn :594-682
RNG order inside one reference learn(): np.random.choice(max_mem, B, replace=False) :340 -> Normal.sample (B,A) :611 ->
Normal.rsample (B,A) :629 -> (use_hint and learn_counter % 10 == 0) Normal.sample (B,A) :662.
Differences from the calibration agent (oracle/calib_ref.py) that this file pins: flat metadata of M values, log-sigma heads,
the KLD hint WITHOUT clamping (:636-641), and the dual update `rho += admm_rho * g` computed from a FRESH no-grad sample of
the UPDATED actor in training mode (BatchNorm running statistics move a third time in that step, :659-669).
Parity status: pinned against the unmodified reference modules by tests/golden/demix_*.npz (tests/golden/make_golden_demix.py).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .calib_ref import EPS, LOG_SQRT_2PI, _init_layer, _state, conv_out, split, trunk


def init_critic(input_dims, n_actions, M, inner, use_influence):
    mods = OrderedDict()
    lin = inner
    convs = ()
    if use_influence:
        w, h = input_dims[1], input_dims[2]
        mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
        mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        lin = conv_out(conv_out(conv_out(w))) * conv_out(conv_out(conv_out(h))) * 32 + inner
        convs = ("conv1", "conv2", "conv3")
    mods["fc1"] = nn.Linear(M + n_actions, 128)
    mods["fc2"] = nn.Linear(128, inner)
    mods["head"] = nn.Linear(lin, 1)
    for k in convs + ("fc1", "fc2"):
        _init_layer(mods[k])
    _init_layer(mods["head"], 0.003)
    return _state(mods)


def init_actor(input_dims, n_actions, M, inner, use_influence):
    mods = OrderedDict()
    lin = inner
    convs = ()
    if use_influence:
        w, h = input_dims[1], input_dims[2]
        mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
        mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        lin = conv_out(conv_out(conv_out(w))) * conv_out(conv_out(conv_out(h))) * 32 + inner
        convs = ("conv1", "conv2", "conv3")
    mods["fc11"] = nn.Linear(M, 128)
    mods["fc12"] = nn.Linear(128, inner)
    mods["fc21"] = nn.Linear(lin, 128)
    mods["fc22mu"] = nn.Linear(128, n_actions)
    mods["fc22logsigma"] = nn.Linear(128, n_actions)
    for k in convs + ("fc11", "fc12", "fc21"):
        _init_layer(mods[k])
    _init_layer(mods["fc22mu"], 0.003)
    _init_layer(mods["fc22logsigma"], 0.003)
    return _state(mods)


def critic_forward(p, b, img, meta, action, use_influence, train=True):
    """ref CriticNetwork.forward :416-426"""
    z = F.relu(F.linear(torch.cat((meta, action), 1), p["fc1.weight"], p["fc1.bias"]))
    z = F.relu(F.linear(z, p["fc2.weight"], p["fc2.bias"]))
    if use_influence:
        x = trunk(p, b, img, F.relu, train)
        return F.linear(torch.cat((x, z), 1), p["head.weight"], p["head.bias"])
    return F.linear(z, p["head.weight"], p["head.bias"])


def actor_forward(p, b, img, meta, use_influence, train=True):
    """ref ActorNetwork.forward :484-496"""
    z = F.relu(F.linear(torch.flatten(meta, start_dim=1), p["fc11.weight"], p["fc11.bias"]))
    z = F.relu(F.linear(z, p["fc12.weight"], p["fc12.bias"]))
    if use_influence:
        x = trunk(p, b, img, F.elu, train)
        x = F.elu(F.linear(torch.cat((x, z), 1), p["fc21.weight"], p["fc21.bias"]))
    else:
        x = F.elu(F.linear(z, p["fc21.weight"], p["fc21.bias"]))
    mu = F.linear(x, p["fc22mu.weight"], p["fc22mu.bias"])
    logsigma = torch.clamp(F.linear(x, p["fc22logsigma.weight"], p["fc22logsigma.bias"]), min=-20, max=2)
    return mu, logsigma


def sample_normal(p, b, img, meta, eps, use_influence, max_action=1, train=True, keep=None):
    """ref sample_normal :498-518 with the N(0,1) draw injected (Normal.sample / rsample both equal loc + randn * scale)"""
    mu, logsigma = actor_forward(p, b, img, meta, use_influence, train)
    sigma = logsigma.exp()
    if keep is not None:
        keep["sigma"] = sigma.detach().clone()
    u = mu + eps * sigma
    at = torch.tanh(u)
    action = at * torch.tensor(max_action)
    logp = -((u - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - LOG_SQRT_2PI
    logp = logp - torch.log(max_action * (1 - at.pow(2)) + EPS)
    return action, logp.sum(1, keepdim=True)


def kld_loss(action, hint):
    """ref learn()::kld_loss :636-641 (no clamping)"""
    a = 0.5 * action + 0.5 + EPS
    h = 0.5 * hint + 0.5 + EPS
    return h * (torch.log(h) - torch.log(a))


class DemixReplay:
    def __init__(self, cap, input_dims, M, A, use_influence):
        self.mem_size, self.mem_cntr, self.use_influence = cap, 0, use_influence
        shp = tuple(input_dims) if use_influence else (1, 1, 1)
        self.img = np.zeros((cap, *shp), np.float32)
        self.img2 = np.zeros((cap, *shp), np.float32)
        self.meta = np.zeros((cap, M), np.float32)
        self.meta2 = np.zeros((cap, M), np.float32)
        self.action = np.zeros((cap, A), np.float32)
        self.hint = np.zeros((cap, A), np.float32)
        self.reward = np.zeros(cap, np.float32)
        self.done = np.zeros(cap, bool)

    def store(self, img, meta, action, reward, img2, meta2, done, hint):
        i = self.mem_cntr % self.mem_size
        if self.use_influence:
            self.img[i], self.img2[i] = img, img2
        self.meta[i], self.meta2[i] = meta, meta2
        self.action[i], self.hint[i], self.reward[i], self.done[i] = action, hint, reward, done
        self.mem_cntr += 1

    def gather(self, idx):
        return (self.img[idx], self.meta[idx], self.action[idx], self.reward[idx], self.img2[idx], self.meta2[idx], self.done[idx], self.hint[idx])


class DemixSacRef:
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001, M=30, reward_scale=2,
                 alpha=0.1, hint_threshold=0.1, admm_rho=1.0, use_hint=False, use_influence=True, inner=16):
        self.gamma, self.tau, self.B, self.A, self.M = gamma, tau, int(batch_size), int(n_actions), int(M)
        self.ui, self.inner = use_influence, inner
        self.max_action = 1
        # construction order of the reference agent (:171-177): actor, critic_1, critic_2, targets 1, 2
        self.actor, self.actor_b = split(init_actor(input_dims, n_actions, M, inner, use_influence))
        self.critic_1, self.critic_1_b = split(init_critic(input_dims, n_actions, M, inner, use_influence))
        self.critic_2, self.critic_2_b = split(init_critic(input_dims, n_actions, M, inner, use_influence))
        self.target_critic_1, self.target_critic_1_b = split(init_critic(input_dims, n_actions, M, inner, use_influence))
        self.target_critic_2, self.target_critic_2_b = split(init_critic(input_dims, n_actions, M, inner, use_influence))
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_1 = torch.optim.Adam(list(self.critic_1.values()), lr=lr_c)
        self.opt_2 = torch.optim.Adam(list(self.critic_2.values()), lr=lr_c)
        self.alpha = torch.tensor(alpha)
        self.use_hint, self.hint_threshold, self.admm_rho = use_hint, hint_threshold, admm_rho
        self.rho = torch.tensor(0.0)
        self.zero = torch.tensor(0.)
        self.learn_counter = 0
        self.replay = DemixReplay(max_mem_size, input_dims, M, n_actions, use_influence)
        self._soft(1.0)
        self.last = {}

    def _soft(self, tau):
        with torch.no_grad():
            for onl, tgt in ((self.critic_1, self.target_critic_1), (self.critic_2, self.target_critic_2)):
                for k in tgt:
                    tgt[k].copy_(tau * onl[k] + (1 - tau) * tgt[k])

    def choose_action(self, img, meta, eps):
        """ref choose_action :583-592: actor in eval mode, no reparameterisation"""
        with torch.no_grad():
            i = torch.as_tensor(img)[None, ] if self.ui else 0
            a, _ = sample_normal(self.actor, self.actor_b, i, torch.as_tensor(meta)[None, ], torch.as_tensor(eps), self.ui, self.max_action, train=False)
        return a.numpy()[0]

    def _gfun(self, actions, hint):
        return torch.max(self.zero, (kld_loss(actions, hint) - self.hint_threshold).mean()).pow(2)

    def learn_on_batch(self, batch, eps_next, eps_cur, eps_dual=None):
        img, meta, a, r, img2, meta2, done, hint = (torch.as_tensor(x) for x in batch)
        r, done = r.unsqueeze(1), done.unsqueeze(1)
        eps_next, eps_cur = torch.as_tensor(eps_next), torch.as_tensor(eps_cur)
        ui = self.ui
        out = {}
        with torch.no_grad():
            a2, logp2 = sample_normal(self.actor, self.actor_b, img2, meta2, eps_next, ui, self.max_action)
            m = torch.min(critic_forward(self.target_critic_1, self.target_critic_1_b, img2, meta2, a2, ui),
                          critic_forward(self.target_critic_2, self.target_critic_2_b, img2, meta2, a2, ui)) - self.alpha * logp2
            m[done] = 0.0
            y = r + self.gamma * m
        q1 = critic_forward(self.critic_1, self.critic_1_b, img, meta, a, ui)
        q2 = critic_forward(self.critic_2, self.critic_2_b, img, meta, a, ui)
        l1, l2 = F.mse_loss(q1, y), F.mse_loss(q2, y)
        self.opt_1.zero_grad()
        self.opt_2.zero_grad()
        (l1 + l2).backward()
        out.update(y=y, q1=q1.detach().clone(), q2=q2.detach().clone(), critic_1_loss=l1.item(), critic_2_loss=l2.item())
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_1.items())
        out["grad_c2"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_2.items())
        self.opt_1.step()
        self.opt_2.step()

        act, logp = sample_normal(self.actor, self.actor_b, img, meta, eps_cur, ui, self.max_action, keep=out)
        qp1 = critic_forward(self.critic_1, self.critic_1_b, img, meta, act, ui)
        qp2 = critic_forward(self.critic_2, self.critic_2_b, img, meta, act, ui)
        out.update(q1_pi=qp1.detach().clone(), q2_pi=qp2.detach().clone())
        cv = torch.min(qp1, qp2)
        actor_loss = (self.alpha * logp - cv).mean()
        if self.use_hint:
            g = self._gfun(act, hint)
            actor_loss = actor_loss + 0.5 * self.admm_rho * g * g + self.rho * g
            out["gfun"] = g.item()
        self.opt_a.zero_grad()
        act.retain_grad()
        actor_loss.backward()
        out.update(actor_loss=actor_loss.item(), pi=act.detach().clone(), logp=logp.detach().clone(), d_pi=act.grad.clone())
        out["grad_a"] = OrderedDict((k, v.grad.clone()) for k, v in self.actor.items())
        self.opt_a.step()
        if self.learn_counter % 10 == 0 and self.use_hint:
            with torch.no_grad():
                act2, _ = sample_normal(self.actor, self.actor_b, img, meta, torch.as_tensor(eps_dual), ui, self.max_action)
                self.rho = self.rho + self.admm_rho * self._gfun(act2, hint)
        self.learn_counter += 1
        self._soft(self.tau)
        self.last = out
        return out

    def learn(self):
        if self.replay.mem_cntr < self.B:
            return None
        idx = np.random.choice(min(self.replay.mem_cntr, self.replay.mem_size), self.B, replace=False)
        e1, e2 = torch.randn(self.B, self.A), torch.randn(self.B, self.A)
        e3 = torch.randn(self.B, self.A) if (self.use_hint and self.learn_counter % 10 == 0) else None
        return self.learn_on_batch(self.replay.gather(idx), e1, e2, e3)
