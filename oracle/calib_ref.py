"""CPU oracle for the calibration conv-SAC agent update.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Restates calibration/calib_sac.py on CPU torch with every random draw injected:
  init_layer :21-24 (U(-sc, sc), sc = 1/sqrt(weight.size(0)) = OUT channels/features, 0.003 for last layers)
  CriticNetwork :90-144  conv(1->16->32->32, k5 s2) + BatchNorm2d + ReLU | fc1/fc2 on [sky, action] (call-site order) | head
  ActorNetwork  :153-231 conv trunk + BatchNorm2d + ELU | fc11/fc12 on sky (ReLU) | fc21 (ELU) | mu, sigma=clamp(.,1e-6,1)
  Agent.learn   :300-378
RNG order inside one reference learn(): np.random.choice(max_mem, B, replace=False) :60 -> Normal.sample (B,A) :318
-> Normal.rsample (B,A) :340.  Both draws equal loc + randn * scale bit for bit, so N(0,1) noise is injected.
Behaviour kept on purpose: every network stays in train() mode inside learn(), so BatchNorm uses BATCH statistics in
all eight forward passes (the no-grad ones included) and updates its running statistics each time (momentum 0.1,
unbiased variance); only choose_action runs the actor in eval() mode (:285-295).  The target is r + gamma*(min Q' -
alpha*logp') with NO reward scale (:320-323; `self.scale` is unused).  The soft update walks parameters() only, the
targets' BatchNorm buffers are never copied (:271-276).  With use_hint the penalty is g = max(0, mean(KLD) - thr)^2 and
the loss adds admm_rho/2*g^2 + rho*g (:343-365); rho += admm_rho*g every 10th call, recomputed from the SAME actions.

The arithmetic lives in torch (conv2d, batch_norm, Linear, Normal, Adam), which the reference does not pin; this
oracle uses the torch installed here (2.11) on CPU like the reference.

Parity status: pinned against the reference by tests/golden/calib_*.npz.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

EPS = 1e-6          # calib_sac.py:19
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))
BN_KEYS = ("running_mean", "running_var", "num_batches_tracked")


def conv_out(size, k=5, s=2):
    return (size - (k - 1) - 1) // s + 1


def _init_layer(layer, sc=None):
    sc = sc or 1.0 / np.sqrt(layer.weight.data.size()[0])
    nn.init.uniform_(layer.weight.data, -sc, sc)
    nn.init.uniform_(layer.bias.data, -sc, sc)


def _state(mods):
    """state_dict of an ordered {name: module} (reference key order: module order, params then buffers per module)."""
    sd = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


def init_critic(input_dims, n_actions, M):
    """Draw-for-draw equivalent of constructing the reference CriticNetwork (:91-131)."""
    w, h = input_dims[1], input_dims[2]
    cw, ch = conv_out(conv_out(conv_out(w))), conv_out(conv_out(conv_out(h)))
    mods = OrderedDict()
    mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
    mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["fc1"] = nn.Linear(n_actions + (5 + 2) * (M + 1), 128)
    mods["fc2"] = nn.Linear(128, 16)
    mods["head"] = nn.Linear(cw * ch * 32 + 16, 1)
    for k in ("conv1", "conv2", "conv3", "fc1", "fc2"):
        _init_layer(mods[k])
    _init_layer(mods["head"], 0.003)
    return _state(mods)


def init_actor(input_dims, n_actions, M):
    """Draw-for-draw equivalent of constructing the reference ActorNetwork (:154-196)."""
    w, h = input_dims[1], input_dims[2]
    cw, ch = conv_out(conv_out(conv_out(w))), conv_out(conv_out(conv_out(h)))
    mods = OrderedDict()
    mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
    mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["fc11"] = nn.Linear((5 + 2) * (M + 1), 128)
    mods["fc12"] = nn.Linear(128, 16)
    mods["fc21"] = nn.Linear(cw * ch * 32 + 16, 128)
    mods["fc22mu"] = nn.Linear(128, n_actions)
    mods["fc22sigma"] = nn.Linear(128, n_actions)
    for k in ("conv1", "conv2", "conv3", "fc11", "fc12", "fc21"):
        _init_layer(mods[k])
    _init_layer(mods["fc22mu"], 0.003)
    _init_layer(mods["fc22sigma"], 0.003)
    return _state(mods)


def split(sd):
    """(trainable parameters as leaf tensors, BatchNorm buffers)"""
    p = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in sd.items() if k.rsplit(".", 1)[1] not in BN_KEYS)
    b = OrderedDict((k, v.clone()) for k, v in sd.items() if k.rsplit(".", 1)[1] in BN_KEYS)
    return p, b


def trunk(p, b, x, act, train):
    for l in (1, 2, 3):
        x = F.conv2d(x, p[f"conv{l}.weight"], p[f"conv{l}.bias"], stride=2)
        if train:
            b[f"bn{l}.num_batches_tracked"] += 1
        x = F.batch_norm(x, b[f"bn{l}.running_mean"], b[f"bn{l}.running_var"], p[f"bn{l}.weight"], p[f"bn{l}.bias"],
                         training=train, momentum=0.1, eps=1e-5)
        x = act(x)
    return torch.flatten(x, start_dim=1)


def critic_forward(p, b, img, action, sky, train=True):
    """ref CriticNetwork.forward :133-144.  Its signature is forward(x=img, y, z) with "y is the action, z the sky" in the
    comment, but EVERY call site passes (img, sky, action) (:319-320,:325-326,:342-343), so fc1 sees cat((sky, action))."""
    x = trunk(p, b, img, F.relu, train)
    y = torch.cat((torch.flatten(sky, start_dim=1), torch.flatten(action, start_dim=1)), 1)
    y = F.relu(F.linear(y, p["fc1.weight"], p["fc1.bias"]))
    y = F.relu(F.linear(y, p["fc2.weight"], p["fc2.bias"]))
    return F.linear(torch.cat((x, y), 1), p["head.weight"], p["head.bias"])


def actor_forward(p, b, img, sky, train=True):
    """ref ActorNetwork.forward :198-210"""
    x = trunk(p, b, img, F.elu, train)
    z = F.relu(F.linear(torch.flatten(sky, start_dim=1), p["fc11.weight"], p["fc11.bias"]))
    z = F.relu(F.linear(z, p["fc12.weight"], p["fc12.bias"]))
    x = F.elu(F.linear(torch.cat((x, z), 1), p["fc21.weight"], p["fc21.bias"]))
    mu = F.linear(x, p["fc22mu.weight"], p["fc22mu.bias"])
    sigma = torch.clamp(F.linear(x, p["fc22sigma.weight"], p["fc22sigma.bias"]), min=EPS, max=1)
    return mu, sigma


def sample_normal(p, b, img, sky, eps, max_action=1, train=True, keep=None):
    """ref sample_normal :212-231 with the N(0,1) draw injected"""
    mu, sigma = actor_forward(p, b, img, sky, train)
    if keep is not None:
        keep["sigma"] = sigma.detach().clone()
    u = mu + eps * sigma
    action = torch.tanh(u) * torch.tensor(max_action)
    logp = -((u - mu) ** 2) / (2 * sigma ** 2) - torch.log(sigma) - LOG_SQRT_2PI
    logp = logp - torch.log(1 - action.pow(2) + EPS)
    return action, logp.sum(1, keepdim=True)


def kld_loss(action, hint):
    """ref learn()::kld_loss :345-351"""
    a = torch.clamp(0.5 * action + 0.5 + EPS, min=EPS, max=1.0)
    h = torch.clamp(0.5 * hint + 0.5 + EPS, min=EPS, max=1.0)
    return h * (torch.log(h) - torch.log(a))


class CalibReplay:
    def __init__(self, cap, input_dims, M, A):
        self.mem_size, self.mem_cntr = cap, 0
        self.img = np.zeros((cap, *input_dims), np.float32)
        self.sky = np.zeros((cap, M + 1, 7), np.float32)
        self.img2 = np.zeros((cap, *input_dims), np.float32)
        self.sky2 = np.zeros((cap, M + 1, 7), np.float32)
        self.action = np.zeros((cap, A), np.float32)
        self.hint = np.zeros((cap, A), np.float32)
        self.reward = np.zeros(cap, np.float32)
        self.done = np.zeros(cap, bool)

    def store(self, img, sky, action, reward, img2, sky2, done, hint):
        i = self.mem_cntr % self.mem_size
        self.img[i], self.sky[i], self.img2[i], self.sky2[i] = img, sky, img2, sky2
        self.action[i], self.hint[i], self.reward[i], self.done[i] = action, hint, reward, done
        self.mem_cntr += 1

    def gather(self, idx):
        return (self.img[idx], self.sky[idx], self.action[idx], self.reward[idx], self.img2[idx], self.sky2[idx],
                self.done[idx], self.hint[idx])


class CalibSacRef:
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001, M=3, reward_scale=2,
                 alpha=0.1, hint_threshold=0.1, admm_rho=1.0, use_hint=False):
        self.gamma, self.tau, self.B, self.A, self.M = gamma, tau, int(batch_size), int(n_actions), int(M)
        self.max_action = 1
        # construction order of the reference Agent (:247-253): actor, critic_1, critic_2, targets 1, 2
        self.actor, self.actor_b = split(init_actor(input_dims, n_actions, M))
        self.critic_1, self.critic_1_b = split(init_critic(input_dims, n_actions, M))
        self.critic_2, self.critic_2_b = split(init_critic(input_dims, n_actions, M))
        self.target_critic_1, self.target_critic_1_b = split(init_critic(input_dims, n_actions, M))
        self.target_critic_2, self.target_critic_2_b = split(init_critic(input_dims, n_actions, M))
        self.opt_a = torch.optim.Adam(list(self.actor.values()), lr=lr_a)
        self.opt_1 = torch.optim.Adam(list(self.critic_1.values()), lr=lr_c)
        self.opt_2 = torch.optim.Adam(list(self.critic_2.values()), lr=lr_c)
        self.alpha = torch.tensor(alpha)
        self.use_hint, self.hint_threshold, self.admm_rho = use_hint, hint_threshold, admm_rho
        self.rho = torch.tensor(0.0)
        self.zero = torch.tensor(0.)
        self.learn_counter = 0
        self.replay = CalibReplay(max_mem_size, input_dims, M, n_actions)
        self._soft(1.0)
        self.last = {}

    def _soft(self, tau):
        with torch.no_grad():
            for onl, tgt in ((self.critic_1, self.target_critic_1), (self.critic_2, self.target_critic_2)):
                for k in tgt:
                    tgt[k].copy_(tau * onl[k] + (1 - tau) * tgt[k])

    def choose_action(self, img, sky, eps):
        """ref choose_action :285-295: actor in eval mode (running statistics), no reparameterisation"""
        with torch.no_grad():
            a, _ = sample_normal(self.actor, self.actor_b, torch.as_tensor(img)[None, ], torch.as_tensor(sky)[None, ],
                                 torch.as_tensor(eps), self.max_action, train=False)
        return a.numpy()[0]

    def _gfun(self, actions, hint):
        return torch.max(self.zero, (kld_loss(actions, hint) - self.hint_threshold).mean()).pow(2)

    def learn_on_batch(self, batch, eps_next, eps_cur):
        img, sky, a, r, img2, sky2, done, hint = (torch.as_tensor(x) for x in batch)
        r, done = r.unsqueeze(1), done.unsqueeze(1)
        eps_next, eps_cur = torch.as_tensor(eps_next), torch.as_tensor(eps_cur)
        out = {}
        with torch.no_grad():
            a2, logp2 = sample_normal(self.actor, self.actor_b, img2, sky2, eps_next, self.max_action)
            m = torch.min(critic_forward(self.target_critic_1, self.target_critic_1_b, img2, a2, sky2),
                          critic_forward(self.target_critic_2, self.target_critic_2_b, img2, a2, sky2)) - self.alpha * logp2
            m[done] = 0.0
            y = r + self.gamma * m
        q1 = critic_forward(self.critic_1, self.critic_1_b, img, a, sky)
        q2 = critic_forward(self.critic_2, self.critic_2_b, img, a, sky)
        l1, l2 = F.mse_loss(q1, y), F.mse_loss(q2, y)
        self.opt_1.zero_grad()
        self.opt_2.zero_grad()
        (l1 + l2).backward()
        out.update(y=y, q1=q1.detach().clone(), q2=q2.detach().clone(), critic_1_loss=l1.item(), critic_2_loss=l2.item())
        out["grad_c1"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_1.items())
        out["grad_c2"] = OrderedDict((k, v.grad.clone()) for k, v in self.critic_2.items())
        self.opt_1.step()
        self.opt_2.step()

        act, logp = sample_normal(self.actor, self.actor_b, img, sky, eps_cur, self.max_action, keep=out)
        qp1 = critic_forward(self.critic_1, self.critic_1_b, img, act, sky)
        qp2 = critic_forward(self.critic_2, self.critic_2_b, img, act, sky)
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
                self.rho = self.rho + self.admm_rho * self._gfun(act, hint)
        self.learn_counter += 1
        self._soft(self.tau)
        self.last = out
        return out

    def learn(self):
        if self.replay.mem_cntr < self.B:
            return None
        idx = np.random.choice(min(self.replay.mem_cntr, self.replay.mem_size), self.B, replace=False)
        e1, e2 = torch.randn(self.B, self.A), torch.randn(self.B, self.A)
        return self.learn_on_batch(self.replay.gather(idx), e1, e2)
