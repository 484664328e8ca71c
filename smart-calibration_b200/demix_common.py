"""Shared body of the two demixing SAC drop-ins (demixing_rl/demix_sac.py and demixing_fuzzy/demix_sac.py of the reference are
the same agent up to `use_influence` and INNER_DIM, see `diff` of the two files): same constructor signatures, methods,
attribute names, state_dict keys / shapes and checkpoint files, running on the sc_calib_* group of libsmartcal_b200.so with
the network-family flags of include/smartcal_b200.h (SC_CALIB_LOGSIGMA | SC_CALIB_DEMIX_HINT | SC_CALIB_FLAT_META, plus
SC_CALIB_NO_TRUNK | SC_CALIB_SIDE32 for the MLP-only agent).

Reference behaviour kept (line numbers: demixing_rl/demix_sac.py; the fuzzy file is the same code shifted):
  * networks :372-527 -- metadata network fc(1|11) -> ReLU -> fc(2|12) -> ReLU, critic fc1 sees cat((metadata, action)),
    policy heads fc22mu / fc22logsigma with sigma = exp(clamp(., -20, 2)), tanh-squashed sample, log-prob with the
    log(1 - a^2 + 1e-6) correction; BatchNorm2d of the conv trunk in training mode inside learn(), eval mode in choose_action;
  * learn :594-682 -- no reward scale, MSE critic losses, soft update over parameters() only, KLD hint penalty WITHOUT clamping,
    and every 10th call a fresh no-grad sample from the UPDATED actor for `rho += admm_rho * g`;
  * ReplayBuffer :310-370 -- observation dicts {'infmap', 'metadata'}.
"""
import os
import pickle
from collections import OrderedDict

import numpy as np
import torch as T
import torch.nn as nn

from . import _lib as L
from .calib_core import CalibDevice
from .calibration.calib_sac import _NetView, _conv_out, init_layer

EPS = 1e-6


def initial_state(kind, input_dims, n_actions, M, inner, use_influence):
    """state_dict of a freshly constructed reference network: same modules in the same order, same init_layer calls"""
    mods = OrderedDict()
    if use_influence:
        mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
        mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
        w, h = input_dims[1], input_dims[2]
        lin = _conv_out(_conv_out(_conv_out(w))) * _conv_out(_conv_out(_conv_out(h))) * 32 + inner
        convs = ("conv1", "conv2", "conv3")
    else:
        lin, convs = inner, ()
    if kind == "critic":
        mods["fc1"], mods["fc2"], mods["head"] = nn.Linear(M + n_actions, 128), nn.Linear(128, inner), nn.Linear(lin, 1)
        plain, last = convs + ("fc1", "fc2"), ("head",)
    else:
        mods["fc11"], mods["fc12"], mods["fc21"] = nn.Linear(M, 128), nn.Linear(128, inner), nn.Linear(lin, 128)
        mods["fc22mu"], mods["fc22logsigma"] = nn.Linear(128, n_actions), nn.Linear(128, n_actions)
        plain, last = convs + ("fc11", "fc12", "fc21"), ("fc22mu", "fc22logsigma")
    for k in plain:
        init_layer(mods[k])
    for k in last:
        init_layer(mods[k], 0.003)
    sd = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


class DemixReplayBuffer(object):
    """reference ReplayBuffer (:310-370); the arrays live in HBM and are read back on demand"""

    def __init__(self, max_size, input_shape, n_actions, M, use_influence=True, filename='replaymem_sac.model', _dev=None):
        if _dev is None:
            raise RuntimeError("ReplayBuffer is owned by a DemixingAgent in the B200 build")
        self._dev, self.mem_size, self.M, self.use_influence = _dev, max_size, M, use_influence
        self._input_shape = tuple(input_shape)
        self.filename = filename

    @property
    def mem_cntr(self):
        return self._dev.counters()[0]

    def _arr(self, name, shape=None):
        a = self._dev.buffer(name).cpu().numpy()
        return a.reshape((self.mem_size, *shape)) if shape is not None else a.reshape(self.mem_size, -1)

    state_memory_img = property(lambda s: s._arr("state_memory_img", s._input_shape))
    new_state_memory_img = property(lambda s: s._arr("new_state_memory_img", s._input_shape))
    state_memory_sky = property(lambda s: s._arr("state_memory_sky"))
    new_state_memory_sky = property(lambda s: s._arr("new_state_memory_sky"))
    action_memory = property(lambda s: s._arr("action_memory"))
    hint_memory = property(lambda s: s._arr("hint_memory"))
    reward_memory = property(lambda s: s._arr("reward_memory").reshape(-1))
    terminal_memory = property(lambda s: s._arr("terminal_memory").reshape(-1) != 0)

    def store_transition(self, state, action, reward, state_, done, hint):
        r = float(reward.item()) if hasattr(reward, "item") else float(reward)
        img = state['infmap'] if self.use_influence else None
        img2 = state_['infmap'] if self.use_influence else None
        self._dev.store(img, state['metadata'], action, r, img2, state_['metadata'], done, hint)

    def sample_buffer(self, batch_size):
        batch = np.random.choice(min(self.mem_cntr, self.mem_size), batch_size, replace=False)
        states = {'infmap': self.state_memory_img[batch] if self.use_influence else 0, 'metadata': self.state_memory_sky[batch]}
        states_ = {'infmap': self.new_state_memory_img[batch] if self.use_influence else 0, 'metadata': self.new_state_memory_sky[batch]}
        return states, self.action_memory[batch], self.reward_memory[batch], states_, self.terminal_memory[batch], self.hint_memory[batch]

    def _names(self):
        n = ["state_memory_sky", "new_state_memory_sky", "action_memory", "hint_memory", "reward_memory", "terminal_memory"]
        return (["state_memory_img", "new_state_memory_img"] if self.use_influence else []) + n

    def save_checkpoint(self):
        snap = {n: getattr(self, n) for n in self._names()}
        snap.update(mem_size=self.mem_size, mem_cntr=self.mem_cntr)
        with open(self.filename, 'wb') as f:
            pickle.dump(snap, f)

    def load_checkpoint(self):
        with open(self.filename, 'rb') as f:
            snap = pickle.load(f)
        get = (lambda k: snap[k]) if isinstance(snap, dict) else (lambda k: snap.__dict__[k])
        for n in self._names():
            src = np.asarray(get(n), dtype=np.float32)
            dst = self._dev.buffer(n)
            dst.copy_(T.as_tensor(src.reshape(dst.shape[0], -1)[:, :dst.shape[1]]).to(dst.device))
        self._dev.set_counters(get("mem_cntr"), self._dev.counters()[1])


class DemixCriticNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, name, M, use_influence=True, _dev=None, _net=L.NET_CRITIC_1, _shapes_=None):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by a DemixingAgent in the B200 build")
        self.input_dims, self.n_actions, self.use_influence = input_dims, n_actions, use_influence
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_sac_critic.model'))


class DemixActorNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, max_action, name, M, use_influence=True, _dev=None, _net=L.NET_ACTOR, _shapes_=None):
        if _dev is None:
            raise RuntimeError("ActorNetwork is owned by a DemixingAgent in the B200 build")
        self.input_dims, self.n_actions, self.max_action, self.reparam_noise = input_dims, n_actions, max_action, EPS
        self.use_influence = use_influence
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_sac_actor.model'))


class DemixingAgentBase():
    """reference DemixingAgent (:529-719).  `inner` = INNER_DIM of the metadata network (16 demixing_rl, 32 demixing_fuzzy)."""

    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size, tau, M, reward_scale, alpha, hint_threshold,
                 admm_rho, name_prefix, use_hint, use_influence, inner, rng=None, device=None):
        self.use_influence = use_influence
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.max_action = 1
        self.scale = reward_scale                      # unused by learn(), like the reference (:606)
        self.use_hint, self.hint_threshold, self.admm_rho = use_hint, hint_threshold, admm_rho
        self.rng = rng or os.environ.get("SMARTCAL_RNG", "reference")
        variant = L.SC_CALIB_LOGSIGMA | L.SC_CALIB_DEMIX_HINT | L.SC_CALIB_FLAT_META
        if not use_influence:
            variant |= L.SC_CALIB_NO_TRUNK
        if inner == 32:
            variant |= L.SC_CALIB_SIDE32
        elif inner != 16:
            raise ValueError("INNER_DIM must be 16 or 32")
        self._dev = dev = CalibDevice(batch_size, max_mem_size, n_actions, M, gamma, tau, lr_a, lr_c, alpha, use_hint=use_hint,
                                      hint_threshold=hint_threshold, admm_rho=admm_rho, img=int(input_dims[1]), device=device,
                                      rng_seed=int(T.initial_seed()) & 0xffffffffffff, variant=variant)
        self.replaymem = DemixReplayBuffer(max_mem_size, input_dims, n_actions, M, use_influence=use_influence, _dev=dev)
        # construction order of the reference (:171-177): actor, critic_1, critic_2, target critics 1, 2
        sd_a = initial_state("actor", input_dims, n_actions, M, inner, use_influence)
        sh_a = OrderedDict((k, tuple(v.shape)) for k, v in sd_a.items())
        self.actor = DemixActorNetwork(lr_a, input_dims, n_actions, self.max_action, name_prefix + 'a_eval', M, use_influence, _dev=dev,
                                       _net=L.NET_ACTOR, _shapes_=sh_a)
        self.actor.load_state_dict(sd_a)
        crit = []
        for name, net in (('q_eval_1', L.NET_CRITIC_1), ('q_eval_2', L.NET_CRITIC_2), ('q_target_1', L.NET_TARGET_CRITIC_1),
                          ('q_target_2', L.NET_TARGET_CRITIC_2)):
            sd = initial_state("critic", input_dims, n_actions, M, inner, use_influence)
            sh = OrderedDict((k, tuple(v.shape)) for k, v in sd.items())
            c = DemixCriticNetwork(lr_c, input_dims, n_actions, name_prefix + name, M, use_influence, _dev=dev, _net=net, _shapes_=sh)
            c.load_state_dict(sd)
            crit.append(c)
        self.critic_1, self.critic_2, self.target_critic_1, self.target_critic_2 = crit
        self.alpha = T.tensor(alpha, requires_grad=False, device=dev.device)
        self.learn_alpha = False
        self.learn_counter = 0
        self.update_network_parameters(self.target_critic_1, self.critic_1, tau=1.)
        self.last_stats = None

    @property
    def rho(self):
        return T.tensor(self.last_stats["rho"] if self.last_stats else 0.0)

    def update_network_parameters(self, target_model=None, origin_model=None, tau=None):
        """:573-578 (parameters only).  Both target critics move together on the device."""
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal, hint):
        self.replaymem.store_transition(state, action, reward, state_, terminal, hint)

    def choose_action(self, observation, evaluation_episode=False):
        """:583-592: eval-mode actor, Normal.sample (one (1,A) randn draw from the torch stream)"""
        eps = T.randn(1, self.n_actions).numpy() if self.rng == 'reference' else None
        img = np.asarray(observation['infmap'], np.float32) if self.use_influence else None
        return self._dev.act(img, np.asarray(observation['metadata'], np.float32), eps)

    def learn(self, want_stats=False):
        dev = self._dev
        cnt = dev.counters()[0]
        if cnt < self.batch_size:
            return
        dev.set_batch(np.random.choice(min(cnt, dev.C), self.batch_size, replace=False))
        dual = self.use_hint and self.learn_counter % 10 == 0
        if self.rng == 'reference':
            e1 = T.randn(self.batch_size, self.n_actions).numpy()
            e2 = T.randn(self.batch_size, self.n_actions).numpy()
            e3 = T.randn(self.batch_size, self.n_actions).numpy() if dual else None
            st = dev.learn(e1, e2, eps_dual=e3, want_stats=want_stats or self.use_hint)
        else:
            st = dev.learn(want_stats=want_stats or self.use_hint)
        self.learn_counter += 1
        if st is not None:
            self.last_stats = st
            if self.use_hint:
                print(f'AC {st["actor_loss"]} {st["gfun"]}')          # :657
        return st if want_stats else None

    def save_models(self):
        self.actor.save_checkpoint()
        self.critic_1.save_checkpoint()
        self.critic_2.save_checkpoint()
        self.replaymem.save_checkpoint()

    def load_replaybuffer(self):
        self.replaymem.load_checkpoint()
        self.update_network_parameters(tau=1.)

    def load_models(self):
        self.actor.load_checkpoint()
        self.critic_1.load_checkpoint()
        self.critic_2.load_checkpoint()
        self.replaymem.load_checkpoint()
        self.update_network_parameters(tau=1.)

    def load_models_for_eval(self):
        self.actor.load_checkpoint()
        self.critic_1.load_checkpoint()
        self.critic_2.load_checkpoint()

    def print(self):
        print(self.actor)
        print(self.critic_1)
