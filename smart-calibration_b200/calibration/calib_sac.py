"""Drop-in for the reference module calibration/calib_sac.py (conv SAC agent for the calibration environment), running
on libsmartcal_b200.so.  Same constructor signature, methods, attribute names, state_dict keys/shapes and checkpoint
files (./a_eval_sac_actor.model, ./q_eval_{1,2}_sac_critic.model, replaymem_sac.model; calib_sac.py:43,130,195,381-399).

Reference behaviour kept (see oracle/calib_ref.py for line references): BatchNorm2d runs in training mode in all eight
forward passes of learn() and updates its running statistics each time; choose_action uses eval mode; the critic's fc1
sees cat((sky, action)) because every call site passes (img, sky, action); the target has no reward scale; the soft
update walks parameters() only; the hint penalty is the KLD form with `rho += admm_rho * g` every 10th call.
Initial weights are drawn on the host with torch, module by module in the reference's construction order, so the global
torch RNG stream is consumed exactly like the reference constructor consumes it.
"""
import os
import pickle
from collections import OrderedDict

import numpy as np
import torch as T
import torch.nn as nn

try:
    from .. import _lib as L
    from ..calib_core import CalibDevice
except ImportError:      # imported as a top-level module (sys.path points at this directory, like the reference)
    import smartcal_b200
    from smartcal_b200 import _lib as L
    from smartcal_b200.calib_core import CalibDevice

mydevice = T.device('cuda') if T.cuda.is_available() else T.device('cpu')
EPS = 1e-6
_BN = ("running_mean", "running_var", "num_batches_tracked")


def init_layer(layer, sc=None):
    """calib_sac.py:21-24"""
    sc = sc or 1. / np.sqrt(layer.weight.data.size()[0])
    T.nn.init.uniform_(layer.weight.data, -sc, sc)
    T.nn.init.uniform_(layer.bias.data, -sc, sc)


def _conv_out(size, kernel_size=5, stride=2):
    return (size - (kernel_size - 1) - 1) // stride + 1


def _initial_state(kind, input_dims, n_actions, M):
    """state_dict of a freshly constructed reference network: same modules created in the same order, same init_layer calls"""
    w, h = input_dims[1], input_dims[2]
    lin = _conv_out(_conv_out(_conv_out(w))) * _conv_out(_conv_out(_conv_out(h))) * 32 + 16
    mods = OrderedDict()
    mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
    mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    if kind == "critic":
        mods["fc1"], mods["fc2"], mods["head"] = nn.Linear(n_actions + 7 * (M + 1), 128), nn.Linear(128, 16), nn.Linear(lin, 1)
        plain, last = ("conv1", "conv2", "conv3", "fc1", "fc2"), ("head",)
    else:
        mods["fc11"], mods["fc12"], mods["fc21"] = nn.Linear(7 * (M + 1), 128), nn.Linear(128, 16), nn.Linear(lin, 128)
        mods["fc22mu"], mods["fc22sigma"] = nn.Linear(128, n_actions), nn.Linear(128, n_actions)
        plain, last = ("conv1", "conv2", "conv3", "fc11", "fc12", "fc21"), ("fc22mu", "fc22sigma")
    for k in plain:
        init_layer(mods[k])
    for k in last:
        init_layer(mods[k], 0.003)
    sd = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


class _NetView:
    """state_dict()/load_state_dict()/parameters()-style access to one network living in the device arena"""

    def __init__(self, dev, net, keys_shapes, checkpoint_file):
        self._dev, self._net, self._shapes, self.checkpoint_file = dev, net, keys_shapes, checkpoint_file
        self.device = dev.device
        self.training = True

    def _t(self, k, slot=L.SLOT_PARAM):
        return self._dev.tensor(self._net, k, slot).reshape(self._shapes[k])

    def state_dict(self):
        return OrderedDict((k, self._t(k).detach().clone()) for k in self._shapes)

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._shapes if k not in sd]
        if strict and missing:
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}")
        for k in self._shapes:
            if k in sd:
                dst = self._t(k)
                dst.copy_(T.as_tensor(sd[k]).to(device=dst.device, dtype=dst.dtype).reshape(dst.shape))

    def named_parameters(self):
        return [(k, self._t(k)) for k in self._shapes if k.rsplit(".", 1)[1] not in _BN]

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def adam_state(self):
        return OrderedDict((k, (self._t(k, L.SLOT_ADAM_M).clone(), self._t(k, L.SLOT_ADAM_V).clone())) for k, _ in self.named_parameters())

    def grads(self):
        return OrderedDict((k, self._t(k, 3).clone()) for k, _ in self.named_parameters())

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def save_checkpoint(self):
        T.save(OrderedDict((k, v.cpu()) for k, v in self.state_dict().items()), self.checkpoint_file)

    def load_checkpoint(self):
        self.load_state_dict(T.load(self.checkpoint_file))

    def load_checkpoint_for_eval(self):
        self.load_state_dict(T.load(self.checkpoint_file, map_location=T.device('cpu')))


class ReplayBuffer(object):
    """calib_sac.py:26-88; the eight arrays live in HBM and are read back on demand"""

    def __init__(self, max_size, input_shape, M, n_actions, _dev=None):
        if _dev is None:
            raise RuntimeError("ReplayBuffer is owned by an Agent in the B200 build")
        self._dev, self.mem_size, self.M = _dev, max_size, M
        self._input_shape = tuple(input_shape)
        self.filename = 'replaymem_sac.model'

    @property
    def mem_cntr(self):
        return self._dev.counters()[0]

    def _arr(self, name, shape=None, dt=None):
        a = self._dev.buffer(name).cpu().numpy()
        a = a.reshape((self.mem_size, *shape)) if shape is not None else a.reshape(self.mem_size, -1)
        return a.astype(dt) if dt is not None else a

    state_memory_img = property(lambda s: s._arr("state_memory_img", s._input_shape))
    new_state_memory_img = property(lambda s: s._arr("new_state_memory_img", s._input_shape))
    state_memory_sky = property(lambda s: s._arr("state_memory_sky", (s.M + 1, 7)))
    new_state_memory_sky = property(lambda s: s._arr("new_state_memory_sky", (s.M + 1, 7)))
    action_memory = property(lambda s: s._arr("action_memory"))
    hint_memory = property(lambda s: s._arr("hint_memory"))
    reward_memory = property(lambda s: s._arr("reward_memory").reshape(-1))
    terminal_memory = property(lambda s: s._arr("terminal_memory").reshape(-1) != 0)

    def store_transition(self, state, action, reward, state_, done, hint):
        r = float(reward.item()) if hasattr(reward, "item") else float(reward)
        self._dev.store(state['img'], state['sky'], action, r, state_['img'], state_['sky'], done, hint)

    def sample_buffer(self, batch_size):
        max_mem = min(self.mem_cntr, self.mem_size)
        batch = np.random.choice(max_mem, batch_size, replace=False)
        states = {'img': self.state_memory_img[batch], 'sky': self.state_memory_sky[batch]}
        states_ = {'img': self.new_state_memory_img[batch], 'sky': self.new_state_memory_sky[batch]}
        return states, self.action_memory[batch], self.reward_memory[batch], states_, self.terminal_memory[batch], self.hint_memory[batch]

    def save_checkpoint(self):
        snap = {n: getattr(self, n) for n in ("state_memory_img", "state_memory_sky", "new_state_memory_img", "new_state_memory_sky",
                                              "action_memory", "hint_memory", "reward_memory", "terminal_memory")}
        snap.update(mem_size=self.mem_size, mem_cntr=self.mem_cntr)
        with open(self.filename, 'wb') as f:
            pickle.dump(snap, f)

    def load_checkpoint(self):
        with open(self.filename, 'rb') as f:
            snap = pickle.load(f)
        get = (lambda k: snap[k]) if isinstance(snap, dict) else (lambda k: getattr(snap, k))
        for n in ("state_memory_img", "state_memory_sky", "new_state_memory_img", "new_state_memory_sky", "action_memory", "hint_memory",
                  "reward_memory", "terminal_memory"):
            src = np.asarray(get(n), dtype=np.float32)
            dst = self._dev.buffer(n)
            dst.copy_(T.as_tensor(src.reshape(dst.shape[0], -1)[:, :dst.shape[1]]).to(dst.device))
        self._dev.set_counters(get("mem_cntr"), self._dev.counters()[1])


def _shapes(kind, input_dims, n_actions, M):
    return OrderedDict((k, tuple(v.shape)) for k, v in _initial_state(kind, input_dims, n_actions, M).items())


class CriticNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, name, M, _dev=None, _net=L.NET_CRITIC_1, _shapes_=None):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_sac_critic.model'))


class ActorNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, max_action, name, M, _dev=None, _net=L.NET_ACTOR, _shapes_=None):
        if _dev is None:
            raise RuntimeError("ActorNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions, self.max_action, self.reparam_noise = input_dims, n_actions, max_action, EPS
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_sac_actor.model'))


class Agent():
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, M=3, reward_scale=2, alpha=0.1, hint_threshold=0.1, admm_rho=1.0, name_prefix='',
                 use_hint=False, rng=None, device=None):
        assert 2 * M >= n_actions
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.max_action = 1
        self.scale = reward_scale                      # unused by learn(), like the reference (calib_sac.py:258)
        self.use_hint, self.hint_threshold, self.admm_rho = use_hint, hint_threshold, admm_rho
        self.rng = rng or os.environ.get("SMARTCAL_RNG", "reference")
        self._dev = dev = CalibDevice(batch_size, max_mem_size, n_actions, M, gamma, tau, lr_a, lr_c, alpha, use_hint=use_hint,
                                      hint_threshold=hint_threshold, admm_rho=admm_rho, img=int(input_dims[1]), device=device,
                                      rng_seed=int(T.initial_seed()) & 0xffffffffffff)
        self.replaymem = ReplayBuffer(max_mem_size, input_dims, M, n_actions, _dev=dev)
        # construction order of the reference (calib_sac.py:247-253): actor, critic_1, critic_2, target critics 1, 2
        sd_a = _initial_state("actor", input_dims, n_actions, M)
        sh_a = OrderedDict((k, tuple(v.shape)) for k, v in sd_a.items())
        self.actor = ActorNetwork(lr_a, input_dims, n_actions, self.max_action, 'a_eval', M, _dev=dev, _net=L.NET_ACTOR, _shapes_=sh_a)
        self.actor.load_state_dict(sd_a)
        crit = []
        for name, net in (('q_eval_1', L.NET_CRITIC_1), ('q_eval_2', L.NET_CRITIC_2), ('q_target_1', L.NET_TARGET_CRITIC_1),
                          ('q_target_2', L.NET_TARGET_CRITIC_2)):
            sd = _initial_state("critic", input_dims, n_actions, M)
            sh = OrderedDict((k, tuple(v.shape)) for k, v in sd.items())
            c = CriticNetwork(lr_c, input_dims, n_actions, name, M, _dev=dev, _net=net, _shapes_=sh)
            c.load_state_dict(sd)
            crit.append(c)
        self.critic_1, self.critic_2, self.target_critic_1, self.target_critic_2 = crit
        self.alpha = T.tensor(alpha, requires_grad=False, device=mydevice)
        self.learn_counter = 0
        self.update_network_parameters(self.target_critic_1, self.critic_1, tau=1.)
        self.update_network_parameters(self.target_critic_2, self.critic_2, tau=1.)
        self.last_stats = None

    @property
    def rho(self):
        return T.tensor(self.last_stats["rho"] if self.last_stats else 0.0)

    def update_network_parameters(self, target_model=None, origin_model=None, tau=None):
        """calib_sac.py:271-276 (parameters only).  Both target critics move together on the device."""
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal, hint):
        self.replaymem.store_transition(state, action, reward, state_, terminal, hint)

    remember = store_transition

    def choose_action(self, observation):
        """calib_sac.py:285-295: eval-mode actor, Normal.sample (one (1,A) randn draw from the torch stream)"""
        eps = T.randn(1, self.n_actions).numpy() if self.rng == 'reference' else None
        return self._dev.act(np.asarray(observation['img'], np.float32), np.asarray(observation['sky'], np.float32), eps)

    def learn(self, want_stats=False):
        dev = self._dev
        cnt = dev.counters()[0]
        if cnt < self.batch_size:
            return
        dev.set_batch(np.random.choice(min(cnt, dev.C), self.batch_size, replace=False))
        if self.rng == 'reference':
            e1 = T.randn(self.batch_size, self.n_actions).numpy()
            e2 = T.randn(self.batch_size, self.n_actions).numpy()
            st = dev.learn(e1, e2, want_stats=want_stats or self.use_hint)
        else:
            st = dev.learn(want_stats=want_stats or self.use_hint)
        self.learn_counter += 1
        if st is not None:
            self.last_stats = st
            if self.use_hint:
                print(f'AC {st["actor_loss"]} {st["gfun"]}')          # calib_sac.py:358
        return st if want_stats else None

    def save_models(self):
        self.actor.save_checkpoint()
        self.critic_1.save_checkpoint()
        self.critic_2.save_checkpoint()
        self.replaymem.save_checkpoint()

    def load_models(self):
        self.actor.load_checkpoint()
        self.critic_1.load_checkpoint()
        self.critic_2.load_checkpoint()
        self.replaymem.load_checkpoint()
        self.update_network_parameters(tau=1.)

    def load_models_for_eval(self):
        self.actor.load_checkpoint_for_eval()
        self.critic_1.load_checkpoint_for_eval()
        self.critic_2.load_checkpoint_for_eval()

    def print(self):
        print(self.actor)
        print(self.critic_1)
