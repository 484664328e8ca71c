"""Drop-in for the reference module demixing_rl/demix_td3.py (conv TD3 agent of the demixing environment, :355-640, with its
prioritized replay :23-167 and networks :231-352): `from demix_td3 import DemixingAgent` (demixing_rl/main_td3.py:2) keeps working
with this directory on sys.path.  The agent runs on the sc_calib_* group of libsmartcal_b200.so with SC_CALIB_TD3 |
SC_CALIB_FLAT_META (include/smartcal_b200.h): same constructor signature, methods, attribute names, state_dict keys / shapes and
checkpoint files (./a_eval_td3_actor.model, ./a_target_td3_actor.model, ./q_{eval,target}_{1,2}_td3_critic.model,
prioritized_replaymem_td3.model).

Reference behaviour kept (oracle/demix_td3_ref.py lists the line references; it is bit-identical to the unmodified module,
tests/golden/make_golden_demix_td3.py):
  * `prioritized` is True (:374); the sum tree lives in HBM as the reference's flat float64 heap; new transitions enter with the
    largest leaf priority (100 while the tree is empty); learn() draws B stratified samples, refreshes the priorities from
    0.5 (|q1 - y| + |q2 - y|) BEFORE the critic step, and weights both losses by mean(is_weights) (the (B,) x (B,1) broadcast of PER.mse);
  * target actor / target critics run with their BatchNorm layers in eval mode, the online critics never leave train mode (the
    priority pre-pass and every critic_1 forward of the actor update move their running statistics);
  * the actor and all three targets move every `update_actor_interval`-th call; update_network_parameters blends every state_dict
    entry, BatchNorm buffers included; with use_hint the actor takes five Adam steps inside the adaptive-ADMM loop;
  * choose_action: `warmup` calls of pure noise (float64, shape (n_actions,)), then eval-mode actor + noise, clamped (float32, shape
    (1, n_actions)) -- both branches draw from numpy's global stream exactly like the reference.
Not provided: the stand-alone ReplayBuffer class (unreachable in the reference, `prioritized` is hard-wired) and host-side
PER.sample_buffer / batch_update (sampling and priority updates happen inside learn() on the device tree).
"""
import os
import pickle
from collections import OrderedDict

import numpy as np
import torch as T
import torch.nn as nn

import smartcal_b200  # noqa: F401
from smartcal_b200 import _lib as L
from smartcal_b200.calib_core import CalibDevice
from smartcal_b200.calibration.calib_sac import _NetView, _conv_out, init_layer  # noqa: F401
from smartcal_b200.demix_common import initial_state as _initial_state_sac


def _initial_state(kind, input_dims, n_actions, M):
    """state_dict of a freshly constructed reference network (:232-271 / :292-331): same modules in the same order, same init_layer
    calls, so the global torch stream is consumed exactly like the reference constructor consumes it"""
    if kind == "critic":
        return _initial_state_sac("critic", input_dims, n_actions, M, 16, True)       # identical module list to demix_sac.py's critic
    mods = OrderedDict()
    mods["conv1"], mods["bn1"] = nn.Conv2d(1, 16, kernel_size=5, stride=2), nn.BatchNorm2d(16)
    mods["conv2"], mods["bn2"] = nn.Conv2d(16, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    mods["conv3"], mods["bn3"] = nn.Conv2d(32, 32, kernel_size=5, stride=2), nn.BatchNorm2d(32)
    w, h = input_dims[1], input_dims[2]
    lin = _conv_out(_conv_out(_conv_out(w))) * _conv_out(_conv_out(_conv_out(h))) * 32 + 16
    mods["fc11"], mods["fc12"], mods["fc21"], mods["fc22"] = nn.Linear(M, 128), nn.Linear(128, 16), nn.Linear(lin, 128), nn.Linear(128, n_actions)
    for k in ("conv1", "conv2", "conv3", "fc11", "fc12", "fc21"):
        init_layer(mods[k])
    init_layer(mods["fc22"], 0.003)
    sd = OrderedDict()
    for name, m in mods.items():
        for k, v in m.state_dict().items():
            sd[f"{name}.{k}"] = v.detach().clone()
    return sd


class _TreeView(object):
    """reference SumTree surface (demix_sac.py:37-149) over the HBM heap: `tree` (float64 array, root 0, leaves at [C-1, 2C-1)),
    `capacity`, `data_pointer`, `total_priority`, len()"""

    def __init__(self, per):
        self._per = per
        self.capacity = per.mem_size

    @property
    def tree(self):
        return self._per._dev.buffer("tree").cpu().numpy().reshape(-1)

    @property
    def data_pointer(self):
        return self._per.mem_cntr % self.capacity

    @property
    def data_length(self):
        return len(self)

    @property
    def total_priority(self):
        return float(self._per._dev.buffer("tree")[0, 0].item())

    def __len__(self):
        return min(self._per.mem_cntr, self.capacity)


class PER(object):
    """reference PER (:23-167); arrays and tree live in HBM"""
    epsilon = 0.01
    alpha = 0.6
    beta_increment_per_sampling = 1e-4
    absolute_error_upper = 100.

    def __init__(self, capacity, input_shape, n_actions, M, _dev=None):
        if _dev is None:
            raise RuntimeError("PER is owned by a DemixingAgent in the B200 build")
        self._dev, self.mem_size, self.M = _dev, capacity, M
        self._input_shape = tuple(input_shape)
        self.tree = _TreeView(self)
        self.filename = 'prioritized_replaymem_td3.model'

    @property
    def mem_cntr(self):
        return self._dev.counters()[0]

    @property
    def beta(self):
        return self._dev.per_beta()

    def __len__(self):
        return len(self.tree)

    def is_full(self):
        return len(self.tree) >= self.tree.capacity

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

    def store_transition(self, state, action, reward, state_, done, hint, error=None):
        """:60-78 with error=None (the only call site, :436): priority = largest leaf, or absolute_error_upper for an empty tree"""
        if error is not None:
            raise NotImplementedError("explicit store-time errors are not used by demix_td3.DemixingAgent (:436)")
        r = float(reward.item()) if hasattr(reward, "item") else float(reward)
        self._dev.store(state['infmap'], state['metadata'], action, r, state_['infmap'], state_['metadata'], done, hint)

    def sample_buffer(self, batch_size):
        raise RuntimeError("the B200 build samples inside DemixingAgent.learn() (stratified draw on the device tree)")

    def batch_update(self, idxs, errors):
        raise RuntimeError("the B200 build refreshes priorities inside DemixingAgent.learn()")

    def mse(self, expected, targets, is_weights):
        """:137-140 (kept for callers that evaluate it on their own tensors)"""
        td_error = expected - targets
        weighted_squared_error = is_weights * td_error * td_error
        return T.sum(weighted_squared_error) / T.numel(weighted_squared_error)

    _NAMES = ("state_memory_img", "new_state_memory_img", "state_memory_sky", "new_state_memory_sky", "action_memory", "hint_memory",
              "reward_memory", "terminal_memory")

    def save_checkpoint(self):
        snap = {n: getattr(self, n) for n in self._NAMES}
        snap.update(mem_size=self.mem_size, mem_cntr=self.mem_cntr, tree=self.tree.tree, beta=self.beta)
        with open(self.filename, 'wb') as f:
            pickle.dump(snap, f)

    @staticmethod
    def read_checkpoint(filename):
        """host side of load_checkpoint: reads a file written by save_checkpoint above (a dict) OR by the reference's
        PER.save_checkpoint (:142-145: the pickled PER object, whose `tree` is a demix_sac.SumTree carrying the raw float64 heap;
        `beta` / `mem_cntr` are instance attributes only once they were written).  Only __dict__ entries are looked at, never the
        device-backed properties of the classes the pickle resolves to."""
        with open(filename, 'rb') as f:
            snap = pickle.load(f)
        d = snap if isinstance(snap, dict) else snap.__dict__
        out = {n: np.asarray(d[n], dtype=np.float32) for n in PER._NAMES}
        tree = d["tree"]
        tree = tree if isinstance(tree, np.ndarray) else tree.__dict__["tree"]
        out["tree"] = np.asarray(tree, np.float64).reshape(-1)
        out["mem_cntr"] = int(d.get("mem_cntr", 0))
        out["beta"] = float(d.get("beta", 0.4))
        return out

    def load_checkpoint(self):
        d = self.read_checkpoint(self.filename)
        for n in self._NAMES:
            dst = self._dev.buffer(n)
            dst.copy_(T.as_tensor(d[n].reshape(dst.shape[0], -1)[:, :dst.shape[1]]).to(dst.device))
        dst = self._dev.buffer("tree")
        dst.copy_(T.as_tensor(d["tree"].reshape(dst.shape)).to(dst.device))
        self._dev.set_counters(d["mem_cntr"], self._dev.counters()[1])
        self._dev.set_per_beta(d["beta"])

    def normalize_reward(self):
        """:163-167"""
        n = min(self.mem_cntr, self.mem_size)
        if n == 0:
            return
        buf = self._dev.buffer("reward_memory")
        r = buf[:n, 0]
        mu, sigma = r.mean(), r.std(unbiased=False)
        buf[:n, 0] = (r - mu) / sigma


class CriticNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, name, M, _dev=None, _net=L.NET_CRITIC_1, _shapes_=None):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by a DemixingAgent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_td3_critic.model'))


class ActorNetwork(_NetView):
    def __init__(self, lr, input_dims, n_actions, name, M, _dev=None, _net=L.NET_ACTOR, _shapes_=None):
        if _dev is None:
            raise RuntimeError("ActorNetwork is owned by a DemixingAgent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, _shapes_, os.path.join('./', name + '_td3_actor.model'))


class DemixingAgent():
    """reference DemixingAgent (:355-640)"""

    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, M=30, update_actor_interval=2, warmup=1000, noise=0.1, use_hint=False, rng=None, device=None):
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.max_action, self.min_action = 1, -1
        self.learn_step_cntr, self.time_step = 0, 0
        self.warmup, self.update_actor_interval = warmup, update_actor_interval
        self.prioritized = True
        self.noise, self.use_hint = noise, use_hint
        self._admm_rho = 0.1                            # :400
        self.Nadmm, self.adaptive_admm, self.corr_min = 5, True, 0.5
        self.rng = rng or os.environ.get("SMARTCAL_RNG", "reference")
        self._dev = dev = CalibDevice(batch_size, max_mem_size, n_actions, M, gamma, tau, lr_a, lr_c, 0.0, use_hint=use_hint,
                                      hint_threshold=0.0, admm_rho=self._admm_rho, img=int(input_dims[1]), device=device,
                                      rng_seed=int(T.initial_seed()) & 0xffffffffffff, variant=L.SC_CALIB_TD3 | L.SC_CALIB_FLAT_META,
                                      update_actor_interval=update_actor_interval)
        self.replaymem = PER(max_mem_size, input_dims, n_actions, M, _dev=dev)
        # construction order of the reference (:380-389): actor, critic_1, critic_2, target actor, target critics 1, 2
        nets = []
        for kind, name, net in (("actor", 'a_eval', L.NET_ACTOR), ("critic", 'q_eval_1', L.NET_CRITIC_1), ("critic", 'q_eval_2', L.NET_CRITIC_2),
                                ("actor", 'a_target', L.NET_TARGET_ACTOR), ("critic", 'q_target_1', L.NET_TARGET_CRITIC_1),
                                ("critic", 'q_target_2', L.NET_TARGET_CRITIC_2)):
            sd = _initial_state(kind, input_dims, n_actions, M)
            sh = OrderedDict((k, tuple(v.shape)) for k, v in sd.items())
            cls = ActorNetwork if kind == "actor" else CriticNetwork
            v = cls(lr_a if kind == "actor" else lr_c, input_dims, n_actions, name, M, _dev=dev, _net=net, _shapes_=sh)
            v.load_state_dict(sd)
            nets.append(v)
        self.actor, self.critic_1, self.critic_2, self.target_actor, self.target_critic_1, self.target_critic_2 = nets
        self.update_network_parameters(tau=1.)
        self.last_stats = None

    @property
    def admm_rho(self):
        return self._admm_rho

    @admm_rho.setter
    def admm_rho(self, v):
        self._admm_rho = float(v)
        self._dev.set_scalars(0.0, self._admm_rho)

    def update_network_parameters(self, tau=None):
        """:408-432: every state_dict entry of the three targets, BatchNorm buffers included"""
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal, hint):
        self.replaymem.store_transition(state, action, reward, state_, terminal, hint)

    def choose_action(self, observation):
        """:439-456"""
        if self.time_step < self.warmup:
            mu = np.random.normal(scale=self.noise, size=(self.n_actions,))
            nz = np.random.normal(scale=self.noise, size=(self.n_actions,)).astype(np.float32)
            out = np.clip(mu + nz.astype(np.float64), self.min_action, self.max_action)
        else:
            nz = np.random.normal(scale=self.noise, size=(self.n_actions,)).astype(np.float32)
            out = self._dev.act(np.asarray(observation['infmap'], np.float32), np.asarray(observation['metadata'], np.float32), nz)[None, ]
        self.time_step += 1
        return out

    def learn(self, want_stats=False):
        """:459-603"""
        dev = self._dev
        if dev.counters()[0] < self.batch_size:
            return
        if self.rng == 'reference':
            u01 = np.random.random_sample(self.batch_size)      # np.random.uniform(a, b) = a + (b - a) * random_sample(), B times (:100)
            st = dev.learn_td3(u01, np.random.normal(scale=0.2), want_stats=want_stats)
        else:
            st = dev.learn_td3(want_stats=want_stats)
        self.learn_step_cntr += 1
        if st is not None:
            self.last_stats = st
        return st if want_stats else None

    def _all(self):
        return (self.actor, self.target_actor, self.critic_1, self.critic_2, self.target_critic_1, self.target_critic_2)

    def save_models(self):
        for n in self._all():
            n.save_checkpoint()
        self.replaymem.save_checkpoint()

    def load_models(self):
        for n in self._all():
            n.load_checkpoint()
        self.replaymem.load_checkpoint()
        self.update_network_parameters(tau=1.)

    def load_models_for_eval(self):
        self.actor.load_checkpoint()
        self.critic_1.load_checkpoint()
        self.critic_2.load_checkpoint()

    def print(self):
        print(self.actor)
        print(self.critic_1)
