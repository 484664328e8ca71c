"""Drop-in for the reference module elasticnet/enet_sac.py: same class names, constructor signatures, methods,
attribute names and checkpoint files -- but learn()/choose_action()/replay/PER run as sm_100a CUDA kernels in
libsmartcal_b200.so on an HBM-resident agent (see include/smartcal_b200.h for the C ABI each method binds).

Put this directory on sys.path instead of the reference's `elasticnet/` and `main_sac.py`, `enetenv.py`,
`distributed_per_sac.py`-style callers run unchanged:   from enet_sac import Agent, ReplayBuffer, PER, ActorNetwork

Deviations from the reference, all deliberate and documented in DESIGN.md:
  * `prioritized=True` is honoured (reference flag is dead, enet_sac.py:488-490): PER is wired the way
    enet_td3.py wires it (sample -> IS weights, priorities from 0.5(|q1-y|+|q2-y|) before the critic step,
    PER.mse broadcast quirk).  `prioritized=False` is the reference-identical path.
  * extra keyword `rng=` ('reference' | 'device'), default from $SMARTCAL_RNG or 'reference' (host draws from
    numpy's / torch's global generators in the reference's order).
"""
import copyreg
import pickle

import numpy as np
import torch as T

try:
    from . import _host as H
except ImportError:          # imported as a top-level module (directory placed on sys.path like the reference's)
    import _host as H

L = H.L


def is_power_of_2(n):
    return n != 0 and (n & (n - 1)) == 0


class ReplayBuffer(object):
    """Stand-alone: a host ring like the reference's (what distributed actors fill and ship).  Owned by an Agent
    (`_dev` given): the six arrays live in HBM and the attributes below read them back on demand."""

    def __new__(cls, max_size=None, input_shape=None, n_actions=None, _dev=None):
        if cls is ReplayBuffer and _dev is not None:
            return object.__new__(_DeviceReplayBuffer)
        return object.__new__(cls)

    def __init__(self, max_size, input_shape, n_actions, _dev=None):
        self.mem_size = max_size
        self.mem_cntr = 0
        S = int(input_shape[0])
        for name, shape, dt in (("state_memory", (max_size, S), np.float32), ("new_state_memory", (max_size, S), np.float32),
                                ("action_memory", (max_size, n_actions), np.float32), ("reward_memory", (max_size,), np.float32),
                                ("terminal_memory", (max_size,), bool), ("hint_memory", (max_size, n_actions), np.float32)):
            setattr(self, name, np.zeros(shape, dtype=dt))
        self.filename = 'replaymem_sac.model'

    def store_transition(self, state, action, reward, state_, done, hint):
        i = self.mem_cntr % self.mem_size
        self.state_memory[i] = H.flat_state(state)
        self.new_state_memory[i] = H.flat_state(state_)
        self.action_memory[i], self.reward_memory[i] = action, H._scalar(reward)
        self.terminal_memory[i], self.hint_memory[i] = done, hint
        self.mem_cntr += 1

    def sample_buffer(self, batch_size):
        pick = np.random.choice(min(self.mem_cntr, self.mem_size), batch_size, replace=False)
        return tuple(getattr(self, n)[pick] for n in ("state_memory", "action_memory", "reward_memory",
                                                      "new_state_memory", "terminal_memory", "hint_memory"))

    def save_checkpoint(self):
        with open(self.filename, 'wb') as f:
            pickle.dump(self, f)

    def load_checkpoint(self):
        with open(self.filename, 'rb') as f:
            other = pickle.load(f)
        self.mem_size, self.mem_cntr = other.mem_size, other.mem_cntr
        for n in H._ARRAYS:
            setattr(self, n, getattr(other, n))


class _DeviceReplayBuffer(H.DeviceReplayMixin, ReplayBuffer):
    def __init__(self, max_size, input_shape, n_actions, _dev=None):
        self._dev = _dev
        self.mem_size = max_size
        self.filename = 'replaymem_sac.model'

    def store_transition(self, state, action, reward, state_, done, hint):
        # deferred: fused into the next learn() (sc_agent_store_learn); anything that looks at the replay flushes it first
        self._dev.store(self._dev.pack(H.flat_state(state), action, H._scalar(reward), H.flat_state(state_), bool(done), hint), defer=True)

    def store_transition_from_buffer(self, state, action, reward, state_, done, hint):
        self._dev.store(self._dev.pack(state, action, H._scalar(reward), state_, bool(done), hint), defer=True)

    def sample_buffer(self, batch_size):
        pick = np.random.choice(min(self.mem_cntr, self.mem_size), batch_size, replace=False)
        return tuple(getattr(self, n)[pick] for n in ("state_memory", "action_memory", "reward_memory",
                                                      "new_state_memory", "terminal_memory", "hint_memory"))

    def __reduce__(self):
        # pickles as a plain host ReplayBuffer through the standard protocol (class name + __dict__ only), so the
        # reference's enet_sac.ReplayBuffer.load_checkpoint reads the file as well
        return (copyreg._reconstructor, (ReplayBuffer, object, None), self._host_state())

    def load_checkpoint(self):
        with open(self.filename, 'rb') as f:
            other = pickle.load(f)
        st = H.plain_state(other)
        self._upload_arrays(st)
        m, l, d, b = self._dev.counters()
        self._dev.set_counters(st["mem_cntr"], l, st["mem_cntr"] % self.mem_size, b)


class _HostSumTree(object):
    """what a device SumTree pickles as: class SumTree with the reference's attribute set (tree, capacity, data_pointer,
    data_length) in its __dict__ -- readable by the reference's enet_sac.SumTree and by SumTree below (plain attributes win
    over the device properties when no device is attached)"""

    def __reduce__(self):
        return (copyreg._reconstructor, (SumTree, object, None), dict(self.__dict__))


class SumTree():
    """Read/write window on the device-resident float64 sum tree (reference SumTree, enet_sac.py:82-200)."""

    def __init__(self, capacity, _dev=None):
        self.capacity = int(capacity)
        assert is_power_of_2(self.capacity), "Capacity must be power of 2."
        self._dev = _dev
        self._host = None if _dev is not None else np.zeros(2 * self.capacity - 1)

    def _plain(self, name):
        """attribute of an instance rebuilt by pickle (no device behind it)"""
        d = self.__dict__
        return d.get(name) if d.get("_dev") is None and name in d else None

    @property
    def tree(self):
        if self.__dict__.get("_dev") is None:
            return self.__dict__.get("_host") if self.__dict__.get("_host") is not None else self.__dict__.get("tree")
        return self._dev.buffer("tree")[:, 0].cpu().numpy()

    @tree.setter
    def tree(self, values):
        if self._dev is None:
            self._host = np.asarray(values, np.float64)
        else:
            v = self._dev.buffer("tree")
            v.copy_(T.as_tensor(np.asarray(values, np.float64)).reshape(-1, 1).to(v.device))

    @property
    def data_pointer(self):
        return self._plain("data_pointer") if self.__dict__.get("_dev") is None else self._dev.counters()[2]

    @property
    def data_length(self):
        return self._plain("data_length") if self.__dict__.get("_dev") is None else min(self._dev.counters()[0], self.capacity)

    def __len__(self):
        return self.data_length

    @property
    def total_priority(self):
        return float(self.tree[0])


class PER(H.DeviceReplayMixin):
    """Prioritised replay on the device sum tree (reference PER, enet_sac.py:203-346)."""
    epsilon = 0.01
    alpha = 0.6
    beta_increment_per_sampling = 1e-4
    absolute_error_upper = 100.

    def __init__(self, capacity, input_shape, n_actions, _dev=None):
        assert is_power_of_2(int(capacity)), "Capacity must be power of 2."
        if _dev is None:       # stand-alone PER: private arena, batch size fixed at first sample_buffer call
            self._standalone = (int(capacity), int(input_shape[0]), int(n_actions))
        self._dev = _dev
        self.mem_size = capacity
        self.tree = SumTree(capacity, _dev=_dev) if _dev is not None else None
        self.filename = 'prioritized_replaymem_sac.model'
        self.last_tree_idx = None

    def _ensure(self, batch_size):
        if self._dev is None:
            C, S, A = self._standalone
            self._dev = H.DeviceAgent(L.SC_KIND_SAC, S, A, batch_size, C, 0.99, 0.005, 1e-3, 1e-3, prioritized=True,
                                      per_clip=self.absolute_error_upper)
            self.tree = SumTree(C, _dev=self._dev)
            for args in getattr(self, "_pending", []):
                self._store(*args)
        return self._dev

    @property
    def beta(self):
        if self.__dict__.get("_dev") is None:
            return self.__dict__.get("beta", 0.4)
        return self._dev.counters()[3]

    def __len__(self):
        return len(self.tree)

    def is_full(self):
        return len(self.tree) >= self.tree.capacity

    def _store(self, s, a, r, s2, done, hint, error):
        dev = self._dev
        if dev is None:        # stand-alone and batch size not known yet: queue
            self.__dict__.setdefault("_pending", []).append((s, a, r, s2, done, hint, error))
            return
        packed = dev.pack(s, a, H._scalar(r), s2, bool(done), hint)
        if error is None:
            dev.store(packed, 1, 0.0, defer=True)
        else:
            p = min((abs(error) + self.epsilon) ** self.alpha, self.absolute_error_upper)   # enet_sac.py:242, same objects
            dev.store(packed, 3 if isinstance(error, T.Tensor) else 2, H._scalar(p), defer=True)

    def store_transition(self, state, action, reward, state_, done, hint, error=None):
        self._store(H.flat_state(state), action, reward, H.flat_state(state_), done, hint, error)

    def store_transition_from_buffer(self, state, action, reward, state_, done, hint, error=None):
        self._store(state, action, reward, state_, done, hint, error)

    def sample_buffer(self, batch_size):
        dev = self._ensure(batch_size)
        if batch_size != dev.B:
            raise ValueError(f"this PER was bound to batch_size={dev.B}")
        idxs, w = dev.per_sample(np.random.random_sample(batch_size), want_outputs=True)
        self.last_tree_idx = idxs
        data = idxs - (dev.C - 1)
        out = tuple(getattr(self, n)[data] for n in ("state_memory", "action_memory", "reward_memory",
                                                     "new_state_memory", "terminal_memory", "hint_memory"))
        return out + (idxs, w)

    def batch_update(self, idxs, errors):
        return self._dev.per_update(idxs, errors)

    def mse(self, expected, targets, is_weights):
        td_error = expected - targets
        weighted_squared_error = is_weights * td_error * td_error
        return T.sum(weighted_squared_error) / T.numel(weighted_squared_error)

    def __reduce__(self):
        # class PER / SumTree with the reference's attribute names in their __dict__ (arrays, counters, tree / capacity /
        # data_pointer / data_length): the reference's PER.load_checkpoint reads the file, and so does ours -- the loaders
        # below only look at __dict__, never at the device-backed properties
        st = self._host_state()
        tree = _HostSumTree()
        tree.__dict__.update(capacity=self.tree.capacity, tree=self.tree.tree.copy(), data_pointer=int(self.tree.data_pointer),
                             data_length=int(self.tree.data_length))
        st["tree"] = tree
        st["beta"] = float(self.beta)
        return (copyreg._reconstructor, (PER, object, None), st)

    def save_checkpoint(self):
        with open(self.filename, 'wb') as f:
            pickle.dump(self, f)

    def load_checkpoint(self):
        """reads files written by this class AND by the reference's PER.save_checkpoint (a pickled reference PER whose
        SumTree carries a raw `tree` ndarray)"""
        with open(self.filename, 'rb') as f:
            other = pickle.load(f)
        st = H.plain_state(other)
        self._upload_arrays(st)
        tree = H.plain_state(st["tree"])
        self.tree.tree = np.asarray(tree["tree"] if "tree" in tree else tree["_host"], np.float64)
        m, l, d, b = self._dev.counters()
        self._dev.set_counters(st["mem_cntr"], l, int(tree.get("data_pointer", st["mem_cntr"] % self.mem_size)), float(st.get("beta", b)))


class CriticNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, name, _dev=None, _net=L.NET_CRITIC_1):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by an Agent in the B200 build (its weights live in the agent's HBM arena)")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, H.critic_keys(), './' + name + '_sac_critic.model')

    def forward(self, state, action):
        raise NotImplementedError("critic forward runs fused inside Agent.learn() (sc_agent_learn)")


class ActorNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, max_action, name, _dev=None):
        self.input_dims, self.n_actions, self.max_action = input_dims, n_actions, max_action
        self.reparam_noise = 1e-6
        self._private = _dev is None
        if _dev is None:       # stand-alone actor (distributed_per_sac.Actor): private arena, inference only
            _dev = H.DeviceAgent(L.SC_KIND_SAC, int(input_dims[0]), n_actions, 1, 1, 0.99, 0.005, lr, lr, max_action=max_action)
        super().__init__(_dev, L.NET_ACTOR, H.actor_keys(L.SC_KIND_SAC), './' + name + '_sac_actor.model')
        if self._private:
            self.load_state_dict(H.actor_initial_state(int(input_dims[0]), n_actions, L.SC_KIND_SAC))

    def sample_normal(self, state, reparameterize=True):
        """single-state sampling through sc_agent_act (enet_sac.py:446-466); returns (action (1,A), None)."""
        s = T.as_tensor(state).detach().cpu().reshape(-1).numpy().astype(np.float32)
        eps = T.randn(self.n_actions).numpy()
        a = self._dev.act(s, eps)
        return T.as_tensor(a).unsqueeze(0), None


class Agent():
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, reward_scale=2, alpha=0.1, name_prefix='', prioritized=False,
                 use_hint=False, rng=None, device=None):
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.max_action, self.min_action = 1, -1
        self.prioritized, self.use_hint = prioritized, use_hint
        self.rng = rng or H.default_rng_mode()
        S = int(input_dims[0])
        self._dev = dev = H.DeviceAgent(L.SC_KIND_SAC, S, n_actions, batch_size, max_mem_size, gamma, tau, lr_a, lr_c,
                                        reward_scale=reward_scale, alpha=alpha, prioritized=prioritized, use_hint=use_hint,
                                        max_action=self.max_action, device=device,
                                        rng_seed=int(T.initial_seed()) & 0xffffffffffff)
        self.replaymem = (PER(max_mem_size, input_dims, n_actions, _dev=dev) if prioritized
                          else ReplayBuffer(max_mem_size, input_dims, n_actions, _dev=dev))
        # same construction (and therefore generator-consumption) order as the reference, enet_sac.py:493-499
        self.actor = ActorNetwork(lr_a, input_dims, n_actions, self.max_action, name_prefix + 'a_eval', _dev=dev)
        self.actor.load_state_dict(H.actor_initial_state(S, n_actions, L.SC_KIND_SAC))
        self.critic_1 = CriticNetwork(lr_c, input_dims, n_actions, name_prefix + 'q_eval_1', _dev=dev, _net=L.NET_CRITIC_1)
        self.critic_1.load_state_dict(H.critic_initial_state(S, n_actions))
        self.critic_2 = CriticNetwork(lr_c, input_dims, n_actions, name_prefix + 'q_eval_2', _dev=dev, _net=L.NET_CRITIC_2)
        self.critic_2.load_state_dict(H.critic_initial_state(S, n_actions))
        self.target_critic_1 = CriticNetwork(lr_c, input_dims, n_actions, name_prefix + 'q_target_1', _dev=dev, _net=L.NET_TARGET_CRITIC_1)
        H.critic_initial_state(S, n_actions)      # the reference draws (then overwrites) the target inits too
        self.target_critic_2 = CriticNetwork(lr_c, input_dims, n_actions, name_prefix + 'q_target_2', _dev=dev, _net=L.NET_TARGET_CRITIC_2)
        H.critic_initial_state(S, n_actions)
        self._alpha, self._admm_rho = float(alpha), 0.01
        self.scale = reward_scale
        self.learn_alpha = False
        if self.use_hint:
            self.hint_threshold = 0.1
        self.update_network_parameters(tau=1.)
        self.last_stats = None

    # scalars the kernels read live in the library: assignments (agent.alpha = ..., agent.admm_rho = ...) reach the device
    @property
    def alpha(self):
        return T.tensor(self._alpha, device=self._dev.device)

    @alpha.setter
    def alpha(self, v):
        self._alpha = float(v)
        self._dev.set_scalars(self._alpha, float(self.rho), self._admm_rho)

    @property
    def admm_rho(self):
        return self._admm_rho

    @admm_rho.setter
    def admm_rho(self, v):
        self._admm_rho = float(v)
        self._dev.set_scalars(self._alpha, float(self.rho), self._admm_rho)

    @property
    def rho(self):
        """hint multiplier (enet_sac.py:515); it evolves on the device -- read from there"""
        return T.tensor(self._dev.rho())

    @property
    def learn_counter(self):
        return self._dev.n_learned

    def update_network_parameters(self, tau=None):
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal, hint):
        self.replaymem.store_transition(state, action, reward, state_, terminal, hint)

    remember = store_transition     # name used by the tutorial code the reference credits (README.md:5)

    def choose_action(self, observation):
        eps = T.randn(self.n_actions).numpy() if self.rng == 'reference' else None
        return self._dev.act(H.flat_state(observation), eps)

    def learn(self, want_stats=False):
        """Agent.learn() (enet_sac.py:555-628).  want_stats=True returns this step's losses (one synchronisation);
        want_stats='deferred' returns the PREVIOUS step's losses without synchronising, so host-side preparation of the next
        step overlaps the device (every step's losses are still read exactly once)."""
        dev = self._dev
        if dev.n_stored < self.batch_size:
            dev.flush()
            return
        if (want_stats == 'deferred') != getattr(self, '_deferred', False):
            self._deferred = want_stats == 'deferred'
            dev.set_stats_mode(self._deferred)
        B, A = self.batch_size, self.n_actions
        # the reference prints `learn_counter rho` every 100th call when hints are on (enet_sac.py:619-620): only then is a
        # read-back (and its synchronisation) needed
        report = self.use_hint and dev.n_learned % 100 == 0
        if self.rng == 'reference':
            # draws in the reference's order: numpy batch first, then torch noise (enet_sac.py:48/294, :570, :589, :611); the
            # arrays (and a deferred store_transition) travel to the device in the step's one host->device copy
            if self.prioritized:
                kw = dict(per_u01=np.random.random_sample(B))
            else:
                kw = dict(batch_idx=np.random.choice(min(dev.n_stored, dev.C), B, replace=False))
            e1, e2 = T.randn(B, A).numpy(), T.randn(B, A).numpy()
            e3 = T.randn(B, A).numpy() if (self.use_hint and dev.n_learned % 10 == 0) else None
            st = dev.learn(e1, e2, e3, want_stats=want_stats or report, **kw)
        else:
            st = dev.learn(want_stats=want_stats or report)
        if st is not None:
            self.last_stats = st
            if report:
                print(f'{st["learn_counter"] - 1} {st["rho"]}')
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
        self.update_network_parameters(tau=1.)

    def print(self):
        print(self.actor)
        print(self.critic_1)
