"""Drop-in for the reference module elasticnet/enet_ddpg.py (DDPG agent), running on libsmartcal_b200.so.
Same constructor signature, methods, attribute names and checkpoint files (./a_eval_ddpg_actor.model,
./q_eval_ddpg_critic.model, ...; enet_ddpg.py:121,169,305-310).

DDPG has the same actor/critic networks as the TD3 agent and a strictly simpler update, so it runs on the TD3
row-phase program (include/smartcal_b200.h: SC_FLAG_CRITIC_LOSS_SUM): update_actor_interval = 1, target-smoothing
noise 0 (the clamp to [-1,1] of a tanh output is the identity), the sum-of-squares critic loss of enet_ddpg.py:281, and
the second critic slot of the twin-critic kernels mirrors the first (same initial weights, same minibatch, deterministic
kernels => bit-identical forever, so min(Q1,Q2) = Q).  The mirrored critic costs ~25 % of a step; a single-critic
program is listed in DESIGN.md as follow-up.
"""
import numpy as np
import torch as T

try:
    from . import _host as H
    from .enet_sac import ReplayBuffer as _SacReplayBuffer
except ImportError:
    import _host as H
    from enet_sac import ReplayBuffer as _SacReplayBuffer

L = H.L


class OUActionNoise(object):
    """Ornstein-Uhlenbeck exploration noise with memory (enet_ddpg.py:23-43), drawn from the global NumPy stream."""

    def __init__(self, mu, sigma=0.15, theta=.2, dt=1e-2, x0=None):
        self.theta, self.mu, self.sigma, self.dt, self.x0 = theta, mu, sigma, dt, x0
        self.reset()

    def __call__(self):
        drift = self.theta * (self.mu - self.x_prev) * self.dt
        x = self.x_prev + drift + self.sigma * np.sqrt(self.dt) * np.random.normal(size=self.mu.shape)
        self.x_prev = x
        return x

    def reset(self):
        self.x_prev = self.x0 if self.x0 is not None else np.zeros_like(self.mu)

    def __repr__(self):
        return 'OrnsteinUhlenbeckActionNoise(mu={}, sigma={})'.format(self.mu, self.sigma)


class ReplayBuffer(object):
    """enet_ddpg.py:45-92: five arrays (no hints), HBM-resident when owned by an Agent."""

    def __init__(self, max_size, input_shape, n_actions, _dev=None):
        self._inner = _SacReplayBuffer(max_size, input_shape, n_actions, _dev=_dev)
        self._inner.filename = 'replaymem_ddpg.model'
        self._zero_hint = np.zeros(n_actions, np.float32)

    def store_transition(self, state, action, reward, state_, done):
        self._inner.store_transition(state, action, reward, state_, done, self._zero_hint)

    def sample_buffer(self, batch_size):
        s, a, r, s2, d, _ = self._inner.sample_buffer(batch_size)
        return s, a, r, s2, d.astype(np.float32)

    def __getattr__(self, name):        # mem_cntr, mem_size, *_memory, save/load_checkpoint, filename
        if name == '_inner':            # (an object rebuilt by pickle has no _inner yet: do not recurse)
            raise AttributeError(name)
        return getattr(self._inner, name)


class CriticNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, name, _dev=None, _net=L.NET_CRITIC_1):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, H.critic_keys(), './' + name + '_ddpg_critic.model')


class ActorNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, name, _dev=None, _net=L.NET_ACTOR):
        if _dev is None:
            raise RuntimeError("ActorNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, H.actor_keys(L.SC_KIND_TD3), './' + name + '_ddpg_actor.model')


class Agent():
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size=100, tau=0.001, rng=None, device=None):
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.rng = rng or H.default_rng_mode()
        S = int(input_dims[0])
        self._dev = dev = H.DeviceAgent(L.SC_KIND_TD3, S, n_actions, batch_size, max_mem_size, gamma, tau, lr_a, lr_c,
                                        update_actor_interval=1, flags=L.SC_FLAG_CRITIC_LOSS_SUM, device=device,
                                        rng_seed=int(T.initial_seed()) & 0xffffffffffff)
        self.replaymem = ReplayBuffer(max_mem_size, input_dims, n_actions, _dev=dev)
        # construction order of the reference (enet_ddpg.py:202-209): actor, critic, target actor, target critic
        self.actor = ActorNetwork(lr_a, input_dims, n_actions, 'a_eval', _dev=dev, _net=L.NET_ACTOR)
        self.actor.load_state_dict(H.actor_initial_state(S, n_actions, L.SC_KIND_TD3))
        self.critic = CriticNetwork(lr_c, input_dims, n_actions, 'q_eval', _dev=dev, _net=L.NET_CRITIC_1)
        self.critic.load_state_dict(H.critic_initial_state(S, n_actions))
        self.target_actor = ActorNetwork(lr_a, input_dims, n_actions, 'a_target', _dev=dev, _net=L.NET_TARGET_ACTOR)
        H.actor_initial_state(S, n_actions, L.SC_KIND_TD3)      # the reference draws and then overwrites these (:215-238)
        self.target_critic = CriticNetwork(lr_c, input_dims, n_actions, 'q_target', _dev=dev, _net=L.NET_TARGET_CRITIC_1)
        H.critic_initial_state(S, n_actions)
        self._mirror = CriticNetwork(lr_c, input_dims, n_actions, '_mirror', _dev=dev, _net=L.NET_CRITIC_2)
        self._mirror_t = CriticNetwork(lr_c, input_dims, n_actions, '_mirror_t', _dev=dev, _net=L.NET_TARGET_CRITIC_2)
        self._sync_mirror()
        self.noise = OUActionNoise(mu=np.zeros(n_actions))
        self.update_network_parameters(tau=1.)
        self.last_stats = None

    def _sync_mirror(self):
        """second critic slot := critic (parameters and Adam moments), after construction / load_models"""
        self._mirror.load_state_dict(self.critic.state_dict())
        self._mirror_t.load_state_dict(self.target_critic.state_dict())
        for slot in (L.SLOT_ADAM_M, L.SLOT_ADAM_V):
            for k in H.critic_keys():
                self._dev.tensor(L.NET_CRITIC_2, k, slot).copy_(self._dev.tensor(L.NET_CRITIC_1, k, slot))

    def update_network_parameters(self, tau=None):
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal):
        self.replaymem.store_transition(state, action, reward, state_, terminal)

    remember = store_transition

    def choose_action(self, observation):
        """enet_ddpg.py:243-249: actor(s) + OU noise (no clamp)."""
        mu = self._dev.act(H.flat_state(observation), None)
        return (T.tensor(mu) + T.tensor(self.noise(), dtype=T.float)).numpy()

    def learn(self, want_stats=False):
        dev = self._dev
        if dev.n_stored < self.batch_size:
            dev.flush()
            return
        kw = {}
        if self.rng == 'reference':
            kw = dict(batch_idx=np.random.choice(min(dev.n_stored, dev.C), self.batch_size, replace=False))
        st = dev.learn(td3_noise=0.0, want_stats=want_stats, **kw)
        if st is not None:
            st = dict(st)
            st["critic_loss"] = st["critic_1_loss"] * self.batch_size       # sum of squares (enet_ddpg.py:281)
            self.last_stats = st
        return st if want_stats else None

    def save_models(self):
        for n in (self.actor, self.target_actor, self.critic, self.target_critic):
            n.save_checkpoint()
        self.replaymem.save_checkpoint()

    def load_models(self):
        for n in (self.actor, self.target_actor, self.critic, self.target_critic):
            n.load_checkpoint()
        self.replaymem.load_checkpoint()
        self._sync_mirror()

    def load_models_for_eval(self):
        self.actor.load_checkpoint_for_eval()
        self.critic.load_checkpoint_for_eval()
        self._sync_mirror()

    def print(self):
        print(self.actor)
        print(self.critic)
