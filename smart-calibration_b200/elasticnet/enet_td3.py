"""Drop-in for the reference module elasticnet/enet_td3.py (TD3 agent with optional PER), running on
libsmartcal_b200.so.  Same constructor signature, methods, attribute names and checkpoint files
(./a_eval_td3_actor.model, ./q_eval_1_td3_critic.model, ...; enet_td3.py:53,102,151-159).

Kept quirks: one scalar of target-smoothing noise per step, priorities refreshed before the critic step,
(B,)x(B,1) broadcast in PER.mse and the actor loss, actor + all three targets move only every
`update_actor_interval`-th learn(); with use_hint the adaptive-ADMM inner loop of enet_td3.py:310-361 (five actor steps,
penalty re-estimated after the 4th) runs as five actor row-phase + Adam launches with the multipliers kept on the device.
"""
import numpy as np
import torch as T

try:
    from . import _host as H
    from .enet_sac import ReplayBuffer, PER
except ImportError:
    import _host as H
    from enet_sac import ReplayBuffer, PER

L = H.L


class CriticNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, name, _dev=None, _net=L.NET_CRITIC_1):
        if _dev is None:
            raise RuntimeError("CriticNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, H.critic_keys(), './' + name + '_td3_critic.model')


class ActorNetwork(H.NetView):
    def __init__(self, lr, input_dims, n_actions, name, _dev=None, _net=L.NET_ACTOR):
        if _dev is None:
            raise RuntimeError("TD3 ActorNetwork is owned by an Agent in the B200 build")
        self.input_dims, self.n_actions = input_dims, n_actions
        super().__init__(_dev, _net, H.actor_keys(L.SC_KIND_TD3), './' + name + '_td3_actor.model')


class Agent():
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, update_actor_interval=2, warmup=1000, noise=0.1, prioritized=False,
                 use_hint=False, rng=None, device=None):
        self.gamma, self.tau, self.batch_size, self.n_actions = gamma, tau, batch_size, n_actions
        self.max_action, self.min_action = 1, -1
        self.time_step, self.warmup = 0, warmup
        self.update_actor_interval = update_actor_interval
        self.prioritized, self.use_hint = prioritized, use_hint
        self.admm_rho, self.Nadmm, self.adaptive_admm, self.corr_min = 0.1, 5, True, 0.5
        self.noise = noise
        self.rng = rng or H.default_rng_mode()
        S = int(input_dims[0])
        self._dev = dev = H.DeviceAgent(L.SC_KIND_TD3, S, n_actions, batch_size, max_mem_size, gamma, tau, lr_a, lr_c,
                                        prioritized=prioritized, use_hint=use_hint, update_actor_interval=update_actor_interval,
                                        admm_rho=self.admm_rho,
                                        device=device, rng_seed=int(T.initial_seed()) & 0xffffffffffff)
        self.replaymem = (PER(max_mem_size, input_dims, n_actions, _dev=dev) if prioritized
                          else ReplayBuffer(max_mem_size, input_dims, n_actions, _dev=dev))
        # construction order of the reference (enet_td3.py:151-159)
        self.actor = ActorNetwork(lr_a, input_dims, n_actions, 'a_eval', _dev=dev, _net=L.NET_ACTOR)
        self.actor.load_state_dict(H.actor_initial_state(S, n_actions, L.SC_KIND_TD3))
        self.critic_1 = CriticNetwork(lr_c, input_dims, n_actions, 'q_eval_1', _dev=dev, _net=L.NET_CRITIC_1)
        self.critic_1.load_state_dict(H.critic_initial_state(S, n_actions))
        self.critic_2 = CriticNetwork(lr_c, input_dims, n_actions, 'q_eval_2', _dev=dev, _net=L.NET_CRITIC_2)
        self.critic_2.load_state_dict(H.critic_initial_state(S, n_actions))
        self.target_actor = ActorNetwork(lr_a, input_dims, n_actions, 'a_target', _dev=dev, _net=L.NET_TARGET_ACTOR)
        H.actor_initial_state(S, n_actions, L.SC_KIND_TD3)
        self.target_critic_1 = CriticNetwork(lr_c, input_dims, n_actions, 'q_target_1', _dev=dev, _net=L.NET_TARGET_CRITIC_1)
        H.critic_initial_state(S, n_actions)
        self.target_critic_2 = CriticNetwork(lr_c, input_dims, n_actions, 'q_target_2', _dev=dev, _net=L.NET_TARGET_CRITIC_2)
        H.critic_initial_state(S, n_actions)
        self.update_network_parameters(tau=1.)
        self.last_stats = None
        self._admm_rho_sent = self.admm_rho
        self._dev_rng = np.random.default_rng(int(T.initial_seed()) & 0xffffffff)

    @property
    def learn_step_cntr(self):
        return self._dev.n_learned

    def update_network_parameters(self, tau=None):
        self._dev.soft_update(self.tau if tau is None else tau)

    def store_transition(self, state, action, reward, state_, terminal, hint):
        if not self.prioritized:
            self.replaymem.store_transition(state, action, reward, state_, terminal, hint)
        else:   # reward doubles as the initial error (enet_td3.py:205)
            self.replaymem.store_transition(state, action, reward, state_, terminal, hint, reward)

    remember = store_transition

    def choose_action(self, observation):
        """enet_td3.py:207-220: Gaussian warm-up, then actor(s) + N(0,noise), clamped to [-1,1]."""
        if self.time_step < self.warmup:
            mu = T.tensor(np.random.normal(scale=self.noise, size=(self.n_actions,)))
            mu_prime = mu + T.tensor(np.random.normal(scale=self.noise, size=(self.n_actions,)), dtype=T.float)
            out = T.clamp(mu_prime, self.min_action, self.max_action).numpy()
        else:
            nz = np.random.normal(scale=self.noise, size=(self.n_actions,)).astype(np.float32)
            out = self._dev.act(H.flat_state(observation), nz)
        self.time_step += 1
        return out

    def learn(self, want_stats=False):
        dev = self._dev
        if dev.n_stored < self.batch_size:
            dev.flush()
            return
        if self.use_hint and self.admm_rho != self._admm_rho_sent:     # callers set agent.admm_rho after construction (main_td3.py:22)
            dev.set_scalars(0.0, 0.0, self.admm_rho)
            self._admm_rho_sent = self.admm_rho
        if self.rng == 'reference':
            # reference draw order (enet_td3.py:232/294 then :249); the arrays ride in the step's one host->device copy
            if self.prioritized:
                kw = dict(per_u01=np.random.random_sample(self.batch_size))
            else:
                kw = dict(batch_idx=np.random.choice(min(dev.n_stored, dev.C), self.batch_size, replace=False))
            st = dev.learn(td3_noise=np.random.normal(scale=0.2), want_stats=want_stats, **kw)
        else:
            # rng='device': minibatch on the device (Philox); the ONE scalar of target-smoothing noise per step
            # (enet_td3.py:248-249) comes from the agent's own seeded generator, not from numpy's global stream
            st = dev.learn(td3_noise=float(self._dev_rng.normal(scale=0.2)), want_stats=want_stats)
        if st is not None:
            self.last_stats = st
        return st if want_stats else None

    def save_models(self):
        for n in (self.actor, self.target_actor, self.critic_1, self.critic_2, self.target_critic_1, self.target_critic_2):
            n.save_checkpoint()
        self.replaymem.save_checkpoint()

    def load_models(self):
        for n in (self.actor, self.target_actor, self.critic_1, self.critic_2, self.target_critic_1, self.target_critic_2):
            n.load_checkpoint()
        self.replaymem.load_checkpoint()
        self.update_network_parameters(tau=1.)

    def load_models_for_eval(self):
        for n in (self.actor, self.critic_1, self.critic_2):
            n.load_checkpoint_for_eval()

    def print(self):
        print(self.actor)
        print(self.critic_1)
