"""Drop-in for the reference script elasticnet/distributed_per_sac.py: one Learner (rank 0) fed by Actors (ranks >= 1).

Same roles, same per-episode protocol and the same `Learner` / `Actor` method names as the reference
(distributed_per_sac.py:23-152), but the reference's torch.distributed.rpc / TensorPipe-over-TCP star is replaced by
one process per GPU with torch.distributed (NCCL over NVLink 5 / NVSwitch):

  reference (per episode and actor)                         here (per episode, all actors at once)
  ------------------------------------------------------    ---------------------------------------------------------
  actor pulls weights: learner_rref.get_actor_params()      ONE dist.broadcast of the contiguous actor block (1.53 MB)
    -> dict of CPU tensors pickled over TCP  (:84-90,126)      straight from the learner's HBM arena into each actor's
  actor pushes its 100-slot ReplayBuffer, pickled (:144)    ONE dist.gather of packed [100][2S+2A+2] float32 blocks that
                                                               land in the learner's HBM
  learner: store_transition_from_buffer + learn() per       same loop, but store reads the landed block in place
    transition under a threading.Lock (:42-55)                 (sc_replay_store_dev) and learn() is the CUDA path

Launch (one box):  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 distributed_per_sac.py --episodes 10
The reference's CLI (--rank/--world-size/--learner-addr/--learner-port) is accepted too and maps onto the same
environment variables.  `backend='gloo'` runs the identical protocol on CPU tensors (used by the CPU tests with stub
agents; the real agents need a GPU).
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

try:
    from . import _host as H
except ImportError:
    import _host as H

LEARNER_NAME = 'learner'
ACTOR_NAME = 'actor{}'
TRANSITIONS_PER_EPISODE = 100      # 10 epochs x 10 steps, distributed_per_sac.py:130-141


def packed_len(S, A):
    return 2 * S + 2 * A + 2


def pack_buffer(buf, n, S, A):
    """reference ReplayBuffer (host arrays) -> [n][2S+2A+2] float32 rows [s | s2 | a | hint | r | done]"""
    out = np.zeros((n, packed_len(S, A)), np.float32)
    out[:, :S] = buf.state_memory[:n]
    out[:, S:2 * S] = buf.new_state_memory[:n]
    out[:, 2 * S:2 * S + A] = buf.action_memory[:n]
    out[:, 2 * S + A:2 * S + 2 * A] = buf.hint_memory[:n]
    out[:, 2 * S + 2 * A] = buf.reward_memory[:n]
    out[:, 2 * S + 2 * A + 1] = buf.terminal_memory[:n].astype(np.float32)
    return out


class Exchange:
    """the two collectives of an episode; tensors live on `device` (cuda for nccl, cpu for gloo).

    Upload path: ONE all_gather_into_tensor of fixed [n][T] blocks into a buffer allocated once (every rank contributes a
    block, the learner's is a placeholder).  Round 1 used dist.gather, which NCCL runs as world-1 point-to-point receives
    into freshly allocated tensors on rank 0: at 8 ranks that serialised fan-in (first use of 7 peer channels, 4 of them on
    the far NUMA node) cost 13 ms per episode against 0.27 ms at 4 ranks.  all_gather is a ring / NVLS collective whose cost
    does not depend on who needs the result, and its channels are set up once by warm()."""

    def __init__(self, device):
        self.device = device
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        # gloo cannot all_gather CUDA tensors: stage through the host (1-GPU test boxes run both roles on one device)
        self.stage_cpu = dist.get_backend() == 'gloo' and torch.device(device).type == 'cuda'
        self._gathered = None

    def broadcast_actor(self, flat_params):
        if self.stage_cpu:
            h = flat_params.detach().cpu()
            dist.broadcast(h, src=0)
            if self.rank != 0:
                flat_params.copy_(h.to(flat_params.device))
            return flat_params
        dist.broadcast(flat_params, src=0)
        return flat_params

    def _buffer(self, block):
        shape = (self.world,) + tuple(block.shape)
        if self._gathered is None or tuple(self._gathered.shape) != shape or self._gathered.dtype != block.dtype:
            self._gathered = torch.empty(shape, dtype=block.dtype, device='cpu' if self.stage_cpu else block.device)
        return self._gathered

    def gather_blocks(self, block):
        """block: [n][T] on every rank (the learner contributes a placeholder).  Returns the actors' blocks on rank 0 (views
        of the persistent landing buffer, valid until the next call), None elsewhere."""
        out = self._buffer(block)
        src = block.detach().cpu() if self.stage_cpu else block.contiguous()
        try:
            dist.all_gather_into_tensor(out.view(-1), src.reshape(-1))
        except (RuntimeError, NotImplementedError):      # backends without the flat variant
            dist.all_gather([out[k] for k in range(self.world)], src)
        if self.rank != 0:
            return None
        land = out.to(block.device) if self.stage_cpu else out
        return [land[k] for k in range(1, self.world)]

    def warm(self, n_params, rows, row_len, dtype=torch.float32):
        """create every channel of both collectives before the first timed episode"""
        flat = torch.zeros(n_params, dtype=dtype, device=self.device)
        blk = torch.zeros((rows, row_len), dtype=dtype, device=self.device)
        for _ in range(2):
            self.broadcast_actor(flat)
            self.gather_blocks(blk)
        if torch.device(self.device).type == 'cuda':
            torch.cuda.synchronize()


class SyntheticENetEnv:
    """stand-in with the observation/step contract of enetenv.ENetEnv (obs dict {'eig': (N,), 'A': (N*M,)}, step ->
    (obs, reward, done, hint, info)) for benchmarks and tests; the real env is the reference's own enetenv.py"""

    def __init__(self, M=20, N=20, provide_hint=True, seed=0):
        self.M, self.N, self.provide_hint = M, N, provide_hint
        self.rs = np.random.RandomState(seed)

    def _obs(self):
        return {'eig': torch.from_numpy(self.rs.standard_normal(self.N).astype(np.float32)),
                'A': torch.from_numpy(self.rs.standard_normal(self.N * self.M).astype(np.float32))}

    def reset(self):
        return self._obs()

    def step(self, action):
        reward = torch.tensor(float(self.rs.standard_normal()))
        hint = self.rs.uniform(-1, 1, 2).astype(np.float32)
        if self.provide_hint:
            return self._obs(), reward, False, hint, {}
        return self._obs(), reward, False, {}


def make_env(M, N, provide_hint=True, synthetic=False, seed=0):
    if not synthetic:
        try:
            from enetenv import ENetEnv        # the reference's environment, unchanged, when it is on sys.path
            return ENetEnv(M, N, provide_hint=provide_hint)
        except ImportError:
            pass
    return SyntheticENetEnv(M, N, provide_hint, seed)


class Learner:
    def __init__(self, world_size, N=20, M=20, agent=None, exchange=None, batch_size=64, max_mem_size=1024):
        self.N, self.M, self.world_size = N, M, world_size
        if agent is None:
            try:
                from .enet_sac import Agent
            except ImportError:
                from enet_sac import Agent
            # same hyper-parameters as the reference Learner (distributed_per_sac.py:38-40)
            agent = Agent(gamma=0.99, batch_size=batch_size, n_actions=2, tau=0.005, max_mem_size=max_mem_size,
                          input_dims=[N + N * M], lr_a=1e-3, lr_c=1e-3, reward_scale=N, prioritized=True, use_hint=True)
        self.agent = agent
        self.ex = exchange
        self.learn_steps = 0

    def get_actor_params(self):
        """the contiguous actor parameter block in HBM (reference returned a dict of CPU copies, :84-90)"""
        return self.agent.actor.flat()

    def download_replaybuffer(self, obs_id, block, n=None):
        """ingest one actor's packed transitions: store + one learn() per transition (reference :42-57)"""
        n = block.shape[0] if n is None else n
        print(f'Learner: update replaybuffer from {obs_id}')
        dev = getattr(self.agent, "_dev", None)
        for i in range(n):
            if dev is not None:
                dev.store_dev(block, i, 1 if dev.prioritized else 0, 0.0)
            else:                              # stub agents in the CPU tests
                self.agent.store_packed(block[i])
            self.agent.learn()
            self.learn_steps += 1
        print(f'Learner: replaybuffer size {self.agent.replaymem.mem_cntr}')

    def run_episode(self, episode, empty_block):
        print(f'Learner: episode {episode}')
        self.ex.broadcast_actor(self.get_actor_params())
        blocks = self.ex.gather_blocks(empty_block)
        for k, blk in enumerate(blocks):
            self.download_replaybuffer(k + 1, blk)
        if episode % 10 == 0:
            self.agent.save_models()

    def run_episodes(self, max_episodes, S=None, A=2, device=None):
        S = S or (self.N + self.N * self.M)
        device = device or self.ex.device
        empty = torch.zeros((TRANSITIONS_PER_EPISODE, packed_len(S, A)), dtype=torch.float32, device=device)
        for episode in range(max_episodes):
            self.run_episode(episode, empty)


class Actor:
    def __init__(self, N, M, input_dims, n_actions, max_mem_size=100, env=None, actor=None, exchange=None, rank=None,
                 synthetic_env=False):
        self.id = dist.get_rank() if rank is None else rank
        print(f'Actor: id {self.id}')
        self.N, self.M, self.S, self.A = N, M, int(input_dims[0]), n_actions
        self.env = env if env is not None else make_env(M, N, True, synthetic_env, seed=self.id)
        try:
            from .enet_sac import ActorNetwork, ReplayBuffer
        except ImportError:
            from enet_sac import ActorNetwork, ReplayBuffer
        self.actor = actor if actor is not None else ActorNetwork(lr=1e-3, input_dims=input_dims, n_actions=n_actions,
                                                                  max_action=1, name='a_' + str(self.id) + '_eval')
        self.replaymem = ReplayBuffer(max_mem_size, input_dims, n_actions)
        self.ex = exchange

    def choose_action(self, observation):
        state = torch.cat((observation['eig'], observation['A']))
        actions, _ = self.actor.sample_normal(state, reparameterize=False)
        return actions.cpu().detach().numpy()[0]

    def run_observations(self, quiet=False):
        """one episode: refresh weights, 10 epochs x 10 env steps, upload the buffer (reference :123-146)"""
        self.ex.broadcast_actor(self.actor.flat())
        for epoch in range(10):
            observation = self.env.reset()
            for ci in range(10):
                action = self.choose_action(observation)
                observation_, reward, done, hint, info = self.env.step(action)
                self.replaymem.store_transition(observation, action, reward, observation_, done, hint)
                observation = observation_
                if not quiet:
                    print(f'Actor: {self.id} {epoch} {ci} {reward}')
        n = min(self.replaymem.mem_cntr, self.replaymem.mem_size)
        block = torch.from_numpy(pack_buffer(self.replaymem, TRANSITIONS_PER_EPISODE, self.S, self.A)).to(self.ex.device)
        self.ex.gather_blocks(block)
        self.replaymem.mem_cntr = 0
        return n


def run_process(rank, world_size, addr, port, episodes=1000, backend=None, synthetic_env=False, quiet=False):
    os.environ.setdefault('MASTER_ADDR', addr)
    os.environ.setdefault('MASTER_PORT', str(port))
    backend = backend or ('nccl' if torch.cuda.is_available() else 'gloo')
    if backend == 'nccl':
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', rank % max(torch.cuda.device_count(), 1))))
        device = torch.device('cuda', torch.cuda.current_device())
    else:
        device = torch.device('cpu')
    dist.init_process_group(backend, rank=rank, world_size=world_size)
    ex = Exchange(device)
    N = M = 20
    if rank == 0:
        print(f'invoking learner at {addr}:{port}')
        learner = Learner(world_size, N=N, M=M, exchange=ex)
        learner.run_episodes(episodes)
    else:
        print(f'invoking actor {rank}')
        actor = Actor(N, M, [N + N * M], 2, exchange=ex, rank=rank, synthetic_env=synthetic_env)
        for _ in range(episodes):
            actor.run_observations(quiet=quiet)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Elastic net regression tuning with distributed PER (NCCL)',
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('--rank', default=int(os.environ.get('RANK', 0)), type=int, metavar='r', help='rank of this process')
    parser.add_argument('--world-size', default=int(os.environ.get('WORLD_SIZE', 2)), type=int, metavar='w',
                        help='number of processes, one learner and actors')
    parser.add_argument('--learner-addr', default=os.environ.get('MASTER_ADDR', '127.0.0.1'), type=str, metavar='l',
                        help='learner (rank 0) address')
    parser.add_argument('--learner-port', default=os.environ.get('MASTER_PORT', '59999'), type=str, metavar='p',
                        help='learner (rank 0) port')
    parser.add_argument('--episodes', default=1000, type=int)
    parser.add_argument('--synthetic-env', action='store_true', help='use the synthetic stand-in instead of enetenv.ENetEnv')
    args = parser.parse_args()
    run_process(rank=args.rank, world_size=args.world_size, addr=args.learner_addr, port=args.learner_port,
                episodes=args.episodes, synthetic_env=args.synthetic_env)
