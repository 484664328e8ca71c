"""CPU (gloo, world_size 2): the actor->learner exchange protocol of distributed_per_sac.py with stub agents -- weight
broadcast reaches the actor, 100 packed transitions per episode land on the learner in order, one learn() per
transition, packed layout round-trips.  The CUDA agents are exercised by the -m gpu tests and bench.py."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class StubNet:
    def __init__(self, n, value):
        self._flat = torch.full((n,), float(value))

    def flat(self):
        return self._flat

    def sample_normal(self, state, reparameterize=True):
        return torch.tanh(state[:2]).unsqueeze(0), None


class StubMem:
    mem_cntr = 0


class StubAgent:
    def __init__(self, nparams):
        self.actor = StubNet(nparams, 3.25)
        self.replaymem = StubMem()
        self.rows, self.learn_calls, self.saved = [], 0, 0

    def store_packed(self, row):
        self.rows.append(row.clone())
        self.replaymem.mem_cntr += 1

    def learn(self):
        self.learn_calls += 1

    def save_models(self):
        self.saved += 1


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import smartcal_b200  # noqa: F401
    from smartcal_b200.elasticnet import distributed_per_sac as D
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ex = D.Exchange(torch.device("cpu"))
    N = M = 4
    S, A, NP = N + N * M, 2, 37
    episodes = 2
    if rank == 0:
        agent = StubAgent(NP)
        learner = D.Learner(world, N=N, M=M, agent=agent, exchange=ex)
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            learner.run_episodes(episodes, S=S, A=A)
        rows = torch.stack(agent.rows).numpy()
        np.save(os.path.join(outdir, "learner_rows.npy"), rows)
        np.save(os.path.join(outdir, "learner_meta.npy"), np.array([agent.learn_calls, learner.learn_steps, agent.saved]))
    else:
        actor_net = StubNet(NP, 0.0)
        env = D.SyntheticENetEnv(M, N, True, seed=5)
        actor = D.Actor(N, M, [S], A, env=env, actor=actor_net, exchange=ex, rank=rank)
        sent = []
        orig = D.pack_buffer

        def spy(buf, n, S_, A_):
            out = orig(buf, n, S_, A_)
            sent.append(out.copy())
            return out
        D.pack_buffer = spy
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(episodes):
                assert actor.run_observations(quiet=True) == 100
        np.save(os.path.join(outdir, "actor_rows.npy"), np.concatenate(sent))
        np.save(os.path.join(outdir, "actor_flat.npy"), actor_net.flat().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_actor_learner_exchange_gloo_world2():
    port = 29500 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        lrows, arows = np.load(os.path.join(d, "learner_rows.npy")), np.load(os.path.join(d, "actor_rows.npy"))
        meta = np.load(os.path.join(d, "learner_meta.npy"))
        flat = np.load(os.path.join(d, "actor_flat.npy"))
    assert lrows.shape == (200, 2 * 20 + 2 * 2 + 2)
    assert np.array_equal(lrows, arows), "transitions did not arrive intact / in order"
    assert meta[0] == 200 and meta[1] == 200, "one learn() per ingested transition (distributed_per_sac.py:47-55)"
    assert meta[2] == 1, "save_models every 10th episode starting with episode 0"
    assert np.all(flat == 3.25), "actor did not receive the learner's weights"


def test_pack_buffer_layout():
    sys.path.insert(0, ROOT)
    import smartcal_b200  # noqa: F401
    from smartcal_b200.elasticnet import distributed_per_sac as D
    from smartcal_b200.elasticnet.enet_sac import ReplayBuffer
    S, A = 6, 2
    rb = ReplayBuffer(8, [S], A)
    for i in range(3):
        o = {"eig": torch.full((2,), float(i)), "A": torch.full((4,), float(i) + 0.5)}
        rb.store_transition(o, np.array([0.1 * i, -0.1 * i], np.float32), 2.0 * i, o, i == 1, np.array([9.0, 8.0], np.float32))
    p = D.pack_buffer(rb, 3, S, A)
    assert p.shape == (3, D.packed_len(S, A))
    assert np.allclose(p[2, :S], [2, 2, 2.5, 2.5, 2.5, 2.5]) and np.allclose(p[2, 2 * S:2 * S + A], [0.2, -0.2])
    assert np.allclose(p[:, 2 * S + 2 * A], [0, 2, 4]) and np.allclose(p[:, 2 * S + 2 * A + 1], [0, 1, 0])
