"""GPU, 2 processes: the reference's Learner / Actor roles (distributed_per_sac.py:23-152) with REAL agents -- the learner's
prioritized SAC agent with hints, the actor's stand-alone ActorNetwork on its own HBM arena -- exchanging over
torch.distributed: NCCL with one GPU per rank when the box has two, otherwise gloo with both roles on cuda:0 (gloo cannot
move CUDA tensors for these collectives, so Exchange stages through the host there; the agents and kernels are the same)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir, nccl):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import contextlib
    import io
    import smartcal_b200  # noqa: F401
    from smartcal_b200.elasticnet import distributed_per_sac as D
    torch.cuda.set_device(rank if nccl else 0)
    device = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl" if nccl else "gloo", rank=rank, world_size=world)
    ex = D.Exchange(device)
    N = M = 4
    S, A = N + N * M, 2
    episodes = 2
    torch.manual_seed(10 + rank)
    np.random.seed(10 + rank)
    if rank == 0:
        learner = D.Learner(world, N=N, M=M, exchange=ex, batch_size=64, max_mem_size=1024)       # real enet_sac.Agent(prioritized, hints)
        ex.warm(learner.get_actor_params().numel(), D.TRANSITIONS_PER_EPISODE, D.packed_len(S, A))
        sent_flat = []
        orig = ex.broadcast_actor

        def spy(flat):
            sent_flat.append(flat.detach().cpu().numpy().copy())
            return orig(flat)
        ex.broadcast_actor = spy
        with contextlib.redirect_stdout(io.StringIO()):
            learner.run_episodes(episodes, S=S, A=A)
        mem = learner.agent.replaymem
        np.savez(os.path.join(outdir, "learner.npz"), state=mem.state_memory, new_state=mem.new_state_memory, action=mem.action_memory,
                 hint=mem.hint_memory, reward=mem.reward_memory, done=mem.terminal_memory, mem_cntr=mem.mem_cntr,
                 learn_counter=learner.agent.learn_counter, learn_steps=learner.learn_steps, flat=np.stack(sent_flat),
                 tree_total=mem.tree.total_priority, finite=np.isfinite(learner.agent.actor.flat().cpu().numpy()).all())
    else:
        env = D.SyntheticENetEnv(M, N, True, seed=5)
        actor = D.Actor(N, M, [S], A, env=env, exchange=ex, rank=rank)                              # real stand-alone ActorNetwork
        ex.warm(actor.actor.flat().numel(), D.TRANSITIONS_PER_EPISODE, D.packed_len(S, A))
        sent, got = [], []
        orig = D.pack_buffer

        def spy(buf, n, S_, A_):
            out = orig(buf, n, S_, A_)
            sent.append(out.copy())
            return out
        D.pack_buffer = spy
        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(episodes):
                assert actor.run_observations(quiet=True) == 100
                got.append(actor.actor.flat().detach().cpu().numpy().copy())
        np.savez(os.path.join(outdir, "actor.npz"), rows=np.concatenate(sent), flat=np.stack(got))
    dist.barrier()
    dist.destroy_process_group()


def test_learner_and_actor_classes_with_real_agents_world2():
    nccl = torch.cuda.device_count() >= 2
    port = 29700 + (os.getpid() % 2000)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d, nccl), nprocs=2, join=True)
        lz, az = np.load(os.path.join(d, "learner.npz")), np.load(os.path.join(d, "actor.npz"))
        rows, S, A = az["rows"], 20, 2
        assert rows.shape == (200, 2 * S + 2 * A + 2)
        # the learner's PER holds the actor's 200 transitions in arrival order (distributed_per_sac.py:47-55)
        assert int(lz["mem_cntr"]) == 200
        assert np.array_equal(lz["state"][:200], rows[:, :S]) and np.array_equal(lz["new_state"][:200], rows[:, S:2 * S])
        assert np.array_equal(lz["action"][:200], rows[:, 2 * S:2 * S + A]) and np.array_equal(lz["hint"][:200], rows[:, 2 * S + A:2 * S + 2 * A])
        assert np.array_equal(lz["reward"][:200], rows[:, 2 * S + 2 * A]) and not lz["done"][:200].any()
        # one learn() call per ingested transition; it is a no-op until the buffer holds a batch (enet_sac.py:556)
        assert int(lz["learn_steps"]) == 200 and int(lz["learn_counter"]) == 200 - 63
        assert float(lz["tree_total"]) > 0 and bool(lz["finite"])
        # each episode the actor's network holds exactly the block the learner broadcast (trained in between: they differ)
        assert np.array_equal(az["flat"], lz["flat"])
        assert not np.array_equal(lz["flat"][0], lz["flat"][1])
