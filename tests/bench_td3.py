"""BASELINE.json configs[3]: elastic-net TD3, batch 1024, prioritized, capacity 4096 -- device-side steps/s (CUDA events, warm L2)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smartcal_b200.elasticnet import enet_td3, enet_ddpg
from smartcal_b200.synth import make_transitions, pack
S, A, C = 420, 2, 4096
for name, B in (("td3", 1024), ("td3", 256), ("ddpg", 64)):
    torch.manual_seed(0); np.random.seed(0)
    if name == "td3":
        agent = enet_td3.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005,
                               update_actor_interval=2, warmup=100, noise=0.1, prioritized=True, rng="device")
    else:
        agent = enet_ddpg.Agent(gamma=0.99, lr_a=1e-4, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.001, rng="device")
    tr = make_transitions(100, C, S, A, "uniform", 0.0)
    agent._dev.fill(pack(tr), None)
    dev = agent._dev
    for _ in range(20):
        agent.learn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        agent.learn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{name} B={B}: {n / ms * 1e3:.0f} learn steps/s ({ms / n * 1e3:.0f} us/step), launches/step {dev.launch_count() / (n + 20):.1f}")
