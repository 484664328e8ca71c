"""ncu target: one critic-only and one actor-update learn() of the demixing conv TD3 agent at the demixing_rl/main_td3.py shape (batch 64,
use_hint=True), eager launches (SMARTCAL_GRAPH=0 so that every kernel shows up in the launch list)."""
import os, sys
os.environ["SMARTCAL_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smartcal_b200.demixing_rl import demix_td3
from smartcal_b200.synth import make_demix_transitions
M, A, B, C = 20, 5, 64, 128
torch.manual_seed(0); np.random.seed(0)
agent = demix_td3.DemixingAgent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3, lr_c=1e-3,
                                update_actor_interval=2, warmup=200, noise=0.1, use_hint=True, rng="device")
tr = make_demix_transitions(100, B, M, A, 128, 0.0)
for i in range(B):
    agent.store_transition({"infmap": tr["img"][i], "metadata": tr["meta"][i]}, tr["action"][i], float(tr["reward"][i]),
                           {"infmap": tr["img2"][i], "metadata": tr["meta2"][i]}, False, tr["hint"][i])
for _ in range(2):
    agent.learn()
torch.cuda.synchronize()
print("launches", agent._dev.launch_count())
