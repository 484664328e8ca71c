"""ncu driver: a few learn() steps of the demixing_fuzzy MLP agent at main_sac.py's shape (n_meta 32, 128 actions, batch 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smartcal_b200.demixing_fuzzy import demix_sac
from smartcal_b200.synth import make_demix_transitions
M, A, B, C = 32, 128, 256, 1024
torch.manual_seed(0)
agent = demix_sac.DemixingAgent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], n_meta=M, lr_a=3e-4,
                                lr_c=3e-4, alpha=0.03, use_hint=False, use_influence=False, rng="device")
tr = make_demix_transitions(100, C, M, A, 0, 0.0)
dev = agent._dev
for name, key in (("state_memory_sky", "meta"), ("new_state_memory_sky", "meta2"), ("action_memory", "action"), ("hint_memory", "hint")):
    dev.buffer(name).copy_(torch.as_tensor(tr[key].reshape(C, -1)).to(dev.device))
dev.set_counters(C, 0)
rs = np.random.RandomState(1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    dev.set_batch(rs.choice(C, B, replace=False))
    dev.learn()
torch.cuda.synchronize()
