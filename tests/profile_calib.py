"""Tiny driver for ncu: a few learn() steps of the calibration conv-SAC workload (BASELINE.json configs[2], B=256, M=10)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smartcal_b200.calibration import calib_sac
from smartcal_b200.synth import make_calib_transitions
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M, B, C = 10, 256, 512
A = 2 * M
torch.manual_seed(0); np.random.seed(0)
agent = calib_sac.Agent(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], M=M, lr_a=1e-3, lr_c=1e-3,
                        reward_scale=M, alpha=0.03, use_hint=False, rng="device")
tr = make_calib_transitions(100, C, M, A, 128, 0.0)
dev = agent._dev
for name, key in (("state_memory_img", "img"), ("new_state_memory_img", "img2"), ("state_memory_sky", "sky"), ("new_state_memory_sky", "sky2"),
                  ("action_memory", "action"), ("hint_memory", "hint")):
    dev.buffer(name).copy_(torch.as_tensor(tr[key].reshape(C, -1)).to(dev.device))
dev.set_counters(C, 0)
rs = np.random.RandomState(1)
import time
for i in range(steps):
    if i == steps // 2:
        torch.cuda.synchronize(); t0 = time.perf_counter()
    dev.set_batch(rs.choice(C, B, replace=False))
    dev.learn()
torch.cuda.synchronize()
if steps >= 4:
    print(f"{(time.perf_counter() - t0) / (steps - steps // 2) * 1e3:.3f} ms/step (second half of {steps} steps, wall clock)")
print("done", dev.launch_count())
