"""Times the UNMODIFIED reference's Agent.learn() (/root/reference/elasticnet/enet_sac.py, CPU) next to the oracle port
(oracle/sac_ref.py) on this container's host cores, same synthetic replay, batch 256, uniform replay (the reference's SAC never
uses PER).  The reference does not travel to the GPU box, so bench.py attaches this file's result to its cpu_baseline as
`port_vs_reference`.   python tools/measure_port_vs_reference.py  ->  profiles/r2_port_vs_reference.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
S, A, B, C = 420, 2, 256, 4096
HP = dict(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[S], batch_size=B, n_actions=A, max_mem_size=C, tau=0.005, reward_scale=20, alpha=0.03)
from smartcal_b200.synth import make_transitions  # noqa: E402

tr = make_transitions(100, C, S, A, "normal", 0.0)
threads = min(8, os.cpu_count() or 1)
torch.set_num_threads(threads)
res = {}

sys.path.insert(0, "/root/reference/elasticnet")
import enet_sac as ref  # noqa: E402  (the unmodified reference)
torch.manual_seed(0)
np.random.seed(0)
ag = ref.Agent(**HP, prioritized=False, use_hint=False)
m = ag.replaymem
m.state_memory[:], m.new_state_memory[:], m.action_memory[:], m.hint_memory[:] = tr["state"], tr["new_state"], tr["action"], tr["hint"]
m.reward_memory[:], m.terminal_memory[:] = tr["reward"], False
m.mem_cntr = C
for _ in range(10):
    ag.learn()
t0 = time.perf_counter()
n = 150
for _ in range(n):
    ag.learn()
res["reference_steps_per_s"] = n / (time.perf_counter() - t0)

from oracle.sac_ref import SacRef  # noqa: E402
torch.manual_seed(0)
np.random.seed(0)
orc = SacRef(**HP, prioritized=False)
for i in range(C):
    orc.store_flat(tr["state"][i], tr["action"][i], float(tr["reward"][i]), tr["new_state"][i], False, tr["hint"][i])
for _ in range(10):
    orc.learn()
t0 = time.perf_counter()
for _ in range(n):
    orc.learn()
res["port_steps_per_s"] = n / (time.perf_counter() - t0)
res["port_over_reference"] = res["port_steps_per_s"] / res["reference_steps_per_s"]
res["how"] = (f"unmodified /root/reference/elasticnet/enet_sac.Agent.learn() vs oracle/sac_ref.SacRef.learn(), S=420 A=2 batch 256 uniform replay, "
              f"{threads} torch threads, {n} timed steps after 10 warm-up, build container ({os.cpu_count()} vCPU, torch {torch.__version__} CPU)")
json.dump(res, open(os.path.join(ROOT, "profiles", "r2_port_vs_reference.json"), "w"), indent=1)
print(res)
