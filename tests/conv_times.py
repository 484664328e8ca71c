"""Timeline of CTA 0 of the tcgen05 conv2 forward kernel (SMARTCAL_STAMPS=1): producer / MMA / gather warp 0 / epilogue warp 0."""
import os, sys
os.environ["SMARTCAL_STAMPS"] = "1"
os.environ["SMARTCAL_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from smartcal_b200.calibration import calib_sac
from smartcal_b200.synth import make_calib_transitions
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
for _ in range(2):
    dev.set_batch(rs.choice(C, B, replace=False))
    dev.learn()
torch.cuda.synchronize()
ts = dev.buffer("debug_ts").view(torch.int64).cpu().numpy().reshape(-1)
names = {1: "producer: issue W", 2: "mma: stage full", 3: "mma: issued", 4: "gather: loop top", 5: "gather: split done", 6: "gather: slot free",
         7: "gather: arrived", 8: "epilogue: acc done", 9: "epilogue: stored"}
evs = []
for role in range(4):
    n = int(ts[256 * role])
    evs += [(int(e) >> 8, int(e) & 255, role) for e in ts[256 * role + 1:256 * role + 1 + n]]
evs.sort()
t0 = evs[0][0]
for t, tag, role in evs[:int(sys.argv[1]) if len(sys.argv) > 1 else 140]:
    print(f"  {(t - t0) / 1e3:8.2f}  {'    ' * role}{names.get(tag, tag)}")
