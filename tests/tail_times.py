"""Stage timeline of k_tail (critic phase) from SMARTCAL_STAMPS marks: mean over the 64 CTAs of agent 0, microseconds."""
import os, sys
os.environ["SMARTCAL_STAMPS"] = "1"
os.environ["SMARTCAL_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
for _ in range(6):
    dev.learn()
torch.cuda.synchronize()
ts = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[2400:2400 + 128].reshape(64, 16).astype(np.float64)
names = {0: "start", 1: "rows + params staged", 2: "actor fc3+LN+heads", 3: "policy head", 10: "fc21", 11: "LN21", 12: "fc22 (4 nets)", 4: "LN22 (branches done)",
         5: "critic q", 6: "y / dq / stores", 7: "dcat + LN bwd 12/22", 8: "stores + w22^T x2", 9: "LN bwd 21 + stores"}
order = [0, 1, 2, 3, 10, 11, 12, 4, 5, 6, 7, 8, 9]
prev = None
for i in order:
    if prev is not None:
        d = (ts[:, i] - ts[:, prev]) / 1e3
        print(f"  {names[i]:26s} {d.mean():7.2f}  (min {d.min():6.2f} max {d.max():6.2f})")
    prev = i
print(f"  total per CTA {((ts[:, 9] - ts[:, 0]) / 1e3).mean():.2f} us; first start -> last end {(ts[:, 9].max() - ts[:, 0].min()) / 1e3:.2f} us")
