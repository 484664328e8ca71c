"""Quick A/B timer: warm learn() steps of the bench workload under the current SMARTCAL_* environment (no L2 flush)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
for _ in range(30):
    dev.learn()
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(3):
    e0.record()
    for _ in range(n):
        dev.learn()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / n * 1e3)
print(f"{os.environ.get('SMARTCAL_GEOM', 'default'):8s} KS={os.environ.get('SMARTCAL_DENSE_KS', 'auto'):5s} {best:8.1f} us/step (warm L2)")
