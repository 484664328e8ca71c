"""Quick A/B timer: learn() steps of the bench workload with the L2 flushed between steps (bench.py's `value` condition) and
back to back, under the current SMARTCAL_* environment."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
cold = min(bench.time_agent(dev.learn, n, 30, flush) for _ in range(2))
warm = min(bench.time_agent(dev.learn, n, 30, None) for _ in range(2))
print(f"PREFETCH={os.environ.get('SMARTCAL_PREFETCH', '1')}  flushed {cold * 1e3:7.1f} us/step   warm {warm * 1e3:7.1f} us/step")
