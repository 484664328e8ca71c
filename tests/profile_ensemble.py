"""ncu driver: a few ensemble steps (E agents, bench workload)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from smartcal_b200.core import DeviceEnsemble
E = int(sys.argv[1]) if len(sys.argv) > 1 else 8
agents = [bench.make_agent(100 + i, "device")[0] for i in range(E)]
ens = DeviceEnsemble([a._dev for a in agents])
for _ in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    ens.learn()
torch.cuda.synchronize()
print("done")
