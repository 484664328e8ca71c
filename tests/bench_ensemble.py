"""Aggregate learn() throughput of E agents stepped by one launch sequence (DeviceEnsemble), bench workload, warm L2.
usage: python tests/bench_ensemble.py [E ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from smartcal_b200.core import DeviceEnsemble

Es = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8]
for E in Es:
    agents = [bench.make_agent(100 + i, "device")[0] for i in range(E)]
    ens = DeviceEnsemble([a._dev for a in agents])
    for _ in range(20):
        ens.learn()
    torch.cuda.synchronize()
    n = 200
    best = 1e9
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            ens.learn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    print(f"E={E:3d}  {best:9.1f} us per ensemble step  {best / E:8.1f} us per agent-step  {E / best * 1e6:10.0f} agent-steps/s", flush=True)
    ens.close()
    for a in agents:
        a._dev.close()
