"""Tiny driver for ncu: a few learn() steps of the bench workload (S=420, A=2, B=256, PER C=4096, device RNG)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
agent, _ = bench.make_agent(100, "device")
for _ in range(steps):
    agent._dev.learn()
torch.cuda.synchronize()
print("done", agent._dev.launch_count())
