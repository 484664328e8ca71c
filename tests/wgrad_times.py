"""Timeline of CTA 0 of k_wgrad_tc (critic phase), SMARTCAL_STAMPS=1: producer / MMA / epilogue marks."""
import os, sys
os.environ["SMARTCAL_STAMPS"] = "1"
os.environ["SMARTCAL_GRAPH"] = "0"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
for _ in range(5):
    dev.learn()
torch.cuda.synchronize()
tl = dev.buffer("debug_ts").view(torch.int64).cpu().numpy().reshape(-1)[2304 * 8:2304 * 8 + 192]
names = {1: "producer: issue stage", 2: "mma: stage landed", 3: "mma: issued", 6: "epilogue: accumulator done", 7: "epilogue: tile in shared memory",
         8: "epilogue: Adam done"}
evs = []
for who in range(3):
    n = int(tl[64 * who])
    evs += [(int(e) >> 8, int(e) & 255, who) for e in tl[64 * who + 1:64 * who + 1 + n]]
evs.sort()
t0 = evs[0][0]
for t, tag, who in evs:
    print(f"  {(t - t0) / 1e3:8.2f}  {'    ' * who}{names.get(tag, tag)}")
