"""Timeline of CTA 0 of the tcgen05 dense forward kernel (stage 1 of the critic phase), SMARTCAL_STAMPS=1, plus a CUDA-event
timing of the phase groups with the tensor-core stages on and off."""
import os, sys
os.environ["SMARTCAL_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
for _ in range(5):
    dev.learn()
torch.cuda.synchronize()
tl = dev.buffer("debug_ts").view(torch.int64).cpu().numpy().reshape(-1)[2880 * 8:2880 * 8 + 192]
names = {1: "producer: issue stage", 2: "mma: stage landed", 3: "mma: issued", 4: "split: raw landed", 5: "split: lo written",
         6: "epilogue: accumulator done", 8: "epilogue: tmem read, vectors ready", 9: "epilogue: CTA stats done",
         10: "epilogue: cluster exchange done", 7: "epilogue: stored"}
evs = []
for who in range(3):
    n = int(tl[64 * who])
    evs += [(int(e) >> 8, int(e) & 255) for e in tl[64 * who + 1:64 * who + 1 + n]]
evs.sort()
print("timeline of CTA 0 of k_dense_tc, critic stage 1 (us since first event):")
if evs:
    t0 = evs[0][0]
    for t, tag in evs:
        print(f"  {(t - t0) / 1e3:8.2f}  {names.get(tag, tag)}")
M = (1 << 56) - 1
ts = dev.buffer("debug_ts").view(torch.int64).cpu().numpy() & M
def rows(lo, n):
    r = ts[lo:lo + n].astype(np.float64)
    return r[r[:, 0] > 0]
spans = []
for name, lo in [("critic dense 1", 1280), ("critic dense 2", 1408), ("actor dense 1", 1792), ("actor dense 2", 1920)]:
    r = rows(lo, 128)
    spans.append((name, r[:, 0].min(), r[:, 1].min(), r[:, 1].max(), r[:, 2].max(), r[:, 3].max(), len(r)))
t0 = spans[0][1]
print("in-situ spans (us since the first CTA of critic dense 1 entered): entry | dependency met (min..max) | accumulators done | end")
for name, e, d0, d1, acc, end, n in spans:
    print(f"  {name:16s} {n:4d} CTAs  entry {(e - t0) / 1e3:7.1f}  dep {(d0 - t0) / 1e3:7.1f}..{(d1 - t0) / 1e3:7.1f}  acc {(acc - t0) / 1e3:7.1f}  end {(end - t0) / 1e3:7.1f}")

