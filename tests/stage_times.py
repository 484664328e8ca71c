"""Stage-boundary timestamps of the critic row-phase kernel (SMARTCAL_STAMPS=1): where does a CTA spend its time?"""
import os, sys
os.environ["SMARTCAL_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, bench
agent, _ = bench.make_agent(100, "device")
dev = agent._dev
for _ in range(5):
    dev.learn()
torch.cuda.synchronize()
# the actor phase also stamps slot 0 only; run the critic phase alone by reading after a full step (slots 1..5 are critic-only)
ts = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()
ncta = int((ts[:512, 1] != 0).sum())
ts = ts[:ncta].astype(np.float64)
names = ["stage1 (5x fc1/fc11)", "stage2 (5x fc2/fc12)", "tail", "cluster barrier", "dgrad"]
t0 = ts[:, 1:].min()
print("CTAs", ncta)
for i, n in enumerate(names):
    d = (ts[:, i + 2] - ts[:, i + 1]) if i > 0 else (ts[:, 1] - ts[:, 1] * 0)
for i, n in enumerate(names):
    a, b = (ts[:, i + 1], ts[:, i + 2]) if i > 0 else (None, None)
print("per-CTA durations (us): mean / min / max")
d = (ts[:, 6] - ts[:, 0]) / 1e3
print(f"  {'gather + barrier':24s} {d.mean():8.1f} {d.min():8.1f} {d.max():8.1f}")
d = (ts[:, 1] - ts[:, 6]) / 1e3
print(f"  {'stage1 (5x fc1/fc11)':24s} {d.mean():8.1f} {d.min():8.1f} {d.max():8.1f}")
print(f"  first CTA start -> last CTA end: {(ts[:, 5].max() - ts[:, 0].min()) / 1e3:.1f} us; CTA start spread {(ts[:, 0].max() - ts[:, 0].min()) / 1e3:.1f} us")
for i, n in enumerate(names[1:], start=1):
    d = (ts[:, i + 1] - ts[:, i]) / 1e3
    print(f"  {n:24s} {d.mean():8.1f} {d.min():8.1f} {d.max():8.1f}")
t2 = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[512:512 + ncta].astype(np.float64)
print("inside stage 2: tiles | barrier1 | finalize | barrier2 (us, mean/min/max)")
prev = ts[:, 1]
for i, n in enumerate(["gemm tiles", "cluster barrier 1", "LN finalize", "cluster barrier 2"]):
    d = (t2[:, i] - prev) / 1e3
    print(f"  {n:24s} {d.mean():8.1f} {d.min():8.1f} {d.max():8.1f}")
    prev = t2[:, i]
pr = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[1024:1024 + ncta].astype(np.float64) / 1.965e3
print("thread-0 cycle accumulators over stages 1+2 (us @1.965 GHz, mean over CTAs):")
for i, n in enumerate(["wait full barrier", "mma (FFMA loop)", "park+combine (2 syncs)", "tile epilogue (stats, stores)", "LN combine", "LN combine+normalise"]):
    print(f"  {n:32s} {pr[:, i].mean():8.1f}")
print("  (stage1 starts at the actor-phase stamp of slot 0, so it is derived from the kernel time minus the rest)")

ta = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[1536:1536 + ncta].astype(np.float64)
print("actor row-phase per-CTA durations (us): mean / min / max")
for i, n in enumerate(["stage1 (fc1 + 2x fc11)", "stage2 (fc2 + 2x fc12)", "tail", "cluster barrier", "dgrad"]):
    d = (ta[:, i + 1] - ta[:, i]) / 1e3
    print(f"  {n:24s} {d.mean():8.1f} {d.min():8.1f} {d.max():8.1f}")
print(f"  first CTA start -> last CTA end: {(ta[:, 5].max() - ta[:, 0].min()) / 1e3:.1f} us; CTA start spread {(ta[:, 0].max() - ta[:, 0].min()) / 1e3:.1f} us")
print(f"  critic kernel end -> actor kernel start (wgrad in between): {(ta[:, 0].min() - ts[:, 5].max()) / 1e3:.1f} us")

pt = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[2560:2560 + ncta].astype(np.float64) / 1.965e3
print("critic tail phases, thread 0 (us, mean over CTAs):")
for i, n in enumerate(["stage rows from L2", "actor fc3 + LN + heads", "policy head", "action branches x4 (fc21, LN, fc22, LN)", "q heads",
                       "TD target, errors, cache stores", "d cat, LN bwd (12, 22)", "fc22^T, LN bwd 21, stores"]):
    print(f"  {n:40s} {pt[:, i].mean():8.1f}")

tl = dev.buffer("debug_ts").view(torch.int64).cpu().numpy()[2048:2048 + 64].reshape(-1)
n = int(tl[0]); ev = tl[1:1 + n]
names = {1: "tile start", 2: "k-loop done", 3: "combined", 4: "epilogue done", 5: "stage tiles done", 6: "barrier 1 passed", 7: "LN done", 8: "barrier 2 passed"}
print("timeline of CTA 0 / thread 0 through the dense stages (us since first event):")
t0 = ev[0] >> 8
prev = t0
for e in ev:
    t, tag = e >> 8, int(e & 255)
    print(f"  {(t - t0) / 1e3:8.2f}  (+{(t - prev) / 1e3:5.2f})  {names.get(tag, tag)}")
    prev = t


tl = dev.buffer("debug_ts").view(torch.int64).cpu().numpy().reshape(-1)[2304 * 8:2304 * 8 + 192]
names = {1: "producer: issue stage", 2: "mma: stage ready (lo done)", 3: "mma: issued", 4: "split: raw landed", 5: "split: lo written", 6: "epilogue: accumulator done",
         7: "epilogue: tile in smem", 8: "epilogue: Adam done"}
evs = []
for who in range(3):
    n = int(tl[64 * who])
    evs += [(int(e) >> 8, int(e) & 255) for e in tl[64 * who + 1:64 * who + 1 + n]]
evs.sort()
print("timeline of CTA 0 of the tcgen05 critic wgrad kernel (us since first event):")
if evs:
    t0 = evs[0][0]
    for t, tag in evs:
        print(f"  {(t - t0) / 1e3:8.2f}  {names.get(tag, tag)}")
