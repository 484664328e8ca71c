"""diagnostic: per-tensor gradient error of the calibration conv-SAC path vs the oracle (one step)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import test_gpu_calib as TC
use_hint = len(sys.argv) > 1 and sys.argv[1] == "hint"
M, A, B, C, seed = 3, 6, 8, 16, 21
agent, orc, tr = TC.build(M, A, B, C, seed, use_hint, 0.2)
dev = agent._dev
rs = np.random.RandomState(seed + 1)
idx = rs.choice(C, B, replace=False)
e1, e2 = rs.standard_normal((B, A)).astype(np.float32), rs.standard_normal((B, A)).astype(np.float32)
dev.set_batch(idx)
st = dev.learn(e1, e2, want_stats=True)
res = orc.learn_on_batch(orc.replay.gather(idx), e1, e2)
print("q1_pi dev", dev.buffer("q1_pi").cpu().numpy().ravel()); print("q1_pi orc", res["q1_pi"].numpy().ravel()); print("q2_pi dev", dev.buffer("q2_pi").cpu().numpy().ravel()); print("q2_pi orc", res["q2_pi"].numpy().ravel())
print("d_pi dev", dev.buffer("d_pi").cpu().numpy()[:3]); print("d_pi orc (total, incl. logp path)", res["d_pi"].numpy()[:3])
pi = dev.buffer("pi").cpu().numpy().astype(np.float64); glp = 0.03 / B
tot = dev.buffer("d_pi").cpu().numpy() + glp * 2 * pi / (1 - pi * pi + 1e-6)
d = res["d_pi"].numpy()
print("total d_pi: max abs err", np.max(np.abs(tot - d)), "max", np.max(np.abs(d)), "critic part orc - dev (row 0)", (d - glp * 2 * pi / (1 - pi * pi + 1e-6))[0], dev.buffer("d_pi").cpu().numpy()[0])
du_exp = d * (1 - pi * pi)
print("expected colsum(du) from oracle d_pi:", du_exp.sum(0), " oracle fc22mu.bias grad:", res["grad_a"]["fc22mu.bias"].numpy(), " device:", agent.actor.grads()["fc22mu.bias"].cpu().numpy().ravel())
print("device d_mu row0", dev.buffer("d_mu").cpu().numpy()[0], "expected", du_exp[0])
print("losses", st["critic_1_loss"], res["critic_1_loss"], st["actor_loss"], res["actor_loss"])
for nm, gkey in (("critic_1", "grad_c1"), ("actor", "grad_a")):
    got = getattr(agent, nm).grads()
    for k, g in res[gkey].items():
        gd = got[k].cpu().numpy().reshape(-1).astype(np.float64); gr = g.numpy().reshape(-1).astype(np.float64)
        sc = np.max(np.abs(gr))
        ratio = float(np.dot(gd, gr) / (np.dot(gr, gr) + 1e-300))
        print(f"{nm:9s} {k:18s} max|g| {sc:10.3e}  max err/scale {np.max(np.abs(gd-gr))/(sc+1e-300):9.2e}  projection ratio {ratio:.6f}")
