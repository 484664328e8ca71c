"""How the seeds of the demixing-TD3 parity cases were chosen (tests/golden/make_golden_demix_td3.py, tests/test_gpu_zz_demix_td3.py).

At batch 8 a conv critic has ~10^6 ReLU pre-activations per pass; now and then one of them that carries a large share of the
gradient lies within an ulp of zero, and the reference's own float32 gradient is then percent-level away from its float64
evaluation (seed 42, critic_2: conv1.weight 0.7 %, conv2.weight 4 %).  A second float32 implementation lands on either side of such
an edge, so those cases cannot pin anything.  This script replays the test's flow on CPU (oracle/demix_td3_ref.py, teacher-forced)
next to a float64 shadow of every step and prints the worst float32-vs-float64 gradient difference relative to the tensor's largest
gradient; the committed cases use seeds whose worst value is at rounding level.

    python tools/screen_demix_td3_seeds.py 41 42 43 ...       (CPU, ~10 s per seed and variant)"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smartcal_b200  # noqa: E402,F401
from oracle.demix_td3_ref import DemixTd3Ref  # noqa: E402
from smartcal_b200.synth import make_demix_transitions  # noqa: E402


def worst_rel(g32, g64):
    w = 0.0
    for k in g32:
        if k.startswith("conv") and k.endswith(".bias"):
            continue
        s = g64[k].abs().max().item()
        w = max(w, (g32[k].double() - g64[k]).abs().max().item() / (s + 1e-30))
    return w


def screen(seed, use_hint, flow, steps=4, M=20, A=5, B=8, C=16):
    hp = dict(gamma=0.99, batch_size=B, n_actions=A, tau=0.005, max_mem_size=C, input_dims=[1, 128, 128], lr_a=1e-3, lr_c=1e-3,
              update_actor_interval=2, warmup=4, noise=0.1, use_hint=use_hint)
    tr = make_demix_transitions(seed + 100, C, M, A, 128, 0.2)
    torch.manual_seed(seed)
    np.random.seed(seed)
    orc = DemixTd3Ref(**hp, M=M)
    for i in range(C):
        orc.store(tr["img"][i], tr["meta"][i], tr["action"][i], float(tr["reward"][i]), tr["img2"][i], tr["meta2"][i], bool(tr["done"][i]), tr["hint"][i])
    rs = np.random.RandomState(seed + 1)
    if flow == "golden":
        np.random.seed(seed + 1)
    worst = 0.0
    for step in range(steps):
        if flow == "golden":
            u01, noise = np.random.random_sample(B), np.random.normal(scale=0.2)
        else:
            u01, noise = rs.random_sample(B), float(rs.normal(scale=0.2)) * (3.0 if step == 2 else 1.0)
        sh = orc.double_shadow()
        idxs, data, w = orc.per.sample(u01)
        sh.per.sample(u01)
        batch = orc.replay.gather(data)
        r32 = orc.learn_on_batch(batch, noise, w, idxs)
        r64 = sh.learn_on_batch(batch, noise, w, idxs)
        ws = [worst_rel(r32["grad_c1"], r64["grad_c1"]), worst_rel(r32["grad_c2"], r64["grad_c2"])]
        if r32["actor_step"]:
            ws.append(worst_rel(r32["admm"][0]["grad_a"], r64["admm"][0]["grad_a"]) if use_hint else worst_rel(r32["grad_a"], r64["grad_a"]))
        worst = max(worst, *ws)
    return worst


if __name__ == "__main__":
    for seed in [int(x) for x in sys.argv[1:]] or [41, 42]:
        for flow in ("golden", "oracle-test"):
            print(f"seed {seed:4d} {flow:12s} plain {screen(seed, False, flow):.2e}   hints {screen(seed, True, flow):.2e}", flush=True)
