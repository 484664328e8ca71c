"""Writes tests/golden/ref_prioritized_replaymem_td3.model with the UNMODIFIED reference's PER.save_checkpoint
(/root/reference/demixing_rl/demix_td3.py:142-145; small arrays: input_shape [1, 8, 8], capacity 8) plus an .npz of the values a loader
must reproduce.  Run in the build container (the reference does not travel to the GPU box):   python tests/golden/make_ref_td3_per_pickle.py"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
stub = types.ModuleType("demixingenv")            # demix_td3.py:9 imports the environment (casacore / astropy); the PER does not use it
stub.DemixingEnv = type("DemixingEnv", (), {})
sys.modules["demixingenv"] = stub
sys.path.insert(0, "/root/reference/demixing_rl")
import demix_td3 as ref  # noqa: E402  (the reference module)

np.random.seed(6)
M, A, C = 20, 5, 8
per = ref.PER(C, [1, 8, 8], A, M)
for i in range(11):          # wraps the ring
    s = {"infmap": np.random.randn(1, 8, 8).astype(np.float32), "metadata": np.random.randn(M).astype(np.float32)}
    s2 = {"infmap": np.random.randn(1, 8, 8).astype(np.float32), "metadata": np.random.randn(M).astype(np.float32)}
    a, h = np.random.uniform(-1, 1, A).astype(np.float32), np.random.uniform(-1, 1, A).astype(np.float32)
    per.store_transition(s, a, float(i), s2, i % 4 == 0, h, None if i % 3 == 0 else float(np.random.uniform(0, 2)))
per.sample_buffer(4)         # moves beta (an instance attribute from here on)
os.chdir(HERE)
per.filename = "ref_prioritized_replaymem_td3.model"
per.save_checkpoint()
np.savez(os.path.join(HERE, "ref_prioritized_replaymem_td3.npz"), tree=per.tree.tree, state_memory_img=per.state_memory_img,
         new_state_memory_sky=per.new_state_memory_sky, reward_memory=per.reward_memory, terminal_memory=per.terminal_memory,
         mem_cntr=per.mem_cntr, beta=per.beta)
print("wrote", per.filename, "mem_cntr", per.mem_cntr, "beta", per.beta)
