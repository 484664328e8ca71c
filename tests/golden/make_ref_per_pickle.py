"""Writes tests/golden/ref_prioritized_replaymem_sac.model with the UNMODIFIED reference's PER.save_checkpoint
(/root/reference/elasticnet/enet_sac.py:331-333) plus an .npz of the values a loader must reproduce.  Run in the build
container (the reference does not travel to the GPU box):   python tests/golden/make_ref_per_pickle.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/elasticnet")
import enet_sac as ref  # noqa: E402  (the reference module)

np.random.seed(5)
S, A, C = 24, 2, 16
per = ref.PER(C, [S], A)
for i in range(21):          # wraps the ring; errors given and None (max-leaf) both occur
    s, s2 = np.random.randn(S).astype(np.float32), np.random.randn(S).astype(np.float32)
    a, h = np.random.uniform(-1, 1, A).astype(np.float32), np.random.uniform(-1, 1, A).astype(np.float32)
    err = None if i % 5 == 0 else float(np.random.uniform(0, 2))
    per.store_transition_from_buffer(s, a, float(i), s2, i % 4 == 0, h, err)
os.chdir(HERE)
per.filename = "ref_prioritized_replaymem_sac.model"
per.save_checkpoint()
np.savez(os.path.join(HERE, "ref_prioritized_replaymem_sac.npz"), tree=per.tree.tree, state_memory=per.state_memory,
         reward_memory=per.reward_memory, mem_cntr=per.mem_cntr, data_pointer=per.tree.data_pointer)
print("wrote", per.filename, "mem_cntr", per.mem_cntr, "data_pointer", per.tree.data_pointer)
