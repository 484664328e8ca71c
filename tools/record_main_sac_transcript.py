"""Config 1 of BASELINE.json as a transcript: runs the UNMODIFIED reference loop /root/reference/elasticnet/main_sac.py (argv:
--episodes 16 --steps 5 --seed 0) with the reference's own ENetEnv (enetenv.py, LBFGS solves and all) and the reference's own
Agent on CPU, and records every call the loop makes on the Agent API -- constructor kwargs, choose_action(observation) with the
returned action, store_transition(...) arguments exactly as the environment produced them (dicts of CPU tensors, 0-dim tensor
rewards), learn() -- plus the torch generator state in front of the constructor and of the first choose_action calls.
tests/test_gpu_config1.py replays the transcript against the B200 facade on the GPU box, where neither the reference nor its
environment exist.   python tools/record_main_sac_transcript.py   ->   tests/golden/main_sac_transcript.npz"""
import contextlib
import importlib.util
import io
import os
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/elasticnet"
os.environ["CUDA_VISIBLE_DEVICES"] = ""
sys.path.insert(0, os.path.join(ROOT, "tests", "stubs"))     # gymnasium stand-in
sys.path.insert(1, REF)

spec = importlib.util.spec_from_file_location("ref_enet_sac", os.path.join(REF, "enet_sac.py"))
ref = importlib.util.module_from_spec(spec)
sys.modules["ref_enet_sac"] = ref          # (main_sac.py pickles the replay buffer at episode 0: the class must be importable by name)
spec.loader.exec_module(ref)

LOG = dict(obs_eig=[], obs_A=[], action=[], rng=[], st_eig=[], st_A=[], st_action=[], st_reward=[], st_eig2=[], st_A2=[], st_done=[],
           st_hint=[], reward_is_tensor=[], learn_calls=0)
NRNG = 24


class RecordingAgent(ref.Agent):
    def __init__(self, *a, **kw):
        LOG["ctor_kwargs"] = {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in kw.items()}
        LOG["ctor_rng"] = torch.get_rng_state().numpy().copy()
        assert not a, "main_sac.py constructs the Agent with keywords only"
        super().__init__(**kw)

    def choose_action(self, observation):
        if len(LOG["rng"]) < NRNG:
            LOG["rng"].append(torch.get_rng_state().numpy().copy())
        out = super().choose_action(observation)
        assert isinstance(observation["eig"], torch.Tensor) and observation["eig"].device.type == "cpu"
        LOG["obs_eig"].append(observation["eig"].numpy().copy()); LOG["obs_A"].append(observation["A"].numpy().copy())
        LOG["action"].append(np.asarray(out).copy())
        return out

    def store_transition(self, state, action, reward, state_, terminal, hint):
        LOG["st_eig"].append(state["eig"].numpy().copy()); LOG["st_A"].append(state["A"].numpy().copy())
        LOG["st_action"].append(np.asarray(action).copy())
        LOG["reward_is_tensor"].append(isinstance(reward, torch.Tensor))
        LOG["st_reward"].append(float(reward))
        LOG["st_eig2"].append(state_["eig"].numpy().copy()); LOG["st_A2"].append(state_["A"].numpy().copy())
        LOG["st_done"].append(bool(terminal)); LOG["st_hint"].append(np.asarray(hint).copy())
        return super().store_transition(state, action, reward, state_, terminal, hint)

    def learn(self):
        LOG["learn_calls"] += 1
        return super().learn()


shim = types.ModuleType("enet_sac")
shim.__dict__.update({k: v for k, v in ref.__dict__.items() if not k.startswith("__")})
shim.Agent = RecordingAgent
sys.modules["enet_sac"] = shim

with tempfile.TemporaryDirectory() as d:
    os.chdir(d)
    sys.argv = ["main_sac.py", "--episodes", "16", "--steps", "5", "--seed", "0"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        runpy.run_path(os.path.join(REF, "main_sac.py"), run_name="__main__")
    os.chdir(ROOT)

n = len(LOG["st_reward"])
assert n == 80 and LOG["learn_calls"] == 80 and len(LOG["action"]) == 80
kw = LOG["ctor_kwargs"]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "main_sac_transcript.npz"),
                    ctor_keys=np.array(sorted(kw.keys())), ctor_vals=np.array([repr(kw[k]) for k in sorted(kw.keys())]),
                    ctor_rng=LOG["ctor_rng"], rng=np.stack(LOG["rng"]), obs_eig=np.stack(LOG["obs_eig"]), obs_A=np.stack(LOG["obs_A"]),
                    action=np.stack(LOG["action"]), st_eig=np.stack(LOG["st_eig"]), st_A=np.stack(LOG["st_A"]), st_action=np.stack(LOG["st_action"]),
                    st_reward=np.array(LOG["st_reward"], np.float64), st_eig2=np.stack(LOG["st_eig2"]), st_A2=np.stack(LOG["st_A2"]),
                    st_done=np.array(LOG["st_done"]), st_hint=np.stack(LOG["st_hint"]), reward_is_tensor=np.array(LOG["reward_is_tensor"]),
                    stdout_tail=np.array(buf.getvalue().splitlines()[-3:]))
print("recorded", n, "transitions; ctor kwargs:", kw)
print(buf.getvalue().splitlines()[-1])
