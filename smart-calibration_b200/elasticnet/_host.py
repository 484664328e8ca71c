"""Host-side pieces shared by the enet_sac / enet_td3 facades: reference-order weight initialisation, state-dict
views of the HBM-resident networks, and the replay facades.  No arithmetic of the update path lives here."""
import math
import os
import pickle
import sys
from collections import OrderedDict

import numpy as np
import torch as T

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
import smartcal_b200  # noqa: E402  (alias of the hyphenated package directory)
from smartcal_b200 import _lib as L  # noqa: E402
from smartcal_b200.core import DeviceAgent  # noqa: E402

H1, H2, H3, HA1, HA2 = 512, 256, 128, 128, 64


def default_rng_mode():
    """'reference': minibatch indices / PER uniforms / N(0,1) noise are drawn on the host from numpy's and
    torch's GLOBAL generators in the reference's order (same seeds -> same draws as the reference).
    'device': Philox streams inside the CUDA library, no per-step host traffic."""
    return os.environ.get("SMARTCAL_RNG", "reference")


# ---- initialisation in the reference's draw order -------------------------------------------------------
def _linear_default(out_f, in_f):
    lin = T.nn.Linear(in_f, out_f)          # consumes the CPU generator exactly like the reference constructor
    return lin.weight.data, lin.bias.data


def _uniform_fanout(w, b, sc=None):
    sc = sc or 1. / np.sqrt(w.size()[0])    # enet_sac.py:19 -- size()[0] is the OUT dimension
    T.nn.init.uniform_(w, -sc, sc)
    T.nn.init.uniform_(b, -sc, sc)


def critic_initial_state(S, A):
    """state_dict of a freshly constructed reference CriticNetwork (enet_sac.py:357-374), draw for draw."""
    sd = OrderedDict()
    dims = OrderedDict([("fc11", (H1, S)), ("fc12", (H2, H1)), ("fc21", (HA1, A)), ("fc22", (HA2, HA1)), ("fc3", (1, H2 + HA2))])
    for k, (o, i) in dims.items():
        sd[k + ".weight"], sd[k + ".bias"] = _linear_default(o, i)
    for k, n in (("bn11", H1), ("bn12", H2), ("bn21", HA1), ("bn22", HA2)):
        sd[k + ".weight"], sd[k + ".bias"] = T.ones(n), T.zeros(n)
    for k, sc in (("fc11", None), ("fc12", None), ("fc21", None), ("fc21", None), ("fc3", 0.003)):
        _uniform_fanout(sd[k + ".weight"], sd[k + ".bias"], sc)
    return sd


def actor_initial_state(S, A, kind):
    """state_dict of a fresh reference ActorNetwork (SAC enet_sac.py:415-428, TD3 enet_td3.py:86-97)."""
    heads = ["fc4mu", "fc4logsigma"] if kind == L.SC_KIND_SAC else ["fc4"]
    sd = OrderedDict()
    dims = OrderedDict([("fc1", (H1, S)), ("fc2", (H2, H1)), ("fc3", (H3, H2))] + [(h, (A, H3)) for h in heads])
    for k, (o, i) in dims.items():
        sd[k + ".weight"], sd[k + ".bias"] = _linear_default(o, i)
    for k, n in (("bn1", H1), ("bn2", H2), ("bn3", H3)):
        sd[k + ".weight"], sd[k + ".bias"] = T.ones(n), T.zeros(n)
    for k in dims:
        _uniform_fanout(sd[k + ".weight"], sd[k + ".bias"], 0.003 if k.startswith("fc4") else None)
    return sd


def critic_keys():
    return ["fc11.weight", "fc11.bias", "fc12.weight", "fc12.bias", "fc21.weight", "fc21.bias", "fc22.weight",
            "fc22.bias", "fc3.weight", "fc3.bias", "bn11.weight", "bn11.bias", "bn12.weight", "bn12.bias",
            "bn21.weight", "bn21.bias", "bn22.weight", "bn22.bias"]


def actor_keys(kind):
    heads = (["fc4mu.weight", "fc4mu.bias", "fc4logsigma.weight", "fc4logsigma.bias"] if kind == L.SC_KIND_SAC
             else ["fc4.weight", "fc4.bias"])
    return ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias"] + heads + \
           ["bn1.weight", "bn1.bias", "bn2.weight", "bn2.bias", "bn3.weight", "bn3.bias"]


def flat_state(observation):
    """cat(eig, A) of the gym observation dict (enet_sac.py:40,548) as float32 numpy."""
    if isinstance(observation, dict):
        v = T.cat((T.as_tensor(observation['eig']).reshape(-1), T.as_tensor(observation['A']).reshape(-1)))
    else:
        v = T.as_tensor(observation).reshape(-1)
    return v.detach().cpu().numpy().astype(np.float32, copy=False)


def _scalar(x):
    if isinstance(x, T.Tensor):
        return float(x.detach().cpu().reshape(-1)[0])
    return float(np.asarray(x).reshape(-1)[0])


# ---- network views ---------------------------------------------------------------------------------------
class NetView:
    """Stands where the reference keeps an nn.Module (agent.actor, agent.critic_1, ...): parameters are zero-copy
    torch views of the HBM arena under the reference state_dict key names."""

    def __init__(self, dev, net_id, keys, checkpoint_file):
        self._dev, self._net, self._keys = dev, net_id, keys
        self.checkpoint_file = checkpoint_file
        self.device = dev.device
        self.training = True

    def _t(self, name, slot=L.SLOT_PARAM):
        v = self._dev.tensor(self._net, name, slot)
        return v[0] if name.endswith(".bias") or name.startswith("bn") else v

    def named_parameters(self):
        for k in self._keys:
            yield k, self._t(k)

    def parameters(self):
        for _, p in self.named_parameters():
            yield p

    def state_dict(self):
        return OrderedDict((k, self._t(k).detach().clone()) for k in self._keys)

    def adam_state(self):
        return OrderedDict((k, (self._t(k, L.SLOT_ADAM_M).clone(), self._t(k, L.SLOT_ADAM_V).clone())) for k in self._keys)

    def load_adam_state(self, moments):
        """moments: {key: (exp_avg, exp_avg_sq)} -> torch.optim.Adam state of the reference optimizer"""
        for k, (m, v) in moments.items():
            self._t(k, L.SLOT_ADAM_M).copy_(T.as_tensor(m).to(self.device))
            self._t(k, L.SLOT_ADAM_V).copy_(T.as_tensor(v).to(self.device))

    def flat(self):
        """the whole contiguous parameter block (one NCCL broadcast moves the network)"""
        return self._dev.tensor(self._net, "*")[0]

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._keys if k not in sd]
        extra = [k for k in sd if k not in self._keys]
        if strict and (missing or extra):
            raise RuntimeError(f"Error(s) in loading state_dict: missing {missing}, unexpected {extra}")
        for k in self._keys:
            if k in sd:
                dst = self._t(k)
                src = T.as_tensor(sd[k]).to(device=dst.device, dtype=dst.dtype)
                if tuple(src.shape) != tuple(dst.shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(src.shape)} vs {tuple(dst.shape)}")
                dst.copy_(src)

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def to(self, *a, **k):
        return self

    def save_checkpoint(self):
        T.save(OrderedDict((k, v.cpu()) for k, v in self.state_dict().items()), self.checkpoint_file)

    def load_checkpoint(self):
        self.load_state_dict(T.load(self.checkpoint_file, map_location='cpu'))

    def load_checkpoint_for_eval(self):
        self.load_state_dict(T.load(self.checkpoint_file, map_location=T.device('cpu')))

    def __repr__(self):
        return f"{type(self).__name__}(sm_100a HBM-resident, keys={len(self._keys)})"


def plain_state(obj):
    """attribute dict of an unpickled replay object (ours, the reference's, or already a dict) WITHOUT going through class
    properties: a PER rebuilt by pickle has no device behind the DeviceReplayMixin properties"""
    if isinstance(obj, dict):
        return obj
    return dict(getattr(obj, "__dict__", {}))


# ---- replay facades ---------------------------------------------------------------------------------------
_ARRAYS = ("state_memory", "new_state_memory", "action_memory", "reward_memory", "terminal_memory", "hint_memory")


class DeviceReplayMixin:
    """attribute surface of reference ReplayBuffer / PER over the HBM-resident arrays"""
    _dev = None

    @property
    def mem_cntr(self):
        if self.__dict__.get("_dev") is None:          # an instance rebuilt by pickle: plain attributes
            return self.__dict__.get("mem_cntr", 0)
        return self._dev.n_stored

    @mem_cntr.setter
    def mem_cntr(self, v):
        m, l, d, b = self._dev.counters()
        self._dev.set_counters(int(v), l, d, b)

    def _arr(self, name):
        if self.__dict__.get("_dev") is None:
            return self.__dict__.get(name)
        v = self._dev.buffer(name)
        a = v.cpu().numpy()
        if name in ("reward_memory",):
            return a[:, 0].copy()
        if name == "terminal_memory":
            return a[:, 0].astype(bool)
        return a.copy()

    state_memory = property(lambda self: self._arr("state_memory"))
    new_state_memory = property(lambda self: self._arr("new_state_memory"))
    action_memory = property(lambda self: self._arr("action_memory"))
    hint_memory = property(lambda self: self._arr("hint_memory"))
    reward_memory = property(lambda self: self._arr("reward_memory"))
    terminal_memory = property(lambda self: self._arr("terminal_memory"))

    def _upload_arrays(self, src):
        src = plain_state(src)
        for name in _ARRAYS:
            a = np.asarray(src[name])
            v = self._dev.buffer(name)
            t = T.as_tensor(a.astype(np.uint8) if name == "terminal_memory" else a.astype(np.float32))
            v.copy_(t.reshape(v.shape[0], -1).to(v.device))

    def _host_state(self):
        st = {name: getattr(self, name) for name in _ARRAYS}
        st["mem_size"], st["mem_cntr"], st["filename"] = self.mem_size, self.mem_cntr, self.filename
        return st
