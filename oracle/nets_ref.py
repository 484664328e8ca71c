"""CPU oracle for the elastic-net Actor/Critic MLPs.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file.  The product path never does.

Functional (dict-of-tensors) restatement of
  elasticnet/enet_sac.py:18-21    init_layer  (U(-sc,sc), sc = 1/sqrt(weight.size(0)) i.e. OUT features)
  elasticnet/enet_sac.py:352-394  CriticNetwork  (two branches -> concat -> fc3)
  elasticnet/enet_sac.py:407-466  ActorNetwork + sample_normal (tanh-squashed Gaussian)
  elasticnet/enet_td3.py:81-112   TD3 ActorNetwork (single tanh head)
The arithmetic lives in torch (Linear = addmm, LayerNorm eps 1e-5 biased variance, ELU alpha 1),
a third-party dependency the reference does not pin; this oracle uses the torch installed here
(2.11) on CPU exactly as the reference does.

Parity status: pinned against the reference by tests/golden/sac_*.npz / td3_*.npz.
"""
import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

H1, H2, H3 = 512, 256, 128       # ref enet_sac.py:357-358,415-417
HA1, HA2 = 128, 64               # ref enet_sac.py:360-361
LOG_SQRT_2PI = math.log(math.sqrt(2 * math.pi))


def critic_shapes(S, A):
    """state_dict order of reference CriticNetwork."""
    return OrderedDict([
        ("fc11.weight", (H1, S)), ("fc11.bias", (H1,)),
        ("fc12.weight", (H2, H1)), ("fc12.bias", (H2,)),
        ("fc21.weight", (HA1, A)), ("fc21.bias", (HA1,)),
        ("fc22.weight", (HA2, HA1)), ("fc22.bias", (HA2,)),
        ("fc3.weight", (1, H2 + HA2)), ("fc3.bias", (1,)),
        ("bn11.weight", (H1,)), ("bn11.bias", (H1,)),
        ("bn12.weight", (H2,)), ("bn12.bias", (H2,)),
        ("bn21.weight", (HA1,)), ("bn21.bias", (HA1,)),
        ("bn22.weight", (HA2,)), ("bn22.bias", (HA2,)),
    ])


def actor_shapes(S, A, kind="sac"):
    heads = ([("fc4mu.weight", (A, H3)), ("fc4mu.bias", (A,)),
              ("fc4logsigma.weight", (A, H3)), ("fc4logsigma.bias", (A,))]
             if kind == "sac" else [("fc4.weight", (A, H3)), ("fc4.bias", (A,))])
    return OrderedDict([
        ("fc1.weight", (H1, S)), ("fc1.bias", (H1,)),
        ("fc2.weight", (H2, H1)), ("fc2.bias", (H2,)),
        ("fc3.weight", (H3, H2)), ("fc3.bias", (H3,)),
        *heads,
        ("bn1.weight", (H1,)), ("bn1.bias", (H1,)),
        ("bn2.weight", (H2,)), ("bn2.bias", (H2,)),
        ("bn3.weight", (H3,)), ("bn3.bias", (H3,)),
    ])


def _default_linear(out_f, in_f):
    """Consume the generator exactly like nn.Linear(in_f, out_f) construction does."""
    lin = torch.nn.Linear(in_f, out_f)
    return lin.weight.detach().clone(), lin.bias.detach().clone()


def _fanout_uniform(shape_w, shape_b, sc=None):
    sc = sc or 1.0 / math.sqrt(shape_w[0])
    w = torch.empty(shape_w).uniform_(-sc, sc)
    b = torch.empty(shape_b).uniform_(-sc, sc)
    return w, b


def init_critic(S, A):
    """Draw-for-draw equivalent of constructing reference CriticNetwork (enet_sac.py:357-374):
    five default nn.Linear inits, then init_layer on fc11, fc12, fc21, fc21 (again), fc3(0.003);
    fc22 keeps the nn.Linear default."""
    sh = critic_shapes(S, A)
    p = OrderedDict()
    for name in ("fc11", "fc12", "fc21", "fc22", "fc3"):
        o, i = sh[name + ".weight"]
        p[name + ".weight"], p[name + ".bias"] = _default_linear(o, i)
    for name, sc in (("fc11", None), ("fc12", None), ("fc21", None), ("fc21", None), ("fc3", 0.003)):
        p[name + ".weight"], p[name + ".bias"] = _fanout_uniform(sh[name + ".weight"], sh[name + ".bias"], sc)
    for name in ("bn11", "bn12", "bn21", "bn22"):
        n = sh[name + ".weight"][0]
        p[name + ".weight"], p[name + ".bias"] = torch.ones(n), torch.zeros(n)
    return OrderedDict((k, p[k]) for k in sh)


def init_actor(S, A, kind="sac"):
    """Draw-for-draw equivalent of reference ActorNetwork construction
    (enet_sac.py:415-428 / enet_td3.py:86-97)."""
    sh = actor_shapes(S, A, kind)
    lin = ["fc1", "fc2", "fc3"] + (["fc4mu", "fc4logsigma"] if kind == "sac" else ["fc4"])
    p = OrderedDict()
    for name in lin:
        o, i = sh[name + ".weight"]
        p[name + ".weight"], p[name + ".bias"] = _default_linear(o, i)
    for name in lin:
        sc = 0.003 if name.startswith("fc4") else None
        p[name + ".weight"], p[name + ".bias"] = _fanout_uniform(sh[name + ".weight"], sh[name + ".bias"], sc)
    for name in ("bn1", "bn2", "bn3"):
        n = sh[name + ".weight"][0]
        p[name + ".weight"], p[name + ".bias"] = torch.ones(n), torch.zeros(n)
    return OrderedDict((k, p[k]) for k in sh)


def _block(p, fc, bn, x):
    z = F.linear(x, p[fc + ".weight"], p[fc + ".bias"])
    z = F.layer_norm(z, (z.shape[-1],), p[bn + ".weight"], p[bn + ".bias"], 1e-5)
    return F.elu(z)


def critic_forward(p, state, action):
    """ref CriticNetwork.forward enet_sac.py:382-394."""
    x = _block(p, "fc12", "bn12", _block(p, "fc11", "bn11", state))
    y = _block(p, "fc22", "bn22", _block(p, "fc21", "bn21", action))
    return F.linear(torch.cat((x, y), dim=1), p["fc3.weight"], p["fc3.bias"])


def actor_trunk(p, state):
    return _block(p, "fc3", "bn3", _block(p, "fc2", "bn2", _block(p, "fc1", "bn1", state)))


def sac_actor_forward(p, state):
    """ref ActorNetwork.forward enet_sac.py:436-444."""
    h = actor_trunk(p, state)
    mu = F.linear(h, p["fc4mu.weight"], p["fc4mu.bias"])
    logsigma = torch.clamp(F.linear(h, p["fc4logsigma.weight"], p["fc4logsigma.bias"]), min=-20, max=2)
    return mu, logsigma


def sac_sample(p, state, eps, max_action=1.0, reparam_noise=1e-6):
    """ref ActorNetwork.sample_normal enet_sac.py:446-466 with the N(0,1) draw `eps` injected.
    Normal.sample() and .rsample() both evaluate loc + eps*scale (verified bit-equal under the
    same generator state), the only difference being whether gradients flow; autograd callers
    wrap the no-grad use in torch.no_grad() exactly as the reference does."""
    mu, logsigma = sac_actor_forward(p, state)
    sigma = logsigma.exp()
    u = mu + eps * sigma
    if u.dim() == 1:
        u = u.unsqueeze(0)
        mu, sigma = mu.unsqueeze(0), sigma.unsqueeze(0)
    t = torch.tanh(u)
    action = t * max_action
    logp = -((u - mu) ** 2) / (2 * sigma ** 2) - sigma.log() - LOG_SQRT_2PI
    logp = logp - torch.log(max_action * (1 - t.pow(2)) + reparam_noise)
    return action, logp.sum(1, keepdim=True)


def td3_actor_forward(p, state):
    """ref TD3 ActorNetwork.forward enet_td3.py:106-112."""
    return torch.tanh(F.linear(actor_trunk(p, state), p["fc4.weight"], p["fc4.bias"]))
