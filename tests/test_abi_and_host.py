"""CPU: the C-ABI library loads and exports every symbol include/smartcal_b200.h declares (no compute without a GPU),
the product path fails loudly without a device, and the host-side logic (reference-order init, packing, state-dict
key tables) behaves."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import smartcal_b200
from smartcal_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "smartcal_b200.h")).read()
    declared = set(re.findall(r"\b(sc_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"sc_agent"}
    lib = L.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert lib.sc_abi_version() == L.SC_ABI_VERSION


def test_config_struct_layout_and_arena_size():
    lib = L.load()
    cfg = L.ScConfig(abi_version=L.SC_ABI_VERSION, kind=0, state_dim=420, n_actions=2, batch_size=256, mem_size=4096, prioritized=1,
                     gamma=0.99, tau=0.005, lr_actor=1e-3, lr_critic=1e-3, reward_scale=20, alpha=0.03, max_action=1, per_clip=100)
    n = lib.sc_agent_arena_bytes(C.byref(cfg))
    # nets+Adam+targets 16.0 MB, replay 14.0 MB, workspace ~15 MB
    assert 30e6 < n < 80e6
    bad = L.ScConfig(abi_version=L.SC_ABI_VERSION, kind=0, state_dim=420, n_actions=2, batch_size=256, mem_size=1000, prioritized=1)
    assert lib.sc_agent_arena_bytes(C.byref(bad)) == 0
    assert b"power of 2" in lib.sc_last_error()       # same message as the reference assert (enet_sac.py:99)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from smartcal_b200.elasticnet import enet_sac
    with pytest.raises(L.SmartcalError, match="no CPU fallback"):
        enet_sac.Agent(gamma=0.99, lr_a=1e-3, lr_c=1e-3, input_dims=[24], batch_size=8, n_actions=2)
    lib = L.load()
    cfg = L.ScConfig(abi_version=L.SC_ABI_VERSION, kind=0, state_dim=24, n_actions=2, batch_size=8, mem_size=16)
    h = C.c_void_p()
    assert lib.sc_agent_create(C.byref(cfg), None, 0, None, C.byref(h)) == 2      # SC_ERR_CUDA
    assert b"no CPU fallback" in lib.sc_last_error()


def test_initial_weights_follow_the_reference_draw_order():
    """facade init == golden digests of the real reference's freshly constructed nets (same torch seed)"""
    from smartcal_b200.elasticnet import _host as H
    z = np.load(os.path.join(ROOT, "tests", "golden", "sac_s24_b32.npz"))
    S, A, B, Cc, steps, seed, _ = [int(x) for x in z["cfg"]]
    torch.manual_seed(seed)
    nets = [("actor", H.actor_initial_state(S, A, L.SC_KIND_SAC)), ("critic_1", H.critic_initial_state(S, A)),
            ("critic_2", H.critic_initial_state(S, A))]
    for nm, sd in nets:
        for k, v in sd.items():
            a = v.numpy().astype(np.float64).reshape(-1)
            idx = np.linspace(0, a.size - 1, min(32, a.size)).astype(np.int64)
            got = np.concatenate([[a.sum(), np.abs(a).sum()], a[idx]])
            assert np.array_equal(got, z[f"init/{nm}/{k}/p"]), f"{nm}.{k}"


def test_key_tables_and_packing():
    from smartcal_b200.elasticnet import _host as H
    from smartcal_b200.synth import make_transitions, pack
    assert len(H.critic_keys()) == 18 and len(H.actor_keys(L.SC_KIND_SAC)) == 16 and len(H.actor_keys(L.SC_KIND_TD3)) == 14
    tr = make_transitions(0, 5, 7, 2)
    p = pack(tr)
    assert p.shape == (5, 2 * 7 + 2 * 2 + 2) and p.dtype == np.float32
    assert np.array_equal(p[:, :7], tr["state"]) and np.array_equal(p[:, 14:16], tr["action"])
    obs = {"eig": torch.arange(3.0), "A": torch.arange(4.0)}
    assert np.array_equal(H.flat_state(obs), np.array([0, 1, 2, 0, 1, 2, 3], np.float32))


def test_standalone_host_replaybuffer_matches_reference_shape():
    from smartcal_b200.elasticnet.enet_sac import ReplayBuffer
    rb = ReplayBuffer(8, [5], 2)
    o = {"eig": torch.zeros(2), "A": torch.ones(3)}
    for i in range(11):
        rb.store_transition(o, np.array([i, -i], np.float32), float(i), o, i % 2 == 0, np.zeros(2, np.float32))
    assert rb.mem_cntr == 11 and rb.reward_memory[2] == 10.0 and rb.state_memory.shape == (8, 5)
    np.random.seed(0)
    s, a, r, s2, d, h = rb.sample_buffer(4)
    assert s.shape == (4, 5) and a.shape == (4, 2) and d.dtype == bool


def test_demix_td3_per_reads_reference_written_pickle(monkeypatch):
    """tests/golden/ref_prioritized_replaymem_td3.model was written by the UNMODIFIED reference's demix_td3.PER.save_checkpoint
    (tests/golden/make_ref_td3_per_pickle.py).  With the facade directory on sys.path -- the drop-in arrangement -- the pickle's
    `demix_td3.PER` / `demix_sac.SumTree` resolve to the facade classes, and the host side of PER.load_checkpoint recovers arrays, tree,
    counter and beta from their __dict__s (the device upload that follows is covered by the GPU round-trip test)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "smart-calibration_b200", "demixing_rl"))
    for m in ("demix_td3", "demix_sac"):
        sys.modules.pop(m, None)
    import demix_td3
    g = os.path.join(root, "tests", "golden")
    d = demix_td3.PER.read_checkpoint(os.path.join(g, "ref_prioritized_replaymem_td3.model"))
    z = np.load(os.path.join(g, "ref_prioritized_replaymem_td3.npz"))
    assert np.array_equal(d["tree"], z["tree"]) and d["mem_cntr"] == int(z["mem_cntr"]) == 11 and d["beta"] == float(z["beta"])
    assert np.array_equal(d["state_memory_img"], z["state_memory_img"]) and np.array_equal(d["new_state_memory_sky"], z["new_state_memory_sky"])
    assert np.array_equal(d["reward_memory"], z["reward_memory"]) and np.array_equal(d["terminal_memory"] != 0, z["terminal_memory"])
    for m in ("demix_td3", "demix_sac"):
        sys.modules.pop(m, None)


def test_bench_secondary_objects_cannot_take_the_line_down(monkeypatch):
    """bench.py runs its newest secondary object in a child process with a hard limit and skips secondary objects once its wall-time budget is
    spent: without a GPU the child fails, and the parent gets an {"error": ...} object instead of an exception; past the budget it gets
    {"skipped": ...}"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    if not torch.cuda.is_available():
        r = b.child_object("demix_td3", 2, 120)
        assert isinstance(r, dict) and "error" in r and "value" not in r
    r = b.child_object("demix_td3", 2, 0.01)
    assert "error" in r and "exceeded" in r["error"]
    assert b.over_budget("x") is None
    monkeypatch.setattr(b, "BUDGET_S", -1.0)
    assert "skipped" in b.over_budget("x")


def test_bench_conv_roofline_counts_trunk_passes():
    """bench.py's tensor roofline of the conv agents: 14 trunk-pass equivalents x 33.3 MFLOP x 256 = the 119 GFLOP of SURVEY 8d (config 3)"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(root, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    r = b.conv_roofline(14, 256, 8.0, 845.0, "note")
    assert abs(r["algorithmic_flop"] - 119e9) < 1e9 and abs(r["achieved"] - r["algorithmic_flop"] / 8e-3 / 1e12) < 1e-9 and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / 845.0) < 1e-12
