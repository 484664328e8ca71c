"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the Python facade -> C ABI,
against the CPU oracle on identical seeded inputs and identical random draws.  Tolerances are stated in
tests/parity_utils.py (fp32: activations 2e-5, gradients 2e-4 relative to the tensor max; PER tree indices bit-exact)."""
import numpy as np
import pytest
import torch

import parity_utils as PU

pytestmark = pytest.mark.gpu


def _assert(rep):
    assert rep["ok"], f"{len(rep['failures'])} mismatches, first: {rep['failures'][:8]}"
    assert rep["launches"] > 0, "no CUDA kernels were launched by libsmartcal_b200"


def test_sac_uniform_small():
    _assert(PU.sac_parity_report(S=24, A=2, B=32, C=64, steps=3, seed=7))


def test_sac_ragged_dims():
    # state_dim not a multiple of 4, batch not a multiple of the 32-row tile, 3 actions, a single done-heavy batch
    _assert(PU.sac_parity_report(S=22, A=3, B=40, C=64, steps=2, seed=17))


def test_sac_per_derived_small():
    _assert(PU.sac_parity_report(S=24, A=2, B=32, C=64, steps=3, seed=8, prioritized=True))


def test_sac_hint_with_dual_updates():
    _assert(PU.sac_parity_report(S=22, A=2, B=32, C=64, steps=12, seed=5, use_hint=True, trace=False))


def test_sac_elastic_net_full_size():
    # BASELINE.json configs[0]/[1] shapes: S=420, A=2; batch 256
    _assert(PU.sac_parity_report(S=420, A=2, B=256, C=1024, steps=2, seed=3))


def test_sac_per_elastic_net_full_size_derived():
    _assert(PU.sac_parity_report(S=420, A=2, B=256, C=4096, steps=2, seed=4, prioritized=True))


def test_td3_per_small():
    _assert(PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=4, seed=9, prioritized=True))


def test_td3_uniform_small():
    _assert(PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=2, seed=10, prioritized=False))


def test_td3_per_batch_1024():
    # BASELINE.json configs[3]: elastic-net TD3 batch=1024 prioritized=True
    _assert(PU.td3_parity_report(S=420, A=2, B=1024, C=4096, steps=2, seed=11, prioritized=True))


def test_batch_64_reference_default():
    # BASELINE.json configs[0] batch size (main_sac.py:34)
    _assert(PU.sac_parity_report(S=420, A=2, B=64, C=1024, steps=2, seed=12))


def test_td3_admm_hint_loop():
    # enet_td3.py:310-361 (five inner actor steps, adaptive penalty), admm_rho set after construction like main_td3.py:22
    _assert(PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=4, seed=13, prioritized=True, use_hint=True, admm_rho=1.0))
    _assert(PU.td3_parity_report(S=24, A=2, B=32, C=64, steps=2, seed=14, prioritized=False, use_hint=True))


def test_dense_stage_variants(monkeypatch):
    # the wide forward layers have two implementations: k_dense_tc launches (default; reduction optionally split across
    # CTA pairs) and the mma.sync stages inside k_rowphase (SMARTCAL_DENSE=mma, also what choose_action uses).  Same parity
    # bar for all of them.
    monkeypatch.setenv("SMARTCAL_DENSE", "mma")
    _assert(PU.sac_parity_report(S=24, A=2, B=32, C=64, steps=2, seed=21))
    _assert(PU.sac_parity_report(S=420, A=2, B=64, C=256, steps=2, seed=22, prioritized=True))
    monkeypatch.delenv("SMARTCAL_DENSE")
    for ks in ("1,1", "2,2"):
        monkeypatch.setenv("SMARTCAL_DENSE_KS", ks)
        _assert(PU.sac_parity_report(S=420, A=2, B=64, C=256, steps=2, seed=23, prioritized=True))
        _assert(PU.sac_parity_report(S=56, A=3, B=200, C=256, steps=2, seed=24))     # two row blocks, the second ragged
    monkeypatch.delenv("SMARTCAL_DENSE_KS")
