"""Drop-in for the reference module demixing_rl/demix_sac.py (conv SAC agent of the demixing environment, :372-719, and the
prioritized replay classes :30-308 with absolute_error_upper = 1): `from demix_sac import DemixingAgent` keeps working with
this directory on sys.path.  The agent runs on the sc_calib_* kernels of libsmartcal_b200.so (SC_CALIB_LOGSIGMA |
SC_CALIB_DEMIX_HINT | SC_CALIB_FLAT_META); see smartcal_b200/demix_common.py for the behaviour notes."""
import smartcal_b200  # noqa: F401
from smartcal_b200.demix_common import DemixActorNetwork as ActorNetwork  # noqa: F401
from smartcal_b200.demix_common import DemixCriticNetwork as CriticNetwork  # noqa: F401
from smartcal_b200.demix_common import DemixingAgentBase, DemixReplayBuffer as ReplayBuffer, init_layer  # noqa: F401
from smartcal_b200.elasticnet import enet_sac as _enet

INNER_DIM = 16


def is_power_of_2(n):
    return ((n & (n - 1)) == 0) and n != 0


SumTree = _enet.SumTree


class PER(_enet.PER):
    """demixing_rl/demix_sac.py:151-308: the elastic-net PER with the TD-error clip at 1 (absolute_error_upper = 1., :160) --
    stand-alone use (demix_td3.py binds it): HBM-resident float64 sum tree, bit-exact stratified sampling / priority updates"""
    absolute_error_upper = 1.

    def __init__(self, capacity, input_shape, n_actions, M=None, _dev=None):
        # (the reference keeps image + metadata arrays beside the tree; the tree / sampling / update semantics are what this class pins)
        shape = input_shape if len(input_shape) == 1 else [int(M or 1)]
        super().__init__(capacity, shape, n_actions, _dev=_dev)


class DemixingAgent(DemixingAgentBase):
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, M=30, reward_scale=2, alpha=0.1, hint_threshold=0.1, admm_rho=1.0, name_prefix='', use_hint=False,
                 rng=None, device=None):
        super().__init__(gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size, tau, M, reward_scale, alpha, hint_threshold,
                         admm_rho, name_prefix, use_hint, True, INNER_DIM, rng=rng, device=device)
