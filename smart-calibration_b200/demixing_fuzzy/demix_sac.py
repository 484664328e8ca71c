"""Drop-in for the reference module demixing_fuzzy/demix_sac.py (SAC agent that tunes the fuzzy demixing controller, :96-480):
`from demix_sac import DemixingAgent` keeps working with this directory on sys.path.  With use_influence=False (the default of
demixing_fuzzy/main_sac.py) the agent is a pure MLP -- critic (n_meta + n_actions) -> 128 -> 32 -> 1, actor n_meta -> 128 -> 32 ->
128 -> 2 x n_actions, n_actions = 24 (K-1) + 8 = 128, n_meta = 5K + 2 = 32 -- and runs on the sc_calib_* kernels of
libsmartcal_b200.so without the conv trunk (SC_CALIB_NO_TRUNK | SC_CALIB_SIDE32 | SC_CALIB_LOGSIGMA | SC_CALIB_DEMIX_HINT |
SC_CALIB_FLAT_META).  use_influence=True with INNER_DIM = 32 is not built."""
import smartcal_b200  # noqa: F401
from smartcal_b200.demix_common import DemixActorNetwork as ActorNetwork  # noqa: F401
from smartcal_b200.demix_common import DemixCriticNetwork as CriticNetwork  # noqa: F401
from smartcal_b200.demix_common import DemixingAgentBase, DemixReplayBuffer as ReplayBuffer, init_layer  # noqa: F401

INNER_DIM = 32


class DemixingAgent(DemixingAgentBase):
    def __init__(self, gamma, lr_a, lr_c, input_dims, batch_size, n_actions,
                 max_mem_size=100, tau=0.001, n_meta=30, reward_scale=2, alpha=0.1, hint_threshold=0.1, admm_rho=1.0, name_prefix='',
                 use_hint=False, use_influence=False, rng=None, device=None):
        super().__init__(gamma, lr_a, lr_c, input_dims, batch_size, n_actions, max_mem_size, tau, n_meta, reward_scale, alpha, hint_threshold,
                         admm_rho, name_prefix, use_hint, use_influence, INNER_DIM, rng=rng, device=device)
