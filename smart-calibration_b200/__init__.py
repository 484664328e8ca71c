"""smart-calibration_b200: B200-native (sm_100a) SAC/TD3 agent update path behind the reference's Agent API.

Layout:  csrc/       hand-written CUDA kernels + the C ABI (include/smartcal_b200.h)
         _lib.py     ctypes binding of libsmartcal_b200.so (fails loudly when the library is missing)
         elasticnet/ drop-in modules named like the reference's (enet_sac, enet_td3, distributed_per_sac)
         synth.py    synthetic transitions of the elastic-net shape
"""
__version__ = "0.1.0"
